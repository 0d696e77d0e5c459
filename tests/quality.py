"""Quality metrics for the tolerance-class paths (clustered DDS / CRN): a numpy DXTn block decoder, PSNR
as the reference's image_utils::error_metrics::compute defines it (crnlib/crn_image_utils.cpp:1048-1123:
10*log10(255^2 / mean squared error) over the selected channels) and an LZMA bitrate proxy
(crnlib/crn_dds_comp.cpp:291-303 measures the LZMA-compressed size of the DDS; the same codec settings are
applied here to both sides, so the RATIO is what is compared)."""
import lzma

import numpy as np


def _expand565(c):
    r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
    return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], -1).astype(np.int32)


def decode_color(elems):
    """elems: (n,) uint64 DXT1 elements -> (n,16,4) uint8 RGBA (3-colour index 3 = transparent black)."""
    lo = (elems & 0xFFFF).astype(np.int64); hi = ((elems >> 16) & 0xFFFF).astype(np.int64)
    c0, c1 = _expand565(lo), _expand565(hi)
    four = (lo > hi)[:, None]
    c2 = np.where(four, (c0 * 2 + c1) // 3, (c0 + c1) >> 1)
    c3 = np.where(four, (c1 * 2 + c0) // 3, 0)
    pal = np.stack([c0, c1, c2, c3], 1)                               # (n,4,3)
    sel = ((elems[:, None] >> (32 + 2 * np.arange(16, dtype=np.uint64))) & 3).astype(np.int64)
    rgb = np.take_along_axis(pal, sel[..., None].repeat(3, -1), 1)
    a = np.where((~four) & (sel == 3), 0, 255)
    return np.concatenate([rgb, a[..., None]], -1).astype(np.uint8)


def decode_alpha(elems):
    """elems: (n,) uint64 DXT5 alpha elements -> (n,16) uint8."""
    l = (elems & 0xFF).astype(np.int64); h = ((elems >> 8) & 0xFF).astype(np.int64)
    eight = l > h
    v8 = [l, h] + [(l * (7 - k) + h * k) // 7 for k in range(1, 7)]
    v6 = [l, h] + [(l * (5 - k) + h * k) // 5 for k in range(1, 5)] + [np.zeros_like(l), np.full_like(l, 255)]
    pal = np.where(eight[:, None], np.stack(v8, 1), np.stack(v6, 1))
    sel = ((elems[:, None] >> (16 + 3 * np.arange(16, dtype=np.uint64))) & 7).astype(np.int64)
    return np.take_along_axis(pal, sel, 1).astype(np.uint8)


def decode_blocks(data, fmt):
    """data: packed block bytes; fmt: dxt_format (0 DXT1, 3 DXT5, 4 DXT5A, 5 DXN_XY, 6 DXN_YX) -> (n,16,4)."""
    e = np.frombuffer(data, np.uint64)
    if fmt in (0, 1):
        return decode_color(e)
    if fmt == 3:
        out = decode_color(e[1::2])
        out[..., 3] = decode_alpha(e[0::2])
        return out
    if fmt == 4:
        a = decode_alpha(e)
        out = np.zeros((len(e), 16, 4), np.uint8); out[..., 3] = a
        return out
    x, y = decode_alpha(e[0::2]), decode_alpha(e[1::2])
    out = np.zeros((len(e) // 2, 16, 4), np.uint8)
    out[..., 0] = x if fmt == 5 else y
    out[..., 1] = y if fmt == 5 else x
    out[..., 3] = 255
    return out


def image_to_blocks(img):
    h, w = img.shape[:2]
    bh, bw = (h + 3) // 4, (w + 3) // 4
    ys = np.minimum(np.arange(bh * 4), h - 1); xs = np.minimum(np.arange(bw * 4), w - 1)
    p = img[ys][:, xs]
    return np.ascontiguousarray(p.reshape(bh, 4, bw, 4, 4).transpose(0, 2, 1, 3, 4).reshape(bh * bw, 16, 4))


def psnr(a, b, channels):
    d = a[..., channels].astype(np.float64) - b[..., channels].astype(np.float64)
    mse = float((d * d).mean())
    return 999.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


def lzma_bits(data):
    return 8 * len(lzma.compress(bytes(data), format=lzma.FORMAT_ALONE, preset=9))


def dds_payload(dds):
    """Strips the 128-byte DDS header ('DDS ' + 124-byte DDS_HEADER, no DX10 extension in crnlib's writer)."""
    assert dds[:4] == b"DDS "
    return dds[128:]
