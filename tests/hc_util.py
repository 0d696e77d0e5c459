"""Helpers of the dxt_hc tests: crn_comp's block layout, the reference's dxt_hc::compress through the shim, and a decoder of
dxt_hc's outputs (palettes + indices) back to pixels for the tolerance comparison."""
import ctypes

import numpy as np

import helpers
import quality

P = helpers.P


def hc_layout(face_levels):
    """face_levels[face][level] = (h, w, 4) uint8.  Blocks level-major then face-major, every level padded to multiples of
    8 pixels by edge clamp, weight = min(12, 1.3^level) (crnlib/crn_comp.cpp:706-741).  -> blocks (n,16,4), levels."""
    faces, nlev = len(face_levels), len(face_levels[0])
    blocks, levels, first = [], [], 0
    for l in range(nlev):
        nb = 0
        for f in range(faces):
            img = face_levels[f][l]
            h, w = img.shape[:2]
            ph, pw = (h + 7) & ~7, (w + 7) & ~7
            ys = np.minimum(np.arange(ph), h - 1); xs = np.minimum(np.arange(pw), w - 1)
            b = quality.image_to_blocks(np.ascontiguousarray(img[ys][:, xs]))
            blocks.append(b); nb += len(b)
        bw = ((face_levels[0][l].shape[1] + 7) & ~7) >> 2
        levels.append((first, nb, bw, float(np.float32(min(12.0, np.float32(1.3) ** np.float32(l))))))
        first += nb
    return np.ascontiguousarray(np.concatenate(blocks)), levels


def ref_hc_compress(ref, fmt, blocks, levels, num_faces=1, perceptual=True, codebook_sizes=(3072, 3072, 3072, 3072), deratings=(2.0, 2.0, 3.0),
                    alpha_components=(3, 0), threads=0):
    n = len(blocks)
    lv = np.zeros((len(levels), 4), np.uint32)
    for i, (first, nb, bw, weight) in enumerate(levels):
        lv[i, :3] = (first, nb, bw)
        lv[i, 3] = np.float32(weight).view(np.uint32)
    cs = np.array(codebook_sizes, np.uint32); de = np.array(deratings, np.float32); ac = np.array(alpha_components, np.uint32)
    ei = np.zeros((n, 4), np.uint16); si = np.zeros((n, 4), np.uint16)
    ce = np.zeros(65536, np.uint32); ae = np.zeros(65536, np.uint32); csel = np.zeros(65536, np.uint32); asel = np.zeros(65536, np.uint64)
    sizes = np.zeros(4, np.uint32); enc = np.zeros(n, np.uint8); ti = np.zeros(n, np.uint32)
    ok = ref.ref_hc_compress(int(fmt), n, len(levels), num_faces, P(lv), int(perceptual), P(cs), P(de), P(ac), threads, P(np.ascontiguousarray(blocks)),
                             P(ei), P(si), P(ce), P(ae), P(csel), P(asel), P(sizes), P(enc), P(ti))
    assert ok
    return {"endpoint_indices": ei, "selector_indices": si, "color_endpoints": ce[:sizes[0]].copy(), "alpha_endpoints": ae[:sizes[1]].copy(),
            "color_selectors": csel[:sizes[2]].copy(), "alpha_selectors": asel[:sizes[3]].copy(), "block_encodings": enc, "tile_indices": ti}


def _expand565(c):
    r, g, b = (c >> 11) & 31, (c >> 5) & 63, c & 31
    return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2)], -1).astype(np.int64)


def hc_decode(fmt, out, alpha_components=(3, 0)):
    """Palettes + indices -> (n,16,4) uint8 pixels (channels the format does not carry stay 0 / alpha 255)."""
    ei, si = out["endpoint_indices"].astype(np.int64), out["selector_indices"].astype(np.int64)
    n = len(ei)
    px = np.zeros((n, 16, 4), np.uint8); px[..., 3] = 255
    if fmt in (0, 3):
        ep = out["color_endpoints"][ei[:, 0]].astype(np.int64)
        c0, c1 = _expand565(ep & 0xFFFF), _expand565(ep >> 16)
        pal = np.stack([c0, (c0 * 2 + c1) // 3, (c1 * 2 + c0) // 3, c1], 1)                  # linear selector order
        sel = (out["color_selectors"][si[:, 0]].astype(np.int64)[:, None] >> (2 * np.arange(16))) & 3
        px[..., :3] = np.take_along_axis(pal, sel[..., None].repeat(3, -1), 1)
    na = 1 if fmt in (3, 4) else (2 if fmt in (5, 6) else 0)
    for a in range(na):
        ep = out["alpha_endpoints"][ei[:, 1 + a]].astype(np.int64)
        l, h = ep & 0xFF, (ep >> 8) & 0xFF
        v8 = [l, h] + [(l * (7 - k) + h * k) // 7 for k in range(1, 7)]
        v6 = [l, h] + [(l * (5 - k) + h * k) // 5 for k in range(1, 5)] + [np.zeros_like(l), np.full_like(l, 255)]
        bv = np.where((l > h)[:, None], np.stack(v8, 1), np.stack(v6, 1))
        lin = bv[:, [0, 2, 3, 4, 5, 6, 7, 1]]                                                # g_dxt5_from_linear
        sel = ((out["alpha_selectors"][si[:, 1 + a]][:, None] >> (3 * np.arange(16, dtype=np.uint64))) & np.uint64(7)).astype(np.int64)
        px[..., alpha_components[a]] = np.take_along_axis(lin, sel, 1)
    return px


def index_entropy_bits(out, fmt):
    """Zeroth-order entropy (bits) of the index streams the CRN writer codes: a bitrate proxy that needs no writer."""
    def H(v):
        _, c = np.unique(v, return_counts=True)
        p = c / c.sum()
        return float(-(c * np.log2(p)).sum())
    comps = ([0] if fmt in (0, 3) else []) + ([1] if fmt in (3, 4) else []) + ([1, 2] if fmt in (5, 6) else [])
    bits = H(out["endpoint_indices"][:, 3])
    for c in comps:
        e = out["endpoint_indices"][:, c].astype(np.int64)
        bits += H(np.diff(e, prepend=0)[out["endpoint_indices"][:, 3] == 0]) + H(out["selector_indices"][:, c])
    bits += 32 * len(out["color_endpoints"]) + 16 * len(out["alpha_endpoints"]) + 32 * len(out["color_selectors"]) + 48 * len(out["alpha_selectors"])
    return bits
