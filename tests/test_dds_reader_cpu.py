"""CPU tests of the .dds reader / crn_decompress_dds_to_images path: the CUDA kernels under the SIMT emulator against the
unmodified reference's read_dds + get_unpacked_image(uncook) (oracle/ref_shim.cpp::ref_dds_to_images) and convert_image."""
import ctypes
import struct

import numpy as np
import pytest

import crunch2_b200 as crn
import helpers


def cc(s):
    return struct.unpack("<I", s.encode())[0]


def dds_header(w, h, levels=1, faces=1, fourcc=None, bitcount=0, masks=(0, 0, 0, 0), pf_flags=0, pitch=None, linear=True):
    hd = [0] * 32
    hd[0] = cc("DDS "); hd[1] = 124
    hd[2] = 0x1 | 0x2 | 0x4 | 0x1000 | (0x80000 if linear else 0x8)
    hd[3] = h; hd[4] = w; hd[5] = pitch or 0
    if levels > 1:
        hd[7] = levels; hd[2] |= 0x20000
    hd[19] = 32
    hd[20] = (0x4 if fourcc else 0) | pf_flags
    hd[21] = cc(fourcc) if fourcc else 0
    hd[22] = bitcount
    hd[23:27] = list(masks)
    hd[27] = 0x1000 | ((0x400000 | 0x8) if levels > 1 else 0)
    if faces == 6:
        hd[27] |= 0x8; hd[28] = 0x200 | 0xFC00
    return struct.pack("<32I", *hd)


def block_payload(rng, w, h, levels, faces, bpb, fmt):
    out = []
    for _ in range(faces):
        for l in range(levels):
            lw, lh = max(1, w >> l), max(1, h >> l)
            n = ((lw + 3) // 4) * ((lh + 3) // 4)
            b = rng.integers(0, 256, (n, bpb), dtype=np.uint8)
            if fmt == "DXT1opaque":                     # force color0 > color1: no transparent texel anywhere
                c = b[:, :4].copy().view(np.uint16).reshape(n, 2)
                hi, lo = np.maximum(c[:, 0], c[:, 1]), np.minimum(c[:, 0], c[:, 1])
                hi = np.where(hi == lo, hi + 1, hi).astype(np.uint16)
                b[:, :4] = np.stack([hi, lo], 1).view(np.uint8).reshape(n, 4)
            out.append(b.tobytes())
    return b"".join(out)


def ref_decode(ref, dds, count_px):
    out = np.zeros(count_px * 4, np.uint8)
    desc = (ctypes.c_uint32 * 5)()
    ref.ref_dds_to_images.restype = ctypes.c_int
    ok = ref.ref_dds_to_images(ctypes.c_char_p(dds), ctypes.c_uint32(len(dds)), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(out.size), desc)
    assert ok == 1
    return out, list(desc)


BLOCK_CASES = [("DXT1", 8, "DXT1", 0, 36, 20, 3, 1), ("DXT1", 8, "DXT1opaque", 0, 33, 17, 1, 1), ("DXT3", 16, "", 0, 16, 16, 2, 1), ("DXT5", 16, "", 0, 20, 12, 1, 6),
               ("DXT5", 16, "", "CCxY", 16, 16, 1, 1), ("DXT5", 16, "", "xGxR", 24, 8, 2, 1), ("DXT5", 16, "", "xGBR", 8, 8, 1, 1), ("DXT5", 16, "", "AGBR", 12, 12, 1, 1),
               ("ATI2", 16, "", "A2XY", 16, 20, 2, 1), ("ATI2", 16, "", 0, 16, 16, 1, 1), ("ATI1", 8, "", 0, 9, 7, 1, 1), ("DXT2", 16, "", 0, 8, 8, 1, 1), ("DXT4", 16, "", 0, 8, 8, 1, 1)]


@pytest.mark.parametrize("fourcc,bpb,kind,bitcount,w,h,levels,faces", BLOCK_CASES)
def test_block_dds_to_images_matches_reference(sim, ref, fourcc, bpb, kind, bitcount, w, h, levels, faces):
    rng = np.random.default_rng(abs(hash((fourcc, str(bitcount), w))) % 1000)
    dds = dds_header(w, h, levels, faces, fourcc=fourcc, bitcount=cc(bitcount) if bitcount else 0) + block_payload(rng, w, h, levels, faces, bpb, kind)
    ctx = crn.Context(0, lib=sim)
    imgs, d = ctx.dds_to_images(dds)
    total = sum(i.shape[0] * i.shape[1] for i in imgs)
    want, rd = ref_decode(ref, dds, total)
    # the shim walks faces outermost, levels inside: the same order as index level + levels * face
    got = np.concatenate([i.reshape(-1) for i in imgs])
    assert np.array_equal(got, want)
    assert [d["faces"], d["width"], d["height"], d["levels"], d["pixel_format"]] == rd
    ctx.close()


RAW_CASES = [("A8R8G8B8", 32, (0x00ff0000, 0x0000ff00, 0x000000ff, 0xff000000), 0x41), ("X8R8G8B8", 32, (0x00ff0000, 0x0000ff00, 0x000000ff, 0), 0x40),
             ("R8G8B8", 24, (0xff0000, 0x00ff00, 0x0000ff, 0), 0x40), ("R5G6B5", 16, (0xf800, 0x07e0, 0x001f, 0), 0x40), ("A1R5G5B5", 16, (0x7c00, 0x03e0, 0x001f, 0x8000), 0x41),
             ("A4R4G4B4", 16, (0x0f00, 0x00f0, 0x000f, 0xf000), 0x41), ("L8", 8, (0xff, 0, 0, 0), 0x20000), ("A8L8", 16, (0x00ff, 0, 0, 0xff00), 0x20001), ("A8", 8, (0, 0, 0, 0xff), 0x2),
             ("L8nomask", 8, (0, 0, 0, 0), 0x20000)]


@pytest.mark.parametrize("name,bits,masks,pf", RAW_CASES)
def test_raw_dds_to_images_matches_reference(sim, ref, name, bits, masks, pf):
    rng = np.random.default_rng(bits + len(name))
    w, h, levels = 19, 11, 3
    payload = b"".join(rng.integers(0, 256, max(1, w >> l) * max(1, h >> l) * (bits // 8), dtype=np.uint8).tobytes() for l in range(levels))
    dds = dds_header(w, h, levels, 1, bitcount=bits, masks=masks, pf_flags=pf)
    dds += payload
    ctx = crn.Context(0, lib=sim)
    imgs, d = ctx.dds_to_images(dds)
    want, rd = ref_decode(ref, dds, sum(i.shape[0] * i.shape[1] for i in imgs))
    assert np.array_equal(np.concatenate([i.reshape(-1) for i in imgs]), want)
    assert [d["faces"], d["width"], d["height"], d["levels"], d["pixel_format"]] == rd
    ctx.close()


def test_dds_reader_rejects_bad_files(sim):
    ctx = crn.Context(0, lib=sim)
    good = dds_header(8, 8, fourcc="DXT1") + bytes(8 * 4)
    for bad in (good[:100], b"XXXX" + good[4:], good[:128] + bytes(8), dds_header(0, 8, fourcc="DXT1") + bytes(64), dds_header(8, 8, fourcc="ETC1") + bytes(64),
                dds_header(8, 8, bitcount=12, pf_flags=0x40) + bytes(256), dds_header(8, 8, 9, fourcc="DXT1") + bytes(512)):
        with pytest.raises(crn.CrnGpuError):
            ctx.dds_to_images(bad)
    ctx.close()


@pytest.mark.parametrize("conv", range(1, 10))
def test_convert_pixels_matches_reference(sim, ref, conv):
    """All 65536 (x, y) pairs through regen_z / the YCC matrices, plus random pixels."""
    rng = np.random.default_rng(conv)
    px = rng.integers(0, 256, (300, 256, 4), dtype=np.uint8)
    g = np.arange(65536, dtype=np.uint32)
    if conv in (4,):          # From_xGxR reads (a, g)
        px[:256, :, 3] = (g >> 8).reshape(256, 256); px[:256, :, 1] = (g & 255).reshape(256, 256)
    elif conv in (9,):        # XY_to_XYZ reads (r, g)
        px[:256, :, 0] = (g >> 8).reshape(256, 256); px[:256, :, 1] = (g & 255).reshape(256, 256)
    want = px.copy()
    ref.ref_convert_image(want.ctypes.data_as(ctypes.c_void_p), 256, 300, conv - 1)
    ctx = crn.Context(0, lib=sim)
    got = px.copy()
    ctx.convert_pixels_device(got.ctypes.data, 256, 300, 1024, conv)       # emulator build: "device" memory is host memory
    ctx.synchronize()
    assert np.array_equal(got, want)
    ctx.close()
