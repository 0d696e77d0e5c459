"""The swizzled DXT5 layouts (cCRNFmtDXT5_CCxY / _xGxR / _xGBR / _AGBR, inc/crnlib.h:72-83): the source pixels are "cooked"
(image_utils::convert_image, crnlib/crn_image_utils.cpp:1181-1240) before compression -- create_dds_tex (crn_dds_comp.cpp:90-104) and
crn_comp::compress_init (crn_comp.cpp:440-452) -- and compressed as DXT5 with the perceptual metric off.  Whole files against the
unmodified reference (oracle/_ref): block by block byte for byte, clustered and .crn within the stated tolerance."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
import quality
from bench import mip_chain

SWIZZLED = ["DXT5_CCxY", "DXT5_xGxR", "DXT5_xGBR", "DXT5_AGBR"]
FLAGS_EXACT = 1 | 2 | 8 | 32


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def chain(w, h, seed, n=None):
    c = mip_chain(blockgen.smooth_image(w, h, seed, alpha=True))
    return c if n is None else c[:n]


@pytest.mark.parametrize("fmt", SWIZZLED)
def test_block_by_block_dds_matches_reference_file(simctx, ref, fmt):
    levels = chain(24, 20, 31, 3)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=255, threads=0, flags=FLAGS_EXACT)
    got = simctx.compress_dds([levels], helpers.CRN_FMT[fmt], quality_level=255)
    assert got[:128] == want[:128]              # fourcc DXT5, the layout's tag in dwRGBBitCount (crn_mipmapped_texture.cpp:998-1016)
    assert got == want


def test_dxt5a_of_an_opaque_image_packs_luma(simctx, ref):
    """mip_level::pack_to_dxt (crn_mipmapped_texture.cpp:161-162): an alpha-only format from a source without alpha takes the luma"""
    levels = chain(16, 12, 5, 2)
    for l in levels:
        l[..., 3] = 255
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT["DXT5A"], file_type=1, quality=255, threads=0, flags=FLAGS_EXACT)
    got = simctx.compress_dds([levels], helpers.CRN_FMT["DXT5A"], quality_level=255)
    assert got == want
    assert len(set(got[128:])) > 4              # not the constant-255 block


def decode_dxt5(ctx, payload, levels):
    out, ofs = [], 0
    for img in levels:
        h, w = img.shape[:2]
        nb = ((w + 3) // 4) * ((h + 3) // 4) * 16
        out.append(ctx.unpack_image(3, np.frombuffer(payload[ofs:ofs + nb], np.uint8), w, h))
        ofs += nb
    return out


def cooked(ctx, ref, img, fmt):
    """the reference's own cooking of one image (ref_convert_image), the thing both encoders approximate"""
    import ctypes
    a = np.ascontiguousarray(img).copy()
    conv = {"DXT5_CCxY": 1, "DXT5_xGxR": 3, "DXT5_xGBR": 5, "DXT5_AGBR": 7}[fmt]
    assert ref.ref_convert_image(a.ctypes.data_as(ctypes.c_void_p), a.shape[1], a.shape[0], conv)
    return a


CHANNELS = {"DXT5_CCxY": ([0, 1], [3]), "DXT5_xGxR": ([1], [3]), "DXT5_xGBR": ([1, 2], [3]), "DXT5_AGBR": ([0, 1, 2], [3])}


def psnr(a, b, ch):
    d = np.concatenate([(x.astype(np.float64) - y.astype(np.float64))[..., ch].ravel() for x, y in zip(a, b)])
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("fmt", SWIZZLED)
def test_clustered_dds_within_tolerance(simctx, ref, fmt):
    levels = chain(128, 128, 77, 3)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=128, threads=0, flags=1 | 2 | 8)
    got = simctx.compress_dds([levels], helpers.CRN_FMT[fmt], quality_level=128)
    assert got[:128] == want[:128] and len(got) == len(want)
    truth = [cooked(simctx, ref, l, fmt) for l in levels]
    g, r = decode_dxt5(simctx, got[128:], levels), decode_dxt5(simctx, want[128:], levels)
    for ch in CHANNELS[fmt]:
        a, b = psnr(g, truth, ch), psnr(r, truth, ch)
        assert a >= b - 0.15, (fmt, ch, a, b)      # 20 Ktexel: the small-input allowance of test_qdxt_cpu (0.05 dB is the contract from 64 Ktexel up)
    bg, br = quality.lzma_bits(got[128:]), quality.lzma_bits(want[128:])
    assert bg <= br * 1.02, (bg, br)


@pytest.mark.parametrize("fmt", SWIZZLED)
def test_crn_within_tolerance(simctx, ref, fmt):
    levels = chain(64, 64, 9, 3)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=0, quality=128, threads=0, flags=1 | 2 | 8)
    got, _, _ = simctx.compress_crn([levels], helpers.CRN_FMT[fmt], quality_level=128)
    gi, ri = crn.texture_info(got, lib=simctx._lib), crn.texture_info(want, lib=simctx._lib)
    assert gi == ri
    truth = [cooked(simctx, ref, l, fmt) for l in levels]
    g = decode_dxt5(simctx, simctx.crn_to_dds(got)[128:], levels)
    r = decode_dxt5(simctx, simctx.crn_to_dds(want)[128:], levels)
    assert simctx.crn_to_dds(got)[:128] == simctx.crn_to_dds(want)[:128]
    for ch in CHANNELS[fmt]:
        a, b = psnr(g, truth, ch), psnr(r, truth, ch)
        assert a >= b - 0.15, (fmt, ch, a, b)
    assert len(got) <= len(want) * 1.02, (len(got), len(want))
