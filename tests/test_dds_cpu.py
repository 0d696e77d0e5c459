"""DDS container edge: crn_gpu_dds_header and crn_gpu_crn_to_dds (crn_decompress_crn_to_dds) against the reference's
crn_decompress_crn_to_dds (oracle/_ref), byte for byte; the transcode runs under the SIMT emulator here."""
import ctypes

import numpy as np
import pytest

import crnsynth
import crunch2_b200 as crn
import helpers


@pytest.fixture(scope="module", params=["fast", "exact"])
def simctx(sim, request):
    """Both vector-quantiser flavours (crn_gpu_set_vq_mode): "exact" reproduces the reference's member-order float sums, so the clustered
    output can be compared byte for byte; "fast" (the default) is held to the tolerance only."""
    ctx = crn.Context(0, lib=sim)
    ctx.set_vq_mode(request.param == "exact")
    yield ctx
    ctx.close()


def ref_to_dds(ref, data):
    buf = np.frombuffer(data, np.uint8)
    size = ctypes.c_uint32()
    p = ref.ref_crn_to_dds(helpers.P(buf), len(data), ctypes.byref(size))
    assert p
    d = ctypes.string_at(p, size.value)
    ref.ref_free(ctypes.c_void_p(p))
    return d


CASES = [(fmt, w, h, lv, faces) for fmt in ("DXT1", "DXT5", "DXN_XY", "DXN_YX", "DXT5A") for (w, h, lv, faces) in ((64, 32, None, 1), (20, 12, 1, 1), (16, 16, 3, 6), (8, 8, 1, 6))]


@pytest.mark.parametrize("fmt,w,h,lv,faces", CASES)
def test_crn_to_dds_matches_reference(simctx, ref, fmt, w, h, lv, faces):
    data = crnsynth.synth_crn(w, h, fmt, levels=lv, faces=faces, seed=w + h, n_color_ep=16, n_color_sel=16, n_alpha_ep=16, n_alpha_sel=16)
    want = ref_to_dds(ref, data)
    got = simctx.crn_to_dds(data)
    assert got[:128] == want[:128]
    assert got == want


def test_header_arguments(sim):
    out = (ctypes.c_uint8 * 128)()
    assert sim.crn_gpu_dds_header(0, 64, 64, 7, 1, out) == 0
    assert bytes(out[:4]) == b"DDS "
    assert sim.crn_gpu_dds_header(10, 64, 64, 7, 1, out) != 0     # ETC1: not a format of this path
    assert sim.crn_gpu_dds_header(0, 0, 64, 7, 1, out) != 0
    assert sim.crn_gpu_dds_header(0, 64, 64, 1, 2, out) != 0


def test_crn_to_dds_rejects_garbage(simctx):
    with pytest.raises(crn.CrnGpuError):
        simctx.crn_to_dds(b"not a crn file at all, just some bytes to get past any minimum size check........................")


# ---- crn_compress(cCRNFileTypeDDS): whole files against the reference's ------------------------------------------
FLAGS_EXACT = 1 | 2 | 8 | 32      # perceptual | hierarchical | both block types | endpoint caching off (the thread-independent mode)


def chain(w, h, seed, n=None):
    import blockgen
    from bench import mip_chain
    c = mip_chain(blockgen.smooth_image(w, h, seed, alpha=True))
    return c if n is None else c[:n]


@pytest.mark.parametrize("fmt,w,h,n", [("DXT1", 32, 16, None), ("DXT5", 16, 16, 2), ("DXT3", 16, 8, 1), ("DXN_XY", 12, 20, 2), ("DXT5A", 16, 16, 3)])
def test_compress_dds_block_by_block_matches_reference_file(simctx, ref, fmt, w, h, n):
    levels = chain(w, h, 60 + w, n)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=255, threads=0, flags=FLAGS_EXACT)
    got = simctx.compress_dds([levels], helpers.CRN_FMT[fmt], quality_level=255)
    assert got[:128] == want[:128]
    assert got == want


def test_compress_dds_cubemap_and_dxt1a(simctx, ref):
    faces = [chain(8, 8, 700 + f, 2) for f in range(6)]
    for f in faces:
        f[0][0:4, 0:4, 3] = 0                  # transparent texels: DXT1 becomes DXT1A under cCRNCompFlagDXT1AForTransparency (128)
    want, _, _ = helpers.ref_compress(ref, faces, helpers.CRN_FMT["DXT1"], file_type=1, quality=255, threads=0, flags=FLAGS_EXACT | 128)
    got = simctx.compress_dds(faces, helpers.CRN_FMT["DXT1"], quality_level=255, dxt1a_for_transparency=True)
    assert got == want


def test_compress_dds_clustered_matches_reference_file(simctx, ref):
    """Clustered path: on this input the single-thread reference is reproduced byte for byte (as in test_qdxt_cpu)."""
    levels = chain(64, 64, 1)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT["DXT1"], file_type=1, quality=128, threads=0, flags=1 | 2 | 8)
    got = simctx.compress_dds([levels], helpers.CRN_FMT["DXT1"], quality_level=128)
    assert len(got) == len(want) and got[:128] == want[:128]
    if simctx.vq_exact:
        assert got == want


# ---- crn_compress with a crn_mipmap_params: level 0 in, generated chain, whole file out ----------------------------
def ref_compress_mip_chain(ref, img, fmt, file_type, quality, flags):
    ref.ref_compress_mip_chain.restype = ctypes.c_void_p
    h, w = img.shape[:2]
    size = ctypes.c_uint32()
    img = np.ascontiguousarray(img)
    p = ref.ref_compress_mip_chain(file_type, helpers.CRN_FMT[fmt], w, h, img.ctypes.data_as(ctypes.c_void_p), flags, quality, 4, 0, ctypes.byref(size))
    assert p
    d = ctypes.string_at(p, size.value)
    ref.ref_free(ctypes.c_void_p(p))
    return d


@pytest.mark.parametrize("fmt,w,h,alpha", [("DXT1", 32, 16, False), ("DXT5", 24, 24, True), ("DXT5", 16, 16, False), ("DXT5A", 20, 12, True)])
def test_compress_mip_chain_dds_matches_reference_file(simctx, ref, fmt, w, h, alpha):
    """crn_compress(comp_params, mipmap_params) with the default crn_mipmap_params: generated chain + packer, byte for byte."""
    import blockgen
    img = blockgen.smooth_image(w, h, 90 + w, alpha=True)
    if not alpha:
        img[..., 3] = 255                     # no alpha anywhere: the reference filters three components and writes 255
    want = ref_compress_mip_chain(ref, img, fmt, 1, 255, FLAGS_EXACT)
    got = simctx.compress_mip_chain([img], helpers.CRN_FMT[fmt], "dds", quality_level=255)
    assert got[:128] == want[:128]
    assert got == want


def test_compress_mip_chain_crn_within_tolerance(simctx, ref):
    import blockgen
    from test_crn_compress_cpu import check
    img = blockgen.smooth_image(64, 64, 8, alpha=True)
    want = ref_compress_mip_chain(ref, img, "DXT1", 0, 128, 1 | 2 | 8)
    got = simctx.compress_mip_chain([img], helpers.CRN_FMT["DXT1"], "crn", quality_level=128)
    levels = simctx.generate_mipmaps(img)
    check(ref, "DXT1", [levels], got, want)
    assert crn.texture_info(got, lib=simctx._lib)["levels"] == 7


def test_compress_dds_without_adaptive_tiles(simctx, ref):
    """cCRNCompFlagHierarchical cleared: qdxt1 / qdxt5 take one training vector per block (crn_qdxt1.cpp:370-403, crn_qdxt5.cpp:346-382)."""
    import quality
    levels = chain(64, 64, 3, 3)
    for fmt in ("DXT1", "DXT5"):
        want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=100, threads=0, flags=1 | 8)
        hier, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=100, threads=0, flags=1 | 2 | 8)
        got = simctx.compress_dds([levels], helpers.CRN_FMT[fmt], quality_level=100, hierarchical=False)
        assert want != hier                          # the flag matters on this input
        assert len(got) == len(want) and got[:128] == want[:128]
        if simctx.vq_exact:
            assert got == want
        else:
            gf = 3 if fmt == "DXT5" else 0
            g = [simctx.unpack_image(gf, np.frombuffer(x, np.uint8)[128:128 + 16 * 16 * (16 if gf else 8)], 64, 64) for x in (got, want)]
            for ch in ([0, 1, 2],) + (([3],) if gf else ()):
                a, b = quality.psnr(g[0], levels[0], ch), quality.psnr(g[1], levels[0], ch)
                assert a >= b - 0.25, (fmt, ch, a, b)          # 4 Ktexel level: rounding noise of the fast VQ dominates at this size


def test_compress_dds_bitrate_search_and_retry(simctx, ref):
    """create_compressed_texture's search on a .dds (crn_texture_comp.cpp:120-262): a reachable target, and one above what quality 255
    gives, where the reference clears cCRNCompFlagHierarchical and searches a second time (:232-250)."""
    if not simctx.vq_exact:
        pytest.skip("compared file against file: exact mode only (the fast VQ is covered by the tolerance tests)")
    levels = chain(32, 32, 11, 2)
    for target in (2.0, 7.5):
        want, ref_q, ref_rate = helpers.ref_compress(ref, [levels], helpers.CRN_FMT["DXT1"], file_type=1, bitrate=target, threads=0, want_bitrate=True)
        got, rate, q = simctx.compress_dds([levels], helpers.CRN_FMT["DXT1"], target_bitrate=target)
        assert q == ref_q, (target, q, ref_q)
        assert abs(rate - ref_rate) <= 0.02 * ref_rate, (target, rate, ref_rate)      # liblzma vs the reference's own LZMA coder
        assert got == want, target
