"""crn_compress to .CRN on the device (crn_gpu_compress_crn through the C-ABI: block gather kernel, dxt_hc pipeline, host
writer) against the reference's crn_compress, plus the round trip through our own transcoder.  Tolerance class: PSNR within
0.05 dB, file size (= CRN bitrate) within 1 % of the reference (north_star)."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
from test_crn_compress_cpu import check

pytestmark = pytest.mark.gpu


def chain(w, h, seed, n=None):
    from bench import mip_chain
    c = mip_chain(blockgen.smooth_image(w, h, seed, alpha=True))
    return c if n is None else c[:n]


@pytest.mark.parametrize("name,w,h,q", [("DXT1", 256, 256, 128), ("DXT5", 256, 128, 128), ("DXN_XY", 128, 128, 200), ("DXT5A", 128, 128, 90)])
def test_crn_quality_level_matches_reference(gpu_ctx, ref, name, w, h, q):
    face_levels = [chain(w, h, 40 + w)]
    want, _, _ = helpers.ref_compress(ref, face_levels, helpers.CRN_FMT[name], file_type=0, quality=q, threads=0)
    l0 = gpu_ctx.launch_count
    got, rate, used_q = gpu_ctx.compress_crn(face_levels, helpers.CRN_FMT[name], quality_level=q)
    assert gpu_ctx.launch_count > l0 and used_q == q
    check(ref, name, face_levels, got, want)
    # our transcoder and the reference's decoder agree on our file, level by level
    tex = gpu_ctx.unpack_begin(got)
    mine = tex.unpack_all().tobytes()
    tex.close()
    assert mine == b"".join(b"".join(lv) for lv in helpers.ref_unpack_all(ref, got))


def test_crn_target_bitrate_matches_reference(gpu_ctx, ref):
    face_levels = [chain(256, 256, 77)]
    target = 1.25
    want, ref_q, ref_rate = helpers.ref_compress(ref, face_levels, helpers.CRN_FMT["DXT1"], file_type=0, bitrate=target, threads=0, want_bitrate=True)
    got, rate, q = gpu_ctx.compress_crn(face_levels, helpers.CRN_FMT["DXT1"], target_bitrate=target)
    assert abs(rate - target) <= abs(ref_rate - target) + 0.01, (rate, ref_rate, q, ref_q)
    assert abs(int(q) - int(ref_q)) <= 4, (q, ref_q)
    check(ref, "DXT1", face_levels, got, want, psnr_tol=0.1, size_tol=0.01)


def test_crn_cubemap(gpu_ctx, ref):
    faces = [chain(128, 128, 500 + f) for f in range(6)]
    want, _, _ = helpers.ref_compress(ref, faces, helpers.CRN_FMT["DXT1"], file_type=0, quality=128, threads=0)
    got, _, _ = gpu_ctx.compress_crn(faces, helpers.CRN_FMT["DXT1"], quality_level=128)
    check(ref, "DXT1", faces, got, want)
    tex = gpu_ctx.unpack_begin(got)
    mine = tex.unpack_all().tobytes()
    tex.close()
    assert mine == b"".join(b"".join(lv) for lv in helpers.ref_unpack_all(ref, got))


@pytest.mark.parametrize("fmt,w,h,faces", [("DXT5", 512, 256, 1), ("DXT1", 64, 64, 6), ("DXN_XY", 100, 60, 1)])
def test_crn_to_dds_matches_reference(gpu_ctx, ref, fmt, w, h, faces):
    """crn_decompress_crn_to_dds: byte-identical .dds (header + every level of every face)."""
    import crnsynth
    from test_dds_cpu import ref_to_dds
    data = crnsynth.synth_crn(w, h, fmt, faces=faces, seed=w)
    l0 = gpu_ctx.launch_count
    got = gpu_ctx.crn_to_dds(data)
    assert gpu_ctx.launch_count > l0
    assert got == ref_to_dds(ref, data)


@pytest.mark.parametrize("fmt,w,h", [("DXT1", 256, 256), ("DXT5", 128, 64), ("DXN_YX", 64, 64)])
def test_compress_dds_block_by_block_is_the_reference_file(gpu_ctx, ref, fmt, w, h):
    """crn_compress(cCRNFileTypeDDS) at quality 255, endpoint caching off: the whole .dds, header included, byte for byte."""
    levels = chain(w, h, 9 + w)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmt], file_type=1, quality=255, threads=0, flags=1 | 2 | 8 | 32)
    got = gpu_ctx.compress_dds([levels], helpers.CRN_FMT[fmt], quality_level=255)
    assert got == want


def test_compress_dds_clustered_within_tolerance(gpu_ctx, ref):
    import quality
    from test_qdxt_cpu import assert_within_tolerance
    levels = chain(256, 256, 31)
    want, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT["DXT5"], file_type=1, quality=128, threads=0, flags=1 | 2 | 8)
    got = gpu_ctx.compress_dds([levels], helpers.CRN_FMT["DXT5"], quality_level=128)
    assert len(got) == len(want) and got[:128] == want[:128]
    src = np.concatenate([quality.image_to_blocks(l) for l in levels])
    a, b = quality.decode_blocks(got[128:], 3), quality.decode_blocks(want[128:], 3)
    ps = [(quality.psnr(a, src, c), quality.psnr(b, src, c)) for c in ([0, 1, 2], [3])]
    assert_within_tolerance(ps, quality.lzma_bits(got[128:]), quality.lzma_bits(want[128:]))


def test_compress_mip_chain_is_the_reference_file(gpu_ctx, ref):
    """crn_compress(comp_params, mipmap_params): level 0 in, chain generated on the device, packed, .dds byte for byte."""
    from test_dds_cpu import ref_compress_mip_chain
    img = blockgen.smooth_image(256, 128, 3, alpha=True)
    want = ref_compress_mip_chain(ref, img, "DXT5", 1, 255, 1 | 2 | 8 | 32)
    got = gpu_ctx.compress_mip_chain([img], helpers.CRN_FMT["DXT5"], "dds", quality_level=255)
    assert got == want


def test_bitrate_search_reuses_state_and_buffers(gpu_ctx):
    """the trials of a target-bitrate search share the tile pass / training sets, and the buffer pool serves every trial
    from its size classes -- a second search performs no cudaMalloc at all (crn_gpu_pool_mallocs stands still)."""
    faces = [chain(256, 256, 70)]
    a = gpu_ctx.compress_crn(faces, helpers.CRN_FMT["DXT1"], target_bitrate=1.0)
    m0, l0 = gpu_ctx.pool_mallocs, gpu_ctx.launch_count
    b = gpu_ctx.compress_crn(faces, helpers.CRN_FMT["DXT1"], target_bitrate=1.0)
    assert a == b
    assert gpu_ctx.pool_mallocs == m0, "the pool allocated during a repeated search"
    assert gpu_ctx.launch_count > l0


@pytest.mark.gpu
def test_gpu_writer_orderings_at_full_palette_size(gpu_ctx, monkeypatch):
    """8192-entry palettes: the device orderings (writer_kernels.cuh) against crn_writer.h's host loops, whole file"""
    import blockgen
    from bench import mip_chain
    levels = [np.ascontiguousarray(l) for l in mip_chain(blockgen.smooth_image(1024, 1024, 77, alpha=False))]
    dev, _, _ = gpu_ctx.compress_crn([levels], 0, quality_level=255)
    monkeypatch.setenv("CRN_B200_HOST_ORDER", "1")
    host, _, _ = gpu_ctx.compress_crn([levels], 0, quality_level=255)
    monkeypatch.delenv("CRN_B200_HOST_ORDER")
    assert dev == host
