"""Clustered-DDS quantiser building blocks under the SIMT emulator vs the oracle port."""
import ctypes

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
import quality

P = helpers.P


class MipDesc(ctypes.Structure):
    _fields_ = [("first_block", ctypes.c_uint32), ("block_width", ctypes.c_uint32), ("block_height", ctypes.c_uint32)]


def layout_levels(levels):
    """levels: list of (h,w,4) images -> (blocks (n,16,4), mips [(first, bw, bh)])"""
    blocks, mips, first = [], [], 0
    for img in levels:
        b = quality.image_to_blocks(img)
        bh, bw = (img.shape[0] + 3) // 4, (img.shape[1] + 3) // 4
        mips.append((first, bw, bh)); first += len(b); blocks.append(b)
    return np.ascontiguousarray(np.concatenate(blocks)), mips


def port_training(port, kind, comp, blocks, mips, hierarchical=1):
    n = len(blocks); D = 2 if kind else 6
    vecs = np.zeros((n, D), np.uint8); w = np.zeros(n, np.uint32)
    nch = sum(((bw + 1) // 2) * ((bh + 1) // 2) for _, bw, bh in mips)
    enc = np.zeros(nch, np.uint8)
    m = np.array(mips, np.uint32).ravel()
    port.op_qdxt_training(kind, comp, P(blocks), n, P(m), len(mips), hierarchical, P(vecs), P(w), P(enc))
    return vecs, w, enc


def gpu_training(ctx, kind, comp, blocks, mips):
    n = len(blocks); D = 2 if kind else 6
    vecs = np.zeros((n, D), np.uint8); w = np.zeros(n, np.uint32)
    nch = sum(((bw + 1) // 2) * ((bh + 1) // 2) for _, bw, bh in mips)
    enc = np.zeros(nch, np.uint8)
    arr = (MipDesc * len(mips))(*[MipDesc(*m) for m in mips])
    ctx._check(ctx._lib.crn_gpu_qdxt_training(ctx._ctx, kind, comp, blocks.ctypes.data, n, arr, len(mips), vecs.ctypes.data, w.ctypes.data, enc.ctypes.data))
    ctx.synchronize()
    return vecs, w, enc


@pytest.fixture(scope="module", params=["fast", "exact"])
def simctx(sim, request):
    """Both vector-quantiser flavours (crn_gpu_set_vq_mode): "exact" reproduces the reference's member-order float sums, so the clustered
    output can be compared byte for byte; "fast" (the default) is held to the tolerance only."""
    ctx = crn.Context(0, lib=sim)
    ctx.set_vq_mode(request.param == "exact")
    yield ctx
    ctx.close()


@pytest.mark.parametrize("kind,comp", [(0, 3), (1, 3), (1, 0)])
def test_training_vectors_match_port(simctx, port, kind, comp):
    from bench import mip_chain
    for (w, h, seed, flat) in ((64, 64, 1, False), (40, 24, 2, False), (13, 9, 3, False), (64, 32, 4, True)):
        base = blockgen.flat_image(w, h, seed, tile=8) if flat else blockgen.smooth_image(w, h, seed, alpha=True)
        if flat:
            base[::3, ::2, :3] //= 2
        blocks, mips = layout_levels(mip_chain(base))
        a = gpu_training(simctx, kind, comp, blocks, mips)
        b = port_training(port, kind, comp, blocks, mips)
        assert (a[2] == b[2]).all(), ("encoding", kind, w, h, np.nonzero(a[2] != b[2])[0][:5])
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (kind, w, h)


# ---- the whole clustered-DDS pipeline (crn_gpu_qdxt_init / crn_gpu_qdxt_pack) vs crn_compress(cCRNFileTypeDDS) --------

GPUFMT = dict(DXT1=0, DXT5=3, DXT5A=4, DXN_XY=5, DXN_YX=6)
CHANNELS = {0: ([0, 1, 2],), 3: ([0, 1, 2], [3]), 4: ([3],), 5: ([0, 1],), 6: ([0, 1],)}


def compare_with_reference(ctx, ref, fmtname, levels, q, params=None, flags=1 | 2 | 8):
    """Returns (gpu bytes, ref bytes, [(psnr_gpu, psnr_ref) per channel group], bits_gpu, bits_ref)."""
    dds, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[fmtname], file_type=1, quality=q, threads=0, flags=flags)
    ref_data = np.frombuffer(quality.dds_payload(dds), np.uint8)
    qd = ctx.qdxt_init(GPUFMT[fmtname], levels, params)
    out = qd.pack(q)
    info = qd.info()
    qd.close()
    assert len(out) == len(ref_data)
    src = np.concatenate([quality.image_to_blocks(l) for l in levels])
    f = GPUFMT[fmtname]
    a = quality.decode_blocks(out.tobytes(), f); b = quality.decode_blocks(ref_data.tobytes(), f)
    ps = [(quality.psnr(a, src, c), quality.psnr(b, src, c)) for c in CHANNELS[f]]
    return out, ref_data, ps, quality.lzma_bits(out.tobytes()), quality.lzma_bits(ref_data.tobytes()), info


def assert_within_tolerance(ps, bits_gpu, bits_ref, texels=None):
    """BASELINE.json north_star: RGB/alpha PSNR within 0.05 dB and bitrate within 1 % of the reference.  That bound is stated for the
    benchmark configurations (>= 1 Mtexel).  Below 64 Ktexel a texture has a few dozen endpoint clusters per element and a few KB of LZMA
    output, and ANY change of summation order moves the figures by more than that (the reference's own result moves as much between
    thread counts), so tiny test inputs are held to 0.15 dB / 2 % and the contract's bound applies from 256 x 256 up (texels = None: contract)."""
    small = texels is not None and texels < 65536
    for g, r in ps:
        assert abs(g - r) <= (0.15 if small else 0.05), ps
    assert abs(bits_gpu - bits_ref) <= (0.02 if small else 0.01) * bits_ref, (bits_gpu, bits_ref)


@pytest.mark.parametrize("fmtname,w,h,q,seed", [
    ("DXT1", 64, 64, 128, 1),
    ("DXT5A", 64, 64, 128, 4),
    ("DXN_XY", 64, 64, 128, 5),
    ("DXN_YX", 40, 24, 200, 6),
    ("DXT1", 128, 128, 60, 3),
])
def test_clustered_dds_matches_reference_bytes(simctx, ref, fmtname, w, h, q, seed):
    """With one thread the reference is deterministic; on these inputs its cross-cluster endpoint cache
    (crn_dxt1.cpp:748-763, a thread-schedule dependent heuristic this path does not have) never changes a result,
    and the output is byte-identical."""
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(w, h, seed, alpha=True))
    out, ref_data, ps, bg, br, info = compare_with_reference(simctx, ref, fmtname, levels, q)
    assert_within_tolerance(ps, bg, br, None if simctx.vq_exact else w * h)
    if simctx.vq_exact:
        assert np.array_equal(out, ref_data)
    assert all(k > 0 for k in info["endpoint_clusters"])


def test_clustered_dds_dxt5_within_tolerance(simctx, ref):
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(64, 64, 2, alpha=True))
    out, ref_data, ps, bg, br, info = compare_with_reference(simctx, ref, "DXT5", levels, 128)
    assert_within_tolerance(ps, bg, br, None if simctx.vq_exact else 64 * 64)
    if simctx.vq_exact:
        assert np.array_equal(out.view(np.uint64)[0::2], ref_data.view(np.uint64)[0::2])      # alpha elements identical
    assert info["num_elements"] == 2


def test_clustered_dds_256_fast_mode_meets_the_contract(sim, ref):
    """The default (single-launch) quantiser at the smallest size where the contract's bound is meaningful: 256 x 256 DXT5 + mips."""
    from bench import mip_chain
    ctx = crn.Context(0, lib=sim)
    levels = mip_chain(blockgen.smooth_image(256, 256, 8, alpha=True))
    out, ref_data, ps, bg, br, info = compare_with_reference(ctx, ref, "DXT5", levels, 128)
    assert_within_tolerance(ps, bg, br)
    ctx.close()


def test_clustered_dds_single_level_and_repack(simctx, ref):
    """No mip chain, odd size; pack() twice on one state at two quality levels (crnlib's bitrate search does this)."""
    img = blockgen.smooth_image(52, 36, 9, alpha=True)
    qd = simctx.qdxt_init(GPUFMT["DXT1"], [img])
    for q in (40, 220):
        out = qd.pack(q)
        dds, _, _ = helpers.ref_compress(ref, [[img]], helpers.CRN_FMT["DXT1"], file_type=1, quality=q, threads=0)
        ref_data = np.frombuffer(quality.dds_payload(dds), np.uint8)
        src = quality.image_to_blocks(img)
        a = quality.decode_blocks(out.tobytes(), 0); b = quality.decode_blocks(ref_data.tobytes(), 0)
        assert_within_tolerance([(quality.psnr(a, src, [0, 1, 2]), quality.psnr(b, src, [0, 1, 2]))], quality.lzma_bits(out.tobytes()), quality.lzma_bits(ref_data.tobytes()),
                                None if simctx.vq_exact else 52 * 36)
    qd.close()


def test_clustered_dds_rejects_unsupported(simctx):
    img = blockgen.smooth_image(16, 16, 1, alpha=True)
    with pytest.raises(crn.CrnGpuError):
        simctx.qdxt_init(2, [img])                                   # DXT3 is never clustered
    with pytest.raises(crn.CrnGpuError):
        simctx.qdxt_init(0, [img], crn.PackParams(dxt_quality=7))    # crn_dxt_quality is 0 .. 4


@pytest.mark.parametrize("dxt_quality", [0, 2])
def test_clustered_dds_lower_dxt_quality(sim, ref, dxt_quality):
    """m_dxt_quality below better reaches the per-cluster optimiser through qdxt1_params (crn_qdxt1.cpp:556-575): evaluate_solution_fast"""
    ctx = crn.Context(0, lib=sim)
    ctx.set_vq_mode(True)
    img = blockgen.smooth_image(96, 96, 21, alpha=False)
    qd = ctx.qdxt_init(GPUFMT["DXT1"], [img], crn.PackParams(dxt_quality=dxt_quality))
    out = qd.pack(128)
    qd.close()
    dds, _, _ = helpers.ref_compress(ref, [[img]], helpers.CRN_FMT["DXT1"], file_type=1, quality=128, threads=0, dxt_quality=dxt_quality)
    ref_data = np.frombuffer(quality.dds_payload(dds), np.uint8)
    src = quality.image_to_blocks(img)
    a = quality.decode_blocks(out.tobytes(), 0); b = quality.decode_blocks(ref_data.tobytes(), 0)
    assert_within_tolerance([(quality.psnr(a, src, [0, 1, 2]), quality.psnr(b, src, [0, 1, 2]))], quality.lzma_bits(out.tobytes()), quality.lzma_bits(ref_data.tobytes()), 96 * 96)
    ctx.close()
