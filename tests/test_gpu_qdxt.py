"""GPU parity of the clustered-DDS path through the C-ABI: the vector quantiser against the reference's
clusterizer / threaded_clusterizer, and crn_gpu_qdxt_init / crn_gpu_qdxt_pack against crn_compress(cCRNFileTypeDDS)
(oracle/_ref) within the stated tolerance (PSNR 0.05 dB, LZMA size 1 %)."""
import numpy as np
import pytest
import torch

import blockgen
import crunch2_b200 as crn
import helpers
import quality
from test_qdxt_cpu import GPUFMT, assert_within_tolerance, compare_with_reference
from test_vq_cpu import agreement, make_vectors, ref_clusterize

pytestmark = pytest.mark.gpu


@pytest.fixture()
def exact_ctx(gpu_ctx):
    """the member-order vector quantiser (crn_gpu_set_vq_mode(1)) for the tests that compare cluster assignments with the reference's"""
    gpu_ctx.set_vq_mode(True)
    yield gpu_ctx
    gpu_ctx.set_vq_mode(False)


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed", [
    (6, 2000, 65535, 200, False, 2), (2, 1500, 65535, 100, False, 3), (16, 2000, 300, 0, True, 6),
    (6, 60000, 65535, 6000, False, 21), (2, 80000, 65535, 2000, False, 22), (16, 50000, 4000, 0, True, 23), (6, 400000, 65535, 25000, False, 24),
])
def test_gpu_vq_fast_within_tolerance(gpu_ctx, ref, dims, n, max_size, retrieve, threaded, seed):
    """default builder (vq_fast.cuh: warp / CTA / thread-block-cluster per node, one launch per frontier): same cluster count (+- 1 %),
    quantisation error within 1 % of the reference's, deterministic from run to run"""
    from test_vq_cpu import distortion
    vecs, w = make_vectors(dims, n, seed, max_weight=2048 if dims == 16 else 8)
    co_r, k_r, _ = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    l0 = gpu_ctx.launch_count
    co_g, k_g, _ = gpu_ctx.vq_clusterize(torch.from_numpy(vecs).cuda(), torch.from_numpy(w.view(np.int32)).cuda(), n, dims, max_size, retrieve, threaded)
    launches = gpu_ctx.launch_count - l0
    assert abs(k_g - k_r) <= max(1, k_r // 100)
    assert distortion(vecs, w, co_g) <= distortion(vecs, w, co_r) * (1.02 if dims == 16 else 1.01)      # 16-D uniform noise has no structure: near-ties everywhere
    assert launches <= 200, launches
    co_g2, _, _ = gpu_ctx.vq_clusterize(torch.from_numpy(vecs).cuda(), torch.from_numpy(w.view(np.int32)).cuda(), n, dims, max_size, retrieve, threaded)
    assert np.array_equal(co_g, co_g2)


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed", [
    (6, 2000, 65535, 200, False, 2),
    (2, 1500, 65535, 100, False, 3),
    (16, 2000, 300, 0, True, 6),
])
def test_gpu_vq_exact_small(exact_ctx, ref, dims, n, max_size, retrieve, threaded, seed):
    gpu_ctx = exact_ctx
    vecs, w = make_vectors(dims, n, seed)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    co_g, k_g, cb_g = gpu_ctx.vq_clusterize(torch.from_numpy(vecs).cuda(), torch.from_numpy(w.view(np.int32)).cuda(), n, dims, max_size, retrieve, threaded)
    assert k_g == k_r and (cb_r is None or cb_g == cb_r)
    assert np.array_equal(co_g, co_r)


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed", [
    (6, 60000, 65535, 6000, False, 21),
    (2, 80000, 65535, 2000, False, 22),
    (16, 50000, 4000, 0, True, 23),
])
def test_gpu_vq_large_exact(exact_ctx, ref, dims, n, max_size, retrieve, threaded, seed):
    gpu_ctx = exact_ctx
    vecs, w = make_vectors(dims, n, seed, max_weight=2048 if dims == 16 else 8)
    co_r, k_r, _ = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    co_g, k_g, _ = gpu_ctx.vq_clusterize(torch.from_numpy(vecs).cuda(), torch.from_numpy(w.view(np.int32)).cuda(), n, dims, max_size, retrieve, threaded)
    assert k_g == k_r
    assert np.array_equal(co_g, co_r), agreement(co_g, co_r)
    # run-to-run determinism
    co_g2, _, _ = gpu_ctx.vq_clusterize(torch.from_numpy(vecs).cuda(), torch.from_numpy(w.view(np.int32)).cuda(), n, dims, max_size, retrieve, threaded)
    assert np.array_equal(co_g, co_g2)


@pytest.mark.parametrize("fmtname,w,h,q,seed", [
    ("DXT1", 256, 256, 128, 1),
    ("DXT5", 256, 256, 128, 2),
    ("DXT5A", 256, 128, 64, 3),
    ("DXN_XY", 200, 120, 200, 4),
    ("DXT1", 512, 512, 30, 5),
])
def test_gpu_clustered_dds_within_tolerance(gpu_ctx, ref, fmtname, w, h, q, seed):
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(w, h, seed, alpha=True))
    out, ref_data, ps, bg, br, info = compare_with_reference(gpu_ctx, ref, fmtname, levels, q)
    assert_within_tolerance(ps, bg, br, w * h)


def test_gpu_clustered_dds_device_pixels_and_determinism(gpu_ctx):
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(256, 256, 7, alpha=True))
    dev = [torch.from_numpy(np.ascontiguousarray(l)).cuda() for l in levels]
    qa = gpu_ctx.qdxt_init(GPUFMT["DXT5"], dev); a = qa.pack(128); qa.close()
    qb = gpu_ctx.qdxt_init(GPUFMT["DXT5"], levels); b = qb.pack(128)
    assert np.array_equal(a, b)
    out = torch.zeros(qb.size, dtype=torch.uint8, device="cuda")
    qb.pack(128, out=out)
    assert np.array_equal(out.cpu().numpy(), b)
    qb.close()
