"""dxt_hc pipeline on the device against the reference's dxt_hc::compress (oracle/_ref), through the C ABI."""
import numpy as np
import pytest

import blockgen
import hc_util
from test_hc_cpu import FMTS, assert_tolerance, compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(FMTS))
def test_gpu_hc_matches_reference(gpu_ctx, ref, name):
    from bench import mip_chain
    fmt = FMTS[name]
    img = blockgen.smooth_image(256, 192, 21, alpha=True)
    l0 = gpu_ctx.launch_count
    blocks, g, r, ac = compare(gpu_ctx, ref, fmt, [mip_chain(img)], (512, 512, 256, 512))
    assert gpu_ctx.launch_count - l0 >= 10
    assert np.array_equal(g["block_encodings"], r["block_encodings"])
    assert np.array_equal(g["tile_indices"], r["tile_indices"])
    assert_tolerance(fmt, blocks, g, r, ac)


def test_gpu_hc_cubemap(gpu_ctx, ref):
    from bench import mip_chain
    faces = [mip_chain(blockgen.smooth_image(64, 64, 300 + f, alpha=True)) for f in range(6)]
    blocks, g, r, ac = compare(gpu_ctx, ref, 0, faces, (256, 256, 64, 64))
    assert np.array_equal(g["tile_indices"], r["tile_indices"])
    assert_tolerance(0, blocks, g, r, ac)


def test_gpu_hc_multithreaded_reference(gpu_ctx, ref):
    """The reference's own result moves with its helper-thread count (alternative sub-trees, crn_tree_clusterizer.h:61-79,
    :127-147); the device result must stay within the same tolerance of the 16-thread reference too."""
    from bench import mip_chain
    img = blockgen.smooth_image(256, 256, 33, alpha=True)
    blocks, g, r, ac = compare(gpu_ctx, ref, 3, [mip_chain(img)], (1024, 1024, 256, 1024), threads=15)
    assert_tolerance(3, blocks, g, r, ac, psnr_tol=0.1, bits_tol=0.02)


def test_gpu_hc_device_input(gpu_ctx, ref):
    import torch
    img = blockgen.smooth_image(64, 64, 44, alpha=True)
    blocks, levels = hc_util.hc_layout([[img]])
    a = gpu_ctx.hc_compress(3, blocks, levels, codebook_sizes=(64, 64, 32, 64))
    b = gpu_ctx.hc_compress(3, torch.from_numpy(blocks).cuda(), levels, codebook_sizes=(64, 64, 32, 64))
    for k in ("endpoint_indices", "selector_indices", "color_endpoints", "alpha_selectors"):
        assert np.array_equal(a[k], b[k]), k      # deterministic, host and device input alike


def test_gpu_hc_large_codebooks_uniform(gpu_ctx, ref):
    """8192-entry selector codebooks (the reference's maximum), uniform colour metric, thread-block-cluster tree kernels
    (training sets above 8192 vectors)."""
    from bench import mip_chain
    img = blockgen.smooth_image(512, 512, 55, alpha=True)
    blocks, levels = hc_util.hc_layout([mip_chain(img)])
    g = gpu_ctx.hc_compress(3, blocks, levels, perceptual=False, codebook_sizes=(4096, 8192, 2048, 8192))
    r = hc_util.ref_hc_compress(ref, 3, blocks, levels, perceptual=False, codebook_sizes=(4096, 8192, 2048, 8192))
    assert np.array_equal(g["tile_indices"], r["tile_indices"])
    assert_tolerance(3, blocks, g, r, (3, 0))
    assert g["info"]["unique_vectors"][2] > 8192
