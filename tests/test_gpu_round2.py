"""Round-2 functionality on the device, through the product libraries: the same checks as their emulator twins (imported from the CPU test
modules), with the nvcc build of libcrn_b200.so / libcrnlib_b200.so and the unmodified reference (oracle/_ref) as the oracle."""
import os

import numpy as np
import pytest

import crunch2_b200 as crn
import helpers
import test_dds_cpu
import test_dropin_api
import test_swizzled_cpu
from crunch2_b200 import dropin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exact_ctx():
    ctx = crn.Context(0)
    ctx.set_vq_mode(True)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def product_dropin():
    if not os.path.exists(dropin.library_path()):
        pytest.skip("libcrnlib_b200.so not built (needs the reference's headers at build time)")
    return dropin.load()


@pytest.mark.parametrize("fmt", test_swizzled_cpu.SWIZZLED)
def test_gpu_swizzled_block_by_block_files(gpu_ctx, ref, fmt):
    test_swizzled_cpu.test_block_by_block_dds_matches_reference_file(gpu_ctx, ref, fmt)


@pytest.mark.parametrize("fmt", ["DXT5_CCxY", "DXT5_xGBR"])
def test_gpu_swizzled_clustered_and_crn(gpu_ctx, ref, fmt):
    test_swizzled_cpu.test_clustered_dds_within_tolerance(gpu_ctx, ref, fmt)
    test_swizzled_cpu.test_crn_within_tolerance(gpu_ctx, ref, fmt)


def test_gpu_dxt5a_of_an_opaque_image(gpu_ctx, ref):
    test_swizzled_cpu.test_dxt5a_of_an_opaque_image_packs_luma(gpu_ctx, ref)


def test_gpu_non_hierarchical_and_search_retry(exact_ctx, ref):
    test_dds_cpu.test_compress_dds_without_adaptive_tiles(exact_ctx, ref)
    test_dds_cpu.test_compress_dds_bitrate_search_and_retry(exact_ctx, ref)


@pytest.mark.parametrize("case", ["window past the edge", "clamp by scaling", "nearest pow2", "relative", "renormalise mips", "renormalise top mip", "source mips dropped by a resize"])
def test_gpu_mipmap_source_options(product_dropin, ref, case):
    test_dropin_api.test_mipmap_source_options_match_reference_file(product_dropin, ref, case)


# ---- stand-alone device checks of a11 / a18 (tile analysis + training vectors) and a21 (selector re-vote), bit for bit ----------------------
def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kind,comp", [(0, 3), (1, 3), (1, 0)])
def test_gpu_training_vectors_match_port(gpu_ctx, port, kind, comp):
    """crn_gpu_qdxt_training (dxt_fast tile fits, encoding choice, find_representative_colors; crn_qdxt1.cpp:101-341 / crn_qdxt5.cpp:103-344) on the
    device against the port: chunk encodings, training vectors and weights equal."""
    import blockgen
    import test_qdxt_cpu as tq
    from bench import mip_chain
    for (w, h, seed) in ((256, 192, 1), (40, 24, 2), (13, 9, 3)):
        blocks, mips = tq.layout_levels(mip_chain(blockgen.smooth_image(w, h, seed, alpha=True)))
        n = len(blocks); D = 2 if kind else 6
        nch = sum(((bw + 1) // 2) * ((bh + 1) // 2) for _, bw, bh in mips)
        d_blocks = _dev(blocks.view(np.uint8))
        d_vecs, d_w, d_enc = _dev(np.zeros((n, D), np.uint8)), _dev(np.zeros(n, np.int32)), _dev(np.zeros(nch, np.uint8))
        arr = (tq.MipDesc * len(mips))(*[tq.MipDesc(*m) for m in mips])
        gpu_ctx._check(gpu_ctx._lib.crn_gpu_qdxt_training(gpu_ctx._ctx, kind, comp, d_blocks.data_ptr(), n, arr, len(mips), d_vecs.data_ptr(), d_w.data_ptr(), d_enc.data_ptr()))
        gpu_ctx.synchronize()
        want_v, want_w, want_e = tq.port_training(port, kind, comp, blocks, mips)
        assert (d_enc.cpu().numpy() == want_e).all(), (kind, w, h)
        assert (d_vecs.cpu().numpy() == want_v).all() and (d_w.cpu().numpy().view(np.uint32) == want_w).all(), (kind, w, h)


@pytest.mark.parametrize("is_alpha", [False, True])
def test_gpu_selector_revote_matches_restatement(gpu_ctx, port, is_alpha):
    """crn_gpu_optimize_selectors (qdxt1 / qdxt5::optimize_selectors_task, crn_qdxt1.cpp:714-865, crn_qdxt5.cpp:578-687) on the device"""
    import blockgen
    import test_clusters_cpu as tc
    blocks = blockgen.block_family("smooth", 3000, 71)
    img = helpers.blocks_to_image(blocks)
    packed = helpers.port_pack(port, 4 if is_alpha else 0, img, 4, 1, 1)
    elems = packed.view(np.uint64).copy()
    offs, members = tc.make_clusters(3000, [1, 2, 3, 9, 30, 77, 150, 700] + [6] * 200, 13)
    want = tc.revote_reference(blocks, elems, offs, members, is_alpha, threshold=128)
    d_blocks, d_offs, d_mem, d_el = _dev(blocks.view(np.uint8)), _dev(offs.view(np.int32)), _dev(members.view(np.int32)), _dev(elems.view(np.int64))
    gpu_ctx.optimize_selectors("alpha" if is_alpha else "color", d_blocks, 3000, d_offs, d_mem, len(offs) - 1, d_el, 8, 0, crn.PackParams(perceptual=True), component=3)
    gpu_ctx.synchronize()
    got = d_el.cpu().numpy().view(np.uint64)
    assert (got == want).all()
    assert (got != elems).any()
