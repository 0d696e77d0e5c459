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
