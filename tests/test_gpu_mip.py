"""Mip-chain generation on the device, bit-exact against the reference's image_utils::resample (task-pool form)."""
import numpy as np
import pytest

import blockgen
from test_mip_cpu import ref_mips

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(512, 512), (300, 200), (1024, 4)])
def test_gpu_mip_chain_matches_reference(gpu_ctx, ref, size):
    w, h = size
    img = blockgen.smooth_image(w, h, 23, alpha=True)
    l0 = gpu_ctx.launch_count
    got = gpu_ctx.generate_mipmaps(img)
    want = ref_mips(ref, img)
    assert gpu_ctx.launch_count - l0 == 2 * (len(want) - 1)
    assert len(got) == len(want)
    for l, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), "level %d differs" % l


def test_gpu_mip_linear_rgb(gpu_ctx, ref):
    img = blockgen.smooth_image(256, 128, 29, alpha=True)
    got = gpu_ctx.generate_mipmaps(img, filter="lanczos4", filter_scale=1.0, srgb=False, num_comps=3, wrapping=True)
    want = ref_mips(ref, img, filt="lanczos4", scale=1.0, srgb=False, comps=3, wrap=True)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
