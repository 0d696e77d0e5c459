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


def test_gpu_mip_device_api(gpu_ctx, ref):
    """Device-resident form: crn_gpu_generate_mipmaps into one tight buffer, crn_gpu_resample with a padded destination pitch."""
    import torch
    from test_mip_cpu import ref_resample
    w, h = 320, 192
    img = blockgen.smooth_image(w, h, 41, alpha=True)
    d_img = torch.from_numpy(img).cuda()
    sizes = [(max(1, h >> l), max(1, w >> l)) for l in range(1, 9)]        # 320x192 -> 9 levels (the count stops when both halved sizes are <= 1)
    total = sum(a * b * 4 for a, b in sizes)
    d_out = torch.zeros(total, dtype=torch.uint8, device="cuda")
    n = gpu_ctx.generate_mipmaps_device(d_img, w, h, w * 4, d_out, total)
    gpu_ctx.synchronize()
    assert n == 9
    want = ref_mips(ref, img)
    o, host = 0, d_out.cpu().numpy()
    for l, (a, b) in enumerate(sizes, 1):
        assert np.array_equal(host[o:o + a * b * 4].reshape(a, b, 4), want[l]); o += a * b * 4
    pitch = 100 * 4 + 64
    d_dst = torch.zeros(60 * pitch, dtype=torch.uint8, device="cuda")
    gpu_ctx.resample_device(d_img, w, h, w * 4, d_dst, 100, 60, pitch, filter="lanczos4", filter_scale=1.0)
    gpu_ctx.synchronize()
    got = d_dst.cpu().numpy().reshape(60, pitch)
    assert np.array_equal(got[:, :400].reshape(60, 100, 4), ref_resample(ref, img, 100, 60, filt="lanczos4", scale=1.0))
    assert not got[:, 400:].any()
