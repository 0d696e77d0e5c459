"""DXTn -> RGBA unpack kernel on the device, bit-exact against the reference's dxt_image::unpack."""
import numpy as np
import pytest

from test_unpack_cpu import random_blocks, ref_unpack

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4, 5, 6])
def test_gpu_unpack_matches_reference(gpu_ctx, ref, fmt):
    for (w, h) in ((512, 256), (251, 67), (3, 2)):
        blocks = random_blocks(fmt, w, h, 31 * fmt + w)
        l0 = gpu_ctx.launch_count
        got = gpu_ctx.unpack_image(fmt, blocks, w, h)
        assert gpu_ctx.launch_count == l0 + 1
        assert np.array_equal(got, ref_unpack(ref, fmt, blocks, w, h))


def test_gpu_unpack_device_pitch(gpu_ctx, ref):
    import torch
    w, h, pitch = 100, 36, 512
    blocks = random_blocks(3, w, h, 77)
    d_b = torch.from_numpy(blocks).cuda()
    d_o = torch.zeros(h * pitch, dtype=torch.uint8, device="cuda")
    gpu_ctx.unpack_image_device(3, d_b, w, h, d_o, pitch)
    gpu_ctx.synchronize()
    got = d_o.cpu().numpy().reshape(h, pitch)[:, : w * 4].reshape(h, w, 4)
    assert np.array_equal(got, ref_unpack(ref, 3, blocks, w, h))
    assert not d_o.cpu().numpy().reshape(h, pitch)[:, w * 4:].any()          # padding untouched
