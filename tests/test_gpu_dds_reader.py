"""-m gpu: the .dds reader / uncook / mask-extraction kernels on the device against the unmodified reference (see tests/test_dds_reader_cpu.py)."""
import ctypes

import numpy as np
import pytest

from test_dds_reader_cpu import BLOCK_CASES, RAW_CASES, block_payload, cc, dds_header, ref_decode

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fourcc,bpb,kind,bitcount,w,h,levels,faces", BLOCK_CASES)
def test_gpu_block_dds_to_images(gpu_ctx, ref, fourcc, bpb, kind, bitcount, w, h, levels, faces):
    rng = np.random.default_rng(w * 31 + h)
    w, h = w * 16 + 3, h * 16 + 1                      # larger, ragged sizes on the device
    dds = dds_header(w, h, levels, faces, fourcc=fourcc, bitcount=cc(bitcount) if bitcount else 0) + block_payload(rng, w, h, levels, faces, bpb, kind)
    imgs, d = gpu_ctx.dds_to_images(dds)
    want, rd = ref_decode(ref, dds, sum(i.shape[0] * i.shape[1] for i in imgs))
    assert np.array_equal(np.concatenate([i.reshape(-1) for i in imgs]), want)
    assert [d["faces"], d["width"], d["height"], d["levels"], d["pixel_format"]] == rd


@pytest.mark.parametrize("name,bits,masks,pf", RAW_CASES)
def test_gpu_raw_dds_to_images(gpu_ctx, ref, name, bits, masks, pf):
    rng = np.random.default_rng(bits)
    w, h, levels = 301, 203, 4
    payload = b"".join(rng.integers(0, 256, max(1, w >> l) * max(1, h >> l) * (bits // 8), dtype=np.uint8).tobytes() for l in range(levels))
    dds = dds_header(w, h, levels, 1, bitcount=bits, masks=masks, pf_flags=pf) + payload
    imgs, d = gpu_ctx.dds_to_images(dds)
    want, rd = ref_decode(ref, dds, sum(i.shape[0] * i.shape[1] for i in imgs))
    assert np.array_equal(np.concatenate([i.reshape(-1) for i in imgs]), want)
    assert [d["faces"], d["width"], d["height"], d["levels"], d["pixel_format"]] == rd


@pytest.mark.parametrize("conv", range(1, 10))
def test_gpu_convert_pixels(gpu_ctx, ref, conv):
    import torch
    rng = np.random.default_rng(conv)
    px = rng.integers(0, 256, (300, 256, 4), dtype=np.uint8)
    g = np.arange(65536, dtype=np.uint32)
    if conv == 4:
        px[:256, :, 3] = (g >> 8).reshape(256, 256); px[:256, :, 1] = (g & 255).reshape(256, 256)
    elif conv == 9:
        px[:256, :, 0] = (g >> 8).reshape(256, 256); px[:256, :, 1] = (g & 255).reshape(256, 256)
    want = px.copy()
    ref.ref_convert_image(want.ctypes.data_as(ctypes.c_void_p), 256, 300, conv - 1)
    d = torch.from_numpy(px).cuda()
    gpu_ctx.convert_pixels_device(d.data_ptr(), 256, 300, 1024, conv)
    gpu_ctx.synchronize()
    assert np.array_equal(d.cpu().numpy(), want)
