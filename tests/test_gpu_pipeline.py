"""The device-resident chain a crn_compress(comp_params, mipmap_params) replacement runs for a .CRN: level 0 -> mip chain
(crn_gpu_generate_mipmaps) -> 8-pixel padded block gather per level (crn_gpu_blockify) -> CRN quantiser
(crn_gpu_hc_compress on device blocks), with nothing but level 0 crossing PCIe.  Must equal the same chain fed from the host
with the reference's own mip levels, and a clustered .DDS must come out of the same device-resident levels."""
import numpy as np
import pytest

import blockgen
import hc_util
from test_mip_cpu import ref_mips

pytestmark = pytest.mark.gpu


def test_gpu_mips_blockify_hc_device_resident(gpu_ctx, ref):
    import torch
    w, h = 256, 128
    img = blockgen.smooth_image(w, h, 91, alpha=True)
    want_levels = ref_mips(ref, img)                                   # the reference's Kaiser chain (bit-exact with ours)
    blocks_host, levels = hc_util.hc_layout([want_levels])
    d_img = torch.from_numpy(img).cuda()
    sizes = [(max(1, h >> l), max(1, w >> l)) for l in range(1, len(want_levels))]
    total = sum(a * b * 4 for a, b in sizes)
    d_mips = torch.empty(total, dtype=torch.uint8, device="cuda")
    n = gpu_ctx.generate_mipmaps_device(d_img, w, h, w * 4, d_mips, total)
    assert n == len(want_levels)
    d_blocks = torch.empty((len(blocks_host), 16, 4), dtype=torch.uint8, device="cuda")
    first, o = 0, 0
    for l, (lh, lw) in enumerate([(h, w)] + sizes):
        src = d_img if l == 0 else d_mips[o:]
        bx, by = gpu_ctx.blockify(src.data_ptr() if l == 0 else d_mips.data_ptr() + o, lw, lh, lw * 4, d_blocks.data_ptr() + first * 64, 8)
        assert (first, bx * by, bx) == tuple(levels[l][:3])
        first += bx * by
        if l:
            o += lh * lw * 4
    gpu_ctx.synchronize()
    assert np.array_equal(d_blocks.cpu().numpy(), blocks_host)
    cbs = (256, 256, 128, 256)
    dev = gpu_ctx.hc_compress(3, d_blocks, levels, codebook_sizes=cbs)
    host = gpu_ctx.hc_compress(3, blocks_host, levels, codebook_sizes=cbs)
    for k in ("endpoint_indices", "selector_indices", "color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors"):
        assert np.array_equal(dev[k], host[k]), k
    r = hc_util.ref_hc_compress(ref, 3, blocks_host, levels, codebook_sizes=cbs)
    from test_hc_cpu import assert_tolerance
    assert_tolerance(3, blocks_host, dev, r, (3, 0))


def test_gpu_mips_feed_clustered_dds(gpu_ctx, ref):
    """Device-resident mip levels straight into the clustered-DDS quantiser (torch CUDA tensors as levels)."""
    import torch
    w = h = 128
    img = blockgen.smooth_image(w, h, 93, alpha=True)
    levels_host = gpu_ctx.generate_mipmaps(img)
    d_levels = [torch.from_numpy(np.ascontiguousarray(l)).cuda() for l in levels_host]
    qa = gpu_ctx.qdxt_init(3, d_levels); a = qa.pack(128); qa.close()
    qb = gpu_ctx.qdxt_init(3, levels_host); b = qb.pack(128); qb.close()
    assert np.array_equal(a, b)
