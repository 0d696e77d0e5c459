"""GPU parity of the cluster (N-pixel) optimisers through the C-ABI against the oracle port, at cluster
sizes typical of the clustered paths (1 .. 600 member blocks = 16 .. 9600 pixels)."""
import numpy as np
import pytest
import torch

import blockgen
import crunch2_b200 as crn
import helpers
from test_clusters_cpu import Prm, Res, make_clusters, port_color_cluster, P

pytestmark = pytest.mark.gpu
import ctypes  # noqa: E402


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_color_gpu(ctx, port, blocks, offs, members, q, perc, uab, dxt1a=False, check_every=1):
    n_blocks = len(blocks)
    d_blocks, d_offs, d_mem = to_dev(blocks.view(np.uint8)), to_dev(offs.view(np.int32)), to_dev(members.view(np.int32))
    d_out = torch.zeros(n_blocks * 8, dtype=torch.uint8, device="cuda")
    nc = len(offs) - 1
    d_ep = torch.zeros(nc, dtype=torch.int32, device="cuda"); d_err = torch.zeros(nc, dtype=torch.int64, device="cuda")
    ctx.optimize_clusters("color", d_blocks, n_blocks, d_offs, d_mem, nc, int(offs[-1]), d_out, 8, 0,
                          crn.PackParams(dxt_quality=q, perceptual=perc, use_both_block_types=uab), dxt1a=dxt1a, d_endpoints=d_ep, d_error=d_err)
    ctx.synchronize()
    out = d_out.cpu().numpy().reshape(n_blocks, 8); ep = d_ep.cpu().numpy().view(np.uint32); err = d_err.cpu().numpy().view(np.uint64)
    for c in range(0, nc, check_every):
        m = members[offs[c]:offs[c + 1]]
        px = np.ascontiguousarray(blocks[m].reshape(-1, 4))
        pha = int(dxt1a and uab and (px[:, 3] < 128).any())
        lo, hi, e, sel = port_color_cluster(port, px, q, perc, uab, pha)
        assert (int(ep[c]) & 0xffff, int(ep[c]) >> 16, int(err[c])) == (lo, hi, e), (c, len(m))
        packed = (sel.reshape(-1, 16).astype(np.uint64) << (2 * np.arange(16, dtype=np.uint64))).sum(1)
        want = (np.uint64(lo) | (np.uint64(hi) << np.uint64(16)) | (packed << np.uint64(32)))
        got = out[m].copy().view(np.uint64).ravel()
        assert (got == want).all(), (c, len(m))


@pytest.mark.parametrize("family", ["smooth", "noise", "four", "dark"])
def test_gpu_color_clusters(gpu_ctx, port, family):
    blocks = blockgen.block_family(family, 3000, 17)
    sizes = [1, 1, 2, 3, 5, 8, 16, 33, 64, 100, 250, 600] + [4] * 100 + [20] * 40
    offs, members = make_clusters(3000, sizes, 5)
    run_color_gpu(gpu_ctx, port, blocks, offs, members, 4, 1, 0)
    run_color_gpu(gpu_ctx, port, blocks, offs, members, 4, 1, 1, check_every=3)


def test_gpu_color_clusters_smooth_image_tiles(gpu_ctx, port):
    """Clusters of spatially adjacent blocks of one image (what endpoint clustering really groups)."""
    img = blockgen.smooth_image(256, 256, 2048, alpha=True)
    blocks = np.ascontiguousarray(img.reshape(64, 4, 64, 4, 4).transpose(0, 2, 1, 3, 4).reshape(4096, 16, 4))
    offs = np.arange(0, 4097, 16, dtype=np.uint32)
    members = np.arange(4096, dtype=np.uint32)
    run_color_gpu(gpu_ctx, port, blocks, offs, members, 4, 1, 0, check_every=4)


@pytest.mark.parametrize("family", ["smooth", "noise", "gray"])
def test_gpu_alpha_clusters(gpu_ctx, port, family):
    blocks = blockgen.block_family(family, 3000, 23)
    sizes = [1, 2, 7, 30, 100, 700] + [5] * 150
    offs, members = make_clusters(3000, sizes, 9)
    n_blocks = len(blocks); nc = len(offs) - 1
    d_blocks, d_offs, d_mem = to_dev(blocks.view(np.uint8)), to_dev(offs.view(np.int32)), to_dev(members.view(np.int32))
    for comp, q, both in ((3, 4, 1), (0, 4, 0)):
        d_out = torch.zeros(n_blocks * 16, dtype=torch.uint8, device="cuda")
        d_ep = torch.zeros(nc, dtype=torch.int32, device="cuda"); d_err = torch.zeros(nc, dtype=torch.int64, device="cuda")
        gpu_ctx.optimize_clusters("alpha", d_blocks, n_blocks, d_offs, d_mem, nc, int(offs[-1]), d_out, 16, 0,
                                  crn.PackParams(dxt_quality=q, use_both_block_types=both), component=comp, d_endpoints=d_ep, d_error=d_err)
        gpu_ctx.synchronize()
        out = d_out.cpu().numpy().reshape(n_blocks, 16); ep = d_ep.cpu().numpy().view(np.uint32); err = d_err.cpu().numpy().view(np.uint64)
        for c in range(nc):
            m = members[offs[c]:offs[c + 1]]
            px = np.ascontiguousarray(blocks[m].reshape(-1, 4)); n = len(px)
            f = ctypes.c_uint8(); s = ctypes.c_uint8(); e = ctypes.c_uint64(); bt = ctypes.c_uint8(); sel = np.zeros(n, np.uint8)
            port.op_dxt5_optimize(P(px), n, comp, q, both, ctypes.byref(f), ctypes.byref(s), P(sel), ctypes.byref(e), ctypes.byref(bt))
            assert (int(ep[c]) & 0xff, int(ep[c]) >> 8, int(err[c])) == (f.value, s.value, e.value), (family, comp, c, n)
            packed = (sel.reshape(-1, 16).astype(np.uint64) << (3 * np.arange(16, dtype=np.uint64))).sum(1)
            want = np.uint64(f.value) | (np.uint64(s.value) << np.uint64(8)) | (packed << np.uint64(16))
            assert (out[m][:, :8].copy().view(np.uint64).ravel() == want).all()
