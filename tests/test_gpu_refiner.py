"""GPU parity of the dxt_hc building blocks through the C-ABI against the oracle."""
import numpy as np
import pytest
import torch

from test_refiner_cpu import lib_refine, make_clusters, make_codebook_case, port_nearest, port_refine, ref_refine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dxt1,perc,comp,seed,k", [(1, 1, 0, 21, 3000), (1, 0, 0, 22, 500), (0, 1, 3, 23, 3000), (0, 0, 2, 24, 500)])
def test_gpu_refiner_bit_exact(gpu_ctx, port, dxt1, perc, comp, seed, k):
    px, sel, offs, etb = make_clusters(seed, k, dxt1, max_pixels=1500 if k <= 500 else 200)
    import helpers
    ref = helpers.load_ref()
    want = ref_refine(ref, dxt1, perc, comp, px, sel, offs, etb) if ref is not None else port_refine(port, dxt1, perc, comp, px, sel, offs, etb)
    got = lib_refine(gpu_ctx, dxt1, perc, comp, px, sel, offs, etb, to_dev=lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)).cuda(),
                     to_host=lambda t: t.cpu().numpy().view(np.uint64 if t.dtype == torch.int64 else (np.uint32 if t.dtype == torch.int32 else np.uint8)))
    assert got == want


@pytest.mark.parametrize("dims,n,k", [(6, 100000, 8192), (2, 50000, 4096), (6, 1000, 37)])
def test_gpu_nearest_codebook(gpu_ctx, port, dims, n, k):
    v, cb = make_codebook_case(9, dims, n, k)
    want = port_nearest(port, dims, v, cb)
    dv, dc = torch.from_numpy(v).cuda(), torch.from_numpy(cb).cuda()
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    gpu_ctx.nearest_codebook(dims, dv, n, dc, k, out)
    gpu_ctx.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize("kind,perc,comp,with_accum,n,k", [(0, 1, 0, False, 60000, 2700), (0, 0, 0, False, 5000, 8192), (1, 0, 3, False, 60000, 2700), (1, 0, 1, True, 5000, 8192)])
def test_gpu_assign_selectors(gpu_ctx, port, kind, perc, comp, with_accum, n, k):
    from test_refiner_cpu import make_assign_case, port_assign
    blocks, values, accum, codebook = make_assign_case(31 + kind, kind, n, k)
    accum = accum if with_accum else None
    want = port_assign(port, kind, perc, comp, blocks, values, accum, codebook)
    dev = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
    d_blocks, d_values, d_cb = dev(blocks), dev(values), dev(codebook)
    d_accum = dev(accum) if accum is not None else None
    best = torch.zeros(n, dtype=torch.int32, device="cuda"); refined = torch.zeros(k, dtype=torch.int64, device="cuda"); used = torch.zeros(k, dtype=torch.uint8, device="cuda")
    gpu_ctx.assign_selectors("alpha" if kind else "color", d_blocks, n, d_values, d_cb, k, best, refined, used, perceptual=bool(perc), component=comp, d_values_accum=d_accum)
    gpu_ctx.synchronize()
    assert np.array_equal(best.cpu().numpy().view(np.uint32), want[0])
    assert np.array_equal(refined.cpu().numpy().view(np.uint64), want[1])
    assert np.array_equal(used.cpu().numpy(), want[2])
