"""dxt_hc building blocks on the CPU: the oracle port of dxt_endpoint_refiner against the unmodified
reference, and the CUDA kernels under the SIMT emulator against the port, through the C-ABI."""
import ctypes

import numpy as np
import pytest

import crunch2_b200 as crn
import helpers

P = ctypes.c_void_p


def make_clusters(seed, n_clusters, dxt1, max_pixels=200):
    rng = np.random.default_rng(seed)
    px, sel, offs = [], [], [0]
    for c in range(n_clusters):
        n = int(rng.integers(1, max_pixels)) if c else 1
        base = rng.integers(0, 256, (1, 4))
        spread = int(rng.integers(1, 120))
        p = np.clip(base + rng.integers(-spread, spread + 1, (n, 4)), 0, 255).astype(np.uint8)
        if c % 7 == 3:
            p[:] = p[0]                                   # solid cluster
        s = rng.integers(0, 4 if dxt1 else 8, n).astype(np.uint8)
        if c % 5 == 0:
            s[:] = s[0]                                   # one selector value: a degenerate least-squares system
        px.append(p); sel.append(s); offs.append(offs[-1] + n)
    return np.concatenate(px), np.concatenate(sel), np.array(offs, np.uint32), rng.integers(0, 1 << 36, n_clusters).astype(np.uint64)


def port_refine(port, dxt1, perc, comp, px, sel, offs, etb):
    port.op_refine.restype = ctypes.c_int
    out = []
    for c in range(len(offs) - 1):
        a, b = int(offs[c]), int(offs[c + 1])
        lo, hi, er = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64()
        p, s = np.ascontiguousarray(px[a:b]), np.ascontiguousarray(sel[a:b])
        ok = port.op_refine(dxt1, perc, comp, P(p.ctypes.data), b - a, P(s.ctypes.data), ctypes.c_uint64(int(etb[c])), ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(er))
        out.append((lo.value, hi.value, er.value, ok))
    return out


def ref_refine(ref, dxt1, perc, comp, px, sel, offs, etb):
    k = len(offs) - 1
    lo, hi, er, ok = np.zeros(k, np.uint32), np.zeros(k, np.uint32), np.zeros(k, np.uint64), np.zeros(k, np.uint8)
    ref.ref_refine(dxt1, perc, comp, P(px.ctypes.data), P(sel.ctypes.data), P(offs.ctypes.data), k, P(etb.ctypes.data), P(lo.ctypes.data), P(hi.ctypes.data), P(er.ctypes.data), P(ok.ctypes.data))
    return [(int(lo[i]), int(hi[i]), int(er[i]), int(ok[i])) for i in range(k)]


def lib_refine(ctx, dxt1, perc, comp, px, sel, offs, etb, to_dev=lambda a: a, to_host=lambda a: a):
    k = len(offs) - 1
    ep, er, ok = np.zeros(k, np.uint32), np.zeros(k, np.uint64), np.zeros(k, np.uint8)
    d = [to_dev(x) for x in (px, sel, offs, etb, ep, er, ok)]
    get = (lambda x: x.data_ptr()) if hasattr(d[0], "data_ptr") else (lambda x: x.ctypes.data)
    ctx.refine_endpoints(dxt1, get(d[0]), get(d[1]), get(d[2]), k, get(d[4]), get(d[5]), get(d[6]), perceptual=bool(perc), component=comp, d_error_to_beat=get(d[3]))
    ctx.synchronize()
    ep, er, ok = to_host(d[4]), to_host(d[5]), to_host(d[6])
    return [(int(ep[i]) & 0xffff, int(ep[i]) >> 16, int(er[i]), int(ok[i])) for i in range(k)]


@pytest.mark.parametrize("dxt1,perc,comp,seed", [(1, 1, 0, 1), (1, 0, 0, 2), (0, 1, 3, 3), (0, 0, 1, 4)])
def test_port_matches_reference_refiner(port, ref, dxt1, perc, comp, seed):
    px, sel, offs, etb = make_clusters(seed, 300, dxt1)
    assert port_refine(port, dxt1, perc, comp, px, sel, offs, etb) == ref_refine(ref, dxt1, perc, comp, px, sel, offs, etb)


@pytest.mark.parametrize("dxt1,perc,comp,seed", [(1, 1, 0, 11), (1, 0, 0, 12), (0, 1, 3, 13), (0, 0, 0, 14)])
def test_sim_refiner_matches_port(sim, port, dxt1, perc, comp, seed):
    px, sel, offs, etb = make_clusters(seed, 60, dxt1, max_pixels=120)
    ctx = crn.Context(0, lib=sim)
    assert lib_refine(ctx, dxt1, perc, comp, px, sel, offs, etb) == port_refine(port, dxt1, perc, comp, px, sel, offs, etb)
    ctx.close()


def make_codebook_case(seed, dims, n, k):
    rng = np.random.default_rng(seed)
    cb = (rng.integers(0, 256, (k, dims)) / np.float32(255.0)).astype(np.float32)
    v = (rng.integers(0, 256, (n, dims)) / np.float32(255.0)).astype(np.float32)
    v[::7] = cb[rng.integers(0, k, len(v[::7]))]          # exact matches (the reference stops at distance 0)
    cb[k // 2] = cb[k // 3]                               # duplicate entries: the first one must win
    return v, cb


def port_nearest(port, dims, v, cb):
    out = np.zeros(len(v), np.uint32)
    port.op_nearest_codebook(dims, P(v.ctypes.data), len(v), P(cb.ctypes.data), len(cb), P(out.ctypes.data))
    return out


@pytest.mark.parametrize("dims,n,k", [(6, 300, 1500), (2, 200, 300), (6, 5, 1)])
def test_sim_nearest_codebook_matches_port(sim, port, dims, n, k):
    v, cb = make_codebook_case(5, dims, n, k)
    want = port_nearest(port, dims, v, cb)
    # the port against a float32 numpy restatement of the same loop order (the function has no callable counterpart in the reference)
    d = np.zeros((n, k), np.float32)
    for t in range(dims):
        e = cb[None, :, t] - v[:, None, t]
        d = d + e * e
    assert np.array_equal(want, d.argmin(axis=1))
    ctx = crn.Context(0, lib=sim)
    out = np.zeros(n, np.uint32)
    ctx.nearest_codebook(dims, v.ctypes.data, n, cb.ctypes.data, k, out.ctypes.data)
    ctx.synchronize()
    assert np.array_equal(out, want)
    ctx.close()


# ---- a14 pinned against the reference's own task (through the shim's view of dxt_hc) ---------------------------------
@pytest.mark.parametrize("dims,n,max_size,seed", [(6, 1500, 200, 1), (2, 1200, 150, 2), (6, 300, 4000, 3)])
def test_port_nearest_codebook_matches_reference_task(port, ref, dims, n, max_size, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n // 4 + 1, dims))
    v = np.unique(np.clip(np.repeat(base, 4, axis=0)[:n] + rng.integers(-20, 21, (n, dims)), 0, 255), axis=0).astype(np.float32)
    rng.shuffle(v)
    v = (v / np.float32(255.0)).astype(np.float32)
    n = len(v)
    w = rng.integers(1, 40, n).astype(np.uint32)
    cb = np.zeros((max_size + 1, dims), np.float32)
    idx = np.zeros(n, np.uint32)
    ref.ref_hc_nearest_codebook.restype = ctypes.c_uint32
    k = ref.ref_hc_nearest_codebook(dims, P(v.ctypes.data), P(w.ctypes.data), n, max_size, P(cb.ctypes.data), P(idx.ctypes.data))
    assert 0 < k <= max_size
    cb = np.ascontiguousarray(cb[:k])
    assert np.array_equal(port_nearest(port, dims, v, cb), idx)


# ---- a16: selector codebook assignment + re-vote -------------------------------------------------------------------------
def make_assign_case(seed, kind, n, k):
    rng = np.random.default_rng(seed)
    blocks = np.clip(rng.integers(0, 256, (n, 1, 4)) + rng.integers(-40, 41, (n, 16, 4)), 0, 255).astype(np.uint8)
    if kind == 0:
        values = np.clip(blocks[:, :1, :] + rng.integers(-60, 61, (n, 4, 4)), 0, 255).astype(np.uint8)
        codebook = rng.integers(0, 1 << 32, k, dtype=np.uint64)
    else:
        values = np.sort(rng.integers(0, 256, (n, 8)), axis=1).astype(np.uint8)
        codebook = rng.integers(0, 1 << 48, k, dtype=np.uint64)
    codebook[k // 2] = codebook[k // 3]                     # duplicates: the first must win
    accum = np.clip(values.astype(int) + rng.integers(-3, 4, values.shape), 0, 255).astype(np.uint8) if kind else None
    return blocks, values, accum, codebook


def port_assign(port, kind, perc, comp, blocks, values, accum, codebook):
    n, k = len(blocks), len(codebook)
    best, refined, used = np.zeros(n, np.uint32), np.zeros(k, np.uint64), np.zeros(k, np.uint8)
    port.op_assign_selectors(kind, perc, comp, P(blocks.ctypes.data), n, P(values.ctypes.data), P(accum.ctypes.data) if accum is not None else None,
                             P(codebook.ctypes.data), k, P(best.ctypes.data), P(refined.ctypes.data), P(used.ctypes.data))
    return best, refined, used


def revote(errors, kind):
    """create_color/alpha_selector_codebook's tail (crn_dxt_hc.cpp:1488-1503, :1702-1720) on the reference's error tables."""
    V = 8 if kind else 4
    e = errors.reshape(-1, 16, V).astype(np.int64)
    out = np.zeros(len(e), np.uint64)
    for i in range(len(e)):
        sel = 0
        for p in range(16):
            t = e[i, p]
            if kind == 0:
                s03 = 3 if t[3] < t[0] else 0; s12 = 2 if t[2] < t[1] else 1
                s = s12 if t[s12] < t[s03] else s03
            else:
                s07 = 7 if t[7] < t[0] else 0; s12 = 2 if t[2] < t[1] else 1; s34 = 4 if t[4] < t[3] else 3; s56 = 6 if t[6] < t[5] else 5
                s02 = s12 if t[s12] < t[s07] else s07; s36 = s56 if t[s56] < t[s34] else s34
                s = s36 if t[s36] < t[s02] else s02
            sel |= s << (p * (3 if kind else 2))
        out[i] = sel
    return out


@pytest.mark.parametrize("kind,perc,comp,with_accum,seed", [(0, 1, 0, False, 1), (0, 0, 0, False, 2), (1, 0, 3, False, 3), (1, 0, 1, True, 4)])
def test_port_assign_selectors_matches_reference_task(port, ref, kind, perc, comp, with_accum, seed):
    blocks, values, accum, codebook = make_assign_case(seed, kind, 2000, 300)
    accum = accum if with_accum else None
    n, k, V = len(blocks), len(codebook), 8 if kind else 4
    best, used, errors = np.zeros(n, np.uint32), np.zeros(k, np.uint8), np.zeros(k * 16 * V, np.uint32)
    assert ref.ref_hc_assign_selectors(kind, perc, comp, P(blocks.ctypes.data), n, P(values.ctypes.data), P(accum.ctypes.data) if accum is not None else None,
                                       P(codebook.ctypes.data), k, P(best.ctypes.data), P(errors.ctypes.data), P(used.ctypes.data))
    pb, pr, pu = port_assign(port, kind, perc, comp, blocks, values, accum, codebook)
    assert np.array_equal(pb, best) and np.array_equal(pu, used)
    assert np.array_equal(pr, revote(errors, kind))


@pytest.mark.parametrize("kind,perc,comp,with_accum,seed", [(0, 1, 0, False, 5), (0, 0, 0, False, 6), (1, 0, 3, False, 7), (1, 0, 0, True, 8)])
def test_sim_assign_selectors_matches_port(sim, port, kind, perc, comp, with_accum, seed):
    blocks, values, accum, codebook = make_assign_case(seed, kind, 150, 100)
    accum = accum if with_accum else None
    want = port_assign(port, kind, perc, comp, blocks, values, accum, codebook)
    ctx = crn.Context(0, lib=sim)
    n, k = len(blocks), len(codebook)
    best, refined, used = np.zeros(n, np.uint32), np.zeros(k, np.uint64), np.zeros(k, np.uint8)
    ctx.assign_selectors("alpha" if kind else "color", blocks.ctypes.data, n, values.ctypes.data, codebook.ctypes.data, k, best.ctypes.data, refined.ctypes.data, used.ctypes.data,
                         perceptual=bool(perc), component=comp, d_values_accum=accum.ctypes.data if accum is not None else None)
    ctx.synchronize()
    assert np.array_equal(best, want[0]) and np.array_equal(refined, want[1]) and np.array_equal(used, want[2])
    ctx.close()
