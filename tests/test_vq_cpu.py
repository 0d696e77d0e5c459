"""Frontier-batched vector quantiser (vq_kernels.cuh / vq_host.h) under the SIMT emulator vs the reference's
clusterizer<V> / threaded_clusterizer<V> (oracle/_ref).  The device reproduces the reference's float roundings
(exact integer sums below 2^24, member-order float accumulation above and for the covariance), so the cluster
assignment must be IDENTICAL, including on inputs whose sums pass 2^24."""
import ctypes

import numpy as np
import pytest

import crunch2_b200 as crn
import helpers

P = helpers.P


def make_vectors(dims, n, seed, kind="clumpy", max_weight=8):
    rng = np.random.default_rng(seed)
    if dims == 16:      # linear selector vectors 0..7
        if kind == "clumpy":
            base = rng.integers(0, 8, (max(n // 20, 1), 16))
            vecs = np.clip(base[rng.integers(0, len(base), n)] + rng.integers(-1, 2, (n, 16)), 0, 7)
        else:
            vecs = rng.integers(0, 8, (n, 16))
    else:               # endpoint vectors 0..255
        centers = rng.integers(0, 256, (max(n // 30, 1), dims))
        vecs = np.clip(centers[rng.integers(0, len(centers), n)] + rng.normal(0, 6, (n, dims)), 0, 255)
    return np.ascontiguousarray(vecs.astype(np.uint8)), rng.integers(1, max_weight + 1, n).astype(np.uint32)


def ref_clusterize(ref, vecs, w, max_size, retrieve, threaded):
    n, dims = vecs.shape
    fv = np.ascontiguousarray(vecs.astype(np.float32))
    co = np.zeros(n, np.uint32); k = ctypes.c_uint32(); cb = ctypes.c_uint32()
    if threaded:
        assert ref.ref_threaded_clusterizer16(n, P(fv), P(w), max_size, 1, P(co), ctypes.byref(k)) == 1
        return co, k.value, None
    assert ref.ref_clusterizer(dims, n, P(fv), P(w), max_size, retrieve, P(co), ctypes.byref(k), ctypes.byref(cb)) == 1
    return co, k.value, cb.value


def agreement(a, b):
    """fraction of vectors whose cluster has exactly the same member set in both partitions"""
    n = len(a)
    def sig(c):
        order = np.argsort(c, kind="stable")
        bounds = np.flatnonzero(np.diff(c[order])) + 1
        out = np.zeros(n, np.uint64)
        for grp in np.split(order, bounds):
            out[grp] = (np.uint64(len(grp)) << np.uint64(40)) ^ np.uint64(int(grp.min()) * 2654435761 % (1 << 40)) ^ np.uint64(int(grp.sum()) % (1 << 40))
        return out
    return float(np.mean(sig(a) == sig(b)))


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    ctx.set_vq_mode(True)            # the exact member-order builder (vq_kernels.cuh); the default single-launch builder is tested at the end
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def fastctx(sim):
    ctx = crn.Context(0, lib=sim)    # default mode: single-launch frontier splits (vq_fast.cuh), tolerance class
    yield ctx
    ctx.close()


def distortion(vecs, w, co):
    """weighted squared error of every vector against its cluster's weighted centroid: what the quantiser minimises"""
    v = vecs.astype(np.float64); ww = w.astype(np.float64)
    k = int(co.max()) + 1
    sw = np.bincount(co, ww, k)
    cen = np.stack([np.bincount(co, ww * v[:, d], k) / sw for d in range(v.shape[1])], 1)
    return float((ww * ((v - cen[co]) ** 2).sum(1)).sum())


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed", [
    (6, 300, 65535, 40, False, 1),
    (6, 2000, 65535, 200, False, 2),
    (2, 1500, 65535, 100, False, 3),
    (16, 1000, 65535, 100, False, 4),
    (6, 2000, 100, 0, False, 5),        # budget-limited codebook, every leaf retrieved
    (16, 2000, 300, 0, True, 6),        # threaded_clusterizer: 3 PCA divisions + 4 clusterizers
    (16, 500, 100, 0, True, 7),         # below 128 clusters: single clusterizer
    (6, 1, 65535, 10, False, 8),        # single vector
    (2, 64, 65535, 1000, False, 9),     # more clusters asked for than vectors
])
def test_matches_reference_exactly(simctx, ref, dims, n, max_size, retrieve, threaded, seed):
    vecs, w = make_vectors(dims, n, seed)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    co_g, k_g, cb_g = simctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, n, dims, max_size, retrieve, threaded)
    assert k_g == k_r
    if cb_r is not None:
        assert cb_g == cb_r
    assert np.array_equal(co_g, co_r)


def test_duplicates_and_constant_input(simctx, ref):
    # all vectors equal: the root has zero variance and is never split; duplicates make unsplittable nodes
    vecs = np.full((200, 6), 77, np.uint8); w = np.ones(200, np.uint32)
    co_g, k_g, cb_g = simctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, 200, 6, 65535, 16, False)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, 65535, 16, False)
    assert (k_g, cb_g) == (k_r, cb_r) and np.array_equal(co_g, co_r)
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (7, 6)).astype(np.uint8)
    vecs = np.ascontiguousarray(base[rng.integers(0, 7, 500)]); w = rng.integers(1, 4, 500).astype(np.uint32)
    co_g, k_g, cb_g = simctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, 500, 6, 65535, 64, False)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, 65535, 64, False)
    assert (k_g, cb_g) == (k_r, cb_r) and np.array_equal(co_g, co_r)


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed,max_weight", [
    (6, 5000, 65535, 5000, False, 15, 8),
    (16, 6000, 3000, 0, True, 14, 8),
    (6, 24000, 65535, 1500, False, 16, 64),      # centroid sums ~2^27: several binade crossings in the float accumulation
    (2, 20000, 65535, 300, False, 18, 8),        # sums just above 2^24
    (16, 8000, 500, 0, True, 17, 2048),          # selector weights up to 2048: sums pass 2^24 as well
])
def test_large_matches_reference_exactly(simctx, ref, dims, n, max_size, retrieve, threaded, seed, max_weight):
    vecs, w = make_vectors(dims, n, seed, "uniform" if dims == 16 else "clumpy", max_weight)
    co_r, k_r, _ = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    co_g, k_g, _ = simctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, n, dims, max_size, retrieve, threaded)
    assert k_g == k_r
    assert np.array_equal(co_g, co_r)


@pytest.mark.parametrize("dims,n,max_size,retrieve,threaded,seed,max_weight", [
    (6, 300, 65535, 40, False, 1, 8), (6, 2000, 65535, 200, False, 2, 8), (2, 1500, 65535, 100, False, 3, 8), (16, 1000, 65535, 100, False, 4, 8),
    (6, 2000, 100, 0, False, 5, 8), (16, 2000, 300, 0, True, 6, 8), (16, 500, 100, 0, True, 7, 8), (6, 1, 65535, 10, False, 8, 8), (2, 64, 65535, 1000, False, 9, 8),
    (6, 5000, 65535, 5000, False, 15, 8), (16, 6000, 3000, 0, True, 14, 8), (6, 24000, 65535, 1500, False, 16, 64), (2, 20000, 65535, 300, False, 18, 8), (16, 8000, 500, 0, True, 17, 2048),
])
def test_fast_builder_within_tolerance_of_reference(fastctx, ref, dims, n, max_size, retrieve, threaded, seed, max_weight):
    """The default builder runs the same algorithm with sums in a parallel order: the tree may differ where two candidates are within a
    rounding of each other, so it is held to what the clustered path's contract needs -- the same number of clusters (+- 1 %) and a
    quantisation error within 1 % of the reference's (the reference moves by as much with its own thread count)."""
    vecs, w = make_vectors(dims, n, seed, "uniform" if (dims == 16 and n > 2000) else "clumpy", max_weight)
    co_r, k_r, _ = ref_clusterize(ref, vecs, w, max_size, retrieve, threaded)
    co_g, k_g, _ = fastctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, n, dims, max_size, retrieve, threaded)
    assert abs(k_g - k_r) <= max(1, k_r // 100)
    assert co_g.max() == k_g - 1 and len(np.unique(co_g)) == k_g
    if k_r > 1 and n > 1:
        d_r, d_g = distortion(vecs, w, co_r), distortion(vecs, w, co_g)
        assert d_g <= d_r * 1.01 + 1e-6, (d_g, d_r)
        # most vectors sit in clusters with exactly the reference's member set
        assert agreement(co_g, co_r) > 0.5


def test_fast_builder_degenerate_inputs(fastctx, ref):
    vecs = np.full((200, 6), 77, np.uint8); w = np.ones(200, np.uint32)
    co_g, k_g, cb_g = fastctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, 200, 6, 65535, 16, False)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, 65535, 16, False)
    assert (k_g, cb_g) == (k_r, cb_r) and np.array_equal(co_g, co_r)
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (7, 6)).astype(np.uint8)
    vecs = np.ascontiguousarray(base[rng.integers(0, 7, 500)]); w = rng.integers(1, 4, 500).astype(np.uint32)
    co_g, k_g, cb_g = fastctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, 500, 6, 65535, 64, False)
    co_r, k_r, cb_r = ref_clusterize(ref, vecs, w, 65535, 64, False)
    assert k_g == k_r and distortion(vecs, w, co_g) <= distortion(vecs, w, co_r) + 1e-6


def test_fast_vq_piecewise_rounds_give_a_valid_tree(sim, monkeypatch):
    """vq_fast_host.h round(): large frontiers are split in four pieces (piece-major slot order across the partitions) so that the host records
    one piece while the device splits the next.  Lowering the threshold makes a small input take that path; the partition must still be one:
    every vector in exactly one cluster, cluster count as asked, distortion within the fast VQ's allowance of the one-piece result."""
    rng = np.random.default_rng(3)
    n = 6000
    vecs = rng.integers(0, 256, (n, 16), dtype=np.uint8); wts = rng.integers(1, 9, n).astype(np.uint32)
    ctx = crn.Context(0, lib=sim)
    try:
        res = []
        for piecewise in (False, True):
            if piecewise:
                monkeypatch.setenv("CRN_B200_VQ_PIPELINE_MIN", "8")
            idx, k, _ = ctx.vq_clusterize(vecs.ctypes.data, wts.ctypes.data, n, 16, 700, 0, True)
            assert idx.shape == (n,) and k == 700 and len(np.unique(idx)) == 700
            res.append(distortion(vecs, wts, idx))
        monkeypatch.delenv("CRN_B200_VQ_PIPELINE_MIN")
        assert res[1] <= res[0] * 1.02
    finally:
        ctx.close()


_RADIX_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import helpers
import crunch2_b200 as crn
rng = np.random.default_rng(11)
n = 5000
base = rng.integers(0, 256, (900, 6)).astype(np.uint8)          # repeated vectors: equal keys, so the id tie-break is exercised
vecs = np.ascontiguousarray(base[rng.integers(0, 900, n)]); w = rng.integers(1, 5, n).astype(np.uint32)
ctx = crn.Context(0, lib=helpers.load_sim())
co, k, cb = ctx.vq_clusterize(vecs.ctypes.data, w.ctypes.data, n, 6, 65535, 300, False)
ctx.close()
sys.stdout.write("%%d %%d %%s" %% (k, cb, co.tobytes().hex()))
"""


def test_fast_vq_split_ranks_radix_sort_equals_std_sort(sim):
    """VqOrderSim::sort_cands (vq_fast_host.h): the split ranks that retrieve_clusters(max) prunes by are ordered with an LSD radix sort above
    4096 candidates and std::sort below.  CRN_B200_RANK_RADIX_MIN is read once per process, so each setting runs in its own interpreter; the
    pruned clustering must be the same array either way."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for radix_min in ("0", "100000000"):
        env = dict(os.environ, CRN_B200_RANK_RADIX_MIN=radix_min)
        r = subprocess.run([sys.executable, "-c", _RADIX_CHILD % (root, os.path.join(root, "tests"))], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and outs[0].split()[0] == "300"
