"""Cluster (N-pixel) optimisers under the SIMT emulator vs the oracle port: for every cluster the shared
endpoints, the optimiser's error and every member block's packed element must equal what
dxt1_/dxt5_endpoint_optimizer produce on the concatenated pixels (qdxt1/qdxt5::pack_endpoints_task)."""
import ctypes

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers

P = helpers.P


class Prm(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in "quality perceptual pixels_have_alpha use_alpha_blocks alpha_threshold grayscale transparent_for_black force_alpha_blocks".split()]


class Res(ctypes.Structure):
    _fields_ = [("error", ctypes.c_uint64), ("low", ctypes.c_uint16), ("high", ctypes.c_uint16), ("alpha_block", ctypes.c_uint8)]


def make_clusters(n_blocks, sizes, seed):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_blocks)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    assert offs[-1] <= n_blocks
    return offs, perm[:offs[-1]].astype(np.uint32)


def port_color_cluster(port, px, q, perc, uab, pha):
    n = len(px)
    p = Prm(q, perc, pha, uab, 128 if uab else 0, 0, 0, 0); r = Res(); sel = np.zeros(n, np.uint8)
    port.op_dxt1_optimize(P(px), n, ctypes.byref(p), ctypes.byref(r), P(sel))
    return r.low, r.high, r.error, sel


def run_color(ctx, port, blocks, offs, members, q=4, perc=1, uab=0, dxt1a=False):
    n_blocks = len(blocks)
    out = np.zeros((n_blocks, 8), np.uint8)
    ep = np.zeros(len(offs) - 1, np.uint32); err = np.zeros(len(offs) - 1, np.uint64)
    ctx.optimize_clusters("color", blocks.ctypes.data, n_blocks, offs.ctypes.data, members.ctypes.data, len(offs) - 1, int(offs[-1]),
                          out.ctypes.data, 8, 0, crn.PackParams(dxt_quality=q, perceptual=perc, use_both_block_types=uab), dxt1a=dxt1a,
                          d_endpoints=ep.ctypes.data, d_error=err.ctypes.data)
    ctx.synchronize()
    for c in range(len(offs) - 1):
        m = members[offs[c]:offs[c + 1]]
        px = np.ascontiguousarray(blocks[m].reshape(-1, 4))
        pha = int(dxt1a and uab and (px[:, 3] < 128).any())
        lo, hi, e, sel = port_color_cluster(port, px, q, perc, uab, pha)
        assert (int(ep[c]) & 0xffff, int(ep[c]) >> 16) == (lo, hi), (c, len(m))
        if not (pha and (px[:, 3] >= 128).sum() == 0):
            assert int(err[c]) == e, (c, len(m))
        for k, b in enumerate(m):
            s = sel[16 * k:16 * k + 16]
            bits = sum(int(s[i]) << (2 * i) for i in range(16))
            want = np.frombuffer(int(lo | (hi << 16) | (bits << 32)).to_bytes(8, "little"), np.uint8)
            assert (out[b] == want).all(), (c, k)


@pytest.fixture(scope="module", params=["member-order sums", "lane-parallel sums"])
def simctx(sim, request):
    """crn_gpu_set_vq_mode(exact) keeps every float sum of the optimiser in the reference's member order -- the bit-exact contract.  The default
    forms the O(U) sums of clusters with more than 64 unique colours lane-parallel (dxt1_opt.cuh, Dxt1Params::parallel_sums; tolerance class
    by contract); on these inputs the rounding noise changes no decision, so the same equalities are asserted for both."""
    ctx = crn.Context(0, lib=sim)
    ctx.set_vq_mode(request.param == "member-order sums")
    yield ctx
    ctx.close()


@pytest.mark.parametrize("family", ["smooth", "noise", "four", "solid", "dark", "dxt_like"])
def test_color_clusters_match_port(simctx, port, family):
    blocks = blockgen.block_family(family, 120, 61)
    offs, members = make_clusters(120, [1, 2, 3, 5, 8, 13, 40], 7)
    run_color(simctx, port, blocks, offs, members, q=4, perc=1, uab=0)      # DXT5 colour element: hc evaluator
    run_color(simctx, port, blocks, offs, members, q=4, perc=0, uab=1)      # DXT1: both block types
    run_color(simctx, port, blocks, offs, members, q=3, perc=1, uab=0)


def test_color_clusters_dxt1a(simctx, port):
    blocks = blockgen.block_family("alpha_mix", 64, 5)
    offs, members = make_clusters(64, [1, 4, 9, 20], 3)
    run_color(simctx, port, blocks, offs, members, q=4, perc=1, uab=1, dxt1a=True)


@pytest.mark.parametrize("family", ["smooth", "noise", "two", "gray"])
def test_alpha_clusters_match_port(simctx, port, family):
    blocks = blockgen.block_family(family, 200, 33)
    offs, members = make_clusters(200, [1, 2, 7, 30, 100], 9)
    n_blocks = len(blocks)
    for comp, q, both in ((3, 4, 1), (0, 3, 0), (1, 2, 1)):
        out = np.zeros((n_blocks, 16), np.uint8)
        ep = np.zeros(len(offs) - 1, np.uint32); err = np.zeros(len(offs) - 1, np.uint64)
        simctx.optimize_clusters("alpha", blocks.ctypes.data, n_blocks, offs.ctypes.data, members.ctypes.data, len(offs) - 1, int(offs[-1]),
                                 out.ctypes.data, 16, 8, crn.PackParams(dxt_quality=q, use_both_block_types=both), component=comp,
                                 d_endpoints=ep.ctypes.data, d_error=err.ctypes.data)
        simctx.synchronize()
        for c in range(len(offs) - 1):
            m = members[offs[c]:offs[c + 1]]
            px = np.ascontiguousarray(blocks[m].reshape(-1, 4)); n = len(px)
            f = ctypes.c_uint8(); s = ctypes.c_uint8(); e = ctypes.c_uint64(); bt = ctypes.c_uint8(); sel = np.zeros(n, np.uint8)
            port.op_dxt5_optimize(P(px), n, comp, q, both, ctypes.byref(f), ctypes.byref(s), P(sel), ctypes.byref(e), ctypes.byref(bt))
            assert (int(ep[c]) & 0xff, int(ep[c]) >> 8) == (f.value, s.value), (family, comp, c)
            assert int(err[c]) == e.value
            for k, b in enumerate(m):
                bits = sum(int(sel[16 * k + i]) << (3 * i) for i in range(16))
                want = np.frombuffer(int(f.value | (s.value << 8) | (bits << 16)).to_bytes(8, "little"), np.uint8)
                assert (out[b, 8:] == want).all() and (out[b, :8] == 0).all()


# ---- selector re-vote (qdxt1/qdxt5::optimize_selectors_task) ----------------------------------------
def revote_reference(blocks, elems, offs, members, is_alpha, perceptual=1, comp=3, threshold=0):
    """Straight numpy restatement of crn_qdxt1.cpp:714-865 / crn_qdxt5.cpp:578-687 for the test."""
    out = elems.copy()
    for c in range(len(offs) - 1):
        m = members[offs[c]:offs[c + 1]]
        if len(m) <= 1:
            continue
        cats = {0: [], 1: []}
        for b in m:
            e = int(elems[b])
            if is_alpha:
                cats[int((e & 0xff) <= ((e >> 8) & 0xff))].append(b)
            else:
                lo, hi = e & 0xffff, (e >> 16) & 0xffff
                if lo > hi:
                    cats[0].append(b)
                elif not (threshold > 0 and (blocks[b][:, 3] < threshold).any()):
                    cats[1].append(b)
        for cat, bl in cats.items():
            if len(bl) <= 1:
                continue
            ns = 8 if is_alpha else (3 if cat else 4)
            tot = np.zeros((16, ns), np.int64)
            for b in bl:
                e = int(elems[b])
                if is_alpha:
                    l, h = e & 0xff, (e >> 8) & 0xff
                    if l > h:
                        vals = [l, h] + [(l * (7 - k) + h * k) // 7 for k in range(1, 7)]
                    else:
                        vals = [l, h] + [(l * (5 - k) + h * k) // 5 for k in range(1, 5)] + [0, 255]
                    v = blocks[b][:, comp].astype(np.int64)
                    tot += (v[:, None] - np.array(vals)[None, :]) ** 2
                else:
                    lo, hi = e & 0xffff, (e >> 16) & 0xffff
                    def up(c):
                        r, g, bb = (c >> 11) & 31, (c >> 5) & 63, c & 31
                        return np.array([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (bb << 3) | (bb >> 2)])
                    c0, c1 = up(lo), up(hi)
                    pal = [c0, c1, (c0 * 2 + c1) // 3, (c1 * 2 + c0) // 3] if lo > hi else [c0, c1, (c0 + c1) >> 1]
                    px = blocks[b][:, :3].astype(np.int64)
                    w = np.array([8, 25, 1]) if perceptual else np.array([1, 1, 1])
                    for k in range(ns):
                        tot[:, k] += (((px - pal[k][None, :]) ** 2) * w).sum(1)
            best = tot.argmin(1)
            bits = sum(int(best[i]) << ((3 if is_alpha else 2) * i) for i in range(16))
            for b in bl:
                e = int(out[b])
                out[b] = (e & 0xffff) | (bits << 16) if is_alpha else (e & 0xffffffff) | (bits << 32)
    return out


@pytest.mark.parametrize("is_alpha", [False, True])
def test_selector_revote_matches_restatement(simctx, port, is_alpha):
    blocks = blockgen.block_family("smooth", 300, 71)
    img = helpers.blocks_to_image(blocks)
    packed = helpers.port_pack(port, 4 if is_alpha else 0, img, 4, 1, 1)
    elems = packed.view(np.uint64).copy()
    offs, members = make_clusters(300, [1, 2, 3, 9, 30, 77, 150], 13)
    want = revote_reference(blocks, elems, offs, members, is_alpha, threshold=128)   # default params: use_both_block_types keeps the 128 threshold
    got = elems.copy()
    simctx.optimize_selectors("alpha" if is_alpha else "color", blocks.ctypes.data, 300, offs.ctypes.data, members.ctypes.data, len(offs) - 1,
                              got.ctypes.data, 8, 0, crn.PackParams(perceptual=True), component=3)
    simctx.synchronize()
    assert (got == want).all()
    assert (got != elems).any()


def test_color_clusters_many_unique_colours(simctx, port):
    """Clusters with hundreds to ~1600 unique colours: the evaluation colours are re-ordered heaviest-first and candidates are
    endpoints, error and every selector must still equal the port's."""
    blocks = np.concatenate([blockgen.block_family("noise", 70, 11), blockgen.block_family("smooth", 70, 12)])
    offs, members = make_clusters(140, [100, 25, 14], 9)
    run_color(simctx, port, blocks, offs, members, q=4, perc=1, uab=0)      # hc evaluator
    run_color(simctx, port, blocks, offs, members, q=4, perc=0, uab=1)      # both block types
    run_color(simctx, port, blocks, offs, members, q=3, perc=1, uab=1)


@pytest.mark.parametrize("q", [0, 1, 2, 3])
def test_color_clusters_lower_quality_levels_match_port(simctx, port, q):
    """the N-pixel optimiser below uber quality: evaluate_solution_fast for 0-2 (crn_dxt1.cpp:1594-1757), fewer probes / passes (:715-771)"""
    blocks = blockgen.block_family("smooth", 60, 31 + q)
    offs, members = make_clusters(60, [1, 2, 3, 5, 8, 13], 40 + q)
    run_color(simctx, port, blocks, offs, members, q=q, perc=1, uab=0)
    run_color(simctx, port, blocks, offs, members, q=q, perc=0, uab=1)


def test_alpha_cluster_with_a_weight_that_wraps_the_reference_product(simctx, port):
    """crn_dxt5a.cpp:225-228 forms d*d*weight in 32-bit ints: one value carried by more than 33025 pixels can wrap it.  The kernel scores
    candidates from prefix sums only while no weight can do that and falls back to the per-value loop otherwise -- both against the port."""
    rng = np.random.default_rng(5)
    n_blocks = 2400
    blocks = np.zeros((n_blocks, 16, 4), np.uint8)
    blocks[..., 3] = 17                                         # 38400 pixels, ~36000 of them one value
    noisy = rng.choice(n_blocks * 16, 2400, replace=False)
    blocks.reshape(-1, 4)[noisy, 3] = rng.integers(0, 256, 2400)
    offs = np.array([0, n_blocks], np.uint32); members = np.arange(n_blocks, dtype=np.uint32)
    for q, both in ((4, 1), (3, 0)):
        out = np.zeros((n_blocks, 8), np.uint8)
        ep = np.zeros(1, np.uint32); err = np.zeros(1, np.uint64)
        simctx.optimize_clusters("alpha", blocks.ctypes.data, n_blocks, offs.ctypes.data, members.ctypes.data, 1, n_blocks,
                                 out.ctypes.data, 8, 0, crn.PackParams(dxt_quality=q, use_both_block_types=both), component=3,
                                 d_endpoints=ep.ctypes.data, d_error=err.ctypes.data)
        simctx.synchronize()
        px = np.ascontiguousarray(blocks.reshape(-1, 4)); n = len(px)
        f = ctypes.c_uint8(); s = ctypes.c_uint8(); e = ctypes.c_uint64(); bt = ctypes.c_uint8(); sel = np.zeros(n, np.uint8)
        port.op_dxt5_optimize(P(px), n, 3, q, both, ctypes.byref(f), ctypes.byref(s), P(sel), ctypes.byref(e), ctypes.byref(bt))
        assert (int(ep[0]) & 0xff, int(ep[0]) >> 8) == (f.value, s.value)
        assert int(err[0]) == e.value
        got_sel = np.zeros(n, np.uint8)
        bits = out.view(np.uint64).ravel() >> np.uint64(16)
        for i in range(16):
            got_sel[i::16] = (bits >> np.uint64(3 * i)) & np.uint64(7)
        assert (got_sel == sel).all()
