"""dxt_hc pipeline (crn_gpu_hc_compress) under the SIMT emulator against the reference's dxt_hc::compress (oracle/_ref,
single-task tree quantiser).  The pipeline is tolerance-class by contract (float sums in a different order); on these inputs it
reproduces the reference's tiles exactly and its palettes / indices within the stated tolerance."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import hc_util
import quality

FMTS = {"DXT1": 0, "DXT5": 3, "DXT5A": 4, "DXN_XY": 5}


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def channels(fmt):
    return ([0, 1, 2] if fmt in (0, 3) else []), ([3] if fmt in (3, 4) else ([0, 1] if fmt in (5, 6) else []))


def compare(ctx, ref, fmt, face_levels, cbs, perceptual=True, threads=0):
    faces = len(face_levels)
    blocks, levels = hc_util.hc_layout(face_levels)
    ac = (0, 1) if fmt in (5, 6) else (3, 0)
    g = ctx.hc_compress(fmt, blocks, levels, num_faces=faces, perceptual=perceptual, codebook_sizes=cbs, alpha_components=ac)
    r = hc_util.ref_hc_compress(ref, fmt, blocks, levels, num_faces=faces, perceptual=perceptual, codebook_sizes=cbs, alpha_components=ac, threads=threads)
    return blocks, g, r, ac


def assert_tolerance(fmt, blocks, g, r, ac, psnr_tol=0.05, bits_tol=0.01):
    """north_star tolerance: PSNR within 0.05 dB, bitrate (index-entropy + palette bits proxy) within 1 %."""
    pg, pr = hc_util.hc_decode(fmt, g, ac), hc_util.hc_decode(fmt, r, ac)
    rgb, al = channels(fmt)
    for ch in (rgb, al):
        if ch:
            a, b = quality.psnr(pg, blocks, ch), quality.psnr(pr, blocks, ch)
            assert a >= b - psnr_tol, (ch, a, b)
    bg, br = hc_util.index_entropy_bits(g, fmt), hc_util.index_entropy_bits(r, fmt)
    assert bg <= br * (1 + bits_tol), (bg, br)


@pytest.mark.parametrize("name", sorted(FMTS))
def test_hc_matches_reference_small(simctx, ref, name):
    from bench import mip_chain
    fmt = FMTS[name]
    img = blockgen.smooth_image(64, 48, 11, alpha=True)
    blocks, g, r, ac = compare(simctx, ref, fmt, [mip_chain(img)[:3]], (48, 48, 24, 48))
    # tile determination is integer + double scoring: identical
    assert np.array_equal(g["block_encodings"], r["block_encodings"])
    assert np.array_equal(g["tile_indices"], r["tile_indices"])
    assert_tolerance(fmt, blocks, g, r, ac)
    # structure of the outputs
    for k in ("color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors"):
        assert abs(len(g[k]) - len(r[k])) <= max(1, len(r[k]) // 50), k
        assert len(np.unique(g[k])) == len(g[k]), k + " has duplicates"
    assert g["endpoint_indices"][:, 3].max() <= 2


def test_hc_cubemap_uniform_metric(simctx, ref):
    """Six faces (the boustrophedon tile order restarts per face), two levels, uniform colour metric."""
    faces = [[blockgen.smooth_image(16, 16, 100 + f, alpha=True), blockgen.smooth_image(8, 8, 200 + f, alpha=True)] for f in range(6)]
    blocks, g, r, ac = compare(simctx, ref, 0, faces, (32, 32, 16, 16), perceptual=False)
    assert np.array_equal(g["block_encodings"], r["block_encodings"])
    assert np.array_equal(g["tile_indices"], r["tile_indices"])
    assert_tolerance(0, blocks, g, r, ac)


def test_hc_flat_image(simctx, ref):
    """Solid 8x8 tiles: zero-variance tree nodes, single-entry clusters, equal endpoints."""
    img = np.zeros((32, 32, 4), np.uint8)
    rng = np.random.default_rng(5)
    img[:] = rng.integers(0, 256, (4, 1, 4, 1, 4), dtype=np.uint8).repeat(8, 1).repeat(8, 3).reshape(32, 32, 4)
    for fmt in (0, 3):
        blocks, g, r, ac = compare(simctx, ref, fmt, [[img]], (16, 16, 16, 16))
        assert np.array_equal(g["block_encodings"], r["block_encodings"])
        assert_tolerance(fmt, blocks, g, r, ac)


def test_hc_rejects_bad_params(simctx):
    blocks = np.zeros((6, 16, 4), np.uint8)
    with pytest.raises(crn.CrnGpuError):
        simctx.hc_compress(0, blocks, [(0, 6, 3, 1.0)])          # odd block width
    with pytest.raises(crn.CrnGpuError):
        simctx.hc_compress(2, np.zeros((4, 16, 4), np.uint8), [(0, 4, 2, 1.0)])   # DXT3 is not a dxt_hc format


def test_blockify_matches_crn_comp_layout(simctx):
    """a24 for the CRN path: crn_comp::quantize_images' gather (levels padded to 8 pixels, edge clamp) == hc_layout."""
    for (w, h) in ((20, 12), (8, 8), (3, 5)):
        img = blockgen.smooth_image(w, h, 3 * w + h, alpha=True)
        want, levels = hc_util.hc_layout([[img]])
        got = np.zeros_like(want)
        bx, by = simctx.blockify(img.ctypes.data, w, h, w * 4, got.ctypes.data, 8)
        simctx.synchronize()
        assert (bx, by) == (((w + 7) & ~7) >> 2, ((h + 7) & ~7) >> 2) and bx == levels[0][2]
        assert np.array_equal(got, want)


def test_hc_cta_per_cluster_equals_warp_per_cluster(simctx, monkeypatch):
    """Large clusters (>= kClusterCoopMinBlocks member blocks) are optimised by a whole CTA (cluster_kernels.cuh, dxt1_optimize_clusters_cta_kernel):
    the candidate errors are integer sums split over the warps, so every output must equal the one-warp-per-cluster kernel's, byte for byte."""
    from bench import mip_chain
    img = blockgen.smooth_image(192, 160, 5, alpha=True)
    blocks, levels = hc_util.hc_layout([mip_chain(img)[:2]])
    outs = []
    for no_coop in ("", "1"):
        if no_coop:
            monkeypatch.setenv("CRN_B200_NO_COOP", "1")
        outs.append(simctx.hc_compress(0, blocks, levels, codebook_sizes=(6, 64, 6, 64)))
    monkeypatch.delenv("CRN_B200_NO_COOP")
    g, w = outs
    sizes = np.bincount(g["endpoint_indices"][:, 0].astype(np.int64))
    assert sizes.max() >= 128, sizes            # the cooperative path really ran
    for k in ("color_endpoints", "color_selectors", "endpoint_indices", "selector_indices", "block_encodings", "tile_indices"):
        assert np.array_equal(g[k], w[k]), k
