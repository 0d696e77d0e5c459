"""CPU tests of the CUDA kernels themselves: the same .cu/.cuh sources compiled by g++ against the
SIMT emulator (tests/cusim) must reproduce the oracle bit for bit.  This is test infrastructure -- it
exists because the build container has no GPU -- and is not a product path."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
from golden.make_golden import case_image

GOLD = helpers.golden("pack_golden.json")["cases"]


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("idx", range(0, len(GOLD), 3))
def test_sim_matches_golden(simctx, idx):
    c = GOLD[idx]
    packed = simctx.pack_image(c["fmt"], case_image(c), crn.PackParams(dxt_quality=c["q"], perceptual=c["perc"], use_both_block_types=c["both"]))
    assert packed[:64].tobytes().hex() == c["head"]
    assert helpers.sha(packed) == c["sha256"]


@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4, 5, 6])
def test_sim_matches_port_on_images(simctx, port, fmt):
    """ragged sizes (edge clamping), 1x1, non-multiple-of-4"""
    for (w, h, seed) in ((1, 1, 1), (3, 5, 2), (9, 4, 3), (33, 17, 4)):
        img = blockgen.smooth_image(w, h, seed, alpha=True)
        a = simctx.pack_image(fmt, img)
        b = helpers.port_pack(port, fmt, img)
        assert (a == b).all(), (fmt, w, h)


def test_sim_grayscale_and_uniform_metrics(simctx, port):
    img = helpers.blocks_to_image(blockgen.block_family("smooth", 40, 77))
    a = simctx.pack_image(0, img, crn.PackParams(perceptual=False))
    b = helpers.port_pack(port, 0, img, 4, 0, 1)
    assert (a == b).all()


@pytest.mark.parametrize("q", [0, 1, 2])
@pytest.mark.parametrize("fmt", [0, 1, 3])
def test_sim_low_quality_levels_match_port_and_reference(simctx, port, q, fmt):
    check_low_quality(simctx, port, q, fmt)


def check_low_quality(simctx, port, q, fmt):
    """crn_dxt_quality superfast / fast / normal (evaluate_solution_fast, crn_dxt1.cpp:1594-1757; fewer probes and passes,
    no median4 / lattice below normal, :715-771, :905): every block family, perceptual and uniform metrics, both block types on and off."""
    ref = helpers.load_ref()
    for kind, seed in (("smooth", 11), ("noise", 12), ("dxt_like", 13), ("dark", 14), ("saturated", 15), ("alpha_mix", 16), ("gray", 17)):
        img = helpers.blocks_to_image(blockgen.block_family(kind, 48, seed))
        for perc, both in ((1, 1), (0, 1), (1, 0)):
            a = simctx.pack_image(fmt, img, crn.PackParams(dxt_quality=q, perceptual=bool(perc), use_both_block_types=bool(both)))
            b = helpers.port_pack(port, fmt, img, q, perc, both)
            bad = helpers.mismatching_blocks(a, b, helpers.bytes_per_block(fmt))
            assert bad.size == 0, (kind, q, fmt, perc, both, bad[:5])
            if ref is not None:
                assert (helpers.ref_pack(ref, fmt, img, q, perc, both) == b).all()


def test_sim_low_quality_grayscale_sampling(simctx, port):
    img = blockgen.smooth_image(40, 24, 5, alpha=False)
    for q in (0, 1, 2):
        a = simctx.pack_image(0, img, crn.PackParams(dxt_quality=q, perceptual=False, grayscale_sampling=True))
        pp = helpers.port_pack_ex(port, 0, img, q, 0, 1, grayscale=1) if hasattr(helpers, "port_pack_ex") else None
        if pp is not None:
            assert (a == pp).all()


@pytest.mark.parametrize("q", [1, 3, 4])
@pytest.mark.parametrize("fmt", [0, 1])
def test_sim_transparent_indices_for_black(simctx, port, q, fmt):
    check_transparent_for_black(simctx, port, q, fmt)


def check_transparent_for_black(simctx, port, q, fmt):
    """cCRNCompFlagUseTransparentIndicesForBlack (try_alpha_as_black_optimization, crn_dxt1.cpp:2001-2079): blocks that mix
    near-black and other colours are retried with black as the transparent index; the better of the two encodings is kept."""
    ref = helpers.load_ref()
    rng = np.random.default_rng(40 + q)
    blocks = blockgen.block_family("smooth", 60, 21 + q).copy()
    mask = rng.random(blocks.shape[:2]) < 0.3                      # ~30 % of the pixels go near black
    blocks[mask, :3] = rng.integers(0, 5, (int(mask.sum()), 3), dtype=np.uint8)
    blocks[:5, :, :3] = 2                                           # all-dark blocks: not retried
    if fmt == 1:
        blocks[5:10, :4, 3] = 0                                     # blocks with real transparency: not retried either
    img = helpers.blocks_to_image(blocks)
    changed = 0
    for perc in (1, 0):
        a = simctx.pack_image(fmt, img, crn.PackParams(dxt_quality=q, perceptual=bool(perc), use_transparent_indices_for_black=True))
        b = helpers.port_pack(port, fmt, img, q, perc, 1, 128, 1)
        bad = helpers.mismatching_blocks(a, b, 8)
        assert bad.size == 0, (q, fmt, perc, bad[:5])
        plain = helpers.port_pack(port, fmt, img, q, perc, 1, 128, 0)
        changed += int((plain != b).any())
        if ref is not None:
            assert (helpers.ref_pack(ref, fmt, img, q, perc, 1, 128, 1) == b).all()
    assert changed, "the flag never changed a block: the test input does not exercise the path"
