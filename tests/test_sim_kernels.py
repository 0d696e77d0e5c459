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
