"""Golden vectors of the reference (tests/golden/misc_golden.json, written by tests/golden/make_misc_golden.py from the
unmodified reference) for the paths that are bit-exact: mip chains, DXTn unpacking, and the tile determination of dxt_hc.
These hold without oracle/_ref; the kernels run under the SIMT emulator."""
import hashlib
import json
import os

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import hc_util

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "misc_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("c", G["mips"], ids=lambda c: "%dx%d-%s" % (c["w"], c["h"], c["filt"]))
def test_mip_chain_golden(simctx, c):
    img = blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True)
    got = simctx.generate_mipmaps(img, filter=c["filt"], filter_scale=c["scale"], srgb=c["srgb"], wrapping=c["wrap"], num_comps=c["comps"])
    assert [sha(l) for l in got[1:]] == c["levels"]


@pytest.mark.parametrize("c", G["unpack"], ids=lambda c: "fmt%d-%dx%d" % (c["fmt"], c["w"], c["h"]))
def test_unpack_golden(simctx, c):
    from test_unpack_cpu import random_blocks
    blocks = random_blocks(c["fmt"], c["w"], c["h"], c["seed"])
    assert sha(simctx.unpack_image(c["fmt"], blocks, c["w"], c["h"])) == c["sha256"]


@pytest.mark.parametrize("c", G["hc"], ids=lambda c: "fmt%d" % c["fmt"])
def test_hc_golden(simctx, c):
    from bench import mip_chain
    img = blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True)
    blocks, levels = hc_util.hc_layout([mip_chain(img)[:3]])
    ac = (0, 1) if c["fmt"] in (5, 6) else (3, 0)
    g = simctx.hc_compress(c["fmt"], blocks, levels, codebook_sizes=tuple(c["cbs"]), alpha_components=ac)
    assert sha(g["block_encodings"]) == c["encodings"] and sha(g["tile_indices"]) == c["tiles"]      # exact by construction
    sizes = [len(g[k]) for k in ("color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors")]
    for a, b in zip(sizes, c["sizes"]):
        assert abs(a - b) <= max(1, b // 50)
    # on these inputs the tolerance-class part reproduces the single-task reference too
    assert sha(g["endpoint_indices"]) == c["endpoint_indices"]
    assert sha(g["color_endpoints"]) == c["color_endpoints"] and sha(g["alpha_endpoints"]) == c["alpha_endpoints"]
