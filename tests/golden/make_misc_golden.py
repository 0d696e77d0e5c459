#!/usr/bin/env python3
"""Generates tests/golden/misc_golden.json from the UNMODIFIED reference (oracle/_ref/liboracle_ref.so): sha256 of
  * every level of image_utils::resample'd mip chains (task-pool resampler, the form crn_compress_mip_chain runs),
  * dxt_image::unpack of seeded random blocks in every format,
  * dxt_hc::compress (0 helper threads): block encodings + tile indices (the exact part) and the palette sizes.
Inputs are regenerated from seeds (tests/blockgen.py, numpy default_rng), never stored.  Run in the build container only:
    python tests/golden/make_misc_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402

import blockgen  # noqa: E402
import hc_util  # noqa: E402
import helpers  # noqa: E402
from test_mip_cpu import ref_mips  # noqa: E402
from test_unpack_cpu import random_blocks, ref_unpack  # noqa: E402

MIP_CASES = [dict(w=64, h=64, seed=81, filt="kaiser", scale=0.9, srgb=True, wrap=False, comps=4),
             dict(w=37, h=19, seed=54, filt="kaiser", scale=0.9, srgb=True, wrap=False, comps=4),
             dict(w=48, h=32, seed=5, filt="lanczos4", scale=1.0, srgb=False, wrap=True, comps=3),
             dict(w=40, h=24, seed=57, filt="mitchell", scale=1.0, srgb=True, wrap=False, comps=4)]
UNPACK_CASES = [dict(fmt=f, w=w, h=h, seed=100 * f + w) for f in range(7) for (w, h) in ((32, 16), (13, 9))]
HC_CASES = [dict(fmt=0, w=64, h=48, seed=11, cbs=(48, 48, 24, 48)), dict(fmt=3, w=64, h=48, seed=11, cbs=(48, 48, 24, 48)),
            dict(fmt=5, w=32, h=32, seed=12, cbs=(32, 32, 32, 32))]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hc_inputs(c):
    from bench import mip_chain
    img = blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True)
    return hc_util.hc_layout([mip_chain(img)[:3]])


def main():
    ref = helpers.load_ref()
    assert ref is not None, "reference library not available"
    out = {"mips": [], "unpack": [], "hc": []}
    for c in MIP_CASES:
        img = blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True)
        levels = ref_mips(ref, img, filt=c["filt"], scale=c["scale"], srgb=c["srgb"], wrap=c["wrap"], comps=c["comps"])
        out["mips"].append(dict(c, levels=[sha(l) for l in levels[1:]]))
    for c in UNPACK_CASES:
        blocks = random_blocks(c["fmt"], c["w"], c["h"], c["seed"])
        out["unpack"].append(dict(c, sha256=sha(ref_unpack(ref, c["fmt"], blocks, c["w"], c["h"]))))
    for c in HC_CASES:
        blocks, levels = hc_inputs(c)
        ac = (0, 1) if c["fmt"] in (5, 6) else (3, 0)
        r = hc_util.ref_hc_compress(ref, c["fmt"], blocks, levels, codebook_sizes=c["cbs"], alpha_components=ac)
        out["hc"].append(dict(c, encodings=sha(r["block_encodings"]), tiles=sha(r["tile_indices"]),
                              sizes=[len(r[k]) for k in ("color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors")],
                              endpoint_indices=sha(r["endpoint_indices"]), color_endpoints=sha(r["color_endpoints"]), alpha_endpoints=sha(r["alpha_endpoints"])))
    with open(os.path.join(HERE, "misc_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["mips"]), "mip,", len(out["unpack"]), "unpack,", len(out["hc"]), "hc cases")


if __name__ == "__main__":
    main()
