#!/usr/bin/env python3
"""Generates tests/golden/pack_golden.json from the UNMODIFIED reference (oracle/_ref/liboracle_ref.so,
built from /root/reference by oracle/Makefile).  Run in the build container only:
    python tests/golden/make_golden.py
The reference publishes no golden vectors of its own, so these are outputs of the
reference itself on deterministic synthetic inputs (tests/blockgen.py): for every case the sha256 of
the packed block stream produced by dxt_image::init with endpoint caching disabled, plus the first
64 bytes so a failure shows what differs.  Inputs are regenerated from seeds, never stored."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import blockgen  # noqa: E402
import helpers  # noqa: E402

CASES = []
for fam in blockgen.FAMILIES:
    for fmt in (0, 1, 3, 4, 5):
        for q, perc, both in ((4, 1, 1), (4, 0, 0), (3, 1, 1)):
            CASES.append(dict(kind="blocks", family=fam, n=48, seed=101, fmt=fmt, q=q, perc=perc, both=both))
for (w, h, seed) in ((64, 64, 2048), (37, 23, 7), (5, 3, 9), (1, 1, 3), (128, 16, 11)):
    for fmt in (0, 2, 3, 6):
        CASES.append(dict(kind="smooth", w=w, h=h, seed=seed, fmt=fmt, q=4, perc=1, both=1))
    CASES.append(dict(kind="flat", w=w, h=h, seed=seed, fmt=0, q=4, perc=1, both=1))


def case_image(c):
    if c["kind"] == "blocks":
        return helpers.blocks_to_image(blockgen.block_family(c["family"], c["n"], c["seed"]))
    if c["kind"] == "smooth":
        return blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True)
    return blockgen.flat_image(c["w"], c["h"], c["seed"])


def main():
    ref = helpers.load_ref()
    assert ref is not None, "reference library not available"
    out = []
    for c in CASES:
        packed = helpers.ref_pack(ref, c["fmt"], case_image(c), c["q"], c["perc"], c["both"])
        d = dict(c)
        d["sha256"] = helpers.sha(packed)
        d["head"] = packed[:64].tobytes().hex()
        out.append(d)
    with open(os.path.join(HERE, "pack_golden.json"), "w") as f:
        json.dump(dict(source="crnlib 1.2.0 unmodified, dxt_image::init, endpoint caching disabled", cases=out), f, indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
