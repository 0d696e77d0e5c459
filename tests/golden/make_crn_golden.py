#!/usr/bin/env python3
"""Generates tests/golden/crn/*.crn and tests/golden/crn_golden.json from the UNMODIFIED reference
(oracle/_ref): small .crn files written by the reference's own compressor (crn_compress, CRN file type)
and, for each, the sha256 of every level/face as decoded by the reference's crnd_unpack_level.
Run in the build container only:  python tests/golden/make_crn_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import blockgen  # noqa: E402
import helpers  # noqa: E402
from bench import mip_chain  # noqa: E402

CASES = [  # name, crn_format, (w, h), faces, mips, quality
    ("dxt1_64_mips", 0, (64, 64), 1, True, 128),
    ("dxt1_q255_48x80", 0, (48, 80), 1, True, 255),
    ("dxt5_72x40_mips", 2, (72, 40), 1, True, 128),
    ("dxn_xy_32", 7, (32, 32), 1, False, 128),
    ("dxn_yx_128x64_mips", 8, (128, 64), 1, True, 96),
    ("dxt5a_20x12_mips", 9, (20, 12), 1, True, 128),
    ("dxt1_cube_32_mips", 0, (32, 32), 6, True, 128),
    ("dxt5_ccxy_64x32", 3, (64, 32), 1, False, 128),
    ("dxt1_5x3", 0, (5, 3), 1, False, 128),
    ("dxt5_cube_16", 2, (16, 16), 6, True, 200),
]


def main():
    ref = helpers.load_ref()
    assert ref is not None
    out = []
    for name, fmt, (w, h), faces, mips, q in CASES:
        imgs = []
        for f in range(faces):
            base = blockgen.smooth_image(w, h, 100 + f, alpha=True)
            imgs.append(mip_chain(base) if mips else [base])
        crn, _, _ = helpers.ref_compress(ref, imgs, fmt, quality=q)
        assert crn is not None, name
        with open(os.path.join(HERE, "crn", name + ".crn"), "wb") as fh:
            fh.write(crn)
        levels = helpers.ref_unpack_all(ref, crn)
        out.append(dict(name=name, format=fmt, width=w, height=h, faces=faces, levels=len(levels), size=len(crn),
                        sha256=[[helpers.sha(__import__("numpy").frombuffer(fc, "uint8")) for fc in lv] for lv in levels]))
    with open(os.path.join(HERE, "crn_golden.json"), "w") as fh:
        json.dump(dict(source="crnlib 1.2.0 unmodified: crn_compress -> crnd_unpack_level", cases=out), fh, indent=0)
    print("wrote", len(out), "files", sum(c["size"] for c in out), "bytes")


if __name__ == "__main__":
    main()
