#!/usr/bin/env python3
"""Generates tests/golden/crn_writer_golden.json from the UNMODIFIED reference (oracle/_ref/liboracle_ref.so):
  * writer cases: the palettes + indices dxt_hc::compress returns (0 helper threads) for a seeded image, and the .crn the
    reference's crn_compress writes from them -- stored whole (a few KB), so the writer back-end (crn_gpu_crn_write) can
    be checked byte for byte where the reference cannot run;
  * DDS headers: the first 128 bytes of the reference's crn_decompress_crn_to_dds per format / shape.
Run in the build container only:   python tests/golden/make_crn_writer_golden.py"""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402

import blockgen  # noqa: E402
import crnsynth  # noqa: E402
import crunch2_b200 as crn  # noqa: E402
import hc_util  # noqa: E402
import helpers  # noqa: E402
from bench import mip_chain  # noqa: E402
from test_crn_writer_cpu import HC_FMT  # noqa: E402
from test_dds_cpu import ref_to_dds  # noqa: E402

WRITER_CASES = [dict(name="DXT1", w=64, h=64, nlev=3, q=128, seed=71), dict(name="DXT5", w=64, h=48, nlev=4, q=128, seed=71),
                dict(name="DXN_XY", w=40, h=24, nlev=2, q=200, seed=47), dict(name="DXT5A", w=32, h=32, nlev=6, q=64, seed=39)]
DDS_CASES = [dict(fmt=f, w=w, h=h, levels=lv, faces=fc) for f in ("DXT1", "DXT5", "DXN_XY", "DXN_YX", "DXT5A") for (w, h, lv, fc) in ((64, 32, None, 1), (20, 12, 1, 1), (16, 16, 3, 6))]


def b64(a):
    return base64.b64encode(np.ascontiguousarray(a).tobytes()).decode()


def main():
    ref = helpers.load_ref()
    lib = crn.load_library()
    out = {"writer": [], "dds_header": []}
    for c in WRITER_CASES:
        face_levels = [mip_chain(blockgen.smooth_image(c["w"], c["h"], c["seed"], alpha=True))[:c["nlev"]]]
        want, _, _ = helpers.ref_compress(ref, face_levels, helpers.CRN_FMT[c["name"]], file_type=0, quality=c["q"], threads=0)
        p = crn.crn_params(helpers.CRN_FMT[c["name"]], c["w"], c["h"], c["nlev"], 1, c["q"], lib=lib)
        hp = crn.crn_hc_params(p, lib=lib)
        blocks, lv = hc_util.hc_layout(face_levels)
        cbs = (hp.color_endpoint_codebook_size, hp.color_selector_codebook_size, hp.alpha_endpoint_codebook_size, hp.alpha_selector_codebook_size)
        r = hc_util.ref_hc_compress(ref, HC_FMT[c["name"]], blocks, lv, perceptual=bool(hp.perceptual), codebook_sizes=cbs,
                                    deratings=(hp.adaptive_tile_color_psnr_derating, hp.adaptive_tile_alpha_psnr_derating, hp.adaptive_tile_color_alpha_weighting_ratio),
                                    alpha_components=tuple(hp.alpha_component_indices), threads=0)
        out["writer"].append(dict(c, codebook_sizes=list(cbs), file=base64.b64encode(want).decode(),
                                  **{k: b64(r[k]) for k in ("endpoint_indices", "selector_indices", "color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors")}))
    for c in DDS_CASES:
        data = crnsynth.synth_crn(c["w"], c["h"], c["fmt"], levels=c["levels"], faces=c["faces"], seed=1, n_color_ep=16, n_color_sel=16, n_alpha_ep=16, n_alpha_sel=16)
        out["dds_header"].append(dict(c, header=base64.b64encode(ref_to_dds(ref, data)[:128]).decode()))
    with open(os.path.join(HERE, "crn_writer_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["writer"]), "writer cases,", len(out["dds_header"]), "DDS headers")


if __name__ == "__main__":
    main()
