"""dxt_hc pipeline at scale: python tools/prof_hc.py SIZE [--fmt DXT1] [--faces 1] [--cb 3072] [--ref] -- times
crn_gpu_hc_compress on a synthetic SIZE x SIZE texture with mips and (optionally) the reference's dxt_hc::compress."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import crunch2_b200 as crn  # noqa: E402
import hc_util  # noqa: E402
import quality  # noqa: E402
import blockgen  # noqa: E402
from bench import mip_chain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("size", type=int)
ap.add_argument("--fmt", default="DXT1")
ap.add_argument("--faces", type=int, default=1)
ap.add_argument("--cb", type=int, default=3072)
ap.add_argument("--ref", action="store_true")
ap.add_argument("--threads", type=int, default=15)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
fmt = {"DXT1": 0, "DXT5": 3, "DXT5A": 4, "DXN": 5}[a.fmt]
faces = [mip_chain(blockgen.smooth_image(a.size, a.size, 3000 + f, alpha=True)) for f in range(a.faces)]
blocks, levels = hc_util.hc_layout(faces)
ac = (0, 1) if fmt == 5 else (3, 0)
cbs = (a.cb,) * 4
ctx = crn.Context(0)
for rep in range(a.reps):
    t = time.time(); g = ctx.hc_compress(fmt, blocks, levels, num_faces=a.faces, codebook_sizes=cbs, alpha_components=ac); dt = time.time() - t
    print("gpu rep %d: %.1f ms, %d blocks, %.2f Mtexel/s, info %s, palettes %s" % (rep, dt * 1e3, len(blocks), len(blocks) * 16 / dt / 1e6, g["info"],
          [len(g[k]) for k in ("color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors")]), flush=True)
pg = hc_util.hc_decode(fmt, g, ac)
ch = [0, 1, 2] if fmt in (0, 3) else ([0, 1] if fmt == 5 else [3])
print("gpu psnr %.3f entropy bits %.0f" % (quality.psnr(pg, blocks, ch), hc_util.index_entropy_bits(g, fmt)))
if a.ref:
    import helpers
    ref = helpers.load_ref()
    t = time.time(); r = hc_util.ref_hc_compress(ref, fmt, blocks, levels, num_faces=a.faces, codebook_sizes=cbs, alpha_components=ac, threads=a.threads); dt = time.time() - t
    pr = hc_util.hc_decode(fmt, r, ac)
    print("ref (%d threads): %.1f ms, %.2f Mtexel/s, psnr %.3f entropy bits %.0f palettes %s" % (a.threads + 1, dt * 1e3, len(blocks) * 16 / dt / 1e6, quality.psnr(pr, blocks, ch),
          hc_util.index_entropy_bits(r, fmt), [len(r[k]) for k in ("color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors")]))
    print("tiles equal:", np.array_equal(g["tile_indices"], r["tile_indices"]))
