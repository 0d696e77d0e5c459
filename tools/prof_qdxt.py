"""Times the clustered-DDS path (init / pack) on the GPU next to the reference's crn_compress on the host cores,
and checks the tolerance on the same input.  Usage: python tools/prof_qdxt.py [size ...] [--fmt DXT5] [--q 128] [--no-ref]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import blockgen  # noqa: E402
import crunch2_b200 as crn  # noqa: E402
import helpers  # noqa: E402
import quality  # noqa: E402
from bench import mip_chain  # noqa: E402

GPUFMT = dict(DXT1=0, DXT5=3, DXT5A=4, DXN_XY=5, DXN_YX=6)
CH = {0: ([0, 1, 2],), 3: ([0, 1, 2], [3]), 4: ([3],), 5: ([0, 1],), 6: ([0, 1],)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sizes", nargs="*", type=int, default=[1024])
    ap.add_argument("--fmt", default="DXT5")
    ap.add_argument("--q", type=int, default=128)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--threads", type=int, default=max(0, min(15, (os.cpu_count() or 1) - 1)))
    a = ap.parse_args()
    if os.environ.get("CRN_B200_LIB"):                       # e.g. the phase-clock profiling build (make -C crunch2_b200/csrc prof)
        import ctypes
        from crunch2_b200 import api
        ctx = crn.Context(0, lib=api._declare(ctypes.CDLL(os.environ["CRN_B200_LIB"])))
    else:
        ctx = crn.Context(0)
    for size in a.sizes:
        levels = mip_chain(blockgen.smooth_image(size, size, 11, alpha=True))
        texels = sum(l.shape[0] * l.shape[1] for l in levels)
        for rep in range(2):
            l0 = ctx.launch_count
            t0 = time.time(); qd = ctx.qdxt_init(GPUFMT[a.fmt], levels); t1 = time.time(); out = qd.pack(a.q); t2 = time.time()
            info = qd.info(); qd.close()
            print(f"gpu {a.fmt} {size}^2 q{a.q} rep{rep}: init {t1 - t0:.3f}s pack {t2 - t1:.3f}s total {t2 - t0:.3f}s = {texels / (t2 - t0) / 1e6:.2f} Mtexel/s launches {ctx.launch_count - l0} {info}", flush=True)
        if a.no_ref:
            continue
        ref = helpers.load_ref()
        t0 = time.time()
        dds, _, _ = helpers.ref_compress(ref, [levels], helpers.CRN_FMT[a.fmt], file_type=1, quality=a.q, threads=a.threads)
        tr = time.time() - t0
        ref_data = quality.dds_payload(dds)
        src = np.concatenate([quality.image_to_blocks(l) for l in levels])
        f = GPUFMT[a.fmt]
        ga = quality.decode_blocks(out.tobytes(), f); rb = quality.decode_blocks(ref_data, f)
        ps = [(quality.psnr(ga, src, c), quality.psnr(rb, src, c)) for c in CH[f]]
        bg, br = quality.lzma_bits(out.tobytes()), quality.lzma_bits(ref_data)
        print(f"ref {a.fmt} {size}^2 q{a.q}: {tr:.2f}s ({a.threads + 1} threads) = {texels / tr / 1e6:.3f} Mtexel/s; psnr gpu/ref " +
              " ".join(f"{x:.3f}/{y:.3f} (d {x - y:+.4f})" for x, y in ps) + f"; lzma bits gpu {bg} ref {br} ({100.0 * (bg - br) / br:+.3f}%)", flush=True)


if __name__ == "__main__":
    main()
