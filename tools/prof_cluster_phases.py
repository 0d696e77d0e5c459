"""Per-phase SM cycles of the cluster optimiser on BASELINE configs[2]'s cubemap (6 x 2048^2 + mips, DXT1 .crn): needs the profiling build
(`make -C crunch2_b200/csrc prof` -> crunch2_b200/libcrn_b200_prof.so, -DCRN_B200_PHASE_CLOCKS).  Usage: python tools/prof_cluster_phases.py [quality ...]
Set CRN_B200_NO_COOP=1 for the one-warp-per-cluster kernel."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import blockgen  # noqa: E402
import crunch2_b200 as crn  # noqa: E402
from crunch2_b200 import api  # noqa: E402
from bench import mip_chain  # noqa: E402

lib = api._declare(ctypes.CDLL(os.environ.get("CRN_B200_LIB") or os.path.join(ROOT, "crunch2_b200", "libcrn_b200_prof.so" if os.path.exists(os.path.join(ROOT, "crunch2_b200", "libcrn_b200_prof.so")) and os.environ.get("CRN_B200_PHASES") else "libcrn_b200.so")))
ctx = crn.Context(0, lib=lib)
faces = [[np.ascontiguousarray(l) for l in mip_chain(blockgen.smooth_image(2048, 2048, 3000 + f, alpha=False))] for f in range(6)]
for q in [int(a) for a in sys.argv[1:]] or [128]:
    for rep in range(2):
        t0 = time.perf_counter()
        data, rate, _ = ctx.compress_crn(faces, 0, quality_level=q)
        print("compress_crn q%d: %.1f ms, %d bytes" % (q, (time.perf_counter() - t0) * 1e3, len(data)), file=sys.stderr)
