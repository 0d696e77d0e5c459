"""Phase trace of crn_gpu_compress_crn at BASELINE configs[2] scale (run with CRN_B200_TRACE=1): upload + block gather,
quantiser, writer per pass.  Usage: CRN_B200_TRACE=1 python tools/prof_crn_compress.py [quality]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import blockgen  # noqa: E402
import crunch2_b200 as crn  # noqa: E402
from bench import mip_chain  # noqa: E402

q = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = crn.Context(0)
import numpy as np  # noqa: E402
faces = [[np.ascontiguousarray(l) for l in mip_chain(blockgen.smooth_image(2048, 2048, 3000 + f, alpha=False))] for f in range(6)]
ctx.compress_crn(faces, 0, quality_level=q)
for _ in range(2):
    t0 = time.perf_counter()
    data, rate, _ = ctx.compress_crn(faces, 0, quality_level=q)
    print("compress_crn q%d: %.1f ms, %d bytes, %.3f bpp" % (q, (time.perf_counter() - t0) * 1e3, len(data), rate), file=sys.stderr)
m0 = ctx.pool_mallocs
t0 = time.perf_counter()
data, rate, ql = ctx.compress_crn(faces, 0, target_bitrate=1.25)
print("search 1.25 bpp: %.1f ms, %d bytes, %.3f bpp, quality %d, pool mallocs during the search %d" % ((time.perf_counter() - t0) * 1e3, len(data), rate, ql, ctx.pool_mallocs - m0), file=sys.stderr)
m0 = ctx.pool_mallocs
t0 = time.perf_counter()
data, rate, ql = ctx.compress_crn(faces, 0, target_bitrate=0.8)
print("search 0.80 bpp: %.1f ms, %d bytes, %.3f bpp, quality %d, pool mallocs during the search %d" % ((time.perf_counter() - t0) * 1e3, len(data), rate, ql, ctx.pool_mallocs - m0), file=sys.stderr)
