#!/usr/bin/env python3
"""Profiling driver: one synthetic 2048x2048 DXT5 .crn (realistic 1.2 bpp stream), all levels, a few launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import crnsynth  # noqa: E402
import crunch2_b200 as crn  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = crn.Context(0)
data = crnsynth.synth_crn(size, size, "DXT5", seed=4, with_crc=False, n_color_ep=4096, n_color_sel=4096, n_alpha_ep=2048, n_alpha_sel=2048, skew=0.1)
tex = ctx.unpack_begin(data)
out = torch.empty(tex.total_size, dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    tex.unpack_all_device(out, tex.total_size)
ctx.synchronize()
print("ok", tex.total_size)
