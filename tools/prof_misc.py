"""Small driver for ncu captures of the unpack and mip kernels: python tools/prof_misc.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import blockgen  # noqa: E402
import crunch2_b200 as crn  # noqa: E402

ctx = crn.Context(0)
w = h = 8192
d_blocks = torch.randint(0, 256, ((w // 4) * (h // 4) * 16,), dtype=torch.uint8, device="cuda")
d_rgba = torch.empty(h * w * 4, dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.unpack_image_device(3, d_blocks, w, h, d_rgba, w * 4)
ctx.synchronize()
img = blockgen.smooth_image(4096, 4096, 77, alpha=True)
ctx.generate_mipmaps(img, max_levels=3)
print("done")
