"""Transcode leg of bench.py alone (CUDA events, L2 flushed), plus SM clock samples.  Usage: python tools/bench_transcode.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import crunch2_b200 as crn  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = crn.Context(0)
dev = torch.device("cuda", 0)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sampler = bench.ClockSampler(0, period=0.05)
sampler.start()
out = bench.run_transcode(ctx, ext, dev, flush, steps, 6459.3)
sampler.stop.set(); sampler.join(timeout=2)
out["clocks"] = sampler.summary()
print(json.dumps(out))
