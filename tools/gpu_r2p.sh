#!/bin/bash
mkdir -p gpurun_out
for w in 1 2 3 4; do
python - <<PY
import sys, json, torch
sys.argv=['bench.py']
import bench
bench.C5_WORKERS=$w
import crunch2_b200 as crn
dev=torch.device('cuda:0'); ctx=crn.Context(0)
r=bench.run_batch_c5(ctx, dev, 0, 1, 96, with_reference=False)
print($w, {k:r[k] for k in ('value','ms_per_texture','workers_per_gpu')})
PY
done
CRN_B200_TRACE=1 python tools/prof_crn_compress.py 2>&1 | grep -E "^search|compress_crn q"
