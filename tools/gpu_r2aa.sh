#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
nproc
CRN_B200_POOL_SPIN_MS=${SPIN:-} timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 --c5-textures 256 --no-cpu-baseline > gpurun_out/r2aa_bench_${N}gpu.json 2> gpurun_out/r2aa_bench_${N}gpu.err; tail -2 gpurun_out/r2aa_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2aa_bench_${N}gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}, d.get('e2e'))
for k in ('batch_c5','dxt_hc_sharded'):
    v=d.get(k)
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,list))})
print((d.get('parity') or {}).get('all_within_tolerance'), (d.get('cpu_baseline') or {}).get('cores'))
PY
