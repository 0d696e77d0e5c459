#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; tail -3 gpurun_out/r2u_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_2gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}, d.get('e2e'))
for k in ('batch_c5','dxt_hc_sharded','crn_compress'):
    v=d.get(k)
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,list))})
print((d.get('parity') or {}).get('all_within_tolerance'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-300
