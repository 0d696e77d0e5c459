#!/bin/bash
TAG=r2t
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print({k:(v.get('within_tolerance') if isinstance(v,dict) else v) for k,v in (d.get('parity') or {}).items()})
for k in ('batch_c5','block_pack','transcode','unpack','mipgen','dxt_hc','crn_compress','dxt_hc_sharded'):
    v=d.get(k)
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,list))})
print(d.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
