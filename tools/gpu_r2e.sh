#!/bin/bash
mkdir -p gpurun_out
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/r2e_trace.log 2>&1; grep -E "^gpu|vq_fast<16> n=|pack:|init:" gpurun_out/r2e_trace.log | tail -16
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(json.dumps(d.get('parity'),indent=1)[:6000])
print(d.get('cpu_baseline'))
print({k:(d[k].get('value') if isinstance(d[k],dict) else None) for k in ('batch_c5','block_pack','transcode','unpack','mipgen','dxt_hc','crn_compress') if k in d})
print(d.get('batch_c5',{}).get('reference'))
PY
