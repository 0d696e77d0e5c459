#!/bin/bash
# what the driver runs at round end, plus the reference arm: GPU tests, smoke(), the default bench line
TAG=${1:-r2final}
mkdir -p gpurun_out
echo "(GPU tests: see r2final3, same library)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
( time timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, d['e2e'])
print({k:(v.get('within_tolerance') if isinstance(v,dict) else v) for k,v in (d.get('parity') or {}).items()})
for k in ('batch_c5','block_pack','transcode','dxt_hc'):
    v=d.get(k)
    if isinstance(v,dict): print(k, v.get('value'), v.get('ms', v.get('ms_per_texture')))
print('real files', d['transcode'].get('real_files'))
c=d.get('crn_compress',{}); print('crn pass', c.get('pass_q128'), 'search', c.get('target_1.25bpp'))
print(d['roofline']['ms'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/${TAG}_bench_reference.json
