#!/bin/bash
TAG=r2z
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, d['e2e'])
print({k:(v.get('within_tolerance') if isinstance(v,dict) else v) for k,v in (d.get('parity') or {}).items()})
for k in ('batch_c5','block_pack','transcode','dxt_hc'):
    v=d.get(k)
    if isinstance(v,dict): print(k, v.get('value'), v.get('ms', v.get('ms_per_texture')))
c=d.get('crn_compress',{}); print('crn pass', c.get('pass_q128'), 'search', c.get('target_1.25bpp'))
print(d['roofline']['ms'], d['roofline']['issue']['frac'], d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/${TAG}_bench_reference.json
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/${TAG}_phase_trace_c2.txt 2>&1; grep -E "^gpu" gpurun_out/${TAG}_phase_trace_c2.txt | cut -c1-110
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_kernel -c 1 -f -o gpurun_out/${TAG}_cluster_opt_c2 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
for r in gpurun_out/${TAG}_*.ncu-rep; do python tools/ncu_summary.py $r > ${r%.ncu-rep}_ncu.txt 2>&1; rm -f $r; done
grep -E "gpu__time|issue_active|thread_inst_executed_per|long_scoreboard|inst_executed.sum|pipe_fma|pipe_alu|dram__bytes" gpurun_out/${TAG}_cluster_opt_c2_ncu.txt
