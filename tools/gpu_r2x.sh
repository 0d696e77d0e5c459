#!/bin/bash
TAG=r2x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, d['e2e'])
print({k:(v.get('within_tolerance') if isinstance(v,dict) else v) for k,v in (d.get('parity') or {}).items()})
for k in ('batch_c5','block_pack','transcode','dxt_hc'):
    v=d.get(k)
    if isinstance(v,dict): print(k, v.get('value'), v.get('ms', v.get('ms_per_texture')))
c=d.get('crn_compress',{}); print('crn pass', c.get('pass_q128'), 'search', c.get('target_1.25bpp'))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_cta_kernel -c 1 -f -o gpurun_out/${TAG}_cluster_opt_cta_c3 python tools/prof_cluster_phases.py 128 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:crn_order_color_kernel -c 1 -f -o gpurun_out/${TAG}_order_color python tools/prof_cluster_phases.py 255 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_kernel -c 1 -f -o gpurun_out/${TAG}_cluster_opt_c2 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
for r in gpurun_out/${TAG}_*.ncu-rep; do python tools/ncu_summary.py $r > ${r%.ncu-rep}_ncu.txt 2>&1; rm -f $r; done
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/${TAG}_phase_trace_c2.txt 2>&1
CRN_B200_TRACE=1 python tools/prof_crn_compress.py > gpurun_out/${TAG}_crn_compress_phase_trace.txt 2>&1; grep -E "^search|compress_crn q" gpurun_out/${TAG}_crn_compress_phase_trace.txt | tail -4
ls gpurun_out/
