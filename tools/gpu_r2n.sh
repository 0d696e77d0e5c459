#!/bin/bash
# round-2 evidence: full bench line, launch list of the headline step, ncu --set full of the dominant kernels
TAG=r2n
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(json.dumps(d.get('roofline'))[:1500])
print({k:(v.get('within_tolerance') if isinstance(v,dict) else v) for k,v in (d.get('parity') or {}).items()})
print(d.get('cpu_baseline'))
print({k:(d[k].get('value') if isinstance(d[k],dict) else None) for k in ('batch_c5','block_pack','transcode','unpack','mipgen','dxt_hc','crn_compress') if k in d})
PY
CRN_BENCH_PROFILING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c2.csv \
   python bench.py --steps 1 --warmup 0 --no-transcode --no-cpu-baseline --no-block-pack > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/sum_launches.py gpurun_out/${TAG}_launches_c2.csv > gpurun_out/${TAG}_launch_shares_c2.txt; head -14 gpurun_out/${TAG}_launch_shares_c2.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_kernel -c 1 -f -o gpurun_out/${TAG}_cluster_opt_c2 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_cta_kernel -c 1 -f -o gpurun_out/${TAG}_cluster_opt_cta_c3 python tools/prof_cluster_phases.py 128 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt5_optimize_clusters_kernel -c 1 -f -o gpurun_out/${TAG}_alpha_cluster_c2 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:vq_fast_split_kernel -c 6 -f -o gpurun_out/${TAG}_vq_fast python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
for r in gpurun_out/${TAG}_*.ncu-rep; do python tools/ncu_summary.py $r > ${r%.ncu-rep}_ncu.txt 2>&1; rm -f $r; done
ls -la gpurun_out/
