#!/bin/bash
# ncu evidence for profiles/: launch lists and full captures of the dominant kernels.  Usage: tools/gpu_profiles.sh TAG
TAG=${1:-r1x}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CRN_BENCH_PROFILING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_$TAG.csv \
   python bench.py --steps 1 --warmup 0 --no-transcode --no-cpu-baseline --no-block-pack > gpurun_out/ncu_bench_$TAG.log 2>&1
python tools/sum_launches.py gpurun_out/launches_c2_$TAG.csv > gpurun_out/launch_shares_c2_$TAG.txt; head -12 gpurun_out/launch_shares_c2_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_transcode_$TAG.csv python tools/prof_transcode.py 8192 > /dev/null 2>&1
python tools/sum_launches.py gpurun_out/launches_transcode_$TAG.csv > gpurun_out/launch_shares_transcode_$TAG.txt; head -8 gpurun_out/launch_shares_transcode_$TAG.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dxt1_optimize_clusters_kernel -c 1 -f -o gpurun_out/cluster_opt_4096_$TAG python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:vq_stream_cov_kernel -c 2 -f -o gpurun_out/vq_cov_$TAG python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"transcode_tables_kernel|transcode_walk_resolve_kernel" -c 2 -f -o gpurun_out/transcode_wide_$TAG python tools/prof_transcode.py 8192 > /dev/null 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
