#!/bin/bash
# One GPU box visit: tests, bench (both arms), phase trace, tolerance at full size, ncu launch list.  Usage: tools/gpu_round.sh TAG
TAG=${1:-r1x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-600 gpurun_out/bench_ref_$TAG.json
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 > gpurun_out/trace_$TAG.log 2>&1; tail -60 gpurun_out/trace_$TAG.log
CRN_BENCH_PROFILING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_$TAG.csv \
   python bench.py --steps 1 --warmup 0 --no-transcode --no-cpu-baseline --no-block-pack > gpurun_out/ncu_bench_$TAG.log 2>&1
python tools/sum_launches.py gpurun_out/launches_c2_$TAG.csv | tee gpurun_out/launch_shares_c2_$TAG.txt
