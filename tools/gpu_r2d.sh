#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_qdxt.py -x -q 2>&1 | tail -4
CRN_B200_TRACE=1 CRN_B200_TRACE_ROUNDS=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 > gpurun_out/r2d_trace.log 2>&1; grep -E "^gpu|^ref|vq_fast<16> n=|pack:|init:" gpurun_out/r2d_trace.log | tail -26
grep "round F" gpurun_out/r2d_trace.log | grep "<16>" | tail -30
timeout 900 python bench.py --steps 5 --warmup 3 --no-transcode --no-hc --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cut -c1-700 gpurun_out/r2d_bench.json
