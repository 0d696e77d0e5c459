#!/bin/bash
# round 2, visit B: gpu tests (drop-in, dds reader, corrupt files on device), transcode batch residency check, baseline bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/bench_transcode_batch.py 1024 1024 DXT5 64 3 2>&1 | tail -1 | tee gpurun_out/r2b_batch_1024.json
timeout 600 python tools/bench_transcode_batch.py 296 1024 DXT5 64 3 2>&1 | tail -1 | tee gpurun_out/r2b_batch_296.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2500 gpurun_out/r2b_bench.json
