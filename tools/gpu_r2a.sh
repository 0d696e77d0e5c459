#!/bin/bash
# round 2, visit A: gpu tests + lane-per-stream batch sweeps
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "1024 1024 DXT5" "4096 512 DXT5" "16384 256 DXT5" "4096 1024 DXT5" "4096 512 DXT1"; do
  set -- $cfg
  timeout 600 python tools/bench_transcode_batch.py $1 $2 $3 64 3 2>&1 | tail -1 | tee -a gpurun_out/r2a_streams.jsonl
done
CRN_B200_STREAMS_MIN=100000000 timeout 600 python tools/bench_transcode_batch.py 1024 1024 DXT5 64 3 2>&1 | tail -1 | tee -a gpurun_out/r2a_streams_off.jsonl
