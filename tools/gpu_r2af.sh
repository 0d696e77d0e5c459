#!/bin/bash
# Pipelined member walk in vq_fast_split: phase trace of the clustered-DDS pass (compare profiles/r2x) + the VQ / DDS GPU tests.
mkdir -p gpurun_out
python tools/prof_qdxt.py 2>&1 | tail -40 > gpurun_out/r2af_prof_qdxt.txt; tail -32 gpurun_out/r2af_prof_qdxt.txt
python -m pytest tests/test_gpu_qdxt.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
