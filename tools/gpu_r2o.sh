#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clusters.py tests/test_gpu_qdxt.py tests/test_gpu_hc.py -x -q 2>&1 | tail -3
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/r2o_trace.log 2>&1; grep -E "^gpu|endpoint optimisation" gpurun_out/r2o_trace.log | tail -5
CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 128 41 255 2>&1 | grep -E "cluster optimiser  |compress_crn q"
