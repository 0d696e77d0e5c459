#!/bin/bash
# round 2, visit C: the single-launch vector quantiser on the device
mkdir -p gpurun_out
python -m pytest tests/test_gpu_qdxt.py tests/test_gpu_dropin.py tests/test_gpu_crn.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -8
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 > gpurun_out/r2c_trace_fast.log 2>&1; grep -E "vq_fast|^gpu|^ref|qdxt" gpurun_out/r2c_trace_fast.log | tail -40
CRN_B200_VQ_EXACT=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref 2>&1 | grep "^gpu" | tail -2
python tools/prof_qdxt.py 1024 --fmt DXT1 --q 128 2>&1 | grep -E "^gpu|^ref" | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-transcode --no-hc > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-1500 gpurun_out/r2c_bench.json
