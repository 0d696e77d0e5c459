#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_qdxt.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -3
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 > gpurun_out/r2q_trace.log 2>&1; grep -E "^gpu|^ref|vq_fast<16> n=|pack:|PSNR|psnr|bits" gpurun_out/r2q_trace.log | tail -16
CRN_B200_VQ_NO_PIPELINE=1 CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref 2>&1 | grep -E "^gpu|vq_fast<16> n=" | tail -4
