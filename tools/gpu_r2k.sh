#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/r2k_trace.log 2>&1; grep -E "^gpu|vq_fast<16> n=|pack:|init:" gpurun_out/r2k_trace.log | tail -16
CRN_B200_TRACE=1 python tools/prof_crn_compress.py > gpurun_out/r2k_crn_trace.log 2>&1; grep -E "compress_crn|^search|^pass|cluster optimiser" gpurun_out/r2k_crn_trace.log | tail -40
