#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
CRN_B200_TRACE=1 python tools/prof_crn_compress.py > gpurun_out/r2g_crn_trace.log 2>&1; grep -E "compress_crn|^search|^pass|cluster optimiser|clusters," gpurun_out/r2g_crn_trace.log | tail -60
CRN_B200_NO_COOP=1 CRN_B200_TRACE=1 python tools/prof_crn_compress.py 41 > gpurun_out/r2g_crn_trace_nocoop_q41.log 2>&1; grep -E "compress_crn|cluster optimiser|clusters," gpurun_out/r2g_crn_trace_nocoop_q41.log | head -12
CRN_B200_TRACE=1 python tools/prof_crn_compress.py 41 > gpurun_out/r2g_crn_trace_q41.log 2>&1; grep -E "compress_crn|cluster optimiser|clusters," gpurun_out/r2g_crn_trace_q41.log | head -12
