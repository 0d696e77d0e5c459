#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
CRN_B200_TRACE=1 python tools/prof_crn_compress.py > gpurun_out/r2f_crn_trace.log 2>&1; grep -E "compress_crn|^search|^pass" gpurun_out/r2f_crn_trace.log | tail -40
