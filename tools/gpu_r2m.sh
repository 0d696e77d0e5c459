#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_crn.py -x -q 2>&1 | tail -4
CRN_B200_TRACE=1 python tools/prof_crn_compress.py > gpurun_out/r2m_crn_trace.log 2>&1; grep -E "compress_crn|^search|^pass|orderings|transitions" gpurun_out/r2m_crn_trace.log | tail -60
