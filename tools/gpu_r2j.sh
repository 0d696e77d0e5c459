#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/r2j_trace.log 2>&1; grep -E "^gpu|vq_fast<16> n=|pack:|init:" gpurun_out/r2j_trace.log | tail -16
for v in w8_o2 w8_o3 w4_o4 w4_o5 w16_o1; do
  echo "== $v"
  CRN_B200_LIB=gpurun_variants/lib_$v.so CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 128 41 2>&1 | grep -E "cluster optimiser  " | awk 'NR%2==0'
done
