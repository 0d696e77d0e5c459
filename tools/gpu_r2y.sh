#!/bin/bash
for v in s1 s2 s4 s8; do
  echo "== $v"
  CRN_B200_LIB=gpurun_variants/lib_$v.so CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 128 41 255 2>&1 | grep -E "cluster optimiser  " | awk 'NR%2==0' | cut -c1-70
done
