#!/bin/bash
for t in 1024 512 256; do
  echo "== threads $t"
  CRN_B200_ORDER_THREADS=$t CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 255 2>&1 | grep -E "orderings on the device|compress_crn q|transitions" | tail -3
done
python -m pytest tests/test_gpu_crn.py -x -q -k "orderings" 2>&1 | tail -2
