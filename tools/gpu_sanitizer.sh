#!/bin/bash
# compute-sanitizer over the hot path at small sizes: memcheck on smoke() and the round-2 / dxt_hc GPU tests, racecheck on smoke()
# (the CTA-per-cluster optimiser's double-buffered shared memory, the tree-split cluster reductions).  Logs -> gpurun_out/<tag>_*.txt
TAG=${1:-r2san}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
( time timeout 240 $S --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_memcheck_smoke.txt 2>&1
echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke ok|real" gpurun_out/${TAG}_memcheck_smoke.txt | cut -c1-160
( time timeout 420 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_hc.py tests/test_gpu_clusters.py -m gpu -x -q ) > gpurun_out/${TAG}_memcheck_tests.txt 2>&1
echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed|real" gpurun_out/${TAG}_memcheck_tests.txt | cut -c1-160
( time timeout 300 $S --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1
echo "racecheck smoke rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok|real" gpurun_out/${TAG}_racecheck_smoke.txt | cut -c1-160
for f in gpurun_out/${TAG}_*.txt; do grep -m 12 -E "Invalid|Race reported|hazard|Error:" $f | cut -c1-200; done
