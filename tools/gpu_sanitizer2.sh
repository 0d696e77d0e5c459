#!/bin/bash
# racecheck (shared-memory hazards) and synccheck over the GPU tests of the hot path, after the two fixes r2san asked for
TAG=${1:-r2san2}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
( time timeout 200 $S --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1
echo "racecheck smoke rc=$?"; grep -E "RACECHECK SUMMARY|real" gpurun_out/${TAG}_racecheck_smoke.txt | cut -c1-160
( time timeout 200 $S --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_synccheck_smoke.txt 2>&1
echo "synccheck smoke rc=$?"; grep -E "ERROR SUMMARY|real" gpurun_out/${TAG}_synccheck_smoke.txt | cut -c1-160
for t in test_gpu_hc test_gpu_clusters test_gpu_transcode test_gpu_round2; do
  ( time timeout 240 $S --tool racecheck --error-exitcode 9 python -m pytest tests/$t.py -m gpu -q ) > gpurun_out/${TAG}_racecheck_$t.txt 2>&1
  echo "racecheck $t rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|real" gpurun_out/${TAG}_racecheck_$t.txt | cut -c1-160
done
for f in gpurun_out/${TAG}_*.txt; do grep -E "Race reported|and (Read|Write) access|Error:" $f | sed 's/(crn::[^)]*)/(...)/g' | sort | uniq -c | sort -rn | head -12 | cut -c1-230; done
