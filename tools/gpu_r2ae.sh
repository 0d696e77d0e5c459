#!/bin/bash
# Pipelined member walk in hc_tree_split: launch list of the crn_compress pass (compare r2ac) + the hc / crn GPU tests.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ae_launches_crn_q128.csv python tools/prof_cluster_phases.py 128 > /dev/null 2>&1
python tools/sum_launches.py gpurun_out/r2ae_launches_crn_q128.csv > gpurun_out/r2ae_launch_shares_crn_q128.txt; head -16 gpurun_out/r2ae_launch_shares_crn_q128.txt
rm -f gpurun_out/r2ae_launches_crn_q128.csv
python tools/prof_crn_compress.py 2>&1 | tail -12
python -m pytest tests/test_gpu_hc.py tests/test_gpu_crn.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
