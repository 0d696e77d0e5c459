#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ac_launches_crn_q128.csv python tools/prof_cluster_phases.py 128 > /dev/null 2>&1
python tools/sum_launches.py gpurun_out/r2ac_launches_crn_q128.csv > gpurun_out/r2ac_launch_shares_crn_q128.txt; head -30 gpurun_out/r2ac_launch_shares_crn_q128.txt
