#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hc_tree_split_kernel -s 18 -c 6 -f -o gpurun_out/r2ad_hc_tree python tools/prof_cluster_phases.py 128 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2ad_hc_tree.ncu-rep > gpurun_out/r2ad_hc_tree_ncu.txt 2>&1; rm -f gpurun_out/r2ad_hc_tree.ncu-rep
grep -E "kernel:|gpu__time|grid_size|block_size|issue_active|warps_active|thread_inst_executed_per|pipe_fp64|pipe_fma|pipe_lsu|dram__bytes_read|barrier|long_scoreboard|short_scoreboard|registers" gpurun_out/r2ad_hc_tree_ncu.txt | cut -c1-150
