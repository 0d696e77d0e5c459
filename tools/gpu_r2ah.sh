#!/bin/bash
# ncu launch list of the configs[1] step on the final tree (two repetitions of tools/prof_qdxt.py at 4096^2 DXT5 q128)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ah_launches_c2.csv python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
python tools/sum_launches.py gpurun_out/r2ah_launches_c2.csv > gpurun_out/r2ah_launch_shares_c2.txt; head -24 gpurun_out/r2ah_launch_shares_c2.txt
rm -f gpurun_out/r2ah_launches_c2.csv
