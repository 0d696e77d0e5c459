#!/bin/bash
# warp-per-node fast VQ kernel at 1 / 12 / 16 CTAs per SM (launch bounds): kernel time from the ncu launch list + wall time of the step
mkdir -p gpurun_out
for v in base v12 v16; do
  lib=crunch2_b200/libcrn_b200.so; [ $v != base ] && lib=crunch2_b200/libcrn_b200_$v.so
  CRN_B200_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vq_fast --csv --log-file gpurun_out/r2ai_$v.csv python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > /dev/null 2>&1
  echo "== $v"; python tools/sum_launches.py gpurun_out/r2ai_$v.csv | grep -E "total|16, 32, 1|6, 32, 1|tiny" | head -5; rm -f gpurun_out/r2ai_$v.csv
  CRN_B200_LIB=$lib python tools/prof_qdxt.py 4096 4096 --fmt DXT5 --q 128 --no-ref 2>/dev/null | grep -E "^gpu" | cut -c1-95
done
