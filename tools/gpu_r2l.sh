#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transcode_(tables|walk|resolve|levels|streams)" -c 16 --csv --log-file gpurun_out/r2l_transcode_batch_launches.csv python tools/bench_transcode_batch.py 1024 1024 DXT5 64 2 > gpurun_out/r2l_batch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l_transcode_batch_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); bi=hdr.index('Block Size')
for r in rows[1:]:
    print(r[ki][:60], r[gi], r[bi], r[vi])
PY
