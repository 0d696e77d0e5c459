#!/bin/bash
# phase trace of the clustered-DDS step (configs[1] size) after the radix sort of the split ranks + the pipelined member walks
mkdir -p gpurun_out
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref > gpurun_out/r2ag_phase_trace_c2.txt 2>&1
grep -E "^gpu" gpurun_out/r2ag_phase_trace_c2.txt | cut -c1-110
grep -E "vq_fast<|endpoint tree  |selector VQ" gpurun_out/r2ag_phase_trace_c2.txt | tail -12 | cut -c1-330
python -m pytest tests/test_gpu_qdxt.py tests/test_gpu_pipeline.py tests/test_gpu_round2.py tests/test_gpu_transcode.py -m gpu -x -q 2>&1 | tail -2
