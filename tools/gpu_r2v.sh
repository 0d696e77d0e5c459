#!/bin/bash
python -m pytest tests/test_gpu_qdxt.py tests/test_gpu_dropin.py tests/test_gpu_crn.py -x -q 2>&1 | tail -3
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 2>&1 | awk '/rep0/{f=1} f' | grep -E "^gpu|^ref|pack:" | cut -c1-230
