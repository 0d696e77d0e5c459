#!/bin/bash
CRN_B200_LIB=crunch2_b200/libcrn_b200_prof.so CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref 2>&1 | awk '/rep0/{f=1} f' | grep -E "phase clocks|^\[crn_b200\]   |endpoint optimisation" | cut -c1-200
