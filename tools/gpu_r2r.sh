#!/bin/bash
CRN_B200_TRACE=1 python tools/prof_qdxt.py 4096 --fmt DXT5 --q 128 --no-ref 2>&1 | awk '/rep0/{f=1} f' | grep -E "^gpu|vq_fast<16> n=|pack:" | cut -c1-330
