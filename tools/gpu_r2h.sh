#!/bin/bash
mkdir -p gpurun_out
CRN_B200_PHASES=1 CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 128 41 > gpurun_out/r2h_phase_coop.log 2>&1; grep -E "compress_crn q|phase clocks|^\[crn_b200\]   |cluster optimiser  " gpurun_out/r2h_phase_coop.log | tail -36
CRN_B200_PHASES=1 CRN_B200_NO_COOP=1 CRN_B200_TRACE=1 python tools/prof_cluster_phases.py 128 41 > gpurun_out/r2h_phase_warp.log 2>&1; grep -E "compress_crn q|phase clocks|^\[crn_b200\]   |cluster optimiser  " gpurun_out/r2h_phase_warp.log | tail -36
