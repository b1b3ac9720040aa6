#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
ncu --clock-control none --set full --import-source on -k regex:fir_decim -s 3 -c 1 -f -o gpurun_out/r02b_ncu_fir python tools/fir_probe.py 8192 127 8 > gpurun_out/r02b_ncu_fir.log 2>&1
tail -2 gpurun_out/r02b_ncu_fir.log
