#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter_fm.py tests/test_gpu_parity_holes.py tests/test_gpu_flowgraphs.py tests/test_gpu_shim.py -q -k "fir or filter or fm" 2>&1 | tail -3
ncu --clock-control none --set full --import-source on -k regex:fir_decim -s 3 -c 1 -f -o gpurun_out/r02b_ncu_fir_pair python tools/fir_probe.py 8192 127 8 > gpurun_out/r02b_ncu_fir_pair.log 2>&1
tail -1 gpurun_out/r02b_ncu_fir_pair.log
