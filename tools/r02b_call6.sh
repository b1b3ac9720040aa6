#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "fft" 2>&1 | tail -3
timeout 600 python tools/fft_large_probe.py 2>&1 | tee gpurun_out/r02b_fft_large_probe5.txt
