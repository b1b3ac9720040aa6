#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity_holes.py -x -q -k "other_sizes or tiled or two_pass or 8192 or fft4096w" 2>&1 | tail -5
for cfg in "4096 16384" "2048 32768" "1024 65536"; do timeout 200 python tools/quick_gpu.py $cfg 2>&1 | grep -v "^$" | head -4; done | tee gpurun_out/r02b_tiled_chain.txt
