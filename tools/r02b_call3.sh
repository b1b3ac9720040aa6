#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/chain_variant_probe.py 2>&1 | tee gpurun_out/r02b_chain_variant_probe.txt
NCU="ncu --clock-control none"
B200_FFT4096_VARIANT=w timeout 300 $NCU --set full --import-source on -k regex:fft4096w -s 3 -c 1 -f -o gpurun_out/r02b_ncu_chain_w python tools/quick_gpu.py > gpurun_out/r02b_ncu_chain_w.log 2>&1
