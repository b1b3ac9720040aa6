#!/usr/bin/env bash
# GPU call: fft4096w_kernel (warp-local first exchange) — bit-identity + timing against the classic kernel, the chain
# parity tests on the new kernel, full ncu captures of both.
set -u
mkdir -p gpurun_out
timeout 300 python tools/chain_variant_probe.py 2>&1 | tee gpurun_out/r02b_chain_variant_probe.txt
B200_FFT4096_VARIANT=w timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity_holes.py tests/test_gpu_viz.py -x -q 2>&1 | tail -5
NCU="ncu --clock-control none"
B200_FFT4096_VARIANT=w timeout 300 $NCU --set full --import-source on -k regex:fft4096w -s 3 -c 1 -f -o gpurun_out/r02b_ncu_chain_w python tools/quick_gpu.py > gpurun_out/r02b_ncu_chain_w.log 2>&1
B200_FFT4096_VARIANT=classic timeout 300 $NCU --set full --import-source on -k regex:fft4096_kernel -s 3 -c 1 -f -o gpurun_out/r02b_ncu_chain_classic python tools/quick_gpu.py > gpurun_out/r02b_ncu_chain_classic.log 2>&1
ls -la gpurun_out/*.ncu-rep
