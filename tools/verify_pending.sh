#!/usr/bin/env bash
# One GPU call: NCO microbenchmark, filter/FM parity (halo sharding test), batched-NCO variant parity + timing.
set -u
mkdir -p gpurun_out
tools/bin/microbench3 | tee gpurun_out/microbench3.txt
python -m pytest tests/test_gpu_filter_fm.py tests/test_gpu_flowgraphs.py -m gpu -q 2>&1 | tail -4
echo "== batched NCO"
B200_FM_NCO_BATCHED=1 python -m pytest tests/test_gpu_filter_fm.py tests/test_gpu_flowgraphs.py -m gpu -q -k "wide or config4" 2>&1 | tail -2
python tools/fm_wide_probe.py
B200_FM_NCO_BATCHED=1 python tools/fm_wide_probe.py
