#!/usr/bin/env bash
# Round-2 (second session) validation on one B200: the whole GPU suite, the per-module table, the bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02b_pytest_gpu.txt
timeout 600 python tools/bench_modules.py 2>&1 | tee gpurun_out/r02b_bench_modules.txt
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 1500 gpurun_out/r02b_bench.json
