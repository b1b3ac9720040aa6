#!/usr/bin/env bash
# Second session of round 2, validation on one B200: the whole GPU suite, the per-module table, the bench line, the ncu
# launch list of the bench command and full captures of the new kernels (written to gpurun_out/, summarised into profiles/).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02b_pytest_gpu.txt
timeout 600 python tools/bench_modules.py > gpurun_out/r02b_bench_modules.txt 2>&1; grep -c "" gpurun_out/r02b_bench_modules.txt
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; head -c 600 gpurun_out/r02b_bench.json; echo
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_launches_bench.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:fft4096 -s 3 -c 1 -f -o gpurun_out/r02b_ncu_chain python tools/quick_gpu.py > gpurun_out/r02b_ncu_chain.log 2>&1
timeout 300 $NCU --cache-control none --set full --import-source on -k regex:fft_cols -s 5 -c 1 -f -o gpurun_out/r02b_ncu_cols python tools/fft_large_probe.py ncu > gpurun_out/r02b_ncu_cols.log 2>&1
timeout 300 $NCU --cache-control none --set full --import-source on -k regex:fft_rows256 -s 5 -c 1 -f -o gpurun_out/r02b_ncu_rows python tools/fft_large_probe.py ncu > gpurun_out/r02b_ncu_rows.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
