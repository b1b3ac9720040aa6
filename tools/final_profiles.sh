#!/usr/bin/env bash
# Round-end ncu pass (one GPU): launch list of the bench step, full captures of the fused kernels.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_fm_wide.csv python tools/fm_wide_probe.py > gpurun_out/launches_fm_wide.log 2>&1
$NCU --set full --import-source on -k regex:fft4096 -s 3 -c 1 -f -o gpurun_out/ncu_chain python tools/quick_gpu.py > gpurun_out/ncu_chain.log 2>&1
$NCU --set full --import-source on -k regex:fft4096 -s 3 -c 1 -f -o gpurun_out/ncu_chain_ci8 python tools/typed_chain_probe.py > gpurun_out/ncu_chain_ci8.log 2>&1
$NCU --set full --import-source on -k regex:fir_decim -s 3 -c 1 -f -o gpurun_out/ncu_fir python tools/fir_probe.py 8192 127 8 > gpurun_out/ncu_fir.log 2>&1
$NCU --set full --import-source on -k regex:fm_narrow_fused -s 7 -c 1 -f -o gpurun_out/ncu_fm python tools/fm_probe.py > gpurun_out/ncu_fm.log 2>&1
python tools/fm_wide_probe.py
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/launches_fm_wide.log
