#!/usr/bin/env bash
# Round-end ncu pass (one GPU): launch list of the bench step, full captures of the hot kernels. Writes gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r02_fm_wide_launches.csv python tools/fm_wide_probe.py > gpurun_out/r02_launches_fm_wide.log 2>&1
$NCU --set full --import-source on -k regex:fft4096 -s 3 -c 1 -f -o gpurun_out/r02_ncu_chain python tools/quick_gpu.py > gpurun_out/r02_ncu_chain.log 2>&1
$NCU --set full --import-source on -k regex:fft4096 -s 30 -c 1 -f -o gpurun_out/r02_ncu_chain_colsum python tools/viz_probe.py > gpurun_out/r02_ncu_chain_colsum.log 2>&1
$NCU --set full --import-source on -k regex:colsum_partial -s 2 -c 1 -f -o gpurun_out/r02_ncu_lineplot python tools/viz_probe.py > gpurun_out/r02_ncu_lineplot.log 2>&1
$NCU --set full --import-source on -k regex:fir_decim -s 3 -c 1 -f -o gpurun_out/r02_ncu_fir python tools/fir_probe.py 8192 127 8 > gpurun_out/r02_ncu_fir.log 2>&1
$NCU --set full --import-source on -k regex:fm_narrow_fused -s 7 -c 1 -f -o gpurun_out/r02_ncu_fm_narrow python tools/fm_probe.py > gpurun_out/r02_ncu_fm_narrow.log 2>&1
$NCU --set full --import-source on -k regex:fm_narrow_fused -s 14 -c 1 -f -o gpurun_out/r02_ncu_fm_narrow_deemph python tools/fm_probe.py > gpurun_out/r02_ncu_fm_narrow_deemph.log 2>&1
python tools/fm_wide_probe.py
python tools/fm_probe.py | tail -6
python tools/viz_probe.py
ls -la gpurun_out/r02_*.ncu-rep
