#!/usr/bin/env bash
# Round-end measurement pass (run on the GPU box via gpurun): tests, smoke, bench (both arms), module table.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python tools/bench_modules.py > gpurun_out/final_bench_modules.txt 2>&1
python tools/typed_chain_probe.py >> gpurun_out/final_bench_modules.txt 2>&1
python tools/fm_probe.py >> gpurun_out/final_bench_modules.txt 2>&1
tail -3 gpurun_out/final_tests.txt; cat gpurun_out/final_smoke.txt | tail -2; tail -c 1500 gpurun_out/final_bench.json; tail -40 gpurun_out/final_bench_modules.txt
