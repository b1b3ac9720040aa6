#!/usr/bin/env bash
# Second session of round 2, validation on one B200: the whole GPU suite, smoke(), the per-module table, the bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02b_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/r02b_smoke.txt
timeout 600 python tools/bench_modules.py > gpurun_out/r02b_bench_modules.txt 2>&1; grep -c "" gpurun_out/r02b_bench_modules.txt
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])
for k,w in d['workloads'].items():
    if isinstance(w,dict): print(k, w.get('value'), w.get('ms_per_step'), (w.get('roofline') or {}).get('frac'), list((w.get('cases') or {}).keys()))
P
