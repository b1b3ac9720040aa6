python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02b_bench_2gpu.json 2> gpurun_out/r02b_bench_2gpu.err
tail -c 3000 gpurun_out/r02b_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['config']['numa'])
print(json.dumps(d['e2e_variants'],indent=0)[:1500])
print(json.dumps(d['workloads']['wideband'],indent=0)[:2500])
PY
