python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
grep -v "Warning\|^\*\|OMP\|JETSTREAM\|NCCL version" gpurun_out/r02_bench_4gpu.err | tail -20
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_4gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['config']['numa'])
for k,v in d['e2e_variants'].items(): print(k, v.get('value'), v.get('error'))
w=d['workloads']['wideband']; print(w.get('error'), w.get('kernel_only'), w.get('with_full_gather'), w.get('with_display_reduction'), w.get('collective'))
PY
