#!/usr/bin/env bash
# GPU call: tiled two-pass FFT (parity + sweep + DRAM traffic), 8192 kernel with the register twiddle, chain w-variant check.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "fft" 2>&1 | tail -3
timeout 600 python tools/fft_large_probe.py 2>&1 | tee gpurun_out/r02b_fft_large_probe2.txt
timeout 300 python tools/chain_variant_probe.py 2>&1 | tee gpurun_out/r02b_chain_variant_probe.txt
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:"fft_cols|fft_rows256|fft_radix" -c 80 --csv --log-file gpurun_out/r02b_tiled_dram.csv python tools/fft_large_probe.py ncu > gpurun_out/r02b_tiled_dram.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02b_tiled_dram.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:50]),{})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[:80]:
    print(i,k,m.get('gpu__time_duration.sum'),m.get('dram__bytes_read.sum'),m.get('dram__bytes_write.sum'),m.get('lts__t_sector_hit_rate.pct'))
P
