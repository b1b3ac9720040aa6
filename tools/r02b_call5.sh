#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "fft" 2>&1 | tail -3
timeout 600 python tools/fft_large_probe.py 2>&1 | tee gpurun_out/r02b_fft_large_probe3.txt
NCU="ncu --clock-control none --cache-control none --replay-mode application"
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"fft_cols|fft_rows256" -c 70 --csv --log-file gpurun_out/r02b_tiled_dram2.csv python tools/fft_large_probe.py ncu > gpurun_out/r02b_tiled_dram2.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02b_tiled_dram2.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[:70]:
    print(i,k,m.get('gpu__time_duration.sum'),m.get('dram__bytes_read.sum'),m.get('dram__bytes_write.sum'))
P
