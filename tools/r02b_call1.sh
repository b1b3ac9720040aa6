#!/usr/bin/env bash
# GPU call: two-pass FFT plan — parity tests, timing sweep, DRAM traffic (L2 residency), full capture of the 8192 kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "fft" 2>&1 | tail -5
timeout 600 python tools/fft_large_probe.py 2>&1 | tee gpurun_out/r02b_fft_large_probe.txt
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum -k regex:"col16|fft_radix" -c 140 --csv --log-file gpurun_out/r02b_twopass_dram.csv python tools/fft_large_probe.py ncu > gpurun_out/r02b_twopass_dram.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:fft_radix_kernel -s 1 -c 1 -f -o gpurun_out/r02b_ncu_fft8192 python tools/fft_large_probe.py ncu > gpurun_out/r02b_ncu_fft8192.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02b_twopass_dram.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[:140]:
    print(i,k,m.get('gpu__time_duration.sum'),m.get('dram__bytes_read.sum'),m.get('dram__bytes_write.sum'),m.get('lts__t_sector_hit_rate.pct'))
P
