#!/usr/bin/env bash
set -u
for R in 8 16 4 2; do
for cfg in "7 64" "5 128" "5 64" "3 128"; do set -- $cfg
  echo "R=$R pair OB=$1 threads=$2: $(B200_FIR_PAIR=1 B200_FIR_OB=$1 B200_FIR_THREADS=$2 python tools/fir_probe.py 8192 127 $R | tail -1)"
done
echo "R=$R classic OB=5 threads=128: $(B200_FIR_PAIR=0 B200_FIR_OB=5 B200_FIR_THREADS=128 python tools/fir_probe.py 8192 127 $R | tail -1)"
done
