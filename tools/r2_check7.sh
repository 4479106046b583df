#!/bin/bash
# per-unit lockstep of the statistics GEMMs (DPA_KSYNC_STATS=-1) vs off, interleaved
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline --no-e2e"
for k in 0 -1 0 -1 0 -1; do
  timeout 300 env DPA_KSYNC_STATS=$k $B 2>&1 | tail -1 > gpurun_out/r2_z_ustats$k.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_z_ustats$k.json').read().strip().splitlines()[-1])
    print('ksync_stats $k', round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.3}, d['loss'])
except Exception as e: print('ksync_stats $k ERR', e, open('gpurun_out/r2_z_ustats$k.json').read()[-600:])
PY
done
for k in 0 -1; do
  timeout 600 env DPA_KSYNC_STATS=$k ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dpa_gemm -s 42 -c 14 --csv --log-file gpurun_out/r2_dram_ustats$k.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/dram_by_class.py gpurun_out/r2_dram_ustats$k.csv
done
timeout 900 env DPA_KSYNC_STATS=-1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
