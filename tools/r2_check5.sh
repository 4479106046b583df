#!/bin/bash
# K-lockstep experiment: step time and DRAM bytes per GEMM class vs DPA_KSYNC
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline --no-e2e"
for k in 0 64 16 0 32 128; do
  timeout 300 env DPA_KSYNC=$k $B 2>&1 | tail -1 > gpurun_out/r2_z_ksync$k.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_z_ksync$k.json').read().strip().splitlines()[-1])
    print('ksync $k', round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.3}, d['loss'])
except Exception as e: print('ksync $k ERR', e, open('gpurun_out/r2_z_ksync$k.json').read()[-600:])
PY
done
for k in 0 32 64; do
  timeout 600 env DPA_KSYNC=$k ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dpa_gemm -s 42 -c 14 --csv --log-file gpurun_out/r2_dram_ksync$k.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/dram_by_class.py gpurun_out/r2_dram_ksync$k.csv
done
timeout 600 env DPA_KSYNC=32 python -m pytest tests/test_gpu_dpa.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
