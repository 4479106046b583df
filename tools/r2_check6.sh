#!/bin/bash
# K-lockstep, second experiment: the statistics GEMMs (DPA_KSYNC_STATS), with the gradient GEMMs at their default (32)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline --no-e2e"
for k in 0 64 32 0 64 16; do
  timeout 300 env DPA_KSYNC_STATS=$k $B 2>&1 | tail -1 > gpurun_out/r2_z_kstats$k.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_z_kstats$k.json').read().strip().splitlines()[-1])
    print('ksync_stats $k', round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.3}, d['loss'])
except Exception as e: print('ksync_stats $k ERR', e, open('gpurun_out/r2_z_kstats$k.json').read()[-600:])
PY
done
for k in 0 64 32; do
  timeout 600 env DPA_KSYNC_STATS=$k ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dpa_gemm -s 42 -c 14 --csv --log-file gpurun_out/r2_dram_kstats$k.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/dram_by_class.py gpurun_out/r2_dram_kstats$k.csv
done
echo "== stats lockstep under the parity suite"; timeout 900 env DPA_KSYNC_STATS=64 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== dense / recompute with stats lockstep"; for a in "--dense" "--backward recompute"; do for k in 0 64; do timeout 300 env DPA_KSYNC_STATS=$k python bench.py $a --steps 10 --warmup 3 --no-cpu-baseline --no-variants --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$a kstats=$k', round(d['ms_per_step'],3), d['loss'])"; done; done
