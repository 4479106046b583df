#!/bin/bash
# round-2 visit 1: parity tests, smoke, the new bench line (variants), tile / chunk experiments
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
B="python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline --no-e2e"
echo "== experiments"
timeout 300 env DPA_GRAD_BN=256 $B 2>&1 | tail -1 > gpurun_out/r2_x_bn256.json
timeout 300 $B --chunk-rows 32768 2>&1 | tail -1 > gpurun_out/r2_x_chunk32k.json
timeout 300 env DPA_GRAD_BN=256 $B --chunk-rows 32768 2>&1 | tail -1 > gpurun_out/r2_x_bn256_chunk32k.json
timeout 300 env DPA_HINT_STATS=21 $B 2>&1 | tail -1 > gpurun_out/r2_x_hint21.json
timeout 300 env DPA_HINT_STATS=20 $B 2>&1 | tail -1 > gpurun_out/r2_x_hint20.json
python - <<'PY'
import json, glob
for f in ['gpurun_out/r2_bench1.json'] + sorted(glob.glob('gpurun_out/r2_x_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.05}, d['launches_per_step'], d['loss'])
    except Exception as e: print(f, 'ERR', e)
PY
echo "== dram traffic per GEMM class (ncu, one step)"
for cfgname in "bn512:" "bn256:DPA_GRAD_BN=256" "hint21:DPA_HINT_STATS=21"; do
  name=${cfgname%%:*}; envs=${cfgname#*:}
  timeout 600 env $envs ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dpa_gemm -s 84 -c 28 --csv --log-file gpurun_out/r2_dram_$name.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/dram_by_class.py gpurun_out/r2_dram_$name.csv
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
