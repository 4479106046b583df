#!/bin/bash
# multi-GPU visit: full single-GPU parity suite, 2-GPU parity, vp runs (overlapped / serialised dH reduction)
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all_$N.log 2>&1; grep -v "co-resident\|symmetric" gpurun_out/r2_pytest_all_$N.log | tail -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for ov in 1 0; do
  timeout 300 env DPA_PEER_OVERLAP=$ov $TR bench.py --gpus $N --config cfg3 --parallelism vp --steps 20 --warmup 5 --no-e2e --no-variants 2>gpurun_out/r2_vp_${N}_ov$ov.err | tail -1 > gpurun_out/r2_vp_${N}_ov$ov.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_vp_${N}_ov$ov.json').read().strip().splitlines()[-1])
    print('vp overlap=$ov', $N, round(d['ms_per_step'],3), round(d['value']), d['loss'], d['step_tc_frac_active_alg'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.02}, d['launches_per_step'])
except Exception as e:
    print('vp ERR', e); print(open('gpurun_out/r2_vp_${N}_ov$ov.err').read()[-1500:])
PY
done
echo "== bench --gpus $N"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -c 1800 gpurun_out/r2_bench_${N}gpu.json; grep -v "co-resident\|symmetric" gpurun_out/r2_bench_${N}gpu.err | tail -5
