#!/bin/bash
# vocab-parallel A/B at N GPUs: multi-GPU parity test, then the cfg-3 vp run with the dH reduction overlapped / serialised
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_$N.log 2>&1; grep -v "co-resident\|symmetric" gpurun_out/r2_pytest_multi_$N.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for ov in 1 0 1 0; do
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
