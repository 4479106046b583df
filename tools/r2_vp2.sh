#!/bin/bash
# vocab-parallel: blocks of the peer reduce kernel vs the dW GEMM it runs beside
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for pb in ${PBS:-0 148 74 32 0 148}; do
  timeout 300 env DPA_PEER_BLOCKS=$pb $TR bench.py --gpus $N --config cfg3 --parallelism vp --steps 20 --warmup 5 --no-e2e --no-variants 2>gpurun_out/r2_vp_${N}_pb$pb.err | tail -1 > gpurun_out/r2_vp_${N}_pb$pb.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_vp_${N}_pb$pb.json').read().strip().splitlines()[-1])
    print('vp peer_blocks=$pb', $N, round(d['ms_per_step'],3), d['step_tc_frac_active_alg'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.1})
except Exception as e:
    print('vp ERR', e); print(open('gpurun_out/r2_vp_${N}_pb$pb.err').read()[-1500:])
PY
done
