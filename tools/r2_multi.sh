#!/bin/bash
# round-2 multi-GPU visit: 2-GPU parity tests, the bench line with its N > 1 variants, vocab-parallel runs
N=${1:-2}
mkdir -p gpurun_out
export DPA_VERBOSE=1
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_$N.log 2>&1; grep -v "co-resident" gpurun_out/r2_pytest_multi_$N.log | tail -60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== bench --gpus $N"; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -c 2500 gpurun_out/r2_bench_${N}gpu.json; tail -5 gpurun_out/r2_bench_${N}gpu.err
for mode in "" "--vp-backward recompute" "--vp-torch-exchange"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 300 $TR bench.py --gpus $N --config cfg3 --parallelism vp $mode --steps 10 --warmup 3 --no-e2e --no-variants 2>gpurun_out/r2_vp_${N}_$tag.err | tail -1 > gpurun_out/r2_vp_${N}_$tag.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_vp_${N}_$tag.json').read().strip().splitlines()[-1])
    print('vp $mode', $N, round(d['ms_per_step'],3), round(d['value']), d['loss'], d['step_tc_frac_active_alg'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.05})
except Exception as e:
    print('vp $mode ERR', e); print(open('gpurun_out/r2_vp_${N}_$tag.err').read()[-1500:])
PY
done
