#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))
for k, v in d['variants'].items(): print(k, {a: b for a, b in v.items() if a not in ('note', 'tol', 'shape')})
PY
grep -v "co-resident\|symmetric\|OMP_NUM\|\*\*\*" gpurun_out/r2_bench_${N}gpu.err | tail -5
