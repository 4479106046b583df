#!/bin/bash
# Round-end validation on the GPU box (what the driver runs, plus a 2-rank data-parallel bench when 2 GPUs are visible).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 300 python bench.py > gpurun_out/final_bench.json 2>gpurun_out/final_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], 'frozen', round(d['variants']['lm_head_frozen']['value']),
      'cpu', round(d['cpu_baseline']['value']), d['clocks'])
PY
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-160
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
echo "== bench 2 GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['parallelism'])"
echo "== multi tests"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -1
fi
