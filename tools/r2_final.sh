#!/bin/bash
# final sanity of the round: what the driver runs on one GPU
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v co-resident | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_last.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], d['step_tc_frac_active_alg'], 'launches', d['launches_per_step'],
      {k: v.get('ms_per_step') for k, v in d['variants'].items()}, 'cpu', round(d['cpu_baseline']['value']), d['clocks'])
PY
tail -3 gpurun_out/r2_bench_last.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
