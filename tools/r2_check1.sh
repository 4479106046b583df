#!/bin/bash
# round-2 visit 1: parity tests, smoke, the new bench line (variants), chunk-rows experiment
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
echo "== bench chunk 32768"; timeout 300 python bench.py --steps 20 --warmup 3 --chunk-rows 32768 --no-variants --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r2_bench_chunk32k.json; python - <<'PY'
import json
for f in ('gpurun_out/r2_bench1.json','gpurun_out/r2_bench_chunk32k.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
