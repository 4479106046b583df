#!/bin/bash
# Round-2 closing validation on one GPU: tests, smoke, the default bench line (all single-GPU variants), the reference arm,
# and the zero-fill A/B (library kernel vs cudaMemsetAsync).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2h_pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 400 python bench.py > gpurun_out/r2h_bench.json 2>gpurun_out/r2h_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], 'launches', d['launches_per_step'], 'gemm_share', d['gemm_share_of_step'], d['clocks'])
print({k: v['ms_per_step'] for k, v in d['kernels'].items() if v['ms_per_step'] > 0.05})
for k, v in d['variants'].items(): print(k, {a: b for a, b in v.items() if a not in ('note', 'tol', 'shape')})
PY
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null > gpurun_out/r2h_bench_reference_arm.json; cut -c1-200 gpurun_out/r2h_bench_reference_arm.json
echo "== zero fill A/B"
for i in 1 2; do for z in 1 0; do
DPA_ZERO_KERNEL=$z timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zero_kernel=$z', round(d['ms_per_step'],3), 'misc', d['kernels']['misc'])"
done; done
