#!/bin/bash
# round-2 visit 3: fp32 fused path, C ABI caller, ncu full capture with the right launch window
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== fp32 cfg1"; for b in auto recompute; do timeout 300 python bench.py --config cfg1 --fp32 --backward $b --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$b', round(d['ms_per_step'],3), d['loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.02})"; done
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 600 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dpa_gemm|dz_kl_elem|dz_pair_elem|gather_rows" -s 57 -c 19 -f -o gpurun_out/r2_prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
