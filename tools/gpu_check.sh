#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, one ncu --set full capture.
# Logs under gpurun_out/.  Usage: tools/gpu_check.sh [quick]
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== bench no-dw"; timeout 600 python bench.py --steps 20 --warmup 3 --no-dw --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nodw.log
[ "$1" = quick ] && exit 0
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches.csv
echo "== ncu full (dominant GEMM kernels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dpa_gemm|dz_kl_elem|dz_pair_elem|gather_rows" -s 108 -c 16 -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
