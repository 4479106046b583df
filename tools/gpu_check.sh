#!/bin/bash
# One GPU-box visit: parity tests, smoke, stage timings, bench, ncu launch list. Logs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== cfg2 stages"; timeout 300 python tools/first_light.py cfg2 nodw 2>&1 | tail -20 | tee gpurun_out/cfg2_nodw.log
timeout 300 python tools/first_light.py cfg2 dw 2>&1 | tail -20 | tee gpurun_out/cfg2_dw.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
