mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dpa_gemm -s 93 -c 11 -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_dpa.py -m gpu -x -q 2>&1 | tail -5
