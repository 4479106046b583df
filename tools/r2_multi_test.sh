#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi_$N.log 2>&1; grep -v "co-resident" gpurun_out/r2_pytest_multi_$N.log | tail -70
