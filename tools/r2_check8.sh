#!/bin/bash
# fp32 inputs: K-lockstep of the split-bf16 GEMMs (DPA_KSYNC_SPLIT) vs off
mkdir -p gpurun_out
for k in 0 32 0 32; do
  timeout 300 env DPA_KSYNC_SPLIT=$k python bench.py --config cfg1 --fp32 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fp32 cfg1 ksync_split=$k', round(d['ms_per_step'],3), d['loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.05})"
done
timeout 600 env DPA_KSYNC_SPLIT=32 python -m pytest tests/test_gpu_fp32.py -m gpu -x -q 2>&1 | tail -2
