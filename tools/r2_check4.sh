#!/bin/bash
# round-2 visit: cluster-size occupancy probe, cfg-5 sweep, compute-sanitizer on the round-2 kernels
mkdir -p gpurun_out
echo "== co-resident clusters of the 256-wide GEMM kernel"; python - <<'PY'
from halva_b200 import _lib
lib = _lib.load()
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")
for c in (1, 2, 4, 8):
    n = lib.dpa_debug_max_clusters(c)
    print("cluster size", c, "-> co-resident clusters", n, "= CTAs", n * c if n > 0 else lib.dpa_last_error())
PY
echo "== cfg5 sweep"; timeout 1200 python tests/sweep_cfg5.py --out gpurun_out/r2_cfg5_sweep.md --pairs 1,4,16,42 > gpurun_out/r2_cfg5_sweep.log 2>&1; tail -22 gpurun_out/r2_cfg5_sweep.md
echo "== sanitizer"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_dpa.py tests/test_gpu_fp32.py -m gpu -x -q -k "synthetic_shapes or dense_equals or poisons or chunk_invariant or second_backward" > gpurun_out/r2_sanitizer.log 2>&1; tail -6 gpurun_out/r2_sanitizer.log
