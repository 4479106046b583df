mkdir -p gpurun_out
echo "== perf"; for v in 0,0,256 0,1,256 1,1,256; do DPA_CTA_GROUP=2 timeout 120 python tools/first_light.py gemm_perf $v 2>&1 | grep "gemm perf\|dpa:"; DPA_CTA_GROUP=1 timeout 120 python tools/first_light.py gemm_perf $v 2>&1 | grep "gemm perf"; done | tee gpurun_out/gemm_perf.log
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
