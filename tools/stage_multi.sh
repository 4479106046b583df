mkdir -p gpurun_out
echo "== multi test"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_trainer_shim.py -m gpu -x -q 2>&1 | tail -15
echo "== bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.log | cut -c1-1500
