mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for cr in 16384 32768; do echo "== chunk $cr"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --chunk-rows $cr 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0})"; done
echo "== recompute"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --backward recompute 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0})"
echo "== dense"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --dense 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0})"
