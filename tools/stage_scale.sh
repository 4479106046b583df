mkdir -p gpurun_out
N=$1
for n in $N; do
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-variants 2>&1 | tail -1 > gpurun_out/scale_$n.json
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 3 --no-variants 2>&1 | tail -1 > gpurun_out/scale_$n.json; fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print($n, round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks'])"
done
