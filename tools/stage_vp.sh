mkdir -p gpurun_out
N=$1; CFG=${2:-cfg3}
for mode in "" "--vp-torch-exchange"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --config $CFG --parallelism vp $mode --steps 10 --warmup 3 --no-e2e --no-variants 2>&1 | tail -1 > gpurun_out/vp_${CFG}_$N$mode.json
python -c "
import json; d=json.load(open('gpurun_out/vp_${CFG}_$N$mode.json')); print('vp$mode', $N, round(d['ms_per_step'],3), round(d['value']), d['loss'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.25})"
done
