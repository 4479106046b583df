# Units of the statistics GEMMs on a small vocabulary shard (V_local = 4000 = cfg-3 x 8): tiles per unit, per-unit lockstep,
# gradient tile width.  One GPU; per-class kernel times of the same batch.
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --config cfg3_shard8 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-variants 2>gpurun_out/r2_tpu_$name.err | tail -1 > gpurun_out/r2_tpu_$name.json
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2_tpu_{n}.json'))
    print(n, round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.05}, 'loss', d.get('loss'))
except Exception as e:
    print(n, 'FAILED', e)
P
}
for i in 1 2; do
run old DPA_TPU=5 DPA_KSYNC_STATS=-1
run new X=1
run new_usync DPA_TPU=1 DPA_KSYNC_STATS=-1
run tpu2 DPA_TPU=2 DPA_KSYNC_STATS=0
run new_bn256 DPA_GRAD_BN=256
done
# cost of the empty row-chunk launches at cfg-2 (3 chunks by capacity, 1 with rows): one chunk that holds every position
for i in 1 2; do
for cr in 0 98304; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --chunk-rows $cr 2>gpurun_out/r2_chunk_$cr.err | tail -1 > gpurun_out/r2_chunk_$cr.json
python -c "
import json; d=json.load(open('gpurun_out/r2_chunk_$cr.json')); print('chunk_rows', $cr, round(d['ms_per_step'],3), d['launches_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.05})"
done; done
