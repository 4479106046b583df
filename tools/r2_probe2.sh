# Round-2 probes on one GPU: cfg-4 per-kernel table (and one chunk holding every active row), the 2-pair share of a strong-scaled
# batch with 512- / 256-wide gradient tiles, the cfg-3 x 8 shard shape with the shipped defaults.
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2_p2_{n}.json'))
    print(n, round(d['ms_per_step'],3), 'tc', d.get('step_tc_frac_active_alg'), 'rows', d.get('rows'), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.05})
except Exception as e:
    print(n, 'FAILED', e, open(f'gpurun_out/r2_p2_{n}.err').read()[-600:])
P
}
run() { name=$1; shift; envs=$1; shift
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants "$@" 2>gpurun_out/r2_p2_$name.err | tail -1 > gpurun_out/r2_p2_$name.json; show $name; }
run cfg4 X=1 --config cfg4
run cfg4_chunk64k X=1 --config cfg4 --chunk-rows 65536
run pairs2_bn512 X=1 --pairs 2
run pairs2_bn256 DPA_GRAD_BN=256 --pairs 2
run pairs2_bn512_b X=1 --pairs 2
run pairs2_bn256_b DPA_GRAD_BN=256 --pairs 2
run shard8 X=1 --config cfg3_shard8
