#!/bin/bash
# round-2 visit 2: fused small kernels, kept-logit slots (dense), async zero fill, chunk 32768 default, ncu full capture
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 1500 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
B="python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline --no-e2e"
echo "== experiments"
timeout 300 env DPA_ASYNC_ZERO=0 $B 2>&1 | tail -1 > gpurun_out/r2_y_synczero.json
timeout 300 $B --chunk-rows 16384 2>&1 | tail -1 > gpurun_out/r2_y_chunk16k.json
timeout 300 $B --graph 2>&1 | tail -1 > gpurun_out/r2_y_graph.json
timeout 300 $B --dense 2>&1 | tail -1 > gpurun_out/r2_y_dense.json
timeout 300 $B --no-dw 2>&1 | tail -1 > gpurun_out/r2_y_nodw.json
timeout 300 $B 2>&1 | tail -1 > gpurun_out/r2_y_default.json
python - <<'PY'
import json, glob
for f in ['gpurun_out/r2_bench2.json'] + sorted(glob.glob('gpurun_out/r2_y_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.03}, d['launches_per_step'], d['loss'], d['gemm_share_of_step'], d.get('cuda_graph'))
    except Exception as e: print(f, 'ERR', e)
PY
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r2_launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dpa_gemm|dz_kl_elem|dz_pair_elem|gather_rows" -s 72 -c 18 -f -o gpurun_out/r2_prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
