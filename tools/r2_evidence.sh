#!/bin/bash
# round-2 evidence visit (1 GPU): parity suite, smoke, the driver's bench line, reference arm, ncu launch list + full capture
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; grep -v "co-resident" gpurun_out/r2_pytest_gpu.log | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 400 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
echo "== fp32 cfg1 / dense / recompute"; for a in "--config cfg1 --fp32" "--dense" "--backward recompute" "--no-dw"; do timeout 300 python bench.py $a --steps 20 --warmup 3 --no-cpu-baseline --no-variants --no-e2e 2>/dev/null | tail -1 > "gpurun_out/r2_bench_$(echo $a | tr -d ' -').json"; python -c "
import json,sys; d=json.loads(open('gpurun_out/r2_bench_$(echo $a | tr -d ' -').json').read()); print('$a', round(d['ms_per_step'],3), d['loss'], d['step_tc_frac_active_alg'], d['launches_per_step'])"; done
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r2_launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dpa_gemm|dz_kl_elem|dz_pair_elem|gather_rows" -s 57 -c 19 -f -o gpurun_out/r2_prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
