#!/bin/bash
# L2 super-block height per GEMM class: step time (A/B, twice) and DRAM bytes per GEMM class (ncu, one step) for each setting.
# Usage (GPU box): bash tools/super_m_sweep.sh "STATS,DH,DW" ...      e.g.  "0,0,0" "64,0,0" "64,8,8"
mkdir -p gpurun_out; REPS=${REPS-"1 2"}
for rep in $REPS; do
for cfg in "$@"; do
  IFS=, read -r a b c <<< "$cfg"
  DPA_SUPER_STATS=$a DPA_SUPER_DH=$b DPA_SUPER_DW=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('time', '$cfg', round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"
done; done
for cfg in "$@"; do
  IFS=, read -r a b c <<< "$cfg"
  DPA_SUPER_STATS=$a DPA_SUPER_DH=$b DPA_SUPER_DW=$c timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dpa_gemm -s 84 -c 28 --csv \
    --log-file gpurun_out/super_$a-$b-$c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > /dev/null 2>&1
  python - "$cfg" gpurun_out/super_$a-$b-$c.csv <<'PY'
import csv, re, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
h = rows[0]; ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.defaultdict(float)
for r in rows[1:]:
    m = re.search(r"(Epi\w+)<", r[ki])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
    acc[(m.group(1) if m else "?", r[mi].split("__")[1].split(".")[0])] += float(r[vi].replace(",", "")) * scale
print("dram", sys.argv[1], {f"{k[0]}:{k[1]}": round(v / 1e9, 2) for k, v in sorted(acc.items())})
PY
done
