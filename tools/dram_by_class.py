#!/usr/bin/env python
"""Sum DRAM bytes / time per GEMM epilogue class from an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum` log (one row per launch and metric)."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
per = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(set)
big = defaultdict(lambda: defaultdict(float))
for r in rows[1:]:
    name = r[ix["Kernel Name"]]
    m = re.search(r"GemmCfg<([^>]*)>.*?(Epi\w+)", name)
    cls = (m.group(2) + "<" + m.group(1).replace(" ", "") + ">") if m else name[:40]
    metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1}.get(unit, 1)
    per[cls][metric] += val * scale
    cnt[cls].add(r[ix["ID"]])
    big[cls][metric] = max(big[cls][metric], val * scale)
print(sys.argv[1])
for cls, m in per.items():
    print(f"  {cls:60s} launches {len(cnt[cls]):3d}  read {m.get('dram__bytes_read.sum', 0) / 1e9:7.2f} GB (max launch {big[cls].get('dram__bytes_read.sum', 0) / 1e9:5.2f})"
          f"  write {m.get('dram__bytes_write.sum', 0) / 1e9:6.2f} GB  time {m.get('gpu__time_duration.sum', 0):8.3f} ms")
