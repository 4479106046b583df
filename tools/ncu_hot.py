#!/usr/bin/env python
"""Top stalled SASS instructions per kernel of an .ncu-rep: python tools/ncu_hot.py rep [N] [kernel-index ...]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = [int(x) for x in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for bi, b in enumerate(blocks):
    if only and bi not in only:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
    hdr = rows[0]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[isamp] or 0) for r in body)
    m = re.search(r"GemmCfg<([^>]*)>, dpa::(\w+)<", b[0])
    print(f"=== kernel {bi}: {m.group(2) + '<' + re.sub(r'\(\w+\)', '', m.group(1)) + '>' if m else b[0][:120]}  samples={tot}")
    idx = {id(r): k for k, r in enumerate(body)}
    for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:N]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{idx[id(r)]:5d} {int(r[isamp]):7d} {100*int(r[isamp])/max(tot,1):5.1f}%  exec={r[iex]:>9}  {r[isrc][:80]:80s} {[s for s in st if s[0]]}")
