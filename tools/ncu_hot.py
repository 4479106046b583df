#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an .ncu-rep: python tools/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[isamp] or 0) for r in body)
print("kernel:", lines[0][:160]); print("total samples", tot)
idx = {id(r): k for k, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:N]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx[id(r)]:5d} {int(r[isamp]):7d} {100*int(r[isamp])/max(tot,1):5.1f}%  exec={r[iex]:>9}  {r[isrc][:90]:90s} {st}")
