#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/).

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/rNN_launches.md "title" "command"

Takes the LAST step of the run (everything from the final meta_bitmap_kernel launch on).
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    m = re.search(r"dpa_gemm\w*_kernel<(?:dpa::)?GemmCfg<([^>]*)>, (?:dpa::)?(\w+)<", name)
    if m:
        return f"dpa_gemm<{m.group(1).replace(' ', '')}>/{m.group(2)}"
    if "at::" in name or "cub::" in name:
        return "torch:" + re.sub(r"\s+", " ", name)[:70]
    m = re.search(r"(\w+_kernel)", name)
    return "dpa::" + m.group(1) if m else name[:70]


def main():
    src, dst, title, cmd = sys.argv[1:5]
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(short(r[ki]), float(r[vi]) / 1e3) for r in rows[1:]]
    starts = [i for i, (n, _) in enumerate(seq) if n == "dpa::meta_bitmap_kernel"]
    step = seq[starts[-1]:] if starts else seq
    tot = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        a = agg.setdefault(n, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] = max(a[2], t)
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nCommand: `{cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache, serialised and at un-capped clocks: compare SHARES with the live\n"
                "CUDA-event shares in bench.py's `kernels` object, not absolutes. Last step of the run:\n\n")
        f.write("| kernel | launches | us total | us max | share |\n|---|---:|---:|---:|---:|\n")
        for n, (c, t, mx) in agg.items():
            f.write(f"| `{n}` | {c} | {t:.1f} | {mx:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"| **step** | {len(step)} | {tot:.1f} | | 100% |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
