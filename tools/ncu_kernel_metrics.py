#!/usr/bin/env python
"""Per kernel class, the largest launch of an `ncu --set full` report -> profiles/kernel_metrics.json (+ a markdown table).

    python tools/ncu_kernel_metrics.py gpurun_out/prof_step.ncu-rep profiles/kernel_metrics.json "label of the capture"
"""
import csv, io, json, re, subprocess, sys

CLASS = {"EpiStatsPair": "fwd_pair", "EpiStatsStore": "stash_ref", "EpiStoreRows": "dH", "EpiAccumF32": "dW", "EpiAccumDW": "dW", "EpiDzPair": "dz_pair",
         "EpiStatsRef": "fwd_ref", "EpiDzRef": "dz_ref"}
GB = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rep, dst, label = sys.argv[1:4]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    best = {}
    for li, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]]
        m = re.search(r"GemmCfg<([^>]*)>, (?:dpa::)?(\w+)<", name)
        if m:
            cls = CLASS.get(m.group(2), m.group(2))
            cfg = re.sub(r"\(\w+\)", "", m.group(1)).replace(" ", "")
        else:
            m2 = re.search(r"(\w+_kernel)", name)
            cls, cfg = (m2.group(1) if m2 else name[:40]), ""
        get = lambda k: float(r[col[k]]) if k in col and r[col[k]] not in ("", "n/a") else None
        ms = get("gpu__time_duration.sum") * T[units[col["gpu__time_duration.sum"]]]
        rd = get("dram__bytes_read.sum") * GB[units[col["dram__bytes_read.sum"]]]
        wr = get("dram__bytes_write.sum") * GB[units[col["dram__bytes_write.sum"]]]
        rec = dict(launch=f"#{li} GemmCfg<{cfg}>" if cfg else f"#{li}", ms=round(ms, 4), dram_bytes=int(rd + wr), dram_read=int(rd), dram_write=int(wr),
                   sm_ghz=get("sm__cycles_elapsed.avg.per_second"),
                   tensor_pipe_active_pct=get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   dram_pct=get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   l2_hit_pct=get("lts__t_sector_hit_rate.pct"), regs=get("launch__registers_per_thread"))
        if cls not in best or ms > best[cls]["ms"]:
            best[cls] = rec
    json.dump(dict(source=label, kernels=best), open(dst, "w"), indent=1)
    md = dst.replace(".json", ".md")
    with open(md, "w") as f:
        f.write(f"# ncu --set full: largest launch per kernel class\n\nSource: {label}\n\n")
        f.write("| class | launch | ms (under ncu) | SM GHz | tensor pipe active % | DRAM read GB | DRAM write GB | DRAM % | L2 hit % | regs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for c, k in best.items():
            f.write(f"| {c} | {k['launch']} | {k['ms']} | {k['sm_ghz']} | {k['tensor_pipe_active_pct']} | {k['dram_read']/1e9:.3f} | {k['dram_write']/1e9:.3f} | {k['dram_pct']} | {k['l2_hit_pct']} | {k['regs']} |\n")
    print(open(md).read())


if __name__ == "__main__":
    main()
