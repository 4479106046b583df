#!/usr/bin/env python
"""Print selected metrics of every kernel in an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__cluster_dim_x",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
extra = [h for h in hdr if pat and pat.search(h)]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("----", re.sub(r"\s+", " ", r[ki])[:150])
    for w in WANT + extra:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w} [{units[i]}] = {r[i]}")
