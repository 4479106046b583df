#!/usr/bin/env python
"""BASELINE.json configs[4]: tokens x phrase-density sweep at the 7B shape, fused CUDA path vs the reference op chain
(oracle/dpa_port.py: the same ATen sequence the reference trainer issues) run on the SAME GPU in bf16.

    python tests/sweep_cfg5.py [--out gpurun_out/cfg5_sweep.md]      (a script, not collected by pytest; it lives under
    tests/ because it runs the oracle's torch port as the comparison arm)

The reference chain materialises fp32 [rows, 32000] tensors several times over and runs out of memory at the top end;
that is recorded as OOM, not skipped.
"""
import argparse
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from halva_b200 import dpa_loss, read_status  # noqa: E402
from halva_b200.synth import CONFIGS, make_inputs  # noqa: E402
from oracle import dpa_port  # noqa: E402


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cfg5_sweep.md"))
    ap.add_argument("--pairs", default="1,2,4,8,16,32,42")
    ap.add_argument("--densities", default="0.05,0.1,0.25,0.5")
    ap.add_argument("--no-dw", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for B in [int(x) for x in args.pairs.split(",")]:
        for dens in [float(x) for x in args.densities.split(",")]:
            shape = dataclasses.replace(CONFIGS["cfg2"], B=B, B_ref=B, density=dens)
            t = make_inputs(shape, seed=1234, device=dev, plant=5.0)
            hp, hr = t["h_pair"].requires_grad_(True), t["h_ref_pol"].requires_grad_(True)
            W = t["W"].requires_grad_(not args.no_dw)

            def ours():
                hp.grad = hr.grad = W.grad = None
                loss, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"], alpha=shape.alpha)
                loss.backward()
                return loss

            st = read_status(ours())
            ms = timed(ours, 10 if B <= 16 else 5)
            loss_ours = ours().item()

            def ref():
                hp.grad = hr.grad = W.grad = None
                loss, _, _ = dpa_port.dpa_loss_torch(hp, t["labels_pair"], t["signs_pair"], hr, t["h_ref_frz"], t["labels_ref"], W, W.detach(), shape.alpha)
                loss.backward()
                return loss

            ms_ref, loss_ref, note = None, None, ""
            try:
                loss_ref = ref().item()
                ms_ref = timed(ref, 3, warm=1)
            except torch.cuda.OutOfMemoryError:
                note = "OOM"
            hp.grad = hr.grad = W.grad = None
            torch.cuda.empty_cache()
            rec = dict(pairs=B, tokens=shape.tokens, density=dens, rows_pair=st["n_pair_active"], rows_ref=st["n_ref_active"], ms=round(ms, 3),
                       mtok_s=round(shape.tokens / ms / 1e3, 3), ref_ms=None if ms_ref is None else round(ms_ref, 2),
                       speedup=None if ms_ref is None else round(ms_ref / ms, 1), loss=loss_ours, ref_loss=loss_ref, note=note)
            rows.append(rec)
            print(json.dumps(rec), flush=True)
            del t, hp, hr, W
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("# cfg-5 sweep: tokens x phrase density at the 7B lm_head shape (H 4096, V 32000, T 2048), one B200, bf16\n\n")
        f.write("`python tests/sweep_cfg5.py` -- `ours` = halva_b200.dpa_loss forward+backward (fused-gradient, lm_head trainable); `reference chain` =\n"
                "the reference's ATen op sequence (oracle/dpa_port.py) on the same GPU. tokens = every fed position (2B*T + B*T).\n\n")
        f.write("| pairs | tokens | density | active rows (pair+ref) | ours ms | ours Mtok/s | reference chain ms | speed-up | rel. loss diff |\n|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            d = "" if r["ref_loss"] is None else f"{abs(r['loss'] - r['ref_loss']) / abs(r['ref_loss']):.1e}"
            f.write(f"| {r['pairs']} | {r['tokens']} | {r['density']} | {r['rows_pair']}+{r['rows_ref']} | {r['ms']} | {r['mtok_s']} | "
                    f"{r['ref_ms'] if r['ref_ms'] is not None else r['note']} | {r['speedup'] if r['speedup'] else '-'} | {d} |\n")
    print(open(args.out).read())


if __name__ == "__main__":
    main()
