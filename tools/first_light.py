"""GPU bring-up helper: each stage runs in its own process (a trap in one stage must not poison the rest).
    python tools/first_light.py all            # run every stage with a timeout, log to gpurun_out/first_light.log
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_gemm(variant):
    import torch
    from halva_b200 import debug_gemm
    a_major, b_major, bn = [int(x) for x in variant.split(",")]
    for (M, N, K) in [(128, 256, 64), (256, 512, 256), (384, 320, 200), (1024, 1024, 4096)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
        B = torch.randn(N, K, generator=g, device="cuda").to(torch.bfloat16)
        ref = A.float() @ B.float().t()
        As = A if a_major == 0 else A.t().contiguous()
        Bs = B if b_major == 0 else B.t().contiguous()
        D = debug_gemm(As, Bs, a_major, b_major, bn)
        torch.cuda.synchronize()
        err = (D - ref).abs().max().item()
        print(f"gemm a_major={a_major} b_major={b_major} bn={bn} M={M} N={N} K={K}: max_abs_err={err:.4g} "
              f"(ref max {ref.abs().max().item():.3g}) {'OK' if err < 2e-3 * ref.abs().max().item() + 1e-3 else 'MISMATCH'}")
        if err > 1.0 and M <= 256:
            bad = (D - ref).abs() > 0.5
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            print("   first bad rows", rows, "cols", cols, "D[0,:4]", D[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())


def stage_gemm_perf(variant):
    import torch
    from halva_b200 import debug_gemm
    a_major, b_major, bn = [int(x) for x in variant.split(",")]
    M = N = K = 8192
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        debug_gemm(A, B, a_major, b_major, bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        debug_gemm(A, B, a_major, b_major, bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm perf {variant} 8192^3: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    e0.record()
    for _ in range(10):
        A @ B.t()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas 8192^3: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def stage_parity(name):
    import dataclasses
    import numpy as np
    import torch
    from halva_b200 import dpa_loss, read_status
    from halva_b200.synth import CONFIGS, make_inputs
    from oracle import dpa_oracle
    t = make_inputs(CONFIGS[name], seed=3, device="cuda")
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)
    for rl, mode in ((True, "bf16"), (False, "fp64")):
        hp.grad = hr.grad = W.grad = None
        loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W,
                                    h_ref_frozen=t["h_ref_frz"], alpha=t["alpha"], round_logits=rl)
        print("status", read_status(loss))
        loss.backward()
        torch.cuda.synchronize()
        f = lambda x: x.float().cpu().numpy()
        o = dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(),
                                f(t["h_ref_pol"]), f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]), f(t["W_ref"]),
                                t["alpha"], mode=mode)
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        print(f"parity[{name},{mode}] loss {loss.item():.7f} vs {o['loss']:.7f} | align {align.item():.7f} vs {o['align']:.7f} | "
              f"div {div.item():.7f} vs {o['div']:.7f}")
        print(f"   rel dH_pair {rel(f(hp.grad), o['dH_pair']):.3e} dH_ref {rel(f(hr.grad), o['dH_ref']):.3e} dW {rel(f(W.grad), o['dW']):.3e}")


def stage_cfg2(arg):
    import torch
    from halva_b200 import dpa_loss, _lib, read_status
    from halva_b200.synth import CONFIGS, make_inputs, active_rows
    need_dw = "dw" in arg
    dense = "dense" in arg
    shape = CONFIGS["cfg2"]
    t = make_inputs(shape, seed=1234, device="cuda")
    print("active rows (pair, ref):", active_rows(t["labels_pair"].cpu(), t["signs_pair"].cpu(), t["labels_ref"].cpu()))
    hp = t["h_pair"].requires_grad_(True)
    hr = t["h_ref_pol"].requires_grad_(True)
    W = t["W"].requires_grad_(need_dw)

    def step():
        hp.grad = hr.grad = W.grad = None
        loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W,
                                    h_ref_frozen=t["h_ref_frz"], alpha=t["alpha"], dense=dense)
        loss.backward()
        return loss
    loss = step()
    torch.cuda.synchronize()
    print("loss", loss.item(), read_status(loss), "max mem GB", torch.cuda.max_memory_allocated() / 1e9)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    prof = _lib.profile_collect()
    print(f"cfg2 need_dw={need_dw} dense={dense}: {ms:.3f} ms/step  {shape.tokens / ms * 1e3 / 1e6:.3f} M tokens/s")
    for k, (t_ms, ln) in prof.items():
        print(f"   {k:10s} {t_ms / n:9.3f} ms/step  launches/step {ln / n:.1f}")


STAGES = dict(gemm=stage_gemm, gemm_perf=stage_gemm_perf, parity=stage_parity, cfg2=stage_cfg2)

PLAN = [("gemm", "0,0,256", 120), ("gemm", "0,0,128", 120), ("gemm", "0,1,256", 120), ("gemm", "1,1,256", 120),
        ("gemm", "1,0,256", 120), ("gemm_perf", "0,0,256", 120), ("gemm_perf", "0,1,256", 120), ("gemm_perf", "1,1,256", 120),
        ("parity", "tiny", 180), ("parity", "small", 180), ("cfg2", "nodw", 300), ("cfg2", "dw", 300)]

if __name__ == "__main__":
    if sys.argv[1] == "all":
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        log = open(os.path.join(ROOT, "gpurun_out", "first_light.log"), "w")
        for name, arg, tmo in PLAN:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, name, arg], capture_output=True, text=True, timeout=tmo)
                out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode else "")
                tag = f"rc={r.returncode}"
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                tag = "TIMEOUT"
            msg = f"===== {name} {arg} [{tag}, {time.time() - t0:.1f}s]\n{out}\n"
            print(msg, flush=True)
            log.write(msg)
            log.flush()
    else:
        STAGES[sys.argv[1]](sys.argv[2] if len(sys.argv) > 2 else "")
