"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the reference itself.

Runs the AST-lifted, unmodified reference compute_loss (+ autograd backward) of
/root/reference/llava/train/halva_trainer.py:392-592 on small seeded inputs and stores the
inputs and the reference's outputs.  Works only in the build container (needs
/root/reference); the fixtures it writes are committed so that the GPU box can check the
oracle and the CUDA path against the reference's real results.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from halva_b200.synth import DPAShape, make_inputs  # noqa: E402
from oracle.lift_reference import run_reference, reference_available  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def _case_inputs(name):
    """name -> (inputs dict of CPU torch tensors, torch dtype)"""
    if name == "tiny_fp32":
        return make_inputs(DPAShape(B=2, T=48, H=64, V=512, n_sys=3, n_img=8, n_q=3, k_max=3, density=0.2),
                           seed=11, dtype=torch.float32), torch.float32
    if name == "tiny_bf16":
        return make_inputs(DPAShape(B=2, T=48, H=64, V=512, n_sys=3, n_img=8, n_q=3, k_max=3, density=0.2),
                           seed=11, dtype=torch.bfloat16), torch.bfloat16
    if name == "small_fp32_fullft":      # W_ref != W, ragged K across samples (log 2 terms)
        return make_inputs(DPAShape(B=3, T=96, H=128, V=1000, n_sys=4, n_img=16, n_q=5, k_max=4, density=0.15,
                                    alpha=0.5), seed=5, dtype=torch.float32, shared_head=False), torch.float32
    if name == "small_bf16":
        return make_inputs(DPAShape(B=3, T=160, H=256, V=1000, n_sys=5, n_img=24, n_q=6, k_max=4, density=0.15,
                                    B_ref=2, T_ref=130), seed=6, dtype=torch.bfloat16), torch.bfloat16
    if name == "k1_alpha0_fp32":         # single phrase, divergence switched off
        return make_inputs(DPAShape(B=2, T=40, H=64, V=300, n_sys=2, n_img=6, n_q=2, k_max=1, density=0.1,
                                    alpha=0.0), seed=3, dtype=torch.float32), torch.float32
    if name == "sparse_ids_fp32":        # phrase ids {2,5,9}: columns follow the sorted-unique rank
        inp = make_inputs(DPAShape(B=2, T=64, H=64, V=400, n_sys=3, n_img=8, n_q=3, k_max=3, density=0.2),
                          seed=9, dtype=torch.float32)
        s = inp["signs_pair"]
        remap = {1: 2, 2: 5, 3: 9}
        out = s.clone()
        for a, b in remap.items():
            out[s == a] = b
        inp["signs_pair"] = out
        return inp, torch.float32
    if name == "no_zero_sign_fp32":      # every pos sign >= 1: U[0] (=1) is still dropped (:412-415)
        inp = make_inputs(DPAShape(B=2, T=32, H=64, V=300, n_sys=2, n_img=4, n_q=2, k_max=2, density=0.3),
                          seed=4, dtype=torch.float32)
        s = inp["signs_pair"].clone()
        s[s == 0] = 1
        s[s == -100] = 1
        # make sure ids 2 and 3 exist in both halves so that K_pos == K_neg == 2
        B = s.shape[0] // 2
        for half in (0, 1):
            rows = slice(half * B, (half + 1) * B)
            lab = inp["labels_pair"][rows]
            valid = (lab[:, 1:] != -100)
            t0 = int(valid[0].nonzero()[0]) + 1
            s[half * B, t0:t0 + 2] = 2
            s[half * B, t0 + 3:t0 + 5] = 3
        inp["signs_pair"] = s
        return inp, torch.float32
    if name in ("norm_fp32", "norm_bf16"):   # final RMSNorm folded in: h_* are the PRE-norm stream
        dt = torch.float32 if name.endswith("fp32") else torch.bfloat16
        inp = make_inputs(DPAShape(B=2, T=64, H=128, V=512, n_sys=3, n_img=8, n_q=3, k_max=3, density=0.2),
                          seed=21, dtype=dt, shared_head=name.endswith("bf16"))
        g = torch.Generator().manual_seed(77)
        for k in ("h_pair", "h_ref_pol", "h_ref_frz"):          # un-normalised residual stream: rows of very different scale
            scale = (0.5 + 3.0 * torch.rand(inp[k].shape[:2] + (1,), generator=g))
            inp[k] = (inp[k].float() * scale).to(dt)
        inp["norm_w"] = (1.0 + 0.2 * torch.randn(128, generator=g)).to(dt)
        inp["norm_w_ref"] = inp["norm_w"] if name.endswith("bf16") else (inp["norm_w"].float() + 0.05 * torch.randn(128, generator=g)).to(dt)
        inp["norm_eps"] = 1e-5
        # keep the logits in the usual range: the normalised rows have unit RMS, the synthetic W assumes |h| ~ 1.5
        inp["W"] = (inp["W"].float() * 1.5).to(dt)
        inp["W_ref"] = inp["W"] if name.endswith("bf16") else (inp["W_ref"].float() * 1.5).to(dt)
        return inp, dt
    raise KeyError(name)


CASES = ["tiny_fp32", "tiny_bf16", "small_fp32_fullft", "small_bf16", "k1_alpha0_fp32", "sparse_ids_fp32",
         "no_zero_sign_fp32", "norm_fp32", "norm_bf16"]


def _np(t):
    return t.detach().float().numpy() if t.is_floating_point() else t.detach().numpy()


def generate(name):
    inp, dtype = _case_inputs(name)
    args = (inp["h_pair"], inp["labels_pair"], inp["signs_pair"], inp["h_ref_pol"], inp["h_ref_frz"],
            inp["labels_ref"], inp["W"], inp["W_ref"])
    torch.manual_seed(0)
    nk = {}
    if "norm_w" in inp:
        nk = dict(norm_w=inp["norm_w"], norm_w_ref=inp["norm_w_ref"], norm_eps=inp["norm_eps"])
    ref = run_reference(*args, alpha=inp["alpha"], **nk)
    ref0 = run_reference(*args, alpha=0.0, need_grad=False, **nk)    # loss at alpha 0 == alignment term
    # the same reference code fed fp32 copies of the same (bf16-valued) tensors: "fp32 maths" comparator
    f32 = [a.float() if a.is_floating_point() else a for a in args]
    nk32 = {k: (v.float() if torch.is_tensor(v) else v) for k, v in nk.items()}
    ref32 = run_reference(*f32, alpha=inp["alpha"], **nk32)
    ref32_0 = run_reference(*f32, alpha=0.0, need_grad=False, **nk32)
    out = dict(
        dtype=np.array("bf16" if dtype == torch.bfloat16 else "fp32"),
        alpha=np.float64(inp["alpha"]),
        h_pair=_np(inp["h_pair"]), labels_pair=_np(inp["labels_pair"]), signs_pair=_np(inp["signs_pair"]),
        h_ref_pol=_np(inp["h_ref_pol"]), h_ref_frz=_np(inp["h_ref_frz"]), labels_ref=_np(inp["labels_ref"]),
        W=_np(inp["W"]), W_ref=_np(inp["W_ref"]),
        ref_loss=np.float64(ref["loss"].item()), ref_align=np.float64(ref0["loss"].item()),
        ref_dH_pair=_np(ref["dH_pair"]), ref_dH_ref=_np(ref["dH_ref"]), ref_dW=_np(ref["dW"]),
        ref32_loss=np.float64(ref32["loss"].item()), ref32_align=np.float64(ref32_0["loss"].item()),
        ref32_dH_pair=_np(ref32["dH_pair"]), ref32_dH_ref=_np(ref32["dH_ref"]), ref32_dW=_np(ref32["dW"]),
    )
    if nk:
        out.update(norm_w=_np(inp["norm_w"]), norm_w_ref=_np(inp["norm_w_ref"]), norm_eps=np.float64(inp["norm_eps"]),
                   ref_d_norm_w=_np(ref["d_norm_w"]), ref32_d_norm_w=_np(ref32["d_norm_w"]))
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name:22s} loss={out['ref_loss']:.7f} align={out['ref_align']:.7f} "
          f"loss32={out['ref32_loss']:.7f}  -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def generate_quirks():
    """Behaviour of the reference on ill-formed phrase ids (SURVEY.md section 3.2 'Quirks')."""
    inp, _ = _case_inputs("tiny_fp32")
    base = (inp["h_pair"], inp["labels_pair"], inp["signs_pair"], inp["h_ref_pol"], inp["h_ref_frz"],
            inp["labels_ref"], inp["W"], inp["W_ref"])
    B = inp["signs_pair"].shape[0] // 2
    res = {}
    # K == 0: no phrase tokens at all -> mean of empty -> NaN
    s = torch.zeros_like(inp["signs_pair"])
    r = run_reference(base[0], base[1], s, *base[3:], alpha=0.4, need_grad=False)
    res["k0_is_nan"] = bool(torch.isnan(r["loss"]))
    # K_pos = 2, K_neg = 3 -> RuntimeError
    s = inp["signs_pair"].clone()
    s[s > 2] = 2
    lab = inp["labels_pair"]
    t0 = int((lab[B, 1:] != -100).nonzero()[0]) + 1
    s[B, t0] = 7
    s[B, t0 + 1] = 8
    s[:B][s[:B] > 2] = 2
    try:
        up = torch.unique(s[:B, 1:].masked_fill(s[:B, 1:] == -100, 0)).numel() - 1
        un = torch.unique(s[B:, 1:].masked_fill(s[B:, 1:] == -100, 0)).numel() - 1
        run_reference(base[0], base[1], s, *base[3:], alpha=0.4, need_grad=False)
        res["kmismatch_raises"] = False
    except RuntimeError:
        res["kmismatch_raises"] = True
    res["kmismatch_kpos_kneg"] = [int(up), int(un)]
    np.savez(os.path.join(GOLDEN_DIR, "quirks.npz"), **{k: np.array(v) for k, v in res.items()})
    print("quirks:", res)


if __name__ == "__main__":
    if not reference_available():
        sys.exit("needs /root/reference (build container only)")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for c in CASES:
        if not only or c in only:
            generate(c)
    if not only:
        generate_quirks()
