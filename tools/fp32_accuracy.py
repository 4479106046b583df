"""Accuracy of the fp32-input path at BASELINE configs[0] size against a torch fp64 restatement on the same GPU, next to
(a) torch's own fp32 op chain (TF32 off: what an fp32 reference run achieves) and (b) the bf16 cast of the same inputs.
Prints relative errors; run under gpurun.  Usage: python tools/fp32_accuracy.py [cfg1|small] [plant]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from halva_b200 import dpa_loss  # noqa: E402
from halva_b200.synth import CONFIGS, make_inputs  # noqa: E402
from test_gpu_fp32 import _torch_fp64  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def ours(t, cast=None, **kw):
    c = (lambda x: x) if cast is None else (lambda x: x.to(cast))
    hp, hr, W = c(t["h_pair"]).clone().requires_grad_(True), c(t["h_ref_pol"]).clone().requires_grad_(True), c(t["W"]).clone().requires_grad_(True)
    loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=c(t["h_ref_frz"]), alpha=t["alpha"], **kw)
    loss.backward()
    f = lambda x: x.double().cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad))


def torch32(t):
    from oracle import dpa_port
    hp, hr, W = t["h_pair"].clone().requires_grad_(True), t["h_ref_pol"].clone().requires_grad_(True), t["W"].clone().requires_grad_(True)
    loss, align, div = dpa_port.dpa_loss_torch(hp, t["labels_pair"], t["signs_pair"], hr, t["h_ref_frz"], t["labels_ref"], W, W.detach(), t["alpha"])
    loss.backward()
    f = lambda x: x.double().cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad))


def report(tag, c, o):
    r = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    print(f"{tag:28s} loss {abs(c['loss'] - o['loss']) / abs(o['loss']):.2e} align {abs(c['align'] - o['align']) / abs(o['align']):.2e} "
          f"div {abs(c['div'] - o['div']) / abs(o['div']):.2e} dH_pair {r(c['dH_pair'], o['dH_pair']):.2e} "
          f"dH_ref {r(c['dH_ref'], o['dH_ref']):.2e} dW {r(c['dW'], o['dW']):.2e}", flush=True)


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    plant = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
    for seed in (99, 7):
        t = make_inputs(CONFIGS[cfg], seed=seed, dtype=torch.float32, device="cuda", plant=plant)
        o = _torch_fp64(t)
        print(f"-- {cfg} seed {seed} plant {plant}: fp64 loss {o['loss']:.8f} align {o['align']:.8f} div {o['div']:.8f}")
        report("ours fp32 inputs (split)", ours(t), o)
        report("torch fp32 op chain", torch32(t), o)
        report("ours, inputs cast to bf16", ours(t, torch.bfloat16, round_logits=False), o)
        tb = {k: (v.to(torch.bfloat16).float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in t.items()}
        ob = _torch_fp64(tb)
        report("ours bf16 vals vs fp64(same)", ours(tb, torch.bfloat16, round_logits=False), ob)
        report("ours split, bf16-valued", ours(tb), ob)
