"""fp32 inputs (BASELINE.json configs[0] is the fp32 case): hidden states and lm_head weights in fp32, every tensor-core
operand a split-bf16 pair (hi + lo, three K segments per GEMM), fp32-accurate logits and dLogits, fp32 gradients.

Stated tolerances (north star: "tighter in fp32"; SURVEY.md section 8d parity gates): <= 1e-5 relative on loss / align,
<= 1e-4 (+ 5e-7 absolute) on the divergence (a KL row is 1e-4..1e-2 while every fp32 logit and log-sum-exp carries ~1e-7
absolute rounding, in the reference's own fp32 chain too: tiny batches with a divergence of ~1e-3 sit on that floor) and
<= 1e-4 on dH / dW -- against the reference's OWN fp32 outputs (tests/golden) and the
fp64 oracle on the same fp32 inputs."""
import dataclasses

import numpy as np
import pytest
import torch

from dpa_testutil import load_golden, rel
from oracle import dpa_oracle

pytestmark = pytest.mark.gpu

TOL_LOSS, TOL_DIV, TOL_GRAD = 1e-5, 1e-4, 1e-4
# (Holds at H = 4096 too because the K loop of the split GEMMs is promoted chunk-wise through registers: the tensor core
# adds every K = 16 step into its fp32 accumulator with TRUNCATION -- tools/tc_accum_probe.py: a same-sign K = 4096 dot
# product comes out 9e-6 low -- which, unpromoted, a sum of ~100 phrase log-probs turns into 3e-5 on the alignment loss.
# Measured at the cfg-1 shape, tools/fp32_accuracy.py: loss 3e-6, dH 6e-7 / 1.5e-5, dW 5e-7; torch's own fp32 op chain
# on the same GPU: 2.5e-6, 3.5e-6 / 1.4e-4, 2.2e-6.)
FP32_GOLDEN = ["tiny_fp32", "small_fp32_fullft", "k1_alpha0_fp32", "sparse_ids_fp32", "no_zero_sign_fp32"]


def _cuda32(g):
    f = lambda k: torch.from_numpy(np.asarray(g[k], np.float32)).cuda()
    ix = lambda k: torch.from_numpy(g[k]).cuda()
    return dict(h_pair=f("h_pair"), labels_pair=ix("labels_pair"), signs_pair=ix("signs_pair"), h_ref_pol=f("h_ref_pol"),
                h_ref_frz=f("h_ref_frz"), labels_ref=ix("labels_ref"), W=f("W"), W_ref=f("W_ref"), alpha=float(g["alpha"]))


def _run(t, need_dW=True, grad_out=1.0, **kw):
    from halva_b200 import dpa_loss
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(need_dW)
    shared = t["W_ref"] is t["W"] or torch.equal(t["W_ref"], t["W"])
    loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"],
                                W_ref=None if shared else t["W_ref"], alpha=t["alpha"], **kw)
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    assert hp.grad.dtype == torch.float32 and hr.grad.dtype == torch.float32 and (W.grad is None or W.grad.dtype == torch.float32)
    f = lambda x: None if x is None else x.double().cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad))


def _oracle(t):
    f = lambda x: x.double().cpu().numpy()
    return dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]),
                               f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]), f(t["W_ref"]), t["alpha"], mode="fp64")


def _check(c, o, tag, tol_loss=TOL_LOSS):
    assert abs(c["loss"] - o["loss"]) <= tol_loss * abs(o["loss"]), (tag, c["loss"], o["loss"])
    assert abs(c["align"] - o["align"]) <= tol_loss * abs(o["align"]), (tag, c["align"], o["align"])
    assert abs(c["div"] - o["div"]) <= TOL_DIV * abs(o["div"]) + 5e-7, (tag, c["div"], o["div"])
    for k in ("dH_pair", "dH_ref", "dW"):
        if o[k] is not None and np.linalg.norm(o[k]) > 0:
            assert rel(c[k], o[k]) <= TOL_GRAD, (tag, k, rel(c[k], o[k]))


@pytest.mark.parametrize("bwd", ["fused", "recompute"])
@pytest.mark.parametrize("name", FP32_GOLDEN)
def test_fp32_golden_vs_reference_and_oracle(name, bwd):
    """The reference's own fp32 execution (inputs NOT rounded to bf16) and the fp64 oracle on the same inputs; both gradient
    strategies (fused: the chunked pass keeps fp32 logits, 4 GEMM units per reference-sample row; recompute: 6)."""
    g = load_golden(name)
    t = _cuda32(g)
    c = _run(t, backward=bwd)
    _check(c, _oracle(t), name + "/oracle")
    ref = dict(loss=float(g["ref_loss"]), align=float(g["ref_align"]), dH_pair=g["ref_dH_pair"], dH_ref=g["ref_dH_ref"], dW=g["ref_dW"])
    assert abs(c["loss"] - ref["loss"]) <= TOL_LOSS * abs(ref["loss"]) and abs(c["align"] - ref["align"]) <= TOL_LOSS * abs(ref["align"])
    for k in ("dH_pair", "dH_ref", "dW"):
        if np.linalg.norm(ref[k]) > 0:
            assert rel(c[k], ref[k]) <= TOL_GRAD, (name, k, rel(c[k], ref[k]))
    o = _oracle(t)
    assert np.all(c["dH_pair"][:, -1] == 0) and np.all(c["dH_pair"][:, :-1][~o["active_pair"]] == 0)
    assert np.all(c["dH_ref"][:, -1] == 0) and np.all(c["dH_ref"][:, :-1][~o["active_ref"]] == 0)


@pytest.mark.parametrize("name", ["tiny_bf16", "small_bf16"])
def test_parity_gate_ii_reference_fp32_on_bf16_values(name):
    """SURVEY.md section 8d, parity gate (ii): the reference executed in fp32 on the SAME bf16-valued inputs (`ref32_*` of the
    bf16 golden cases) -- <= 1e-5 on the loss, <= 1e-4 on the gradients.  Fed as fp32 tensors, those values go through the
    split path exactly (mid / lo parts are zero) and the gradients are not squeezed through a bf16 dLogits operand."""
    g = load_golden(name)
    assert str(g["dtype"]) == "bf16"
    t = _cuda32(g)
    assert torch.equal(t["W"], t["W"].to(torch.bfloat16).float())           # bf16-representable values
    c = _run(t)
    assert abs(c["loss"] - float(g["ref32_loss"])) <= TOL_LOSS * abs(float(g["ref32_loss"]))
    assert abs(c["align"] - float(g["ref32_align"])) <= TOL_LOSS * abs(float(g["ref32_align"]))
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(c[k], g["ref32_" + k]) <= TOL_GRAD, (name, k, rel(c[k], g["ref32_" + k]))


def _synth32(seed=3, **over):
    from halva_b200.synth import CONFIGS, make_inputs
    shape = dataclasses.replace(CONFIGS["small"], **over)
    return make_inputs(shape, seed=seed, dtype=torch.float32, device="cuda"), shape


@pytest.mark.parametrize("over", [dict(), dict(T=129, T_ref=77), dict(V=1003, H=128), dict(V=77, H=64, T=40, T_ref=33),
                                  dict(B=1, B_ref=3, k_max=1), dict(H=1024, V=2048 + 64, T=64, T_ref=64, B=1)])
def test_fp32_synthetic_shapes_vs_oracle(over):
    t, _ = _synth32(**over)
    assert t["h_pair"].dtype == torch.float32
    _check(_run(t), _oracle(t), over)


def test_fp32_beats_the_bf16_cast_and_is_chunk_invariant():
    """What the split buys: casting the same fp32 inputs to bf16 (the only option before) is orders of magnitude further
    from the fp64 result.  Row chunking / dense execution / frozen lm_head / grad_output do not change the result."""
    t, _ = _synth32(seed=7, T=200, T_ref=150)
    o = _oracle(t)
    c = _run(t)
    from test_gpu_dpa import _run_cuda
    tb = {k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in t.items()}
    b = _run_cuda(tb, round_logits=False)
    err32, err16 = abs(c["loss"] - o["loss"]) / abs(o["loss"]), abs(b["loss"] - o["loss"]) / abs(o["loss"])
    g32, g16 = rel(c["dW"], o["dW"]), rel(b["dW"], o["dW"])
    assert err32 <= TOL_LOSS and g32 <= TOL_GRAD
    assert g16 > 20 * g32, (g16, g32, err16, err32)
    r = _run(t, backward="recompute")                    # the two gradient strategies agree to fp32 rounding
    assert abs(r["loss"] - c["loss"]) <= 2e-6 * abs(c["loss"]) and abs(r["div"] - c["div"]) <= 1e-4 * abs(c["div"]) + 5e-7
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(r[k], c[k]) <= 2e-5, (k, rel(r[k], c[k]))
    for kw in (dict(chunk_rows=256), dict(dense=True), dict(dense=True, chunk_rows=256)):
        d = _run(t, **kw)
        assert abs(d["loss"] - c["loss"]) <= 2e-6 * abs(c["loss"]), kw
        for k in ("dH_pair", "dH_ref", "dW"):
            assert rel(d[k], c[k]) <= 2e-6, (kw, k, rel(d[k], c[k]))
    f = _run(t, need_dW=False, grad_out=0.5)
    assert f["dW"] is None and rel(f["dH_ref"], 0.5 * c["dH_ref"]) <= 1e-6 and rel(f["dH_pair"], 0.5 * c["dH_pair"]) <= 1e-6


def test_fp32_eval_and_argument_checks():
    from halva_b200 import dpa_loss
    t, _ = _synth32()
    c = _run(t)
    with torch.no_grad():
        loss, align, div = dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"],
                                    h_ref_frozen=t["h_ref_frz"], alpha=t["alpha"])
    # (evaluation runs the statistics-only forward: KL from the online partial sums; training the chunked pass: KL from the kept logits)
    assert abs(loss.item() - c["loss"]) <= 2e-6 * abs(c["loss"]) and abs(div.item() - c["div"]) <= 1e-4 * abs(c["div"]) + 5e-7
    args = (t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"])
    with pytest.raises(ValueError):
        dpa_loss(*args, h_ref_frozen=t["h_ref_frz"], alpha=0.4, norm_weight=t["W"][0])
    with pytest.raises(TypeError):                       # mixed precisions are refused, not silently cast
        dpa_loss(*args, h_ref_frozen=t["h_ref_frz"].to(torch.bfloat16), alpha=0.4)
    with pytest.raises(TypeError):
        dpa_loss(*args[:5], t["W"].to(torch.bfloat16), h_ref_frozen=t["h_ref_frz"], alpha=0.4)


def _torch_fp64(t):
    """Plain PyTorch fp64 restatement of the op chain on the GPU (autograd backward) for the full-size case."""
    d = lambda x: x.double()
    hp, hr, W = d(t["h_pair"]).requires_grad_(True), d(t["h_ref_pol"]).requires_grad_(True), d(t["W"]).requires_grad_(True)
    B = hp.shape[0] // 2

    def logps(h, w, y):
        z = (h @ w.t())[:, :-1]
        tgt = y[:, 1:].clamp_min(0)
        return torch.gather(z.log_softmax(-1), 2, tgt.unsqueeze(2)).squeeze(2) * (y[:, 1:] != -100), z

    lp, _ = logps(hp, W, t["labels_pair"])
    g = t["signs_pair"][:, 1:].masked_fill(t["signs_pair"][:, 1:] == -100, 0)

    def sums(l, s):
        ids = torch.unique(s, sorted=True)
        return torch.stack([(l * (s == k)).sum(-1) for k in ids[1:]], 1)

    align = torch.log(1 + torch.exp(sums(lp[B:], g[B:]) - sums(lp[:B], g[:B]))).mean()
    _, zp = logps(hr, W, t["labels_ref"])
    with torch.no_grad():
        _, zr = logps(d(t["h_ref_frz"]), d(t["W_ref"]), t["labels_ref"])
    lq, lpp = zr.log_softmax(-1), zp.log_softmax(-1)
    kl = (lq.exp() * (lq - lpp)).sum(-1) * (t["labels_ref"][:, 1:] != -100)
    div = kl.sum() / kl.shape[0]
    loss = align + t["alpha"] * div
    loss.backward()
    f = lambda x: x.cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad))


def test_fp32_cfg1_full_size():
    """BASELINE.json configs[0] at full size (1 pair + 1 reference sample, T 1024, H 4096, V 32000, fp32) on the GPU."""
    from halva_b200.synth import CONFIGS, make_inputs
    t = make_inputs(CONFIGS["cfg1"], seed=99, dtype=torch.float32, device="cuda", plant=5.0)
    c = _run(t)
    _check(c, _torch_fp64(t), "cfg1")
