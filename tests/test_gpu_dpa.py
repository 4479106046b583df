"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from dpa_testutil import GOLDEN_CASES, load_golden, rel
from oracle import dpa_oracle

pytestmark = pytest.mark.gpu

# stated tolerances (BASELINE.json north_star): bf16 rel err <= 1e-3 losses, <= 1e-2 gradients;
# fp32 maths on the same bf16 values: <= 2e-5 on loss / align (gradients still pass through a bf16 dLogits operand).
# The divergence alone gets 2e-4 in fp32 mode: a KL row is ~1e-4..1e-2 while each fp32 log-sum-exp carries ~1e-7
# absolute rounding, in this kernel and in the reference's own fp32 softmax().log() chain alike -- the fp64 oracle
# is more exact than either.  (On the total loss the bound stays 2e-5.)
TOL_LOSS_BF16, TOL_GRAD_BF16, TOL_LOSS_FP32, TOL_DIV_FP32 = 1e-3, 1e-2, 2e-5, 2e-4


def _to_cuda(g):
    bf = lambda k: torch.from_numpy(g[k]).to(torch.bfloat16).cuda()
    ix = lambda k: torch.from_numpy(g[k]).cuda()
    t = dict(h_pair=bf("h_pair"), labels_pair=ix("labels_pair"), signs_pair=ix("signs_pair"), h_ref_pol=bf("h_ref_pol"),
             h_ref_frz=bf("h_ref_frz"), labels_ref=ix("labels_ref"), W=bf("W"), W_ref=bf("W_ref"), alpha=float(g["alpha"]))
    if "norm_w" in g:      # final RMSNorm folded into the op: the hidden states are the pre-norm stream
        t.update(norm_w=bf("norm_w"), norm_w_ref=bf("norm_w_ref"), norm_eps=float(g["norm_eps"]))
    return t


def _run_cuda(t, round_logits=True, need_dW=True, grad_out=1.0, **kw):
    from halva_b200 import dpa_loss
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(need_dW)
    W_ref = W if t["W_ref"] is t["W"] else t["W_ref"]
    nw = None
    if "norm_w" in t:
        nw = t["norm_w"].clone().requires_grad_(True)
        kw = dict(kw, norm_weight=nw, norm_weight_ref=t["norm_w_ref"], norm_eps=t["norm_eps"])
    loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"],
                                W_ref=W_ref.detach() if W_ref is not W else None, alpha=t["alpha"], round_logits=round_logits, **kw)
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    f = lambda x: None if x is None else x.float().cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad),
                d_norm_w=None if nw is None else f(nw.grad))


def _oracle(t, mode):
    f = lambda x: x.float().cpu().numpy()
    nk = {}
    if "norm_w" in t:
        # the normed rows are the bf16 GEMM operand in every mode: keep the norm's own rounding points
        nk = dict(norm_w=f(t["norm_w"]), norm_w_ref=f(t["norm_w_ref"]), norm_eps=t["norm_eps"], norm_mode="bf16")
    return dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]),
                               f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]), f(t["W_ref"]), t["alpha"], mode=mode, **nk)


def _close(a, b, tol):
    return abs(a - b) <= tol * max(abs(b), 1e-6)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_inputs_vs_oracle(name):
    """Every golden case (inputs cast to bf16): CUDA vs oracle, in reference-bf16 mode (fused-gradient forward and
    recompute backward) and in fp32-maths mode."""
    g = load_golden(name)
    t = _to_cuda(g)
    for round_logits, mode, tol, bwd in ((True, "bf16", TOL_LOSS_BF16, "fused"), (True, "bf16", TOL_LOSS_BF16, "recompute"),
                                          (False, "fp64", TOL_LOSS_FP32, "recompute")):
        c = _run_cuda(t, round_logits=round_logits, backward=bwd)
        o = _oracle(t, mode)
        assert _close(c["loss"], o["loss"], tol), (name, mode, bwd, c["loss"], o["loss"])
        assert _close(c["align"], o["align"], tol), (name, mode, bwd, c["align"], o["align"])
        tol_div = TOL_DIV_FP32 if mode == "fp64" else tol
        assert abs(c["div"] - o["div"]) <= tol_div * abs(o["div"]) + 1e-7, (name, mode, bwd, c["div"], o["div"])
        for k in ("dH_pair", "dH_ref", "dW") + (("d_norm_w",) if "norm_w" in t else ()):
            assert rel(c[k], o[k]) <= TOL_GRAD_BF16, (name, mode, bwd, k, rel(c[k], o[k]))
        # structural zeros: last position and rows without contribution get exactly zero gradient
        assert np.all(c["dH_pair"][:, -1] == 0) and np.all(c["dH_ref"][:, -1] == 0)
        assert np.all(c["dH_pair"][:, :-1][~o["active_pair"]] == 0)
        assert np.all(c["dH_ref"][:, :-1][~o["active_ref"]] == 0)


@pytest.mark.parametrize("bwd", ["fused", "recompute"])
@pytest.mark.parametrize("name", ["tiny_bf16", "small_bf16", "norm_bf16"])
def test_golden_reference_outputs_bf16(name, bwd):
    """bf16 golden cases: CUDA vs what the reference itself produced in bf16 (tests/golden, oracle/gen_golden.py)."""
    g = load_golden(name)
    c = _run_cuda(_to_cuda(g), round_logits=True, backward=bwd)
    assert _close(c["loss"], float(g["ref_loss"]), TOL_LOSS_BF16)
    assert _close(c["align"], float(g["ref_align"]), TOL_LOSS_BF16)
    for k in ("dH_pair", "dH_ref", "dW") + (("d_norm_w",) if "norm_w" in g else ()):
        assert rel(c[k], g["ref_" + k]) <= TOL_GRAD_BF16, (k, rel(c[k], g["ref_" + k]))
    if "norm_w" in g:
        return      # the reference's fp32 execution keeps the normed rows in fp32; here they are the bf16 GEMM operand
    # and against the reference's fp32 execution on the same bf16 values
    c32 = _run_cuda(_to_cuda(g), round_logits=False)
    assert _close(c32["loss"], float(g["ref32_loss"]), TOL_LOSS_FP32)
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(c32[k], g["ref32_" + k]) <= TOL_GRAD_BF16


def _synth(name="small", seed=3, **over):
    import dataclasses
    from halva_b200.synth import CONFIGS, make_inputs
    shape = dataclasses.replace(CONFIGS[name], **over)
    t = make_inputs(shape, seed=seed, device="cuda")
    t["h_ref_frz"] = t.pop("h_ref_frz")
    return t, shape


@pytest.mark.parametrize("over", [dict(), dict(T=129, T_ref=77), dict(V=1000 - 8, H=192), dict(B=1, B_ref=3, k_max=1),
                                  dict(H=4096, V=2048, T=64, T_ref=64, B=1), dict(V=1003, H=128), dict(V=77, H=64, T=40, T_ref=33)])
def test_synthetic_shapes_vs_oracle(over):
    """Ragged T, V not a multiple of the tile, H not a multiple of 256, B_ref != B."""
    t, shape = _synth(**over)
    o = _oracle(t, "bf16")
    for bwd in ("fused", "recompute"):
        c = _run_cuda(t, backward=bwd)
        assert _close(c["loss"], o["loss"], TOL_LOSS_BF16) and _close(c["align"], o["align"], TOL_LOSS_BF16), bwd
        assert _close(c["div"], o["div"], TOL_LOSS_BF16), bwd
        for k in ("dH_pair", "dH_ref", "dW"):
            assert rel(c[k], o[k]) <= TOL_GRAD_BF16, (bwd, k, rel(c[k], o[k]))


@pytest.mark.parametrize("tpu", ["1", "2", "5"])
@pytest.mark.parametrize("over", [dict(V=3001, H=128), dict(V=1288, H=64, T=129, T_ref=77)])
def test_column_group_width_invariance(over, tpu, monkeypatch):
    """Work units of the statistics GEMMs: 1 (small vocabulary shards), 2 or 5 (full vocabulary) column tiles per unit -- incl.
    a ragged last group (V=3001: 12 tiles = 5 + 5 + 2) -- give the same losses and gradients, in every input form."""
    monkeypatch.setenv("DPA_TPU", tpu)
    t, shape = _synth(**over)
    o = _oracle(t, "bf16")
    for kw in (dict(backward="fused"), dict(backward="recompute"), dict(backward="fused", dense=True, chunk_rows=256)):
        c = _run_cuda(t, **kw)
        assert _close(c["loss"], o["loss"], TOL_LOSS_BF16) and _close(c["div"], o["div"], TOL_LOSS_BF16), kw
        for k in ("dH_pair", "dH_ref", "dW"):
            assert rel(c[k], o[k]) <= TOL_GRAD_BF16, (kw, k, rel(c[k], o[k]))
    # frozen logits given: the copy kernel produces the partials of the same column groups
    from halva_b200 import dpa_loss
    zr = (t["h_ref_frz"].float() @ t["W_ref"].float().T).to(torch.bfloat16)
    hp = t["h_pair"].clone().requires_grad_(True)
    loss, _, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], ref_logits=zr, alpha=t["alpha"])
    loss.backward()
    assert _close(loss.item(), o["loss"], TOL_LOSS_BF16) and _close(div.item(), o["div"], TOL_LOSS_BF16)
    assert rel(hp.grad.float().cpu().numpy(), o["dH_pair"]) <= TOL_GRAD_BF16
    # fp32 inputs: the split-bf16 tilings share the column groups
    f32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in t.items()}
    hp32 = f32["h_pair"].clone().requires_grad_(True)
    l32, _, _ = dpa_loss(hp32, f32["labels_pair"], f32["signs_pair"], f32["h_ref_pol"], f32["labels_ref"], f32["W"],
                         h_ref_frozen=f32["h_ref_frz"], alpha=t["alpha"])
    l32.backward()
    o64 = _oracle(t, "fp64")
    assert _close(l32.item(), o64["loss"], 1e-5) and rel(hp32.grad.cpu().numpy(), o64["dH_pair"]) <= 1e-4


def test_no_reference_labels_and_single_position():
    """Every reference-sample label ignored -> divergence exactly 0 and no reference-sample gradient (the reference sums an
    all-masked tensor); T = 2 leaves one predicting position per sequence."""
    t, _ = _synth()
    t = dict(t)
    t["labels_ref"] = torch.full_like(t["labels_ref"], -100)
    c = _run_cuda(t)
    o = _oracle(t, "bf16")
    assert c["div"] == 0.0 and o["div"] == 0.0 and _close(c["loss"], o["loss"], TOL_LOSS_BF16)
    assert not c["dH_ref"].any() and rel(c["dH_pair"], o["dH_pair"]) <= TOL_GRAD_BF16
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
    W2 = (0.05 * rn(200, 64)).to(torch.bfloat16)
    t2 = dict(h_pair=rn(4, 2, 64).to(torch.bfloat16), h_ref_pol=rn(2, 2, 64).to(torch.bfloat16), h_ref_frz=rn(2, 2, 64).to(torch.bfloat16),
              labels_pair=torch.tensor([[-100, 5], [-100, 6], [-100, 7], [-100, 8]], device="cuda"),
              # one phrase token per half; the other sequence carries sign 0, else the smallest id (1) would be the one dropped
              signs_pair=torch.tensor([[0, 1], [0, 0], [0, 1], [0, 0]], device="cuda"),
              labels_ref=torch.tensor([[-100, 3], [-100, -100]], device="cuda"), W=W2, W_ref=W2, alpha=0.4)
    c, o = _run_cuda(t2), _oracle(t2, "bf16")
    assert _close(c["loss"], o["loss"], TOL_LOSS_BF16) and rel(c["dW"], o["dW"]) <= TOL_GRAD_BF16


def test_int32_labels_and_frozen_head():
    t, _ = _synth()
    a = _run_cuda(t, need_dW=False)
    t32 = dict(t)
    for k in ("labels_pair", "signs_pair", "labels_ref"):
        t32[k] = t[k].to(torch.int32)
    b = _run_cuda(t32, need_dW=False)
    assert a["dW"] is None and b["dW"] is None
    assert a["loss"] == b["loss"] and np.array_equal(a["dH_pair"], b["dH_pair"])


@pytest.mark.parametrize("bwd", ["fused", "recompute"])
def test_dense_equals_compact_and_chunking_invariant(bwd):
    """Processing every position (what the reference executes) gives the same loss and gradients as the compacted
    path; the row-chunk size does not change the result."""
    t, _ = _synth()
    a = _run_cuda(t, backward=bwd)
    d = _run_cuda(t, dense=True, backward=bwd)
    ch = _run_cuda(t, chunk_rows=256, backward=bwd)
    dch = _run_cuda(t, dense=True, chunk_rows=256, backward=bwd)    # pair rows spill past chunk 0: kept-logits + recompute mix
    assert _close(d["loss"], a["loss"], 1e-6) and _close(ch["loss"], a["loss"], 1e-6) and _close(dch["loss"], a["loss"], 1e-6)
    for k in ("dH_pair", "dH_ref"):
        assert np.array_equal(a[k], d[k]) and np.array_equal(a[k], ch[k]), k
        assert rel(dch[k], a[k]) <= 1e-3, k
    assert rel(d["dW"], a["dW"]) <= 1e-4 and rel(ch["dW"], a["dW"]) <= 2e-3   # fp32 partial sums regroup (K order changes)


def test_fused_matches_recompute():
    """The two gradient strategies see the same bf16 logits: alignment loss identical, divergence equal to fp32
    summation order (online partial sums vs one pass over the kept logits), gradients equal to rounding."""
    t, _ = _synth(seed=11)
    a = _run_cuda(t, backward="fused")
    b = _run_cuda(t, backward="recompute")
    assert a["align"] == b["align"] and _close(a["div"], b["div"], 1e-5) and _close(a["loss"], b["loss"], 1e-6)
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(a[k], b[k]) <= 1e-3, (k, rel(a[k], b[k]))
    with torch.no_grad():                                  # evaluation: statistics only, no gradient work, same loss
        from halva_b200 import dpa_loss
        l, _, _ = dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"],
                           h_ref_frozen=t["h_ref_frz"], alpha=t["alpha"])
    assert _close(l.item(), a["loss"], 1e-6)


@pytest.mark.parametrize("bwd", ["fused", "recompute"])
def test_grad_output_scaling_and_determinism(bwd):
    t, _ = _synth()
    a = _run_cuda(t, backward=bwd)
    b = _run_cuda(t, backward=bwd)
    assert a["loss"] == b["loss"] and np.array_equal(a["dW"], b["dW"]) and np.array_equal(a["dH_pair"], b["dH_pair"])
    s = _run_cuda(t, grad_out=0.25, backward=bwd)       # Trainer divides by gradient_accumulation_steps
    assert rel(s["dH_pair"], 0.25 * a["dH_pair"]) <= 5e-3 and rel(s["dW"], 0.25 * a["dW"]) <= 5e-3


def test_grad_scale_hint():
    """dpa_loss(grad_scale_hint=g): the fused forward forms the gradients for the announced upstream gradient; the backward
    rescales only when the actual one differs.  Right hint, wrong hint and no hint give the same gradients (to the bf16
    rounding of dLogits), all equal to the oracle's for that grad_output; the recompute strategy ignores the hint."""
    t, _ = _synth()
    f = lambda x: x.float().cpu().numpy()
    o = dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]),
                            f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]), f(t["W_ref"]), t["alpha"], mode="bf16",
                            grad_output=0.25)
    runs = dict(none=_run_cuda(t, grad_out=0.25, backward="fused"),
                right=_run_cuda(t, grad_out=0.25, backward="fused", grad_scale_hint=0.25),
                wrong=_run_cuda(t, grad_out=0.25, backward="fused", grad_scale_hint=-3.0),
                recompute=_run_cuda(t, grad_out=0.25, backward="recompute", grad_scale_hint=0.25),
                chunked=_run_cuda(t, grad_out=0.25, backward="fused", grad_scale_hint=0.25, chunk_rows=256, dense=True))
    for name, c in runs.items():
        assert c["loss"] == runs["none"]["loss"], name
        for k in ("dH_pair", "dH_ref", "dW"):
            assert rel(c[k], o[k]) <= TOL_GRAD_BF16, (name, k, rel(c[k], o[k]))
            assert rel(c[k], runs["none"][k]) <= 6e-3, (name, k)
    # the right hint IS the reference's order of operations (the upstream gradient enters before dLogits are rounded to bf16)
    assert np.array_equal(runs["right"]["dH_pair"], runs["recompute"]["dH_pair"]) or \
        rel(runs["right"]["dH_pair"], runs["recompute"]["dH_pair"]) <= 2e-3
    # fp32 inputs take the hint too
    from halva_b200 import dpa_loss
    t32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in t.items()}
    out = {}
    for name, kw in (("none", {}), ("right", dict(grad_scale_hint=0.25)), ("wrong", dict(grad_scale_hint=2.0))):
        hp = t32["h_pair"].clone().requires_grad_(True)
        W = t32["W"].clone().requires_grad_(True)
        loss, _, _ = dpa_loss(hp, t32["labels_pair"], t32["signs_pair"], t32["h_ref_pol"], t32["labels_ref"], W,
                              h_ref_frozen=t32["h_ref_frz"], alpha=t32["alpha"], **kw)
        (loss * 0.25).backward()
        out[name] = (hp.grad.cpu().numpy(), W.grad.cpu().numpy())
    for name in ("right", "wrong"):
        assert rel(out[name][0], out["none"][0]) <= 1e-5 and rel(out[name][1], out["none"][1]) <= 1e-5, name
    with pytest.raises(ValueError):
        _run_cuda(t, grad_scale_hint=0.0)


def test_quirks_match_reference():
    """K == 0 -> NaN (as the reference); K_pos != K_neg -> NaN + status bit (the reference raises RuntimeError)."""
    from halva_b200 import dpa_loss, read_status
    t, _ = _synth()
    z = torch.zeros_like(t["signs_pair"])
    loss, _, _ = dpa_loss(t["h_pair"], t["labels_pair"], z, t["h_ref_pol"], t["labels_ref"], t["W"],
                          h_ref_frozen=t["h_ref_frz"], alpha=0.4)
    assert torch.isnan(loss).item()
    s = t["signs_pair"].clone()
    B = s.shape[0] // 2
    s[B:][s[B:] > 0] = 1
    s[:B][s[:B] > 2] = 2
    tt = dict(t); tt["signs_pair"] = s
    hp = t["h_pair"].clone().requires_grad_(True)
    loss, _, _ = dpa_loss(hp, t["labels_pair"], s, t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
    st = read_status(loss)
    if st["K_pos"] != st["K_neg"]:
        assert torch.isnan(loss).item() and (st["status"] & 2)
        # strict=True: one host sync after the forward, then the reference's own error class (with and without grad)
        with pytest.raises(RuntimeError, match="must match the size"):
            dpa_loss(hp, t["labels_pair"], s, t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4, strict=True)
        with torch.no_grad(), pytest.raises(RuntimeError, match="must match the size"):
            dpa_loss(t["h_pair"], t["labels_pair"], s, t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4, strict=True)
    bad = t["signs_pair"].clone()
    bad[0, -3] = 300
    with pytest.raises(ValueError, match="phrase id"):
        dpa_loss(hp, t["labels_pair"], bad, t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4, strict=True)
    ok, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4, strict=True)
    assert torch.isfinite(ok).item()


def test_rejects_cpu_and_fp32_inputs():
    from halva_b200 import DPAError, dpa_loss
    t, _ = _synth()
    with pytest.raises(DPAError):
        dpa_loss(t["h_pair"].cpu(), t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"],
                 h_ref_frozen=t["h_ref_frz"], alpha=0.4)
    with pytest.raises(TypeError):
        dpa_loss(t["h_pair"].float(), t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"],
                 h_ref_frozen=t["h_ref_frz"], alpha=0.4)


def test_cuda_graph_capture_and_replay():
    """The whole forward+backward is stream-ordered (no host sync, row counts live on the device), so it captures into
    a CUDA graph; replays reproduce the eager result bit for bit, also after the inputs change in place."""
    from halva_b200 import dpa_loss
    t, _ = _synth(seed=21)
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)

    def step():
        hp.grad = hr.grad = W.grad = None
        loss, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"], alpha=t["alpha"])
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        lg = step()
    ghp, gW = hp.grad, W.grad
    for scale in (1.0, 0.5):
        with torch.no_grad():
            hp.mul_(scale)                     # new values in the captured buffers
        g.replay()
        torch.cuda.synchronize()
        got = (lg.item(), ghp.clone(), gW.clone())
        e = _run_cuda(dict(t, h_pair=hp.detach()))
        assert got[0] == e["loss"]
        assert np.array_equal(got[1].float().cpu().numpy(), e["dH_pair"]) and np.array_equal(got[2].float().cpu().numpy(), e["dW"])


@pytest.mark.parametrize("bwd", ["fused", "recompute"])
def test_ill_formed_batch_poisons_gradients(bwd):
    """K_pos != K_neg / phrase id > 255: the reference raises (halva_trainer.py:567); here loss AND every gradient are NaN
    (the device cannot raise) -- a truncated batch must never reach the optimiser as finite-but-wrong gradients."""
    from halva_b200 import dpa_loss, read_status
    t, _ = _synth()
    s = t["signs_pair"].clone()
    B = s.shape[0] // 2
    s[B:][s[B:] > 0] = 1
    s[:B][s[:B] > 2] = 2
    bad = t["signs_pair"].clone()
    bad[0, -3] = 300
    for signs, bit in ((s, 2), (bad, 1)):
        hp = t["h_pair"].clone().requires_grad_(True)
        hr = t["h_ref_pol"].clone().requires_grad_(True)
        W = t["W"].clone().requires_grad_(True)
        loss, _, _ = dpa_loss(hp, t["labels_pair"], signs, hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"], alpha=0.4, backward=bwd)
        st = read_status(loss)
        if not (st["status"] & bit):
            continue                                         # (this synthetic batch happened to have K_pos == K_neg)
        loss.backward()
        assert torch.isnan(loss).item()
        assert torch.isnan(W.grad).any().item() and torch.isnan(hr.grad).any().item()
        active = hp.grad[hp.grad != 0]
        assert active.numel() == 0 or torch.isnan(active).all().item()     # no finite non-zero gradient anywhere


def test_second_backward_and_inplace_edit_are_errors():
    from halva_b200 import dpa_loss
    t, _ = _synth()
    hp = t["h_pair"].clone().requires_grad_(True)
    loss, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second"):
        loss.backward()
    # recompute mode keeps only O(rows) state: a retained graph can be walked again, with the same result
    hp.grad = None
    loss, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4,
                          backward="recompute")
    g1, = torch.autograd.grad(loss, hp, retain_graph=True)
    g2, = torch.autograd.grad(loss, hp)
    assert torch.equal(g1, g2)
    # an in-place edit of an input between forward and backward is caught by autograd's version counters
    h2 = t["h_pair"].clone().requires_grad_(True)
    hh = h2 * 1.0
    loss, _, _ = dpa_loss(hh, t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4,
                          backward="recompute")
    hh.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        loss.backward()


def test_tensor_maps_are_cached_across_calls():
    import ctypes as C
    from halva_b200 import _lib, dpa_loss
    lib = _lib.load()
    t, _ = _synth()

    def counts():
        h, m = C.c_longlong(), C.c_longlong()
        lib.dpa_tmap_cache_stats(C.byref(h), C.byref(m))
        return h.value, m.value

    with torch.no_grad():
        for _ in range(2):                                   # the caching allocator hands the same scratch blocks back
            dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
        torch.cuda.synchronize()
        h0, m0 = counts()
        dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
        h1, m1 = counts()
    assert m1 == m0 and h1 > h0                              # steady state: nothing is re-encoded


def test_mixed_device_inputs_raise_before_any_kernel():
    from halva_b200 import DPAError, dpa_loss
    t, _ = _synth()
    with pytest.raises(DPAError, match="labels_pair"):
        dpa_loss(t["h_pair"], t["labels_pair"].cpu(), t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
    with pytest.raises(DPAError, match="labels_ref"):
        dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"].cpu(), t["W"], h_ref_frozen=t["h_ref_frz"], alpha=0.4)
