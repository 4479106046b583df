"""`ref_logits`: the frozen reference model's logits given instead of its hidden states / lm_head.

The given logits are the bf16 lm_head output of the frozen model (halva_trainer.py:516-520 upcasts them with `.float()`),
so the op must agree with the oracle fed the same logits, and -- up to bf16 rounding flips of individual logits between
cuBLAS and the library's own GEMM -- with the op fed h_ref_frozen / W_ref."""
import numpy as np
import pytest
import torch

from dpa_testutil import load_golden, rel
from oracle import dpa_oracle
from test_gpu_dpa import TOL_GRAD_BF16, TOL_LOSS_BF16, _close, _run_cuda, _synth, _to_cuda

pytestmark = pytest.mark.gpu


def _frozen_logits(t, dtype):
    z = (t["h_ref_frz"].float() @ t["W_ref"].float().T).to(torch.bfloat16)       # what the frozen bf16 model emits
    return z.to(dtype)


def _run_given(t, zr, need_dW=True, grad_out=1.0, **kw):
    from halva_b200 import dpa_loss
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(need_dW)
    loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, ref_logits=zr, alpha=t["alpha"], **kw)
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    f = lambda x: None if x is None else x.float().cpu().numpy()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=f(hp.grad), dH_ref=f(hr.grad), dW=f(W.grad))


def _oracle_given(t, zr):
    f = lambda x: x.float().cpu().numpy()
    return dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]), None,
                               t["labels_ref"].cpu().numpy(), f(t["W"]), None, t["alpha"], mode="bf16", ref_logits=f(zr))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("name", ["tiny_bf16", "small_bf16", "small_fp32_fullft", "k1_alpha0_fp32"])
def test_golden_inputs_given_logits_vs_oracle(name, dtype):
    t = _to_cuda(load_golden(name))
    zr = _frozen_logits(t, dtype)
    c, o = _run_given(t, zr), _oracle_given(t, zr)
    for k in ("loss", "align"):
        assert _close(c[k], o[k], TOL_LOSS_BF16), (k, c[k], o[k])
    assert abs(c["div"] - o["div"]) <= TOL_LOSS_BF16 * abs(o["div"]) + 1e-7, (c["div"], o["div"])
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(c[k], o[k]) <= TOL_GRAD_BF16, (k, rel(c[k], o[k]))
    assert np.all(c["dH_ref"][:, :-1][~o["active_ref"]] == 0) and np.all(c["dH_ref"][:, -1] == 0)


@pytest.mark.parametrize("over", [dict(), dict(T=129, T_ref=77), dict(V=1003, H=128), dict(V=77, H=64, T=40, T_ref=33),
                                  dict(B=1, B_ref=3, k_max=1), dict(V=2560 + 8, H=192)])
def test_given_logits_match_hidden_state_form(over):
    """Same inputs through both forms of the frozen branch; ragged T, odd vocabulary sizes (padded row stride), several
    1280-column groups with a partial last one."""
    t, _ = _synth(**over)
    ref = _run_cuda(t, backward="fused")
    for dtype in (torch.bfloat16, torch.float32):
        c = _run_given(t, _frozen_logits(t, dtype))
        assert _close(c["loss"], ref["loss"], TOL_LOSS_BF16) and _close(c["align"], ref["align"], 1e-6)
        assert abs(c["div"] - ref["div"]) <= TOL_LOSS_BF16 * abs(ref["div"]) + 1e-7
        assert np.array_equal(c["dH_pair"], ref["dH_pair"])                     # the alignment half does not see the frozen branch
        for k in ("dH_ref", "dW"):
            assert rel(c[k], ref[k]) <= TOL_GRAD_BF16, (k, rel(c[k], ref[k]))


def test_given_logits_strided_view_chunks_and_eval():
    """A vocabulary-padded (strided) logits tensor is read in place; small row chunks; no_grad evaluation gives the same
    losses without any gradient buffer; grad_output scaling; lm_head frozen."""
    from halva_b200 import dpa_loss
    t, _ = _synth(T=96, T_ref=80)
    zr = _frozen_logits(t, torch.float32)
    Vg = zr.shape[-1]
    wide = torch.full((*zr.shape[:2], Vg + 24), 7.0, device="cuda")           # junk in the padding columns must not be read
    wide[..., :Vg] = zr
    base = _run_given(t, zr)
    for kw in (dict(), dict(chunk_rows=256), dict(dense=True)):
        c = _run_given(t, wide[..., :Vg], **kw)
        if "dense" in kw:          # other row order: the row sums regroup in fp32
            assert _close(c["loss"], base["loss"], 1e-6) and _close(c["div"], base["div"], 1e-5), kw
        else:
            assert c["loss"] == base["loss"] and c["div"] == base["div"], kw
        assert np.array_equal(c["dH_ref"], base["dH_ref"]), kw
    with torch.no_grad():
        loss, align, div = dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"],
                                    ref_logits=zr, alpha=t["alpha"])
    assert loss.item() == base["loss"] and align.item() == base["align"] and div.item() == base["div"]
    assert not loss.requires_grad
    c2 = _run_given(t, zr, need_dW=False, grad_out=0.5)
    assert c2["dW"] is None
    assert rel(c2["dH_ref"], 0.5 * base["dH_ref"]) <= 4e-3 and rel(c2["dH_pair"], 0.5 * base["dH_pair"]) <= 4e-3


def test_given_logits_argument_checks():
    from halva_b200 import dpa_loss
    t, _ = _synth()
    zr = _frozen_logits(t, torch.bfloat16)
    args = (t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], t["W"])
    with pytest.raises(ValueError):
        dpa_loss(*args, alpha=0.4)                                            # neither form of the frozen branch
    with pytest.raises(ValueError):
        dpa_loss(*args, h_ref_frozen=t["h_ref_frz"], ref_logits=zr, alpha=0.4)
    with pytest.raises(ValueError):
        dpa_loss(*args, ref_logits=zr, alpha=0.4, round_logits=False)
    with pytest.raises(ValueError):
        dpa_loss(*args, ref_logits=zr, alpha=0.4, backward="recompute")
    with pytest.raises(ValueError):
        dpa_loss(*args, ref_logits=zr[:, :-1], alpha=0.4)
    with pytest.raises(TypeError):
        dpa_loss(*args, ref_logits=zr.to(torch.float16), alpha=0.4)
    with pytest.raises(RuntimeError):
        dpa_loss(*args, ref_logits=zr.cpu(), alpha=0.4)
