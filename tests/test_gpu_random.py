"""Randomised parity sweep: many small shapes (tile / chunk / vocabulary-group boundaries, dense mode, both gradient
strategies, int32 ids) of the CUDA path against the numpy oracle in the reference's bf16 numerics."""
import dataclasses

import numpy as np
import pytest
import torch

from dpa_testutil import rel
from oracle import dpa_oracle

pytestmark = pytest.mark.gpu


def _case(i):
    rng = np.random.default_rng(1000 + i)
    H = int(rng.choice([64, 128, 192, 320, 512]))
    V = int(rng.choice([8, 77, 256, 1000, 1280, 1288, 2560, 2600, 3001]))
    T = int(rng.choice([8, 33, 64, 129, 200, 257]))
    T_ref = int(rng.choice([8, 40, 128, 255]))
    B = int(rng.integers(1, 5))
    B_ref = int(rng.integers(1, 4))
    k_max = int(rng.integers(1, 6))
    density = float(rng.choice([0.05, 0.15, 0.3, 0.5]))
    flags = dict(dense=bool(rng.integers(0, 2)), chunk_rows=int(rng.choice([0, 256, 512])),
                 backward=str(rng.choice(["fused", "recompute"])), i32=bool(rng.integers(0, 2)), shared=bool(rng.integers(0, 2)))
    return dict(H=H, V=V, T=T, T_ref=T_ref, B=B, B_ref=B_ref, k_max=k_max, density=density), flags


@pytest.mark.parametrize("i", range(36))
def test_random_shape_vs_oracle(i):
    from halva_b200 import dpa_loss, read_status
    from halva_b200.synth import CONFIGS, make_inputs
    over, fl = _case(i)
    shape = dataclasses.replace(CONFIGS["tiny"], **over)
    t = make_inputs(shape, seed=500 + i, device="cuda", shared_head=fl["shared"], plant=3.0)
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)
    ids = {k: (t[k].to(torch.int32) if fl["i32"] else t[k]) for k in ("labels_pair", "signs_pair", "labels_ref")}
    loss, align, div = dpa_loss(hp, ids["labels_pair"], ids["signs_pair"], hr, ids["labels_ref"], W, h_ref_frozen=t["h_ref_frz"],
                                W_ref=None if fl["shared"] else t["W_ref"], alpha=t["alpha"], dense=fl["dense"],
                                chunk_rows=fl["chunk_rows"], backward=fl["backward"])
    loss.backward()
    st = read_status(loss)
    f = lambda x: x.detach().float().cpu().numpy()
    try:
        o = dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]),
                                f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]), f(t["W_ref"]), t["alpha"], mode="bf16")
    except dpa_oracle.PhraseShapeError:
        assert st["status"] != 0 and torch.isnan(loss).item(), (over, fl, st)     # the reference raises / returns NaN here
        return
    assert st["status"] == 0, (over, fl, st)
    if not np.isfinite(o["loss"]):
        assert not np.isfinite(loss.item())
        return
    tol = 1e-3
    assert abs(loss.item() - o["loss"]) <= tol * abs(o["loss"]), (over, fl, loss.item(), o["loss"])
    assert abs(align.item() - o["align"]) <= tol * abs(o["align"]), (over, fl)
    assert abs(div.item() - o["div"]) <= tol * abs(o["div"]) + 1e-7, (over, fl, div.item(), o["div"])
    for k, g in (("dH_pair", hp.grad), ("dH_ref", hr.grad), ("dW", W.grad)):
        if np.linalg.norm(o[k]) == 0:
            assert not f(g).any(), (over, fl, k)
        else:
            assert rel(f(g), o[k]) <= 1e-2, (over, fl, k, rel(f(g), o[k]))


@pytest.mark.parametrize("i", range(24))
def test_random_shape_fp32_and_given_logits(i):
    """The same kind of sweep for the two other input forms: fp32 hidden states / weights (split-bf16 path, against the fp64
    oracle at the fp32 tolerances) and the frozen model's logits given (against the oracle fed the same logits)."""
    from halva_b200 import dpa_loss, read_status
    from halva_b200.synth import CONFIGS, make_inputs
    over, fl = _case(100 + i)
    shape = dataclasses.replace(CONFIGS["tiny"], **over)
    fp32 = i % 2 == 0
    t = make_inputs(shape, seed=900 + i, device="cuda", shared_head=fl["shared"], plant=3.0,
                    dtype=torch.float32 if fp32 else torch.bfloat16)
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)
    ids = {k: (t[k].to(torch.int32) if fl["i32"] else t[k]) for k in ("labels_pair", "signs_pair", "labels_ref")}
    f = lambda x: x.detach().double().cpu().numpy()
    common = dict(alpha=t["alpha"], dense=fl["dense"], chunk_rows=fl["chunk_rows"])
    okw = {}
    if fp32:
        frozen = dict(h_ref_frozen=t["h_ref_frz"], W_ref=None if fl["shared"] else t["W_ref"])
        mode, tol_loss, tol_div, tol_grad = "fp64", 1e-5, 1e-4, 1e-4
    else:
        zr = (t["h_ref_frz"].float() @ t["W_ref"].float().T).to(torch.bfloat16)
        if i % 4 == 1:
            zr = zr.float()
        frozen = dict(ref_logits=zr)
        okw = dict(ref_logits=f(zr))
        mode, tol_loss, tol_div, tol_grad = "bf16", 1e-3, 1e-3, 1e-2
    loss, align, div = dpa_loss(hp, ids["labels_pair"], ids["signs_pair"], hr, ids["labels_ref"], W, **frozen, **common)
    loss.backward()
    st = read_status(loss)
    try:
        o = dpa_oracle.dpa_loss(f(t["h_pair"]), t["labels_pair"].cpu().numpy(), t["signs_pair"].cpu().numpy(), f(t["h_ref_pol"]),
                                None if okw else f(t["h_ref_frz"]), t["labels_ref"].cpu().numpy(), f(t["W"]),
                                None if okw else f(t["W_ref"]), t["alpha"], mode=mode, **okw)
    except dpa_oracle.PhraseShapeError:
        assert st["status"] != 0 and torch.isnan(loss).item(), (over, fl, st)
        return
    assert st["status"] == 0, (over, fl, st)
    if not np.isfinite(o["loss"]):
        assert not np.isfinite(loss.item())
        return
    assert abs(loss.item() - o["loss"]) <= tol_loss * abs(o["loss"]), (fp32, over, fl, loss.item(), o["loss"])
    assert abs(align.item() - o["align"]) <= tol_loss * abs(o["align"]), (fp32, over, fl)
    assert abs(div.item() - o["div"]) <= tol_div * abs(o["div"]) + 5e-7, (fp32, over, fl, div.item(), o["div"])
    for k, g in (("dH_pair", hp.grad), ("dH_ref", hr.grad), ("dW", W.grad)):
        assert g.dtype == t["W"].dtype
        if np.linalg.norm(o[k]) == 0:
            assert not f(g).any(), (fp32, over, fl, k)
        else:
            assert rel(f(g), o[k]) <= tol_grad, (fp32, over, fl, k, rel(f(g), o[k]))
