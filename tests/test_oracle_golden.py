"""The oracle restatement (oracle/dpa_oracle.py) against the reference's own outputs.

tests/golden/*.npz were produced by oracle/gen_golden.py from the AST-lifted, unmodified
/root/reference/llava/train/halva_trainer.py:392-592 (forward + autograd backward).
"""
import numpy as np
import pytest

from oracle import dpa_oracle
from dpa_testutil import golden_path, norm_kwargs, rel


def _run(g, mode):
    return dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], g["signs_pair"], g["h_ref_pol"], g["h_ref_frz"],
                               g["labels_ref"], g["W"], g["W_ref"], float(g["alpha"]), mode=mode, **norm_kwargs(g))


def test_oracle_matches_reference_fp32_maths(golden):
    """fp64 restatement vs the reference run in fp32 on the same values: <= 1e-5 loss, <= 1e-4 grads."""
    o = _run(golden, "fp64")
    assert abs(o["loss"] - golden["ref32_loss"]) <= 1e-5 * abs(golden["ref32_loss"])
    assert abs(o["align"] - golden["ref32_align"]) <= 1e-5 * abs(golden["ref32_align"])
    assert rel(o["dH_pair"], golden["ref32_dH_pair"]) <= 1e-4
    assert rel(o["dH_ref"], golden["ref32_dH_ref"]) <= 1e-4
    assert rel(o["dW"], golden["ref32_dW"]) <= 1e-4
    if "norm_w" in golden:          # final RMSNorm folded in: gradients w.r.t. the pre-norm stream and the norm weight
        assert rel(o["d_norm_w"], golden["ref32_d_norm_w"]) <= 1e-4
    # structural zeros: last position and inactive rows get exactly no gradient in the reference
    assert np.all(golden["ref32_dH_pair"][:, -1] == 0) and np.all(o["dH_pair"][:, -1] == 0)
    inactive = ~o["active_pair"]
    assert np.all(o["dH_pair"][:, :-1][inactive] == 0)
    assert np.abs(golden["ref32_dH_pair"][:, :-1][inactive]).max() == 0


def test_oracle_bf16_mode_matches_reference_bf16(golden):
    """bf16 emulation (rounded logits / dLogits / outputs) vs the reference run in bf16 as the trainer does."""
    if str(golden["dtype"]) != "bf16":
        pytest.skip("fp32 case")
    o = _run(golden, "bf16")
    assert abs(o["loss"] - golden["ref_loss"]) <= 1e-3 * abs(golden["ref_loss"])
    assert abs(o["align"] - golden["ref_align"]) <= 1e-3 * abs(golden["ref_align"])
    assert rel(o["dH_pair"], golden["ref_dH_pair"]) <= 1e-2
    assert rel(o["dH_ref"], golden["ref_dH_ref"]) <= 1e-2
    assert rel(o["dW"], golden["ref_dW"]) <= 1e-2
    if "norm_w" in golden:
        assert rel(o["d_norm_w"], golden["ref_d_norm_w"]) <= 1e-2


def test_divergence_component(golden):
    o = _run(golden, "fp64")
    a = float(golden["alpha"])
    if a == 0:
        assert o["loss"] == o["align"]
        return
    ref_div = (golden["ref32_loss"] - golden["ref32_align"]) / a
    assert abs(o["div"] - ref_div) <= 2e-4 * abs(ref_div) + 1e-6     # fp32 cancellation in (loss - align)


def test_quirks_recorded_from_reference():
    q = np.load(golden_path("quirks"))
    assert bool(q["k0_is_nan"]) and bool(q["kmismatch_raises"])
    g = dict(np.load(golden_path("tiny_fp32")))
    with pytest.raises(dpa_oracle.PhraseShapeError):
        dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], np.zeros_like(g["signs_pair"]), g["h_ref_pol"],
                            g["h_ref_frz"], g["labels_ref"], g["W"], g["W_ref"], 0.4)


def test_bf16_round_is_rne():
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37.0
    want = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(dpa_oracle.bf16_round(x).astype(np.float32), want)


def test_live_reference_if_present():
    """When /root/reference is mounted (build container), re-run the lift and compare -- guards the fixtures."""
    from oracle.lift_reference import reference_available, run_reference
    if not reference_available():
        pytest.skip("no /root/reference here")
    import torch
    g = dict(np.load(golden_path("tiny_fp32")))
    t = lambda k: torch.from_numpy(g[k])
    r = run_reference(t("h_pair"), t("labels_pair"), t("signs_pair"), t("h_ref_pol"), t("h_ref_frz"),
                      t("labels_ref"), t("W"), t("W_ref"), alpha=float(g["alpha"]))
    assert abs(r["loss"].item() - g["ref_loss"]) < 1e-6
    assert rel(r["dW"].numpy(), g["ref_dW"]) < 1e-5


def test_oracle_given_frozen_logits_equals_hidden_state_form(golden):
    """`ref_logits` (the frozen model's logits handed over, what reference_forward(self.ref_model, ...) returns before its
    shift, halva_trainer.py:516-520) is the same computation as h_ref_frz / W_ref."""
    if "norm_w" in golden:
        pytest.skip("the frozen branch of the norm cases goes through the folded RMSNorm")
    g = golden
    for mode in ("fp64", "bf16"):
        a = _run(g, mode)
        zr = g["h_ref_frz"].astype(np.float64) @ g["W_ref"].astype(np.float64).T
        if mode == "bf16":
            zr = dpa_oracle.bf16_round(zr)
        b = dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], g["signs_pair"], g["h_ref_pol"], None, g["labels_ref"], g["W"], None,
                                float(g["alpha"]), mode=mode, ref_logits=zr)
        assert a["loss"] == b["loss"] and a["div"] == b["div"]
        assert np.array_equal(a["dH_ref"], b["dH_ref"]) and np.array_equal(a["dW"], b["dW"])


@pytest.mark.parametrize("i", range(10))
def test_live_reference_random_sweep(i):
    """Build container only (/root/reference present): the numpy oracle AND the torch port against the lifted, unmodified
    reference on random shapes -- B != B_ref, T != T_ref, 1..5 phrases per sample with gaps in the id range, phrase ids that
    may or may not include 0 positions, every-label-ignored sequences -- in fp32, forward and autograd backward.
    Pins the checker beyond the committed fixtures; the fixtures remain what travels to the GPU box."""
    from oracle.lift_reference import reference_available, run_reference
    if not reference_available():
        pytest.skip("no /root/reference here")
    import torch
    from oracle import dpa_port
    rng = np.random.default_rng(4200 + i)
    B, Br = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    T, Tr = int(rng.integers(6, 40)), int(rng.integers(4, 30))
    H, V = int(rng.choice([16, 48])), int(rng.choice([37, 96, 130]))
    K = int(rng.integers(1, 6))
    ids = np.sort(rng.choice(np.arange(1, 12), size=K, replace=False))           # sparse phrase ids, same set for every sequence
    labels = np.full((2 * B, T), -100, np.int64)
    signs = np.zeros((2 * B, T), np.int64)
    for s in range(2 * B):
        t0 = int(rng.integers(1, max(2, T // 3)))
        labels[s, t0:] = rng.integers(0, V, T - t0)
        signs[s, :t0] = -100 if rng.random() < 0.5 else 0
        pos = rng.permutation(np.arange(t0, T))[:max(K, (T - t0) // 2)]
        for n, p in enumerate(pos):
            signs[s, p] = ids[n % K]                                              # every id present in every half
        if rng.random() < 0.3:
            labels[s, T - 1] = -100
    labels_ref = np.full((Br, Tr), -100, np.int64)
    for s in range(Br):
        if rng.random() < 0.2:
            continue                                                              # a reference sample with no supervised token
        t0 = int(rng.integers(1, Tr - 1))
        labels_ref[s, t0:] = rng.integers(0, V, Tr - t0)
    f32 = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    W = 0.3 * f32(V, H)
    W_ref = W if i % 2 == 0 else W + 0.05 * f32(V, H)
    h_pair, h_ref_pol = f32(2 * B, T, H), f32(Br, Tr, H)
    h_ref_frz = h_ref_pol + 0.2 * f32(Br, Tr, H)
    alpha = float(rng.choice([0.0, 0.2, 0.4, 0.5]))
    lt, st_, lrt = torch.from_numpy(labels), torch.from_numpy(signs), torch.from_numpy(labels_ref)
    r = run_reference(h_pair, lt, st_, h_ref_pol, h_ref_frz, lrt, W, W_ref, alpha=alpha)
    n = lambda t: t.numpy().astype(np.float64)
    o = dpa_oracle.dpa_loss(n(h_pair), labels, signs, n(h_ref_pol), n(h_ref_frz), labels_ref, n(W), n(W_ref), alpha, mode="fp64")
    ref_loss = r["loss"].item()
    assert abs(o["loss"] - ref_loss) <= 1e-5 * abs(ref_loss), (o["loss"], ref_loss)
    for k in ("dH_pair", "dH_ref", "dW"):
        g = r[k].numpy()
        if np.linalg.norm(g) == 0:
            assert not o[k].any(), k
        else:
            assert rel(o[k], g) <= 1e-4, (k, rel(o[k], g))
    # the torch port (the CPU baseline of bench.py and the GPU op-chain comparator) issues the reference's op sequence
    hp, hr, Wp = h_pair.clone().requires_grad_(True), h_ref_pol.clone().requires_grad_(True), W.clone().requires_grad_(True)
    lp, _, _ = dpa_port.dpa_loss_torch(hp, lt, st_, hr, h_ref_frz, lrt, Wp, W_ref if W_ref is not W else Wp, alpha)
    lp.backward()
    assert abs(lp.item() - ref_loss) <= 1e-6 * abs(ref_loss) + 1e-7
    for k, g in (("dH_pair", hp.grad), ("dH_ref", hr.grad), ("dW", Wp.grad)):
        assert torch.allclose(g, r[k], rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("i", range(4))
def test_live_reference_random_sweep_bf16(i):
    """The same idea in the trainer's precision: the lifted reference executed in bf16 (bf16 lm_head output, fp32 softmax
    maths, bf16 autograd GEMMs) against the oracle's bf16 mode -- the north star's bf16 tolerances."""
    from oracle.lift_reference import reference_available, run_reference
    if not reference_available():
        pytest.skip("no /root/reference here")
    import torch
    rng = np.random.default_rng(7700 + i)
    B, Br, T, Tr, H, V, K = 2, int(rng.integers(1, 3)), int(rng.integers(10, 30)), int(rng.integers(8, 24)), 64, int(rng.choice([96, 200])), int(rng.integers(1, 4))
    labels = np.full((2 * B, T), -100, np.int64)
    signs = np.zeros((2 * B, T), np.int64)
    for s in range(2 * B):
        t0 = int(rng.integers(1, T // 3))
        labels[s, t0:] = rng.integers(0, V, T - t0)
        for n, p in enumerate(rng.permutation(np.arange(t0, T))[:max(K, (T - t0) // 2)]):
            signs[s, p] = 1 + n % K
    labels_ref = np.full((Br, Tr), -100, np.int64)
    for s in range(Br):
        t0 = int(rng.integers(1, Tr - 1))
        labels_ref[s, t0:] = rng.integers(0, V, Tr - t0)
    bf = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(torch.bfloat16)
    W = (0.2 * bf(V, H).float()).to(torch.bfloat16)
    h_pair, h_ref_pol = bf(2 * B, T, H), bf(Br, Tr, H)
    h_ref_frz = (h_ref_pol.float() + 0.2 * bf(Br, Tr, H).float()).to(torch.bfloat16)
    lt, st_, lrt = torch.from_numpy(labels), torch.from_numpy(signs), torch.from_numpy(labels_ref)
    r = run_reference(h_pair, lt, st_, h_ref_pol, h_ref_frz, lrt, W, W, alpha=0.4)
    n = lambda t: t.float().numpy().astype(np.float64)
    o = dpa_oracle.dpa_loss(n(h_pair), labels, signs, n(h_ref_pol), n(h_ref_frz), labels_ref, n(W), n(W), 0.4, mode="bf16")
    assert abs(o["loss"] - r["loss"].item()) <= 1e-3 * abs(r["loss"].item())
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(o[k], n(r[k])) <= 1e-2, (k, rel(o[k], n(r[k])))
