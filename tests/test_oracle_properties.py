"""Size-independent properties of the oracle restatement (CPU): the same invariants the full-size GPU tests check on the
CUDA path (tests/test_gpu_fullsize.py), here on the checker itself, plus a finite-difference check of its closed-form
backward -- the oracle's gradients are otherwise only pinned at the golden vectors."""
import numpy as np
import pytest

from dpa_testutil import load_golden, rel
from oracle import dpa_oracle

CASES = ["tiny_fp32", "small_fp32_fullft", "sparse_ids_fp32"]


def _args(g):
    return [g["h_pair"].astype(np.float64), g["labels_pair"], g["signs_pair"], g["h_ref_pol"].astype(np.float64),
            g["h_ref_frz"].astype(np.float64), g["labels_ref"], g["W"].astype(np.float64), g["W_ref"].astype(np.float64), float(g["alpha"])]


@pytest.mark.parametrize("name", CASES)
def test_gradients_scale_linearly_with_grad_output(name):
    a = _args(load_golden(name))
    o1, o2 = dpa_oracle.dpa_loss(*a), dpa_oracle.dpa_loss(*a, grad_output=-2.5)
    assert o1["loss"] == o2["loss"]
    for k in ("dH_pair", "dH_ref", "dW"):
        assert rel(o2[k], -2.5 * o1[k]) <= 1e-14


@pytest.mark.parametrize("name", CASES)
def test_pair_permutation(name):
    """The alignment loss is a mean over (sample, phrase): permuting the sample pairs (pos and neg halves alike) permutes
    the dH rows and leaves the loss, the phrase sums' multiset and dW alone."""
    a = _args(load_golden(name))
    B = a[0].shape[0] // 2
    perm = np.random.default_rng(0).permutation(B)
    idx = np.concatenate([perm, B + perm])
    b = list(a)
    b[0], b[1], b[2] = a[0][idx], a[1][idx], a[2][idx]
    o, p = dpa_oracle.dpa_loss(*a), dpa_oracle.dpa_loss(*b)
    assert abs(o["loss"] - p["loss"]) <= 1e-13 * abs(o["loss"])
    assert rel(p["dH_pair"], o["dH_pair"][idx]) <= 1e-13 and rel(p["dW"], o["dW"]) <= 1e-12
    assert np.allclose(p["P"], o["P"][perm]) and np.allclose(p["N"], o["N"][perm])


@pytest.mark.parametrize("name", CASES)
def test_phrase_ids_enter_only_through_their_rank(name):
    """Columns follow the sorted-unique RANK of the ids (halva_trainer.py:412-417): any strictly increasing relabelling of
    the non-zero ids gives the same result."""
    a = _args(load_golden(name))
    s = a[2].copy()
    pos = s > 0
    s[pos] = 3 * s[pos] + 1
    b = list(a)
    b[2] = s
    o, p = dpa_oracle.dpa_loss(*a), dpa_oracle.dpa_loss(*b)
    assert o["loss"] == p["loss"] and np.array_equal(o["dW"], p["dW"]) and np.array_equal(o["dH_pair"], p["dH_pair"])


def test_ignored_positions_do_not_matter():
    """Hidden rows that predict an ignored label (or no phrase) may hold anything: loss and gradients do not change and
    their own gradient is exactly zero -- the property the row compaction of the CUDA path rests on."""
    a = _args(load_golden("tiny_fp32"))
    o = dpa_oracle.dpa_loss(*a)
    b = list(a)
    hp, hr = a[0].copy(), a[3].copy()
    hp[:, :-1][~o["active_pair"]] = 123.0
    hp[:, -1] = -7.0
    hr[:, :-1][~o["active_ref"]] = -55.0
    b[0], b[3] = hp, hr
    hf = a[4].copy()
    hf[:, :-1][~o["active_ref"]] = 9.0
    b[4] = hf
    p = dpa_oracle.dpa_loss(*b)
    assert p["loss"] == o["loss"] and np.array_equal(p["dW"], o["dW"])
    assert np.array_equal(p["dH_pair"], o["dH_pair"]) and np.array_equal(p["dH_ref"], o["dH_ref"])
    assert not p["dH_pair"][:, :-1][~o["active_pair"]].any() and not p["dH_pair"][:, -1].any()


def test_closed_form_backward_against_finite_differences():
    g = load_golden("tiny_fp32")
    a = _args(g)
    o = dpa_oracle.dpa_loss(*a)
    rng = np.random.default_rng(1)
    eps = 1e-6
    for which, key in ((0, "dH_pair"), (3, "dH_ref"), (6, "dW")):
        d = rng.standard_normal(a[which].shape)
        hi, lo = list(a), list(a)
        hi[which] = a[which] + eps * d
        lo[which] = a[which] - eps * d
        # (W_ref is the frozen model's copy, a separate argument: it does not move with W)
        num = (dpa_oracle.dpa_loss(*hi)["loss"] - dpa_oracle.dpa_loss(*lo)["loss"]) / (2 * eps)
        ana = float((o[key] * d).sum())
        assert abs(num - ana) <= 1e-6 * max(abs(ana), 1e-3), (key, num, ana)
