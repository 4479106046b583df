"""halva_b200.host / dpa_plan_host_spans (native, host-only): the span plan covers every row the loss can read (checked
against the oracle's active masks) and equals a plain-Python restatement of the planning rule."""
import numpy as np
import torch

from dpa_testutil import GOLDEN_CASES, load_golden
from halva_b200.host import contributing_rows, plan_spans
from halva_b200.synth import CONFIGS, make_labels_signs
from oracle import dpa_oracle


def _covered(plan, S, T):
    m = np.zeros((S, T), bool)
    for s, a, b in plan:
        m[s, a:b] = True
    return m


def _spans_py(mask, merge_gap):
    """Plain-Python statement of the planning rule: runs of contributing rows, merged when <= merge_gap rows apart."""
    out = []
    for s in range(mask.shape[0]):
        row = []
        for t in np.flatnonzero(mask[s]).tolist():
            if row and t - row[-1][1] <= merge_gap:
                row[-1] = (row[-1][0], t + 1)
            else:
                row.append((t, t + 1))
        out += [(s, a, b) for a, b in row]
    return out


def test_plan_covers_oracle_active_rows_on_golden():
    for name in GOLDEN_CASES:
        g = load_golden(name)
        o = dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], g["signs_pair"], g["h_ref_pol"], g["h_ref_frz"], g["labels_ref"],
                                g["W"], g["W_ref"], float(g["alpha"]))
        lp, sp, lr = (torch.from_numpy(g[k]) for k in ("labels_pair", "signs_pair", "labels_ref"))
        S, T = g["labels_pair"].shape
        cov = _covered(plan_spans(lp, sp, merge_gap=0), S, T)
        assert not (o["active_pair"] & ~cov[:, :-1]).any(), name            # every active row is transferred
        Sr, Tr = g["labels_ref"].shape
        covr = _covered(plan_spans(lr, None, merge_gap=0), Sr, Tr)
        assert np.array_equal(covr[:, :-1], o["active_ref"]) and not covr[:, -1].any(), name


def test_plan_matches_python_rule_and_merges():
    g = torch.Generator().manual_seed(0)
    shape = CONFIGS["cfg2"]
    labels, signs, labels_ref = make_labels_signs(shape, g)
    for gap in (0, 1, 8, 64):
        assert plan_spans(labels, signs, merge_gap=gap) == _spans_py(contributing_rows(labels, signs).numpy(), gap)
        assert plan_spans(labels_ref, None, merge_gap=gap) == _spans_py(contributing_rows(labels_ref).numpy(), gap)
    p32 = plan_spans(labels.int(), signs.int(), merge_gap=8)                # int32 labels: same plan
    assert p32 == plan_spans(labels, signs, merge_gap=8)
    rp = sum(b - a for _, a, b in p32)
    rr = sum(b - a for _, a, b in plan_spans(labels_ref, None))
    assert rp < 0.15 * labels.numel() and rr < 0.75 * labels_ref.numel()
    assert len(p32) <= 2 * shape.B * shape.k_max and len(plan_spans(labels_ref, None)) == shape.B_ref
    assert int(contributing_rows(labels, signs).sum()) <= rp               # merging only ever adds rows


def test_plan_edge_cases():
    T = 7
    none = torch.full((2, T), -100, dtype=torch.int64)
    assert plan_spans(none) == []
    every = torch.zeros((2, T), dtype=torch.int64)
    assert plan_spans(every, merge_gap=0) == [(0, 0, T - 1), (1, 0, T - 1)]   # the last position never predicts anything
    alt = torch.tensor([[-100, 1, -100, 1, -100, 1, -100]])
    assert plan_spans(alt, merge_gap=0) == [(0, 0, 1), (0, 2, 3), (0, 4, 5)]
    assert plan_spans(alt, merge_gap=1) == [(0, 0, 5)]
    sg = torch.tensor([[0, 0, 2, -100, 0, 3, 0]])
    assert plan_spans(torch.zeros_like(sg), sg, merge_gap=0) == [(0, 1, 2), (0, 4, 5)]
