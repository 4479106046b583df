"""halva_b200.host: the span plan covers every row the loss can read (checked against the oracle's active masks)."""
import numpy as np
import torch

from dpa_testutil import GOLDEN_CASES, load_golden
from halva_b200.host import HostStager, contributing_rows
from halva_b200.synth import CONFIGS, make_labels_signs
from oracle import dpa_oracle


def _covered(plan, S, T):
    m = np.zeros((S, T), bool)
    for s, a, b in plan:
        m[s, a:b] = True
    return m


def test_plan_covers_oracle_active_rows_on_golden():
    for name in GOLDEN_CASES:
        g = load_golden(name)
        o = dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], g["signs_pair"], g["h_ref_pol"], g["h_ref_frz"], g["labels_ref"],
                                g["W"], g["W_ref"], float(g["alpha"]))
        st = HostStager(torch.from_numpy(g["labels_pair"]), torch.from_numpy(g["signs_pair"]), torch.from_numpy(g["labels_ref"]), merge_gap=0)
        S, T = g["labels_pair"].shape
        cov = _covered(st.plan["h_pair"], S, T)
        assert not (o["active_pair"] & ~cov[:, :-1]).any(), name            # every active row is transferred
        Sr, Tr = g["labels_ref"].shape
        covr = _covered(st.plan["h_ref"], Sr, Tr)
        assert np.array_equal(covr[:, :-1], o["active_ref"]) and not covr[:, -1].any(), name


def test_plan_is_small_and_merges():
    g = torch.Generator().manual_seed(0)
    shape = CONFIGS["cfg2"]
    labels, signs, labels_ref = make_labels_signs(shape, g)
    st = HostStager(labels, signs, labels_ref)
    rp, rr = st.rows()
    assert rp < 0.15 * labels.numel() and rr < 0.75 * labels_ref.numel()
    assert len(st.plan["h_pair"]) <= 2 * shape.B * shape.k_max and len(st.plan["h_ref"]) == shape.B_ref
    m = contributing_rows(labels, signs)
    assert int(m.sum()) <= rp                                              # merging only ever adds rows
    dense = HostStager(labels, signs, labels_ref, dense=True)
    assert dense.plan["h_pair"] == []
