"""The torch-CPU port used for the CPU baseline agrees with the golden reference outputs and the numpy oracle."""
import numpy as np
import torch

from dpa_testutil import load_golden, rel
from oracle import dpa_oracle, dpa_port


def test_port_matches_golden(golden):
    g = golden
    if "norm_w" in g:
        import pytest
        pytest.skip("the port starts at the post-norm hidden states")
    dt = torch.bfloat16 if str(g["dtype"]) == "bf16" else torch.float32
    fl = lambda k: torch.from_numpy(g[k]).to(dt)
    ix = lambda k: torch.from_numpy(g[k])
    inp = dict(h_pair=fl("h_pair"), labels_pair=ix("labels_pair"), signs_pair=ix("signs_pair"), h_ref_pol=fl("h_ref_pol"),
               h_ref_frz=fl("h_ref_frz"), labels_ref=ix("labels_ref"), W=fl("W"), W_ref=fl("W_ref"), alpha=float(g["alpha"]))
    r = dpa_port.step_cpu(inp)
    tol_l, tol_g = (2e-3, 1e-2) if dt == torch.bfloat16 else (1e-5, 1e-4)
    assert abs(r["loss"] - float(g["ref_loss"])) <= tol_l * abs(float(g["ref_loss"]))
    assert rel(r["dW"].float().numpy(), g["ref_dW"]) <= tol_g
    assert rel(r["dH_pair"].float().numpy(), g["ref_dH_pair"]) <= tol_g


def test_port_matches_numpy_oracle_on_synth():
    from halva_b200.synth import CONFIGS, make_inputs
    inp = make_inputs(CONFIGS["tiny"], seed=2, dtype=torch.float32)
    r = dpa_port.step_cpu(inp)
    f = lambda k: inp[k].numpy()
    o = dpa_oracle.dpa_loss(f("h_pair"), f("labels_pair"), f("signs_pair"), f("h_ref_pol"), f("h_ref_frz"), f("labels_ref"),
                            f("W"), f("W_ref"), inp["alpha"])
    assert abs(r["loss"] - o["loss"]) <= 1e-5 * abs(o["loss"])
    assert rel(r["dW"].numpy(), o["dW"]) <= 1e-4 and rel(r["dH_ref"].numpy(), o["dH_ref"]) <= 1e-4
