import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_CASES = ["tiny_fp32", "tiny_bf16", "small_fp32_fullft", "small_bf16", "k1_alpha0_fp32",
                "sparse_ids_fp32", "no_zero_sign_fp32", "norm_fp32", "norm_bf16"]


def norm_kwargs(g):
    """Final-RMSNorm operands of a golden case (empty for the cases fed post-norm hidden states)."""
    if "norm_w" not in g:
        return {}
    return dict(norm_w=g["norm_w"], norm_w_ref=g["norm_w_ref"], norm_eps=float(g["norm_eps"]))


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name + ".npz")


def load_golden(name):
    d = dict(np.load(golden_path(name)))
    d["name"] = name
    return d


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
