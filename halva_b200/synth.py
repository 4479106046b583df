"""Seeded synthetic inputs for the DPA-loss path, shaped like the HALVA training batches.

Layout of one sequence (post image-splice, as returned by
prepare_inputs_labels_for_multimodal_signed, /root/reference/llava/model/llava_arch.py:229-394):
  [0, n_sys)            system prompt     label -100, sign 0
  [n_sys, +n_img)       image tokens      label -100, sign -100   (llava_arch.py:322-323)
  question              ~n_q tokens       label -100, sign 0
  response .. T_valid   labels ~ U[0,V)   sign 0 or the 1-based phrase id inside <MASK> spans
                                          (llava/train/train_halva.py:263-335)
  tail padding          label -100, sign -100 (llava_arch.py:346-347)
Correct ("pos") and hallucinated ("neg") responses of a pair carry the same number of phrase
spans; span lengths differ by a token or two.  The reference sample is an ordinary SFT sample.
"""
from __future__ import annotations

import dataclasses
import math

import torch

IGNORE = -100


@dataclasses.dataclass
class DPAShape:
    B: int = 16            # sample pairs per call
    T: int = 2048          # sequence length (post-splice)
    H: int = 4096
    V: int = 32000
    B_ref: int = 0         # reference samples per call (0 -> B)
    T_ref: int = 0         # (0 -> T)
    n_sys: int = 35
    n_img: int = 576
    n_q: int = 30
    density: float = 0.10  # fraction of ALL pair rows that sit inside a phrase span
    k_max: int = 8
    alpha: float = 0.4

    def __post_init__(self):
        self.B_ref = self.B_ref or self.B
        self.T_ref = self.T_ref or self.T

    @property
    def tokens(self) -> int:
        """Sequence positions fed to the op per call (2B*T + B_ref*T_ref)."""
        return 2 * self.B * self.T + self.B_ref * self.T_ref


CONFIGS = {
    # BASELINE.json configs[0..3]; cfg-5 is a sweep over (B, density) on top of cfg2
    "cfg1": DPAShape(B=1, T=1024, H=4096, V=32000, alpha=0.4),
    "cfg2": DPAShape(B=16, T=2048, H=4096, V=32000, alpha=0.4),
    "cfg3": DPAShape(B=16, T=2048, H=5120, V=32000, alpha=0.5),
    "cfg4": DPAShape(B=16, T=4096, H=5120, V=32000, alpha=0.2, n_img=196),
    # one rank's GEMM shapes of cfg3 under 8-way vocabulary sharding (V / 8 columns), without the exchange: kernel probes
    "cfg3_shard8": DPAShape(B=16, T=2048, H=5120, V=4000, alpha=0.5),
    "tiny": DPAShape(B=2, T=48, H=64, V=512, n_sys=3, n_img=8, n_q=3, k_max=3, density=0.2),
    "small": DPAShape(B=2, T=160, H=256, V=1000, n_sys=5, n_img=24, n_q=6, k_max=4, density=0.15),
}


def _layout(shape: DPAShape, T: int, g: torch.Generator):
    prefix = shape.n_sys + shape.n_img + shape.n_q
    prefix = min(prefix, max(1, T // 2))
    lo = max(prefix + 4, int(0.6 * T))
    t_valid = int(torch.randint(lo, T + 1, (1,), generator=g))
    return prefix, t_valid


def make_labels_signs(shape: DPAShape, g: torch.Generator):
    """labels/signs for the 2B pair sequences and labels for the B_ref reference samples."""
    B, T, V = shape.B, shape.T, shape.V
    labels = torch.full((2 * B, T), IGNORE, dtype=torch.int64)
    signs = torch.zeros((2 * B, T), dtype=torch.int64)
    n_sys = min(shape.n_sys, T // 8)
    n_img = min(shape.n_img, T // 3)
    for b in range(B):
        K_b = int(torch.randint(1, shape.k_max + 1, (1,), generator=g))
        for half in (0, 1):
            s = b + half * B
            prefix, t_valid = _layout(shape, T, g)
            labels[s, prefix:t_valid] = torch.randint(0, V, (t_valid - prefix,), generator=g)
            signs[s, n_sys:n_sys + n_img] = IGNORE
            signs[s, t_valid:] = IGNORE
            resp = t_valid - prefix
            # span length so that (sum of spans)/T ~= density, jittered per half by -1..+2 tokens
            want = max(K_b, int(round(shape.density * T)))
            want = min(want, resp - K_b - 1)
            base = max(1, want // K_b)
            gap = max(1, (resp - base * K_b) // (K_b + 1))
            pos = prefix + gap
            for k in range(1, K_b + 1):
                ln = max(1, base + int(torch.randint(-1, 3, (1,), generator=g)))
                end = min(pos + ln, t_valid)
                if pos < end:
                    signs[s, pos:end] = k
                pos = end + gap
    labels_ref = torch.full((shape.B_ref, shape.T_ref), IGNORE, dtype=torch.int64)
    for b in range(shape.B_ref):
        prefix, t_valid = _layout(shape, shape.T_ref, g)
        labels_ref[b, prefix:t_valid] = torch.randint(0, V, (t_valid - prefix,), generator=g)
    return labels, signs, labels_ref


def make_inputs(shape: DPAShape, seed: int = 1234, dtype=torch.bfloat16, device="cpu",
                shared_head: bool = True, hidden_scale: float = 1.5, plant: float = 0.0):
    """W ~ N(0, 0.02^2); h ~ 1.5 N(0,1); h_ref_frozen = h_ref_pol + 0.1 N(0,1);
    W_ref = W (LoRA recipe, lm_head shared and frozen) or W + 0.002 N(0,1) (full fine-tune).
    plant > 0 adds `plant * W[label]` to the hidden row that predicts a valid label, which lifts the target logit by
    about plant * |W_y|^2 -- a trained LM gives its targets high probability, purely random hidden states give every
    target ~1/V, and then long phrases make the reference's literal log(1+exp(N-P)) overflow to +inf."""
    g = torch.Generator().manual_seed(seed)
    labels, signs, labels_ref = make_labels_signs(shape, g)
    dev = torch.device(device)
    gd = torch.Generator(device=dev).manual_seed(seed + 7)

    def randn(*s):
        return torch.randn(*s, generator=gd, device=dev, dtype=torch.float32)

    W = (0.02 * randn(shape.V, shape.H)).to(dtype)
    h_pair = (hidden_scale * randn(2 * shape.B, shape.T, shape.H)).to(dtype)
    h_ref_pol32 = hidden_scale * randn(shape.B_ref, shape.T_ref, shape.H)
    h_ref_frz = (h_ref_pol32 + 0.1 * randn(shape.B_ref, shape.T_ref, shape.H)).to(dtype)
    h_ref_pol = h_ref_pol32.to(dtype)
    del h_ref_pol32
    if plant > 0:
        def planted(h, labels, noise=0.0):
            y = labels[:, 1:].to(dev)
            ok = (y != IGNORE).unsqueeze(-1)
            add = plant * W.float()[y.clamp_min(0)] * ok
            h32 = h.float()
            h32[:, :-1] += add
            return h32.to(dtype)
        h_pair = planted(h_pair, labels)
        h_ref_pol = planted(h_ref_pol, labels_ref)
        h_ref_frz = planted(h_ref_frz, labels_ref)
    W_ref = W if shared_head else (W.float() + 0.002 * randn(shape.V, shape.H)).to(dtype)
    return dict(h_pair=h_pair, labels_pair=labels.to(dev), signs_pair=signs.to(dev),
                h_ref_pol=h_ref_pol, h_ref_frz=h_ref_frz, labels_ref=labels_ref.to(dev),
                W=W, W_ref=W_ref, alpha=shape.alpha)


def active_rows(labels, signs, labels_ref):
    """(#pair rows with a non-zero contribution, #reference-sample rows with one)."""
    y, g = labels[:, 1:], signs[:, 1:]
    pair = ((y != IGNORE) & (g != IGNORE) & (g != 0)).sum().item()
    ref = (labels_ref[:, 1:] != IGNORE).sum().item()
    return int(pair), int(ref)


def algorithmic_flops(shape: DPAShape, need_dW: bool, rows_pair: int | None = None, rows_ref: int | None = None):
    """F_alg = 2HV [ A_pol (1 + 1 + w) + A_frz ]; recompute not counted.
    rows_* None -> dense convention (every fed position)."""
    a_pair = 2 * shape.B * shape.T if rows_pair is None else rows_pair
    a_ref = shape.B_ref * shape.T_ref if rows_ref is None else rows_ref
    unit = 2.0 * shape.H * shape.V
    return unit * ((a_pair + a_ref) * (2 + (1 if need_dW else 0)) + a_ref)
