"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of HALVA's DPA-loss path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module; the product path (``halva_b200``) never does and fails
loudly when its CUDA library is missing.

Parity pin: the reference ships no tests / golden vectors for this path (SURVEY.md
section 8c).  The pin is the reference's own code, AST-lifted and executed in the build
container by ``oracle/gen_golden.py``; its inputs/outputs are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this restatement against
every one of them (and, when /root/reference is present, against a live run).

What is restated (all citations are /root/reference/...):
  * lm_head projection + fp32 upcast        vila/train/transformers_replace/models/llama/modeling_llama.py:1022-1029
  * shifted log-softmax gather              llava/train/halva_trainer.py:392-409   (cal_batch_logp)
  * phrase-level accumulation               llava/train/halva_trainer.py:411-419   (accumulate_logps)
  * pos/neg batching and the one-step shift llava/train/halva_trainer.py:421-489   (concatenated_forward)
  * reference-sample path                   llava/train/halva_trainer.py:491-531   (reference_forward)
  * loss assembly                           llava/train/halva_trainer.py:534-592   (compute_loss)
The backward is the closed form derived in SURVEY.md section 3.2 (autograd of the above).

All arithmetic is float64 unless a bf16 rounding point of the reference's bf16 execution
is being emulated (``mode="bf16"``):
  z  : nn.Linear emits bf16 logits, upcast afterwards            -> round_z
  dz : autograd hands a bf16 dLogits to the two backward GEMMs   -> round_dz
  dH, dW : bf16 GEMM outputs                                     -> round_out
"""
from __future__ import annotations

import numpy as np

IGNORE = -100            # llava/constants.py:7


class PhraseShapeError(RuntimeError):
    """K_pos != K_neg (the reference raises RuntimeError unless one side broadcasts) or K == 0."""


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float -> bfloat16 -> float64 (value-exact emulation)."""
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(np.float32).astype(np.float64)
    return np.where(np.isfinite(f), out, f.astype(np.float64))


def _lse(z):
    m = z.max(-1, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(-1, keepdims=True)))[..., 0]


def phrase_columns(signs_half: np.ndarray):
    """halva_trainer.py:412-417,560: -100 -> 0, U = sorted unique over the half batch,
    U[0] dropped unconditionally.  Returns (U[1:], col) with col[b,t] = column index or -1."""
    g = np.where(signs_half == IGNORE, 0, signs_half)
    U = np.unique(g)
    col = np.full(g.shape, -1, dtype=np.int64)
    for i, u in enumerate(U[1:]):
        col[g == u] = i
    return U[1:], col


def rmsnorm_fwd(x, w, eps, mode):
    """llava/model/language_model/modelling_llama.py:64-69: fp32 variance, x * rsqrt(var + eps) cast back to the input
    dtype, then weight * (.)  -- two bf16 rounding points in the reference's bf16 execution."""
    r = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + eps)
    n = x * r
    if mode == "bf16":
        n = bf16_round(n)
    y = w * n
    if mode == "bf16":
        y = bf16_round(y)
    return y, n, r


def rmsnorm_bwd(dy, x, w, n, r, mode):
    """autograd of rmsnorm_fwd: dn = dy * w (a bf16 tensor in bf16 mode), dx = r dn - x r^3 mean(dn x), dw = sum_rows dy n."""
    dn = dy * w
    if mode == "bf16":
        dn = bf16_round(dn)
    dx = r * dn - x * (r ** 3) * (dn * x).mean(-1, keepdims=True)
    dw = (dy * n).reshape(-1, x.shape[-1]).sum(0)
    if mode == "bf16":
        dx, dw = bf16_round(dx), bf16_round(dw)
    return dx, dw


def dpa_loss(h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frz, labels_ref, W, W_ref, alpha,
             mode="fp64", need_dW=True, grad_output=1.0, norm_w=None, norm_w_ref=None, norm_eps=1e-6, norm_mode=None,
             ref_logits=None):
    """Forward + closed-form backward of compute_loss.

    h_pair [2B,T,H]: first B sequences "pos"(correct), last B "neg"(hallucinated)  (:485-487)
    labels_pair/signs_pair [2B,T] int, post-splice, unshifted; h_ref_pol/h_ref_frz [Br,Tr,H];
    labels_ref [Br,Tr]; W/W_ref [V,H].
    mode: "fp64" exact maths on the given values; "bf16" additionally applies the reference's
    bf16 rounding points (inputs are expected to be bf16-representable already).
    ref_logits [Br,Tr,V]: the frozen model's logits given directly (what reference_forward(self.ref_model, ...) returns
    before its shift, :516-520) instead of h_ref_frz / W_ref, which may then be None.
    Returns dict(loss, align, div, P, N, dH_pair, dH_ref, dW, logp_pair, kl_rows).
    """
    rz = bf16_round if mode == "bf16" else (lambda a: a)
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    if ref_logits is not None:
        h_ref_frz, W_ref = np.zeros_like(h_ref_pol), np.zeros_like(W)       # unused
    h_pair, h_ref_pol, h_ref_frz, W, W_ref = map(f64, (h_pair, h_ref_pol, h_ref_frz, W, W_ref))
    fused_norm = norm_w is not None
    norm_mode = mode if norm_mode is None else norm_mode    # "bf16": the norm's own rounding points even when the
    if fused_norm:        # logits maths is exact (the GEMM operand is a bf16 tensor either way)
        # h_* are the pre-norm residual stream (final RMSNorm of the decoder folded into this path)
        norm_w = f64(norm_w)
        norm_w_ref = norm_w if norm_w_ref is None else f64(norm_w_ref)
        x_pair, x_ref_pol = h_pair, h_ref_pol
        h_pair, n_pair_, r_pair_ = rmsnorm_fwd(x_pair, norm_w, norm_eps, norm_mode)
        h_ref_pol, n_ref_, r_ref_ = rmsnorm_fwd(x_ref_pol, norm_w, norm_eps, norm_mode)
        h_ref_frz, _, _ = rmsnorm_fwd(h_ref_frz, norm_w_ref, norm_eps, norm_mode)
    labels_pair = np.asarray(labels_pair).astype(np.int64)
    signs_pair = np.asarray(signs_pair).astype(np.int64)
    labels_ref = np.asarray(labels_ref).astype(np.int64)
    B2, T, H = h_pair.shape
    B = B2 // 2
    Br, Tr, _ = h_ref_pol.shape
    V = W.shape[0]

    # ---- alignment half (:421-489, :550-568)
    z = rz(h_pair[:, :-1] @ W.T)                                   # last position dropped (:404)
    y = labels_pair[:, 1:]
    valid = (y != IGNORE)
    y0 = np.where(valid, y, 0)                                     # -100 -> index 0 (:406)
    lse = _lse(z)
    logp = (np.take_along_axis(z, y0[..., None], -1)[..., 0] - lse) * valid      # (:407, :556-557)
    g = signs_pair[:, 1:]
    Up, colp = phrase_columns(g[:B])
    Un, coln = phrase_columns(g[B:])
    Kp, Kn = len(Up), len(Un)
    if Kp != Kn or Kp == 0:
        raise PhraseShapeError(f"K_pos={Kp} K_neg={Kn}")
    K = Kp
    P = np.zeros((B, K)); N = np.zeros((B, K))
    for k in range(K):
        P[:, k] = (logp[:B] * (colp == k)).sum(-1)
        N[:, k] = (logp[B:] * (coln == k)).sum(-1)
    x = N - P
    with np.errstate(over="ignore"):
        e = np.exp(x)
        align = np.log(1.0 + e).mean()                             # literal log(1+exp) (:567)

    # ---- divergence half (:571-588)
    zp = rz(h_ref_pol[:, :-1] @ W.T)
    zr = rz(h_ref_frz[:, :-1] @ W_ref.T) if ref_logits is None else rz(f64(ref_logits)[:, :-1])
    vr = (labels_ref[:, 1:] != IGNORE)
    lp, lr = _lse(zp), _lse(zr)
    logq = zr - lr[..., None]
    logpp = zp - lp[..., None]
    q = np.exp(logq)
    kl_rows = (q * (logq - logpp)).sum(-1) * vr
    div = kl_rows.sum() / Br                                       # /batch, not /tokens (:587-588)
    loss = align + alpha * div

    # ---- closed-form backward
    go = float(grad_output)
    sig = (e / (1.0 + e)) / (B * K) * go                           # d align / d x
    col = np.concatenate([colp, coln], 0)
    active = (col >= 0) & valid
    coef = np.zeros((B2, T - 1))
    bi = np.arange(B)[:, None].repeat(T - 1, 1)
    cp = np.where(colp >= 0, sig[bi, np.maximum(colp, 0)], 0.0)
    cn = np.where(coln >= 0, sig[bi, np.maximum(coln, 0)], 0.0)
    coef[:B] = -cp                                                 # pos rows: dL/dlogp = -sigma
    coef[B:] = +cn                                                 # neg rows: dL/dlogp = +sigma
    coef = coef * active
    p = np.exp(z - lse[..., None])
    onehot = np.zeros_like(z)
    np.put_along_axis(onehot, y0[..., None], 1.0, -1)
    dz = coef[..., None] * (onehot - p)
    dzr = (go * alpha / Br) * vr[..., None] * (np.exp(logpp) - q)
    if mode == "bf16":
        dz, dzr = bf16_round(dz), bf16_round(dzr)
    dH_pair = np.zeros_like(h_pair)
    dH_pair[:, :-1] = dz @ W
    dH_ref = np.zeros_like(h_ref_pol)
    dH_ref[:, :-1] = dzr @ W
    dW = None
    if need_dW:
        dW = dz.reshape(-1, V).T @ h_pair[:, :-1].reshape(-1, H) + \
             dzr.reshape(-1, V).T @ h_ref_pol[:, :-1].reshape(-1, H)
    if mode == "bf16":
        dH_pair, dH_ref = bf16_round(dH_pair), bf16_round(dH_ref)
        dW = bf16_round(dW) if dW is not None else None
    d_norm_w = None
    if fused_norm:
        dH_pair, dw1 = rmsnorm_bwd(dH_pair, x_pair, norm_w, n_pair_, r_pair_, norm_mode)
        dH_ref, dw2 = rmsnorm_bwd(dH_ref, x_ref_pol, norm_w, n_ref_, r_ref_, norm_mode)
        d_norm_w = dw1 + dw2
    return dict(loss=loss, align=align, div=div, P=P, N=N, dH_pair=dH_pair, dH_ref=dH_ref, dW=dW, d_norm_w=d_norm_w,
                logp_pair=logp, kl_rows=kl_rows, lse_pair=lse, lse_ref_pol=lp, lse_ref_frz=lr,
                active_pair=active, active_ref=vr, col=col, coef_pair=coef)


def row_stats(h_rows, W, labels=None, mode="fp64"):
    """Per-row (lse, target logit) of z = h W^T for kernel-level unit tests."""
    rz = bf16_round if mode == "bf16" else (lambda a: a)
    z = rz(np.asarray(h_rows, np.float64) @ np.asarray(W, np.float64).T)
    lse = _lse(z)
    zy = None
    if labels is not None:
        zy = np.take_along_axis(z, np.asarray(labels, np.int64)[:, None], -1)[:, 0]
    return lse, zy
