"""TEST / BASELINE INFRASTRUCTURE ONLY -- torch-CPU port of the reference DPA-loss op chain.

Used as (a) the `cpu_baseline` / `--impl reference` arm of bench.py ("kind": "port": /root/reference does
not exist on the GPU box, so the reference's own file cannot be executed there) and (b) a second,
autograd-based check of oracle/dpa_oracle.py.  It issues the same ATen op sequence as the reference
(nn.Linear -> .float() -> log_softmax -> gather; unique + per-phrase masked sums; log(1+exp); two
softmax + log for the KL), so its CPU cost is representative of the reference's CPU path
(/root/reference/llava/train/halva_trainer.py:392-592).  Never imported by halva_b200.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IGNORE = -100


def _token_logps(hidden, weight, labels):
    """lm_head + fp32 upcast (modeling_llama.py:1028-1029) then shifted gather (halva_trainer.py:403-407)."""
    logits = F.linear(hidden, weight).float()
    tgt = labels[:, 1:].clone()
    logits = logits[:, :-1, :]
    tgt[tgt == IGNORE] = 0
    lp = torch.gather(logits.log_softmax(-1), 2, tgt.unsqueeze(2)).squeeze(2)
    return lp, logits


def _phrase_sums(logps, signs):
    """halva_trainer.py:411-419: columns follow the sorted unique ids of the half batch, first id dropped."""
    ids = torch.unique(signs, sorted=True)
    out = logps.new_zeros(signs.size(0), len(ids) - 1)
    for i, k in enumerate(ids[1:]):
        out[:, i] = (logps * (signs == k).float()).sum(-1)
    return out


def dpa_loss_torch(h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frz, labels_ref, W, W_ref, alpha, ref_logits=None):
    """Returns (loss, align, div); differentiable w.r.t. h_pair, h_ref_pol, W.
    ref_logits [B,T,V]: the frozen model's (unshifted) logits given instead of h_ref_frz / W_ref."""
    B = h_pair.shape[0] // 2
    lp, _ = _token_logps(h_pair, W, labels_pair)
    y = labels_pair[:, 1:]
    g = signs_pair[:, 1:].masked_fill(signs_pair[:, 1:] == IGNORE, 0)
    lp = lp * (y != IGNORE).float()
    P = _phrase_sums(lp[:B], g[:B])
    N = _phrase_sums(lp[B:], g[B:])
    align = torch.log(1 + torch.exp(N - P)).mean()
    _, z_pol = _token_logps(h_ref_pol, W, labels_ref)
    with torch.no_grad():
        if ref_logits is not None:
            z_frz = ref_logits.float()[:, :-1, :]
        else:
            _, z_frz = _token_logps(h_ref_frz, W_ref, labels_ref)
    q = F.softmax(z_frz, -1)
    p = F.softmax(z_pol, -1)
    kl = q * (q.log() - p.log()) * (labels_ref[:, 1:] != IGNORE).unsqueeze(-1)
    div = kl.sum() / kl.shape[0]
    return align + alpha * div, align, div


def step_cpu(inputs, need_dW=True):
    """One forward+backward on CPU tensors (a dict from halva_b200.synth.make_inputs). Returns the loss value."""
    hp = inputs["h_pair"].detach().clone().requires_grad_(True)
    hr = inputs["h_ref_pol"].detach().clone().requires_grad_(True)
    W = inputs["W"].detach().clone().requires_grad_(need_dW)
    W_ref = W.detach() if inputs["W_ref"] is inputs["W"] else inputs["W_ref"]
    loss, align, div = dpa_loss_torch(hp, inputs["labels_pair"], inputs["signs_pair"], hr, inputs["h_ref_frz"],
                                      inputs["labels_ref"], W, W_ref, inputs["alpha"])
    loss.backward()
    return dict(loss=loss.item(), align=align.item(), div=div.item(), dH_pair=hp.grad, dH_ref=hr.grad, dW=W.grad)
