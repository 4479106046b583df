"""Host-buffer entry of the DPA-loss path: move only the rows that can contribute.

When the hidden states live in (pinned) host memory, the step is bounded by the host->device link long before the
kernels are (1.07 GB per cfg-2 batch; eight GPUs share the host's memory bandwidth).  Labels and phrase ids are known
on the host, and a hidden row (s, t) is only ever read by the kernels when it predicts a contributing label
(halva_trainer.py:403-407,550-562: label[s,t+1] != -100, and for the pair batch also sign[s,t+1] not in {0, -100}).
Those rows form a few contiguous spans per sequence (response tokens / phrase spans), so the stager issues one
asynchronous copy per span into same-shaped device tensors; every other row of the device tensors is never touched
by the library (row compaction), so it need not be transferred.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

IGNORE = -100


def _spans_of(mask_row: torch.Tensor, merge_gap: int) -> List[Tuple[int, int]]:
    idx = torch.nonzero(mask_row, as_tuple=False).flatten().tolist()
    spans: List[Tuple[int, int]] = []
    for t in idx:
        if spans and t - spans[-1][1] <= merge_gap:
            spans[-1] = (spans[-1][0], t + 1)
        else:
            spans.append((t, t + 1))
    return spans


def contributing_rows(labels: torch.Tensor, signs: torch.Tensor | None = None) -> torch.Tensor:
    """[S, T] bool: hidden row (s, t) predicts a label that can contribute (a superset of the rows the kernels pack:
    the smallest phrase id of a half batch is additionally dropped there, halva_trainer.py:412-415)."""
    m = torch.zeros_like(labels, dtype=torch.bool)
    ok = labels[:, 1:] != IGNORE
    if signs is not None:
        ok = ok & (signs[:, 1:] != 0) & (signs[:, 1:] != IGNORE)
    m[:, :-1] = ok
    return m


class HostStager:
    """Plans and issues the host->device copies of one batch.

    tensors: dict with h_pair [2B,T,H], h_ref_pol / h_ref_frz [Br,Tr,H] (pinned host), labels_pair, signs_pair, labels_ref."""

    def __init__(self, labels_pair, signs_pair, labels_ref, dense: bool = False, merge_gap: int = 8):
        self.dense = dense
        self.plan: Dict[str, List[Tuple[int, int, int]]] = {"h_pair": [], "h_ref": []}
        if not dense:
            mp = contributing_rows(labels_pair.cpu(), signs_pair.cpu())
            mr = contributing_rows(labels_ref.cpu())
            for s in range(mp.shape[0]):
                self.plan["h_pair"] += [(s, a, b) for a, b in _spans_of(mp[s], merge_gap)]
            for s in range(mr.shape[0]):
                self.plan["h_ref"] += [(s, a, b) for a, b in _spans_of(mr[s], merge_gap)]

    def rows(self) -> Tuple[int, int]:
        return (sum(b - a for _, a, b in self.plan["h_pair"]), sum(b - a for _, a, b in self.plan["h_ref"]))

    def stage(self, host: Dict[str, torch.Tensor], dev: Dict[str, torch.Tensor]) -> int:
        """Issue the copies on the current stream (non-blocking, pinned source). Returns the bytes moved."""
        n = 0
        for k in ("labels_pair", "signs_pair", "labels_ref"):
            dev[k].copy_(host[k], non_blocking=True)
            n += host[k].numel() * host[k].element_size()
        if self.dense:
            for k in ("h_pair", "h_ref_pol", "h_ref_frz"):
                dev[k].detach().copy_(host[k], non_blocking=True)
                n += host[k].numel() * host[k].element_size()
            return n
        row_bytes = host["h_pair"].shape[-1] * host["h_pair"].element_size()
        dp, hp = dev["h_pair"].detach(), host["h_pair"]
        for s, a, b in self.plan["h_pair"]:
            dp[s, a:b].copy_(hp[s, a:b], non_blocking=True)
            n += (b - a) * row_bytes
        for name in ("h_ref_pol", "h_ref_frz"):
            d, h = dev[name].detach(), host[name]
            for s, a, b in self.plan["h_ref"]:
                d[s, a:b].copy_(h[s, a:b], non_blocking=True)
                n += (b - a) * row_bytes
        return n
