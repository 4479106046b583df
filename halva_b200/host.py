"""Host-buffer entry of the DPA-loss path: move only the rows that can contribute.

When the hidden states live in (pinned) host memory, the step is bounded by the host->device link long before the
kernels are (1.07 GB per cfg-2 batch; eight GPUs share the host's memory bandwidth).  Labels and phrase ids are known
on the host, and a hidden row (s, t) is only ever read by the kernels when it predicts a contributing label
(halva_trainer.py:403-407,550-562: label[s,t+1] != -100, and for the pair batch also sign[s,t+1] not in {0, -100}).
Those rows form a few contiguous spans per sequence (response tokens / phrase spans), so one asynchronous copy per span
goes into same-shaped device tensors; every other row of the device tensors is never touched by the library (row
compaction), so it need not be transferred.

Planning and issuing are native (libhalva_dpa.so: dpa_plan_host_spans / dpa_stage_host_batch, csrc/dpa_host.cu) and
happen per batch, inside the caller's step: `HostStager.stage(host, dev)` plans from THAT batch's labels every time.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

IGNORE = -100


def contributing_rows(labels: torch.Tensor, signs: torch.Tensor | None = None) -> torch.Tensor:
    """[S, T] bool: hidden row (s, t) predicts a label that can contribute (a superset of the rows the kernels pack:
    the smallest phrase id of a half batch is additionally dropped there, halva_trainer.py:412-415)."""
    m = torch.zeros_like(labels, dtype=torch.bool)
    ok = labels[:, 1:] != IGNORE
    if signs is not None:
        ok = ok & (signs[:, 1:] != 0) & (signs[:, 1:] != IGNORE)
    m[:, :-1] = ok
    return m


def plan_spans(labels: torch.Tensor, signs: Optional[torch.Tensor] = None, merge_gap: int = 8) -> List[Tuple[int, int, int]]:
    """(sequence, t0, t1) spans of contributing rows of HOST labels [S, T] (native single pass, no CUDA call)."""
    lib = _lib.load()
    labels = labels.contiguous()
    if labels.is_cuda or labels.dtype not in (torch.int64, torch.int32):
        raise TypeError("plan_spans takes host int64/int32 labels")
    if signs is not None:
        signs = signs.contiguous()
        if signs.is_cuda or signs.dtype != labels.dtype or signs.shape != labels.shape:
            raise TypeError("signs must be a host tensor of the labels' dtype and shape")
    S, T = labels.shape
    cap = S * (T // 2 + 1)
    buf = (_lib.Span * cap)()
    n = lib.dpa_plan_host_spans(labels.data_ptr(), None if signs is None else signs.data_ptr(), S, T,
                                _lib.DPA_I64 if labels.dtype == torch.int64 else _lib.DPA_I32, int(merge_gap), buf, cap)
    if n < 0:
        raise _lib.DPAError(f"dpa_plan_host_spans failed ({n})")
    return [(buf[i].seq, buf[i].t0, buf[i].t1) for i in range(n)]


class HostStager:
    """Issues the host->device copies of one batch; the span plan is rebuilt from the batch's own labels at every
    `stage()` call (microseconds, native).

    host / dev: dicts with h_pair [2B,T,H], h_ref_pol / h_ref_frz [Br,Tr,H], labels_pair, signs_pair, labels_ref --
    `host` pinned host tensors, `dev` device tensors of the same shapes."""

    def __init__(self, dense: bool = False, merge_gap: int = 8):
        self.dense, self.merge_gap = bool(dense), int(merge_gap)
        self.info = _lib.StageInfo()

    def stage(self, host: Dict[str, torch.Tensor], dev: Dict[str, torch.Tensor]) -> int:
        """Plan + issue on the current CUDA stream (non-blocking when `host` is pinned).  Returns the bytes moved;
        `self.info` holds spans / rows / plan_us / issue_us of this call."""
        lib = _lib.load()
        hp, lp = host["h_pair"], host["labels_pair"]
        for k in ("h_pair", "labels_pair", "signs_pair", "h_ref_pol", "labels_ref"):
            if host[k].is_cuda or not dev[k].is_cuda or host[k].shape != dev[k].shape or host[k].dtype != dev[k].dtype \
                    or not host[k].is_contiguous() or not dev[k].is_contiguous():
                raise ValueError(f"{k}: need a contiguous host tensor and a contiguous device tensor of the same shape and dtype")
        frz = host.get("h_ref_frz")
        S, T, H = hp.shape
        Br, Tr, _ = host["h_ref_pol"].shape
        b = _lib.HostBatch(B=S // 2, T=T, B_ref=Br, T_ref=Tr, H=H,
                           dtype=_lib.DPA_F32 if hp.dtype == torch.float32 else _lib.DPA_BF16,
                           index_dtype=_lib.DPA_I64 if lp.dtype == torch.int64 else _lib.DPA_I32,
                           dense=int(self.dense), merge_gap=self.merge_gap,
                           h_pair=hp.data_ptr(), labels_pair=lp.data_ptr(), signs_pair=host["signs_pair"].data_ptr(),
                           h_ref_pol=host["h_ref_pol"].data_ptr(), h_ref_frz=None if frz is None else frz.data_ptr(),
                           labels_ref=host["labels_ref"].data_ptr(),
                           d_h_pair=dev["h_pair"].data_ptr(), d_labels_pair=dev["labels_pair"].data_ptr(),
                           d_signs_pair=dev["signs_pair"].data_ptr(), d_h_ref_pol=dev["h_ref_pol"].data_ptr(),
                           d_h_ref_frz=None if frz is None else dev["h_ref_frz"].data_ptr(), d_labels_ref=dev["labels_ref"].data_ptr())
        with torch.cuda.device(dev["h_pair"].device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.dpa_stage_host_batch(C.byref(b), st, C.byref(self.info)), "dpa_stage_host_batch")
        return int(self.info.bytes)
