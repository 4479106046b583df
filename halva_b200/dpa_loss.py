"""`dpa_loss`: the fused DPA-loss operator (torch.autograd.Function over the C ABI of libhalva_dpa.so).

Mirrors what HalvaTrainer.compute_loss computes once the decoder has produced final hidden states
(/root/reference/llava/train/halva_trainer.py:534-592): same tensor inputs (hidden states, lm_head
weight, labels, phrase ids "signs", frozen-reference hidden states), same scalar loss and gradients.
The hot path is hand-written sm_100a CUDA behind `include/halva_dpa.h`; there is no PyTorch or CPU
fallback -- if the library is missing or the device is not a B200 the call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Args, DPAError, Shape, Workspace


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_inputs(h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frozen, labels_ref, W, W_ref, ref_logits=None, extra=()):
    if h_pair.dim() != 3 or h_pair.shape[0] % 2:
        raise ValueError("h_pair must be [2B, T, H] (first B correct, last B hallucinated sequences)")
    if tuple(h_pair.shape[:2]) != tuple(labels_pair.shape) or tuple(labels_pair.shape) != tuple(signs_pair.shape):
        # same check and message as the reference raises at halva_trainer.py:398-399
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    if tuple(h_ref_pol.shape[:2]) != tuple(labels_ref.shape):
        raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
    if W.dim() != 2 or W.shape[1] != h_pair.shape[2] or h_ref_pol.shape[2] != h_pair.shape[2]:
        raise ValueError("W must be [V, H] with H matching the hidden states")
    frozen = ()
    if ref_logits is not None:
        if ref_logits.dim() != 3 or tuple(ref_logits.shape[:2]) != tuple(labels_ref.shape) or ref_logits.shape[2] != W.shape[0]:
            raise ValueError("ref_logits must be [B_ref, T_ref, V] (the frozen reference model's logits on the reference sample)")
        if not ref_logits.is_cuda:
            raise DPAError(f"ref_logits is on {ref_logits.device}: the DPA-loss path is CUDA (sm_100a) only, there is no CPU fallback")
        if ref_logits.dtype not in (torch.bfloat16, torch.float32):
            raise TypeError("ref_logits must be bf16 or fp32")
    else:
        if h_ref_pol.shape != h_ref_frozen.shape:
            raise ValueError("Logits (batch and sequence length dim) and labels must have the same shape.")
        if W.shape != W_ref.shape:
            raise ValueError("W / W_ref must be [V, H] with H matching the hidden states")
        frozen = (("h_ref_frozen", h_ref_frozen), ("W_ref", W_ref))
    for name, t in (("h_pair", h_pair), ("h_ref_pol", h_ref_pol), ("W", W)) + frozen:
        if t.dtype not in (torch.bfloat16, torch.float32) or t.dtype != h_pair.dtype:
            raise TypeError(f"{name} has dtype {t.dtype}; hidden states and lm_head weights must all be bf16 (the HALVA recipes "
                            "train in bf16) or all fp32 (split-bf16 tensor-core passes, fp32 gradients)")
    for name, t in (("h_pair", h_pair), ("h_ref_pol", h_ref_pol), ("W", W)) + frozen:
        if not t.is_cuda:
            raise DPAError(f"{name} is on {t.device}: the DPA-loss path is CUDA (sm_100a) only, there is no CPU fallback")
    for name, t in (("labels_pair", labels_pair), ("signs_pair", signs_pair), ("labels_ref", labels_ref)):
        if t.dtype not in (torch.int64, torch.int32):
            raise TypeError(f"{name} must be int64 or int32")
    if not (labels_pair.dtype == signs_pair.dtype == labels_ref.dtype):
        raise TypeError("labels_pair, signs_pair and labels_ref must share one integer dtype")
    # every pointer handed to the kernels must be a pointer on ONE device: a CPU label tensor or a weight on another GPU
    # would otherwise fault inside a kernel (sticky CUDA error) instead of raising here
    every = (("h_pair", h_pair), ("labels_pair", labels_pair), ("signs_pair", signs_pair), ("h_ref_pol", h_ref_pol),
             ("labels_ref", labels_ref), ("W", W)) + frozen + ((("ref_logits", ref_logits),) if ref_logits is not None else ()) + \
        tuple((n, t) for n, t in extra if t is not None)
    for name, t in every:
        if not t.is_cuda:
            raise DPAError(f"{name} is on {t.device}: the DPA-loss path is CUDA (sm_100a) only, there is no CPU fallback")
        if t.device != h_pair.device:
            raise DPAError(f"{name} is on {t.device} but h_pair is on {h_pair.device}: all tensors of one call must share a device")


class _Call:
    """Owns the ctypes argument block and the tensors it points to for one forward/backward pair."""

    def __init__(self, h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frozen, labels_ref, W, W_ref, alpha,
                 round_logits, need_dW, dense, chunk_rows, v_offset, V_global, fused, norm=None, ref_logits=None,
                 loss_only=False, grad_hint=1.0):
        lib = _lib.load()
        self.lib = lib
        dev = h_pair.device
        self.keep = [None if t is None else t.contiguous()
                     for t in (h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frozen, labels_ref, W, W_ref)]
        if ref_logits is not None:             # given instead of h_ref_frozen / W_ref; any row stride whose bytes % 16 == 0
            esz = ref_logits.element_size()
            if ref_logits.stride(2) != 1 or ref_logits.stride(0) != ref_logits.shape[1] * ref_logits.stride(1) or \
                    (ref_logits.stride(1) * esz) % 16 or ref_logits.data_ptr() % 16:
                Vg = ref_logits.shape[2]                               # odd vocabulary sizes: pad the row stride to 16 bytes
                buf = torch.empty(*ref_logits.shape[:2], -(-Vg * esz // 16) * 16 // esz, dtype=ref_logits.dtype, device=dev)
                buf[..., :Vg].copy_(ref_logits)
                ref_logits = buf[..., :Vg]
            self.keep.append(ref_logits)
        hp, lp, sp, hr, hf, lr, w, wr = self.keep[:8]
        B2, T, H = hp.shape
        Br, Tr, _ = hr.shape
        flags = (_lib.FLAG_ROUND_LOGITS if round_logits else 0) | (_lib.FLAG_NEED_DW if need_dW else 0) | \
                (_lib.FLAG_DENSE if dense else 0) | (_lib.FLAG_FUSED_GRAD if fused else 0) | \
                (_lib.FLAG_FUSED_NORM if norm is not None else 0) | (_lib.FLAG_LOSS_ONLY if loss_only else 0)
        # the upstream gradient the caller expects (1 / gradient-accumulation steps): the fused forward forms the gradients
        # for it, and the backward rescales them only if the actual d_loss differs
        self.hint = None
        if fused and not loss_only and float(grad_hint) != 1.0:
            self.hint = torch.full((1,), float(grad_hint), dtype=torch.float32, device=dev)
            flags |= _lib.FLAG_GRAD_HINT
        self.loss_only = loss_only
        self.fused = fused
        self.lib_comm = False
        self.norm = None
        if norm is not None:                  # (weight, frozen model's weight, eps, weight needs grad)
            nw, nwr, eps, need = norm
            self.norm = (nw.contiguous(), nwr.contiguous(), float(eps), bool(need))
            self.keep += [self.norm[0], self.norm[1]]
        self.grads = None
        self.shape = Shape(B=B2 // 2, T=T, B_ref=Br, T_ref=Tr, H=H, V=V_global, V_local=w.shape[0], v_offset=v_offset,
                           dtype=_lib.DPA_F32 if hp.dtype == torch.float32 else _lib.DPA_BF16, index_dtype=_lib.DPA_I64 if lp.dtype == torch.int64 else _lib.DPA_I32,
                           flags=flags, chunk_rows=chunk_rows)
        ws = Workspace()
        _lib.check(lib.dpa_query_workspace(C.byref(self.shape), C.byref(ws)), "dpa_query_workspace")
        self.ws = ws
        self.state = torch.empty(ws.state_bytes, dtype=torch.uint8, device=dev)
        self.losses = torch.empty(4, dtype=torch.float32, device=dev)
        self.status = torch.empty(8, dtype=torch.int32, device=dev)
        self.need_dW = need_dW
        self.args = Args(shape=self.shape, alpha=float(alpha),
                         h_pair=_ptr(hp), labels_pair=_ptr(lp), signs_pair=_ptr(sp), h_ref_pol=_ptr(hr), h_ref_frz=_ptr(hf),
                         labels_ref=_ptr(lr), W=_ptr(w), W_ref=_ptr(wr),
                         state=_ptr(self.state), state_bytes=ws.state_bytes, scratch=None, scratch_bytes=0,
                         losses=_ptr(self.losses), status=_ptr(self.status), d_loss=_ptr(self.hint), dH_pair=None, dH_ref=None, dW=None,
                         norm_w=_ptr(self.norm[0]) if self.norm else None, norm_w_ref=_ptr(self.norm[1]) if self.norm else None,
                         d_norm_w=None, norm_eps=self.norm[2] if self.norm else 0.0,
                         ref_logits=_ptr(ref_logits), ref_logits_ld=0 if ref_logits is None else ref_logits.stride(1),
                         ref_logits_dtype=_lib.DPA_F32 if ref_logits is not None and ref_logits.dtype == torch.float32 else _lib.DPA_BF16)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.state.device).cuda_stream)

    def _with_scratch(self, nbytes):
        scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.state.device)
        self.args.scratch = _ptr(scratch)
        self.args.scratch_bytes = int(nbytes)
        return scratch

    def _alloc_grads(self):
        hp, _, _, hr, _, _, w, _ = self.keep[:8]
        dH_pair, dH_ref = torch.empty_like(hp), torch.empty_like(hr)
        dW = torch.empty_like(w) if self.need_dW else None
        dnw = torch.empty_like(self.norm[0]) if self.norm and self.norm[3] else None
        self.args.dH_pair, self.args.dH_ref, self.args.dW, self.args.d_norm_w = _ptr(dH_pair), _ptr(dH_ref), _ptr(dW), _ptr(dnw)
        return dH_pair, dH_ref, dW, dnw

    def forward(self, vocab_group=None):
        with torch.cuda.device(self.state.device):
            if self.fused:
                # loss and (d_loss = 1) gradients in one pass; backward() only scales.  With a vocab_group this is the
                # library-side exchange (statistics all-gather per segment / chunk, dH through peer memory) inside the one call
                self.lib_comm = vocab_group is not None
                self.grads = None if self.loss_only else self._alloc_grads()
                scratch = self._with_scratch(self.ws.fused_bytes)
                _lib.check(self.lib.dpa_forward(C.byref(self.args), self._stream()), "dpa_forward[fused]")
                del scratch
                self.args.scratch, self.args.scratch_bytes = None, 0
                return
            scratch = self._with_scratch(self.ws.fwd_bytes)
            # vocab-parallel with the library's own NCCL communicator (halva_b200.dist.init_library_comm): the exchange
            # steps happen inside dpa_forward / dpa_backward
            from . import dist as _dist
            self.lib_comm = vocab_group is not None and self.lib.dpa_comm_world() > 1 and _dist.library_comm_matches(vocab_group)
            if vocab_group is None or self.lib_comm:
                _lib.check(self.lib.dpa_forward(C.byref(self.args), self._stream()), "dpa_forward")
            else:
                import torch.distributed as dist
                _lib.check(self.lib.dpa_forward_stats(C.byref(self.args), self._stream()), "dpa_forward_stats")
                p, n = C.c_void_p(), C.c_size_t()
                _lib.check(self.lib.dpa_row_stats_view(C.byref(self.args), C.byref(p), C.byref(n)), "dpa_row_stats_view")
                off = p.value - self.state.data_ptr()
                rec = self.state[off:off + n.value]
                G = dist.get_world_size(vocab_group)
                gathered = torch.empty(G * n.value, dtype=torch.uint8, device=rec.device)
                dist.all_gather_into_tensor(gathered, rec, group=vocab_group)     # the one exchange step of the forward
                _lib.check(self.lib.dpa_forward_finish(C.byref(self.args), C.c_void_p(gathered.data_ptr()), G, self._stream()),
                           "dpa_forward_finish")
                del gathered
            del scratch            # stream-ordered free: safe to recycle after the kernels queued above
        self.args.scratch, self.args.scratch_bytes = None, 0

    def backward(self, d_loss: torch.Tensor):
        with torch.cuda.device(self.state.device):
            d_loss = d_loss.detach().to(torch.float32).reshape(1).contiguous()
            self.args.d_loss = _ptr(d_loss)
            if self.fused:
                if self.grads is None:
                    raise RuntimeError("dpa_loss(backward='fused') forms its gradients during the forward pass and hands them over "
                                       "(scaled in place) at the first backward; a second backward through the same call "
                                       "(retain_graph=True, autograd.grad twice) needs backward='recompute'")
                _lib.check(self.lib.dpa_backward(C.byref(self.args), self._stream()), "dpa_backward[fused]")
                grads, self.grads = self.grads, None
                return grads
            dH_pair, dH_ref, dW, dnw = self._alloc_grads()
            scratch = self._with_scratch(self.ws.bwd_bytes)
            _lib.check(self.lib.dpa_backward(C.byref(self.args), self._stream()), "dpa_backward")
            del scratch
        self.args.scratch, self.args.scratch_bytes = None, 0
        return dH_pair, dH_ref, dW, dnw


class DPALossFn(torch.autograd.Function):
    """(h_pair, h_ref_pol, W, norm_weight) -> (loss, align, div); integer inputs and the frozen branch get no gradient."""

    @staticmethod
    def forward(ctx, h_pair, h_ref_pol, W, norm_w, labels_pair, signs_pair, labels_ref, h_ref_frozen, W_ref, alpha, round_logits,
                need_dW, dense, chunk_rows, vocab_group, v_offset, V_global, fused, norm_w_ref, norm_eps, ref_logits, grad_hint=1.0):
        want_grad = any(ctx.needs_input_grad[:4])
        loss_only = ref_logits is not None and not want_grad        # given logits live on the chunked (fused) forward only
        fused = bool(fused and (want_grad or loss_only))            # no gradient wanted (eval): statistics only
        norm = None if norm_w is None else (norm_w.detach(), norm_w_ref.detach(), norm_eps, ctx.needs_input_grad[3])
        det = lambda t: None if t is None else t.detach()
        call = _Call(h_pair.detach(), labels_pair, signs_pair, h_ref_pol.detach(), det(h_ref_frozen), labels_ref,
                     W.detach(), det(W_ref), alpha, round_logits, need_dW and not loss_only, dense, chunk_rows, v_offset, V_global,
                     fused, norm, det(ref_logits), loss_only, grad_hint)
        call.forward(vocab_group)
        ctx.call = call
        ctx.status_t = call.status
        ctx.vocab_group = vocab_group
        # registered so that autograd's version counters catch an in-place edit of the inputs between forward and backward
        # (the recompute backward reads them again); no copy is made
        ctx.save_for_backward(h_pair, h_ref_pol, W)
        loss, align, div = call.losses[0], call.losses[1], call.losses[2]
        ctx.mark_non_differentiable(align, div, call.status)
        return loss, align, div, call.status

    @staticmethod
    def backward(ctx, g_loss, g_align, g_div, g_status):
        call = ctx.call
        if call is None:
            raise RuntimeError("dpa_loss: backward through the same call a second time (its buffers were released); "
                               "use backward='recompute' and retain_graph=True")
        _ = ctx.saved_tensors                       # raises if an input was modified in place since the forward
        dH_pair, dH_ref, dW, dnw = call.backward(g_loss)
        if ctx.vocab_group is not None and not call.lib_comm:
            import torch.distributed as dist
            dist.all_reduce(dH_pair, group=ctx.vocab_group)      # partial sums over vocab shards
            dist.all_reduce(dH_ref, group=ctx.vocab_group)
        if call.fused:
            ctx.call = None                         # recompute mode keeps its O(rows) state: a retained graph can run again
        return (dH_pair, dH_ref, dW, dnw) + (None,) * 18


def dpa_loss(h_pair: torch.Tensor, labels_pair: torch.Tensor, signs_pair: torch.Tensor,
             h_ref_pol: torch.Tensor, labels_ref: torch.Tensor, W: torch.Tensor, *,
             h_ref_frozen: Optional[torch.Tensor] = None, W_ref: Optional[torch.Tensor] = None,
             ref_logits: Optional[torch.Tensor] = None, alpha: float,
             round_logits: bool = True, need_dW: Optional[bool] = None, dense: bool = False, chunk_rows: int = 0,
             vocab_group=None, v_offset: Optional[int] = None, vocab_size: Optional[int] = None,
             backward: str = "auto", norm_weight: Optional[torch.Tensor] = None,
             norm_weight_ref: Optional[torch.Tensor] = None, norm_eps: float = 1e-6, strict: bool = False,
             grad_scale_hint: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fused DPA loss.  Returns (loss, alignment_loss, divergence) as 0-dim fp32 CUDA tensors; `loss` carries grad.

    h_pair       [2B, T, H] bf16   final hidden states; first B sequences "pos" (correct), last B "neg"
                                   (hallucinated) -- the stacking of halva_trainer.py:434-447,485-487
    labels_pair  [2B, T]  int      post-splice labels, -100 = ignore      (unshifted: the op applies the shift)
    signs_pair   [2B, T]  int      phrase ids: 0 / -100 = none, k >= 1 = k-th <MASK> span
    h_ref_pol    [Br,Tr,H] bf16    policy hidden states on the reference sample
    labels_ref   [Br,Tr]  int
    W            [V, H]   bf16     lm_head weight (or this rank's vocab shard with `vocab_group`)
    h_ref_frozen [Br,Tr,H] bf16    frozen reference model's hidden states on the reference sample
    W_ref        [V, H]            frozen lm_head (default: W itself, the LoRA recipe)
    ref_logits   [Br,Tr,V] bf16/fp32  the frozen model's logits on the reference sample, given INSTEAD of h_ref_frozen /
                                   W_ref (e.g. cached: the frozen pass of halva_trainer.py:576-578 does not depend on the
                                   policy).  fp32 values are expected to be bf16-representable (`.float()` of a bf16
                                   model's logits, :520) and are otherwise rounded to bf16.  Needs round_logits=True and
                                   no vocab_group (served by the chunked "fused" pass; one lm_head GEMM less per row).
    alpha                          loss = align + alpha * div            (halva_trainer.py:590)
    round_logits                   True: round logits to bf16 before the fp32 softmax, exactly as the reference's
                                   bf16 nn.Linear does (modeling_llama.py:1028-1029); False: keep fp32 accumulators
    fp32 inputs                    h_pair / h_ref_pol / h_ref_frozen / W / W_ref may all be fp32 instead (an fp32 model, as in
                                   the reference's CPU-runnable configuration): every operand goes to the tensor cores as a
                                   split-bf16 triple (hi + mid + lo, six K segments per GEMM with the accumulator promoted
                                   to fp32 registers every K = 256), logits and dLogits stay fp32-accurate, gradients come
                                   back in fp32.  round_logits is ignored (an fp32 nn.Linear does not round).  The chunked
                                   ("fused") pass keeps fp32 logits, so it too runs the algorithmic GEMM count.
    need_dW                        default W.requires_grad (False under the shipped LoRA recipe: lm_head frozen)
    dense                          process every fed position instead of compacting to rows with a non-zero
                                   contribution (what the reference executes; benchmarking only)
    norm_weight [H]                when given, h_pair / h_ref_pol / h_ref_frozen are the PRE-norm residual stream and the
                                   decoder's final RMSNorm (weight `norm_weight`, frozen model `norm_weight_ref`, `norm_eps`)
                                   is applied to the active rows inside the op (llava/model/language_model/
                                   modelling_llama.py:55-70, :691); gradients come back w.r.t. the pre-norm stream and,
                                   if it requires grad, norm_weight
    backward                       "recompute": forward keeps only per-row statistics, backward recomputes the logits
                                   chunk-wise (BASELINE north star);  "fused": the forward pass also forms the gradients
                                   (for d_loss = 1) -- reference-sample logits of one row chunk are kept in bf16 and
                                   turned into dLogits in place (executed GEMMs = the algorithmic count: 4 per such row
                                   instead of 6, 3 instead of 4 per pair row) -- and
                                   backward only scales by d_loss;  "auto" = "fused" when round_logits is on, a gradient
                                   is needed and there is no vocab_group (or the library's own communicator spans it:
                                   halva_b200.dist.init_library_comm), else "recompute".
    grad_scale_hint                the upstream gradient `loss.backward()` is EXPECTED to deliver (HF Trainer / DeepSpeed divide
                                   the loss by gradient_accumulation_steps before the backward: 1 / steps).  The "fused" pass
                                   forms its gradients for that value during the forward, so that the backward has nothing
                                   to rescale (otherwise one pass over dH / dW); a wrong hint only costs that pass, the
                                   gradients are the same.  Ignored by the "recompute" strategy (it sees the real value).
    Ill-formed phrase ids (K_pos != K_neg, K == 0, id > 255) or labels make the losses AND every gradient NaN and set
    bits in `read_status(loss)["status"]` / `loss.dpa_status`; the reference raises / returns NaN there (SURVEY.md
    section 3.2) -- a truncated batch can never reach the optimiser as finite-but-wrong gradients.
    strict                         True: read the status word back after the forward (ONE host synchronisation) and raise
                                   like the reference does -- RuntimeError for K_pos != K_neg (the broadcast failure of
                                   halva_trainer.py:567), ValueError for phrase ids / labels out of range; K == 0 stays a
                                   NaN loss, as in the reference.  Off by default: the path is free of host syncs.
    """
    if (ref_logits is None) == (h_ref_frozen is None):
        raise ValueError("give exactly one of h_ref_frozen (+ W_ref) and ref_logits")
    grad_scale_hint = float(grad_scale_hint)
    if grad_scale_hint != grad_scale_hint or abs(grad_scale_hint) == float("inf") or grad_scale_hint == 0.0:
        raise ValueError("grad_scale_hint must be a finite, non-zero number (the expected upstream gradient of the loss)")
    if ref_logits is not None:
        if W_ref is not None:
            raise ValueError("W_ref has no meaning together with ref_logits")
        if not round_logits or vocab_group is not None or backward == "recompute":
            raise ValueError("ref_logits needs round_logits=True, no vocab_group and the 'fused' (chunked) pass")
        if norm_weight is not None and norm_weight_ref is None:
            norm_weight_ref = norm_weight
    else:
        W_ref = W if W_ref is None else W_ref
    if h_pair.dtype == torch.float32 and (ref_logits is not None or norm_weight is not None):
        raise ValueError("fp32 inputs: ref_logits and norm_weight belong to the bf16 execution")
    _check_inputs(h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frozen, labels_ref, W, W_ref, ref_logits,
                  extra=(("norm_weight", norm_weight), ("norm_weight_ref", norm_weight_ref)))
    if need_dW is None:
        need_dW = bool(W.requires_grad)
    if vocab_group is not None:
        import torch.distributed as dist
        if v_offset is None or vocab_size is None:
            # shard sizes may differ (halva_b200.dist.vocab_shard_bounds): derive the offsets from the sizes of all shards
            from .dist import shard_offsets
            off, total = shard_offsets(W.shape[0], vocab_group, W.device)
            v_offset = off if v_offset is None else v_offset
            vocab_size = total if vocab_size is None else vocab_size
    else:
        v_offset, vocab_size = 0, W.shape[0]
    if norm_weight is not None:
        if vocab_group is not None:
            raise ValueError("norm_weight (fused final RMSNorm) is not supported together with vocab_group")
        norm_weight_ref = norm_weight if norm_weight_ref is None else norm_weight_ref
        for name, t in (("norm_weight", norm_weight), ("norm_weight_ref", norm_weight_ref)):
            if not t.is_cuda or t.dtype != torch.bfloat16 or t.numel() != h_pair.shape[2]:
                raise TypeError(f"{name} must be a CUDA bf16 tensor of H elements")
    if backward not in ("auto", "recompute", "fused"):
        raise ValueError("backward must be 'auto', 'recompute' or 'fused'")
    fp32 = h_pair.dtype == torch.float32
    if fp32:
        # fp32 model (BASELINE configs[0]): nn.Linear emits fp32 logits, so there is no bf16 rounding point to reproduce; the
        # chunked pass keeps fp32 logits (the promoted sums of the split-bf16 GEMMs) and emits split-bf16 dLogits
        round_logits = False
        if vocab_group is not None and backward == "auto":
            backward = "recompute"
    lib_vp = False
    if vocab_group is not None:
        from .dist import library_comm_matches
        lib_vp = _lib.load().dpa_comm_world() > 1 and library_comm_matches(vocab_group)
    if backward == "fused" and ((not round_logits and not fp32) or (vocab_group is not None and (not lib_vp or fp32))):
        raise ValueError("backward='fused' needs round_logits=True (or fp32 inputs) and, with a vocab_group, bf16 inputs and the "
                         "library's own communicator over that group (halva_b200.dist.init_library_comm)")
    fused = backward == "fused" or (backward == "auto" and (round_logits or fp32) and (vocab_group is None or lib_vp))
    loss, align, div, status = DPALossFn.apply(h_pair, h_ref_pol, W, norm_weight, labels_pair, signs_pair, labels_ref, h_ref_frozen,
                                               W_ref, float(alpha), bool(round_logits), bool(need_dW), bool(dense), int(chunk_rows),
                                               vocab_group, int(v_offset), int(vocab_size), fused, norm_weight_ref, float(norm_eps),
                                               ref_logits, float(grad_scale_hint))
    loss.dpa_status = status                 # device int32[8]: status bits, K, active rows ... (read_status(loss) decodes it)
    out = (loss, align, div)
    if strict:
        st = status.cpu().tolist()
        if st is not None:
            bits = int(st[0])
            if bits & _lib.STATUS_K_MISMATCH:
                raise RuntimeError(f"The size of tensor a ({int(st[4])}) must match the size of tensor b ({int(st[5])}) at non-singleton "
                                   "dimension 1 (phrase columns of the correct / hallucinated halves, halva_trainer.py:567)")
            if bits & (_lib.STATUS_BAD_SIGN | _lib.STATUS_BAD_LABEL):
                raise ValueError("phrase id outside [0, 255] or label outside [0, V) (status %d)" % bits)
    return out


def read_status(loss: torch.Tensor):
    """Synchronising debug helper: status word / K / active-row counts of the call that produced `loss`."""
    st = getattr(loss, "dpa_status", None)
    if st is None:
        st = getattr(loss.grad_fn, "status_t", None)
    if st is None:
        raise ValueError("not a dpa_loss output")
    s = st.cpu().tolist()
    return dict(status=s[0], K=s[1], n_pair_active=s[2], n_ref_active=s[3], K_pos=s[4], K_neg=s[5])


def debug_gemm(A: torch.Tensor, B: torch.Tensor, a_major: int = 0, b_major: int = 0, bn: int = 256) -> torch.Tensor:
    """Unit-test hook: D = A_logical @ B_logical^T (fp32) through the library's tcgen05 mainloop.
    K-major operands are [M,K] / [N,K]; MN-major operands are stored transposed, [K,M] / [K,N]."""
    lib = _lib.load()
    M, K = (A.shape if a_major == 0 else (A.shape[1], A.shape[0]))
    N = B.shape[0] if b_major == 0 else B.shape[1]
    D = torch.empty(M, N, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(lib.dpa_debug_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, a_major, b_major, bn,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dpa_debug_gemm")
    return D
