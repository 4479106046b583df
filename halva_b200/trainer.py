"""Drop-in replacement for ``HalvaTrainer.compute_loss`` (both flavours of the reference).

The reference computes the DPA loss from full ``[rows, vocab]`` fp32 logits
(/root/reference/llava/train/halva_trainer.py:534-592, /root/reference/vila/train/halva_trainer.py:780-852).
``fused_compute_loss`` performs the same three model passes (stacked correct/hallucinated pair batch, policy on the
reference sample, frozen reference model on the reference sample) but stops each of them at the decoder's final hidden
states and hands those, the post-splice labels / phrase ids and the lm_head weights to :func:`halva_b200.dpa_loss` --
the logits never exist.  Usage:

    from halva_b200.trainer import patch_halva_trainer
    patch_halva_trainer(HalvaTrainer)          # before trainer.train()

What is mirrored from the reference (and where):
  * pos/neg stacking with pad values ids 0 / labels -100 / mask False / signs 0      llava halva_trainer.py:434-447
  * the signed multimodal splice (image rows get label = sign = -100)                llava halva_trainer.py:458-466
  * images duplicated for the two halves                                             llava halva_trainer.py:464
  * reference sample: unsigned splice, policy with grad, ref_model under no_grad    llava halva_trainer.py:571-578
  * VILA flavour: ``model`` argument is used (not ``self.model``), ``ref_images.squeeze(1)``, ``loss_holder`` logging
                                                                                     vila halva_trainer.py:795-848
No CPU / eager fallback: the loss itself always runs through libhalva_dpa.so.
"""
from __future__ import annotations

import contextlib
import importlib
from typing import Any, Dict, Optional, Tuple

import torch

_dpa_loss_mod = importlib.import_module(".dpa_loss", __package__)   # the module (the package re-exports the function)

IGNORE_INDEX = -100          # /root/reference/llava/constants.py:7


# ----------------------------------------------------------------------------------------------- model plumbing
def unwrap_model(model):
    """Peel DDP / DeepSpeed engine (``.module``) and PEFT (``.base_model.model``) wrappers down to the
    LlavaLlamaForCausalLM (LLaVA) or LlavaLlamaModel (VILA) that owns the multimodal splice."""
    seen = 0
    while seen < 8:
        seen += 1
        if hasattr(model, "prepare_inputs_labels_for_multimodal"):
            return model
        if hasattr(model, "module"):
            model = model.module
        elif hasattr(model, "base_model") and hasattr(model.base_model, "model"):
            model = model.base_model.model
        else:
            break
    if not hasattr(model, "prepare_inputs_labels_for_multimodal"):
        raise TypeError(f"{type(model).__name__}: cannot find the LLaVA/VILA model (no prepare_inputs_labels_for_multimodal)")
    return model


def _peel_peft(m):
    while hasattr(m, "base_model") and hasattr(m.base_model, "model") and not hasattr(m, "lm_head"):
        m = m.base_model.model
    return m


def decoder_and_head(mm_model) -> Tuple[Any, torch.nn.Module, str]:
    """-> (decoder stack that returns final, normed hidden states; lm_head module; flavour).
    LLaVA: LlavaLlamaForCausalLM.get_model() / .lm_head  (llava/model/language_model/llava_llama.py:26-40).
    VILA:  the language model sits under ``.llm`` (a LlamaForCausalLM, maybe PEFT-wrapped:
           vila/train/train_halva.py:410-412)."""
    if hasattr(mm_model, "llm"):
        llm = _peel_peft(mm_model.llm)
        inner = llm.get_decoder() if hasattr(llm, "get_decoder") else llm.model
        return inner, llm.lm_head, "vila"
    inner = mm_model.get_model() if hasattr(mm_model, "get_model") else mm_model.model
    return inner, mm_model.lm_head, "llava"


@contextlib.contextmanager
def gathered_weight(*weights):
    """ZeRO-3 partitions lm_head.weight and re-assembles it inside lm_head.forward's hooks; reading the weight directly
    bypasses them, so gather explicitly (the pattern of llava/train/train_halva.py:102-113).  No-op otherwise."""
    parts = [w for w in weights if hasattr(w, "ds_id")]
    if not parts:
        yield
        return
    import deepspeed  # only reachable when DeepSpeed created the partitioned parameters
    with deepspeed.zero.GatheredParameters(parts, modifier_rank=None):
        yield


def stack_pair_batch(inputs: Dict[str, torch.Tensor]):
    """Correct ("pos") sequences first, hallucinated ("neg") second, both padded to the longer of the two:
    ids 0, labels -100, attention False, signs 0  (llava halva_trainer.py:434-447)."""
    ids, neg_ids = inputs["input_ids"], inputs["neg_input_ids"]
    B, T = ids.shape[0], max(ids.shape[1], neg_ids.shape[1])

    def stack(a, b, fill, dtype=None):
        out = torch.full((2 * B, T), fill, dtype=dtype or a.dtype, device=a.device)
        out[:B, :a.shape[1]] = a
        out[B:, :b.shape[1]] = b
        return out

    return (stack(ids, neg_ids, 0),
            stack(inputs["labels"], inputs["neg_labels"], IGNORE_INDEX),
            stack(inputs["attention_mask"], inputs["neg_attention_mask"], False, torch.bool),
            stack(inputs["pos_signs"], inputs["neg_signs"], 0))


@contextlib.contextmanager
def _final_norm_bypassed(inner, bypass: bool):
    """LlamaModel applies `self.norm` as its last step (llava/model/language_model/modelling_llama.py:543,691).  With
    `bypass` the decoder returns the pre-norm residual stream and the norm is applied inside dpa_loss to the rows that
    matter."""
    if not bypass:
        yield None
        return
    norm = inner.norm
    inner.norm = torch.nn.Identity()
    try:
        yield norm
    finally:
        inner.norm = norm


def final_hidden_states(model, input_ids, images, labels, attention_mask, signs=None, skip_final_norm: bool = False):
    """Splice the image features in and run the decoder up to (and including, unless skip_final_norm) its final norm.
    Returns (hidden [S,T',H], labels [S,T'], signs [S,T'] or None) aligned row for row, post-splice."""
    mm = unwrap_model(model)
    inner, _, flavour = decoder_and_head(mm)
    if signs is not None:
        (_, position_ids, attention_mask, _, embeds, labels, signs) = mm.prepare_inputs_labels_for_multimodal_signed(
            input_ids, None, attention_mask, None, labels, images, signs)
    else:
        (_, position_ids, attention_mask, _, embeds, labels) = mm.prepare_inputs_labels_for_multimodal(
            input_ids, None, attention_mask, None, labels, images)
    kw = dict(inputs_embeds=embeds, attention_mask=attention_mask, position_ids=position_ids, use_cache=False, return_dict=True)
    if flavour == "vila" and attention_mask is not None:
        # VILA's patched Llama takes the per-sequence lengths for its varlen flash-attention
        # (vila/model/language_model/llava_llama.py:150-166)
        kw["seqlens_in_batch"] = attention_mask.sum(-1).int()
    with _final_norm_bypassed(inner, skip_final_norm):
        out = inner(**kw)
    hidden = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
    return hidden, labels, signs


# ----------------------------------------------------------------------------------------------- the override
def fused_compute_loss(trainer, model, inputs, return_outputs: bool = False, num_items_in_batch: Optional[int] = None,
                       *, round_logits: bool = True, backward: str = "auto", sync_logging: bool = False,
                       fuse_final_norm: bool = False, strict: bool = False, grad_scale_hint: Optional[float] = None):
    """Same contract as ``HalvaTrainer.compute_loss(model, inputs)``: returns the 0-dim fp32 loss (with grad).

    The LLaVA flavour of the reference ignores its ``model`` argument and calls ``self.model``
    (llava halva_trainer.py:548,573); the VILA flavour uses the argument (vila halva_trainer.py:795,821).  Here the
    argument is used when it wraps a multimodal model, else ``trainer.model``.  ``return_outputs`` is accepted and, like
    in the reference, does not change the return value unless True (then ``(loss, {"align", "div"})``).
    ``num_items_in_batch`` (newer transformers) is accepted and ignored, as the reference would.
    ``fuse_final_norm``: run the decoders without their final RMSNorm and let dpa_loss apply it (and its backward) to
    the contributing rows only.
    ``grad_scale_hint``: the upstream gradient the trainer's backward will deliver; default 1 / ``trainer.args.
    gradient_accumulation_steps`` (``Trainer.training_step`` / the DeepSpeed engine divide the loss by it before the backward),
    so that the gradients formed during the forward need no rescaling pass.  Only a hint: a different actual value is honoured.
    ``strict``: read the op's status word back after the forward (one host sync per step) and raise where the reference
    raises -- ``RuntimeError`` for K_pos != K_neg (llava halva_trainer.py:567), ``ValueError`` for phrase ids above 255 /
    labels outside the vocabulary.  Off by default: an ill-formed batch then yields a NaN loss and NaN gradients (never
    finite-but-wrong ones), which the trainer's own NaN handling sees.
    """
    try:
        policy = unwrap_model(model)
    except TypeError:
        policy = unwrap_model(trainer.model)
    frozen = unwrap_model(trainer.ref_model)
    alpha = float(trainer.loss_alpha)
    _, head, flavour = decoder_and_head(policy)
    _, ref_head, _ = decoder_and_head(frozen)

    ids, labels, mask, signs = stack_pair_batch(inputs)
    images = inputs["images"]
    fn = fuse_final_norm
    h_pair, labels_pair, signs_pair = final_hidden_states(policy, ids, torch.cat([images, images], dim=0), labels, mask, signs, fn)

    ref_images = inputs["ref_images"]
    if flavour == "vila":
        ref_images = ref_images.squeeze(1)                                    # vila halva_trainer.py:760 (no-op unless dim 1 is 1)
    h_ref_pol, labels_ref, _ = final_hidden_states(policy, inputs["ref_input_ids"], ref_images, inputs["ref_labels"],
                                                   inputs["ref_attention_mask"], None, fn)
    # The frozen model's pass does not depend on the policy (halva_trainer.py:576-578): a data pipeline may cache its logits
    # on the reference sample (post-splice, [B, T', V]) and put them into the batch as "ref_logits" -- then neither the
    # frozen decoder nor its lm_head runs here, and the op reads the logits in place.
    ref_logits = inputs.get("ref_logits") if hasattr(inputs, "get") else None
    h_ref_frz = None
    if ref_logits is None:
        with torch.no_grad():
            h_ref_frz, _, _ = final_hidden_states(frozen, inputs["ref_input_ids"], ref_images, inputs["ref_labels"],
                                                  inputs["ref_attention_mask"], None, fn)
    elif tuple(ref_logits.shape[:2]) != tuple(labels_ref.shape):
        raise ValueError(f"inputs['ref_logits'] is {tuple(ref_logits.shape)}: expected the post-splice shape "
                         f"[{labels_ref.shape[0]}, {labels_ref.shape[1]}, V] of the reference sample")
    norm_kw = {}
    if fn:
        n_pol, n_frz = decoder_and_head(policy)[0].norm, decoder_and_head(frozen)[0].norm
        norm_kw = dict(norm_weight=n_pol.weight, norm_weight_ref=n_frz.weight.detach(),
                       norm_eps=float(getattr(n_pol, "variance_epsilon", getattr(n_pol, "eps", 1e-6))))

    W, W_ref = head.weight, ref_head.weight
    with gathered_weight(W, W_ref):
        shared = W_ref is W
        if W.dtype != h_pair.dtype:
            # torch.autocast(bf16) over fp32 master weights: F.linear casts the weight to the activations' dtype
            # (differentiably), and so does the shim; an fp32 model with fp32 activations goes through the op's fp32 path
            W_op, W_ref_op = W.to(h_pair.dtype), (None if shared else W_ref.detach().to(h_pair.dtype))
        else:
            W_op, W_ref_op = W, (None if shared else W_ref.detach())
        frozen_kw = dict(h_ref_frozen=h_ref_frz, W_ref=W_ref_op) if ref_logits is None else dict(ref_logits=ref_logits)
        if grad_scale_hint is None:
            gas = getattr(getattr(trainer, "args", None), "gradient_accumulation_steps", 1)
            grad_scale_hint = 1.0 / gas if isinstance(gas, int) and gas > 1 else 1.0
        loss, align, div = _dpa_loss_mod.dpa_loss(
            h_pair, labels_pair, signs_pair, h_ref_pol, labels_ref, W_op, alpha=alpha, round_logits=round_logits,
            need_dW=bool(W.requires_grad), backward=backward, strict=strict, grad_scale_hint=grad_scale_hint,
            **frozen_kw, **norm_kw)

    holder = getattr(trainer, "loss_holder", None)
    if isinstance(holder, dict) and all(k in holder for k in ("contrastive_loss", "divergence", "loss")):
        # vila halva_trainer.py:844-848 logs python floats (three device syncs per step); by default we keep the
        # device scalars and let the caller sync when it actually reads them
        rec = (lambda t: round(t.item(), 7)) if sync_logging else (lambda t: t.detach())
        holder["contrastive_loss"].append(rec(align))
        holder["divergence"].append(rec(div))
        holder["loss"].append(rec(loss))
    if return_outputs:
        return loss, {"align": align, "div": div}
    return loss


def patch_halva_trainer(trainer_cls, **defaults):
    """Replace ``trainer_cls.compute_loss`` by the fused path (class-level monkey patch; returns the original)."""
    original = trainer_cls.compute_loss

    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        return fused_compute_loss(self, model, inputs, return_outputs, num_items_in_batch, **defaults)

    compute_loss.__doc__ = fused_compute_loss.__doc__
    trainer_cls.compute_loss = compute_loss
    trainer_cls._reference_compute_loss = original
    return original
