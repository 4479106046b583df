"""halva_b200.trainer.fused_compute_loss against the reference's own compute_loss on a stand-in LLaVA/VILA model.

CPU part: the shim's plumbing (pos/neg stacking, signed splice, hidden-state extraction, PEFT/engine unwrapping,
VILA logging) is checked with the loss op replaced by the oracle's torch port -- against the AST-lifted, unmodified
reference compute_loss when /root/reference is mounted.  GPU part: the same shim with the real CUDA op.
"""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import dpa_port

IGNORE = -100


class _Inner(nn.Module):
    """Stand-in decoder stack: token-wise MLP + 'final norm', returns .last_hidden_state like LlamaModel."""

    def __init__(self, H):
        super().__init__()
        self.fc = nn.Linear(H, H)
        self.norm = nn.LayerNorm(H)

    def forward(self, inputs_embeds=None, attention_mask=None, position_ids=None, use_cache=False, return_dict=True,
                seqlens_in_batch=None):
        return types.SimpleNamespace(last_hidden_state=self.norm(torch.tanh(self.fc(inputs_embeds))))


class _RMSNorm(nn.Module):
    """Same arithmetic as the reference's LlamaRMSNorm (modelling_llama.py:55-70)."""

    def __init__(self, H, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(1.0 + 0.1 * torch.randn(H))
        self.variance_epsilon = eps

    def forward(self, x):
        dt = x.dtype
        x = x.to(torch.float32)
        x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * x.to(dt)


class FakeLlava(nn.Module):
    """LlavaLlamaForCausalLM look-alike: embeds ids, splices N_IMG 'image' rows after position 2 (labels and signs
    become -100 there, llava_arch.py:322-323), decoder, lm_head."""
    N_IMG = 3

    def __init__(self, n_tok=50, H=64, V=96):
        super().__init__()
        self.embed = nn.Embedding(n_tok, H)
        self.inner = _Inner(H)
        self.lm_head = nn.Linear(H, V, bias=False)
        self.img_proj = nn.Linear(4, H)

    def get_model(self):
        return self.inner

    def _splice(self, input_ids, attention_mask, labels, images, signs):
        e = self.embed(input_ids)
        S = e.shape[0]
        img = self.img_proj(images.to(e.dtype)).unsqueeze(1).expand(S, self.N_IMG, -1)
        cut = 2
        e = torch.cat([e[:, :cut], img, e[:, cut:]], 1)
        pad = lambda x, v: torch.cat([x[:, :cut], torch.full((S, self.N_IMG), v, dtype=x.dtype, device=x.device), x[:, cut:]], 1)
        labels = pad(labels, IGNORE)
        attention_mask = pad(attention_mask, True)
        signs = None if signs is None else pad(signs, IGNORE)
        return e, attention_mask, labels, signs

    def prepare_inputs_labels_for_multimodal_signed(self, input_ids, position_ids, attention_mask, past_key_values, labels, images, signs):
        e, m, l, s = self._splice(input_ids, attention_mask, labels, images, signs)
        return None, None, m, None, e, l, s

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images):
        e, m, l, _ = self._splice(input_ids, attention_mask, labels, images, None)
        return None, None, m, None, e, l

    def forward(self, inputs_embeds=None, labels=None, attention_mask=None):
        h = self.inner(inputs_embeds=inputs_embeds).last_hidden_state
        return types.SimpleNamespace(logits=self.lm_head(h))            # the reference upcasts itself (.to(float32))


class FakeVila(nn.Module):
    """vila LlavaLlamaModel look-alike: language model under .llm, forward(input_ids, images, labels, attention_mask,
    signs) returning .logits/.labels/.signs (vila/model/language_model/llava_llama.py:99-177)."""

    def __init__(self, base: FakeLlava):
        super().__init__()
        self.base = base
        self.llm = types.SimpleNamespace(lm_head=base.lm_head, get_decoder=lambda: base.inner)
        self.prepare_inputs_labels_for_multimodal_signed = base.prepare_inputs_labels_for_multimodal_signed
        self.prepare_inputs_labels_for_multimodal = base.prepare_inputs_labels_for_multimodal

    def forward(self, input_ids=None, images=None, labels=None, attention_mask=None, signs=None):
        if signs is not None:
            _, _, m, _, e, l, s = self.prepare_inputs_labels_for_multimodal_signed(input_ids, None, attention_mask, None, labels, images, signs)
        else:
            _, _, m, _, e, l = self.prepare_inputs_labels_for_multimodal(input_ids, None, attention_mask, None, labels, images)
            s = None
        out = self.base.forward(inputs_embeds=e)
        out.labels, out.signs = l, s
        return out


def make_inputs(B=2, Tp=12, Tn=14, Tr=11, n_tok=50, V=96, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, hi=n_tok: torch.randint(0, hi, s, generator=g)

    def lab_sign(T, k):
        lab = torch.full((B, T), IGNORE)
        lab[:, 4:] = r(B, T - 4, hi=V)
        sg = torch.zeros(B, T, dtype=torch.int64)
        for b in range(B):
            sg[b, 5:7] = 1
            if k > 1:
                sg[b, 8:10] = 2
        return lab, sg

    lp, sp = lab_sign(Tp, 2)
    ln, sn = lab_sign(Tn, 2)
    lr = torch.full((B, Tr), IGNORE)
    lr[:, 3:] = r(B, Tr - 3, hi=V)
    lr[0, -2:] = IGNORE
    d = dict(images=torch.randn(B, 4, generator=g), ref_images=torch.randn(B, 4, generator=g),
             input_ids=r(B, Tp), labels=lp, attention_mask=torch.ones(B, Tp, dtype=torch.bool), pos_signs=sp,
             neg_input_ids=r(B, Tn), neg_labels=ln, neg_attention_mask=torch.ones(B, Tn, dtype=torch.bool), neg_signs=sn,
             ref_input_ids=r(B, Tr), ref_labels=lr, ref_attention_mask=torch.ones(B, Tr, dtype=torch.bool))
    return {k: v.to(device) for k, v in d.items()}


def _trainer(policy, frozen, alpha=0.4, vila=False):
    t = types.SimpleNamespace(model=policy, ref_model=frozen, loss_alpha=alpha, label_pad_token_id=IGNORE, is_encoder_decoder=False)
    if vila:
        t.loss_holder = {"contrastive_loss": [], "divergence": [], "loss": []}
    return t


def _port_as_op(h_pair, labels_pair, signs_pair, h_ref_pol, labels_ref, W, *, h_ref_frozen=None, W_ref=None, ref_logits=None, alpha, **kw):
    return dpa_port.dpa_loss_torch(h_pair, labels_pair, signs_pair, h_ref_pol, h_ref_frozen, labels_ref, W,
                                   W if W_ref is None else W_ref, alpha, ref_logits=ref_logits)


def _grads(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def _models(seed=0, dtype=torch.float32, device="cpu"):
    torch.manual_seed(seed)
    pol = FakeLlava().to(device=device, dtype=dtype)
    frz = FakeLlava().to(device=device, dtype=dtype)
    frz.load_state_dict(pol.state_dict())
    with torch.no_grad():
        for p in frz.inner.parameters():
            p.add_(0.05 * torch.randn_like(p))
    frz.requires_grad_(False)
    return pol, frz


@pytest.mark.parametrize("wrap", ["plain", "engine+peft"])
def test_shim_matches_lifted_reference_llava(monkeypatch, wrap):
    """Plumbing parity on CPU: reference compute_loss (lifted, unmodified) vs the shim with the loss op = torch port."""
    from oracle.lift_reference import lift_trainer_class, reference_available
    if not reference_available():
        pytest.skip("no /root/reference here")
    from halva_b200 import trainer as shim
    monkeypatch.setattr(shim._dpa_loss_mod, "dpa_loss", _port_as_op)
    pol, frz = _models()
    inputs = make_inputs()
    ref_tr = lift_trainer_class("llava")()
    ref_tr.model, ref_tr.ref_model, ref_tr.loss_alpha, ref_tr.label_pad_token_id, ref_tr.is_encoder_decoder = pol, frz, 0.4, IGNORE, False
    loss_ref = ref_tr.compute_loss(pol, inputs)
    loss_ref.backward()
    g_ref = _grads(pol)
    pol.zero_grad()

    handle = pol
    if wrap == "engine+peft":          # DeepSpeed engine (.module) around a PEFT model (.base_model.model)
        handle = types.SimpleNamespace(module=types.SimpleNamespace(base_model=types.SimpleNamespace(model=pol)))
    loss = shim.fused_compute_loss(_trainer(pol, frz), handle, inputs, num_items_in_batch=7)
    loss.backward()
    g = _grads(pol)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert set(g) == set(g_ref)
    for n in g:
        assert torch.allclose(g[n], g_ref[n], rtol=1e-4, atol=1e-6), n


def test_shim_matches_lifted_reference_vila(monkeypatch):
    from oracle.lift_reference import lift_trainer_class, reference_available
    if not reference_available():
        pytest.skip("no /root/reference here")
    from halva_b200 import trainer as shim
    monkeypatch.setattr(shim._dpa_loss_mod, "dpa_loss", _port_as_op)
    pol_b, frz_b = _models(seed=1)
    pol, frz = FakeVila(pol_b), FakeVila(frz_b)
    inputs = make_inputs(seed=1)
    inputs["ref_images"] = inputs["ref_images"].unsqueeze(1)       # the VILA collator keeps a singleton dim (.squeeze(1) in the reference)
    ref_tr = lift_trainer_class("vila")()
    ref_tr.model, ref_tr.ref_model, ref_tr.loss_alpha, ref_tr.label_pad_token_id, ref_tr.is_encoder_decoder = pol, frz, 0.2, IGNORE, False
    ref_tr.loss_holder = {"contrastive_loss": [], "divergence": [], "loss": []}
    loss_ref = ref_tr.compute_loss(pol, inputs)
    loss_ref.backward()
    g_ref = _grads(pol_b)
    pol_b.zero_grad()
    tr = _trainer(pol, frz, alpha=0.2, vila=True)
    loss = shim.fused_compute_loss(tr, pol, inputs, sync_logging=True)
    loss.backward()
    g = _grads(pol_b)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    for n in g_ref:
        assert torch.allclose(g[n], g_ref[n], rtol=1e-4, atol=1e-6), n
    # the VILA logging keeps working: same three series as the reference appends (vila halva_trainer.py:846-848)
    for k in ("contrastive_loss", "divergence", "loss"):
        assert abs(tr.loss_holder[k][0] - ref_tr.loss_holder[k][0]) <= 2e-6, k


def test_shim_takes_cached_frozen_logits_from_the_batch(monkeypatch):
    """inputs["ref_logits"] (the frozen model's logits on the reference sample, post-splice) replaces the frozen pass:
    same loss and gradients, and the frozen model is never called."""
    from halva_b200 import trainer as shim
    monkeypatch.setattr(shim._dpa_loss_mod, "dpa_loss", _port_as_op)
    pol, frz = _models(seed=2)
    inputs = make_inputs(seed=2)
    base = shim.fused_compute_loss(_trainer(pol, frz), pol, inputs)
    base.backward()
    g0 = _grads(pol)
    pol.zero_grad()
    with torch.no_grad():
        h, _, _ = shim.final_hidden_states(frz, inputs["ref_input_ids"], inputs["ref_images"], inputs["ref_labels"],
                                           inputs["ref_attention_mask"])
        cached = frz.lm_head(h)

    class Untouchable:
        def __getattr__(self, name):
            raise AssertionError("the frozen model must not be used when its logits are given")

    tr = _trainer(pol, frz)
    calls = []
    orig = shim.final_hidden_states
    monkeypatch.setattr(shim, "final_hidden_states", lambda m, *a, **k: (calls.append(m is frz), orig(m, *a, **k))[1])
    loss = shim.fused_compute_loss(tr, pol, dict(inputs, ref_logits=cached))
    loss.backward()
    assert not any(calls), "frozen decoder ran"
    assert abs(loss.item() - base.item()) <= 1e-6 * abs(base.item())
    g = _grads(pol)
    for n in g0:
        assert torch.allclose(g[n], g0[n], rtol=1e-5, atol=1e-7), n
    with pytest.raises(ValueError, match="post-splice"):
        shim.fused_compute_loss(tr, pol, dict(inputs, ref_logits=cached[:, :-1]))


def test_shim_announces_the_gradient_accumulation_scale(monkeypatch):
    """Trainer.training_step / the DeepSpeed engine divide the loss by gradient_accumulation_steps before the backward: the
    shim tells the op, so that the gradients formed during the forward need no rescaling pass.  An explicit value wins."""
    from halva_b200 import trainer as shim
    seen = []

    def op(*a, grad_scale_hint=None, **kw):
        seen.append(grad_scale_hint)
        return _port_as_op(*a, **kw)
    monkeypatch.setattr(shim._dpa_loss_mod, "dpa_loss", op)
    pol, frz = _models()
    inputs = make_inputs()
    tr = _trainer(pol, frz)
    shim.fused_compute_loss(tr, pol, inputs)                                   # no TrainingArguments at all
    tr.args = types.SimpleNamespace(gradient_accumulation_steps=4)             # the reference recipe (src/hallava_7b.sh:22)
    shim.fused_compute_loss(tr, pol, inputs)
    shim.fused_compute_loss(tr, pol, inputs, grad_scale_hint=0.5)
    tr.args.gradient_accumulation_steps = 1
    shim.fused_compute_loss(tr, pol, inputs)
    assert seen == [1.0, 0.25, 0.5, 1.0]


def test_stack_pair_batch_pad_values():
    from halva_b200.trainer import stack_pair_batch
    inputs = make_inputs(Tp=9, Tn=13)
    ids, labels, mask, signs = stack_pair_batch(inputs)
    assert ids.shape == (4, 13) and mask.dtype == torch.bool
    assert (ids[:2, 9:] == 0).all() and (labels[:2, 9:] == IGNORE).all() and (~mask[:2, 9:]).all() and (signs[:2, 9:] == 0).all()
    assert torch.equal(ids[2:], inputs["neg_input_ids"]) and torch.equal(signs[:2, :9], inputs["pos_signs"])


def test_patch_installs_override():
    from halva_b200.trainer import patch_halva_trainer

    class T:
        def compute_loss(self, model, inputs, return_outputs=False):
            return "reference"

    orig = patch_halva_trainer(T)
    assert T._reference_compute_loss is orig and T.compute_loss is not orig
    import inspect
    assert "num_items_in_batch" in inspect.signature(T.compute_loss).parameters


@pytest.mark.gpu
def test_shim_cuda_fused_final_norm():
    """fuse_final_norm: decoder run without its RMSNorm, the op applies it to the contributing rows; same loss and
    parameter gradients (incl. the norm weight) as the unfused shim."""
    from halva_b200 import trainer as shim
    pol, frz = _models(seed=4, dtype=torch.bfloat16, device="cuda")
    for m in (pol, frz):
        m.inner.norm = _RMSNorm(64).to(device="cuda", dtype=torch.bfloat16)
    frz.inner.norm.load_state_dict(pol.inner.norm.state_dict())
    frz.requires_grad_(False)
    inputs = make_inputs(B=3, Tp=40, Tn=37, Tr=33, seed=4, device="cuda")
    tr = _trainer(pol, frz)
    la = shim.fused_compute_loss(tr, pol, inputs)
    la.backward()
    ga = _grads(pol)
    pol.zero_grad()
    lb = shim.fused_compute_loss(tr, pol, inputs, fuse_final_norm=True)
    lb.backward()
    gb = _grads(pol)
    assert isinstance(pol.inner.norm, _RMSNorm)                      # restored after the call
    assert abs(la.item() - lb.item()) <= 1e-3 * abs(la.item())
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
    assert set(ga) == set(gb)
    for n in ga:
        assert rel(gb[n], ga[n]) <= 3e-2, (n, rel(gb[n], ga[n]))


@pytest.mark.gpu
@pytest.mark.parametrize("vila", [False, True])
def test_shim_cuda_vs_port_bf16(vila):
    """The real CUDA op behind the shim vs the reference op chain (torch port) on the same bf16 stand-in model."""
    from halva_b200 import trainer as shim
    pol_b, frz_b = _models(seed=2, dtype=torch.bfloat16, device="cuda")
    pol, frz = (FakeVila(pol_b), FakeVila(frz_b)) if vila else (pol_b, frz_b)
    inputs = make_inputs(B=3, Tp=40, Tn=37, Tr=33, seed=2, device="cuda")
    tr = _trainer(pol, frz, alpha=0.4, vila=vila)
    loss, extra = shim.fused_compute_loss(tr, pol, inputs, return_outputs=True)
    loss.backward()
    g = _grads(pol_b)
    pol_b.zero_grad()
    # reference op chain on the same hidden states
    ids, labels, mask, signs = shim.stack_pair_batch(inputs)
    h_pair, lp, sp = shim.final_hidden_states(pol, ids, torch.cat([inputs["images"]] * 2), labels, mask, signs)
    h_rp, lr, _ = shim.final_hidden_states(pol, inputs["ref_input_ids"], inputs["ref_images"], inputs["ref_labels"], inputs["ref_attention_mask"])
    with torch.no_grad():
        h_rf, _, _ = shim.final_hidden_states(frz, inputs["ref_input_ids"], inputs["ref_images"], inputs["ref_labels"], inputs["ref_attention_mask"])
    l2, a2, d2 = dpa_port.dpa_loss_torch(h_pair, lp, sp, h_rp, h_rf, lr, pol_b.lm_head.weight, frz_b.lm_head.weight, 0.4)
    l2.backward()
    g2 = _grads(pol_b)
    assert abs(loss.item() - l2.item()) <= 2e-3 * abs(l2.item())
    assert abs(extra["align"].item() - a2.item()) <= 2e-3 * abs(a2.item())
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
    for n in g2:
        assert rel(g[n], g2[n]) <= 3e-2, (n, rel(g[n], g2[n]))      # bf16 autograd chain through the stand-in decoder
    if vila:
        assert len(tr.loss_holder["loss"]) == 1 and torch.is_tensor(tr.loss_holder["loss"][0])


def test_zero3_partitioned_lm_head_is_gathered(monkeypatch):
    """Under DeepSpeed ZeRO-3 lm_head.weight is partitioned (carries `ds_id`) and only lm_head.forward's hooks re-assemble
    it; the shim reads the weight directly, so it must wrap the op in deepspeed.zero.GatheredParameters
    (the reference's own pattern, llava/train/train_halva.py:102-113)."""
    import contextlib
    import sys
    from halva_b200 import trainer as shim
    seen = []

    @contextlib.contextmanager
    def gathered(params, modifier_rank=None):
        seen.append([id(p) for p in params])
        yield

    fake = types.ModuleType("deepspeed")
    fake.zero = types.SimpleNamespace(GatheredParameters=gathered)
    monkeypatch.setitem(sys.modules, "deepspeed", fake)
    monkeypatch.setattr(shim._dpa_loss_mod, "dpa_loss", _port_as_op)
    pol, frz = _models(seed=3)
    pol.lm_head.weight.ds_id = 7                       # what ZeRO-3 stamps on a partitioned parameter
    loss = shim.fused_compute_loss(_trainer(pol, frz), pol, make_inputs(seed=3))
    assert torch.isfinite(loss) and seen == [[id(pol.lm_head.weight)]]
    seen.clear()
    del pol.lm_head.weight.ds_id
    shim.fused_compute_loss(_trainer(pol, frz), pol, make_inputs(seed=3))
    assert seen == []                                  # nothing to gather outside ZeRO-3
