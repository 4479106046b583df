#!/usr/bin/env python
"""bench.py -- DPA loss forward+backward throughput on synthetic LLaVA-1.5-7B lm_head shapes.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference op chain on the host CPU (oracle port)

One "step" = one dpa_loss forward + backward over one batch of BASELINE.json configs[1]
(16 sample pairs + 16 reference samples, T 2048, H 4096, V 32000, bf16; per GPU under data parallelism).
Prints ONE JSON line (rank 0).  Besides the headline (`value`, `e2e`, `roofline`, `cpu_baseline`) the line carries
`variants`: the same batch under the dense convention, the reference's op chain on the same GPU, the fp32 cfg-1 case,
and at N > 1 the vocab-parallel cfg-3 run, a multi-GPU parity self-check, the strong-scaling row and cfg-4.
See DESIGN.md "Measurement" for the definitions used here.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DPA loss fwd+bwd tokens/s (7B lm_head shape)"
UNIT = "tokens/s"
PLANT = 5.0     # synthetic hidden states give their targets a realistic (high) probability, see halva_b200/synth.py


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"), hbm_gbs=d["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


def load_traffic(kernel_class):
    """DRAM bytes (read + write) of the largest launch of `kernel_class` from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "kernel_metrics.json")
    if not os.path.exists(p):
        return None, "no ncu capture committed"
    d = json.load(open(p))
    k = d.get("kernels", {}).get(kernel_class)
    if not k:
        return None, f"no {kernel_class} launch in {d.get('source')}"
    return k["dram_bytes"], f"largest {kernel_class} launch in {str(d.get('source'))[:60]}"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[], note="no NVML samples")
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2], sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    power_w_max=round(max(self.power), 1) if self.power else None, samples=len(s))


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_sample_inputs(config="cfg2"):
    """Bounded sample of the benchmarked workload for the host-CPU arms: ONE sample pair + ONE reference sample of the
    configuration's own shape (cfg-2: T 2048, H 4096, V 32000) -- 1/16 of the batch a GPU processes per step -- in the
    trainer's precision (bf16 weights / hidden states, fp32 softmax maths, exactly the reference's op chain)."""
    import torch
    from halva_b200.synth import CONFIGS, make_inputs
    shape = dataclasses.replace(CONFIGS[config], B=1, B_ref=1)
    # the CPU's bf16 matmul is the reference's own precision; hosts without AMX / AVX512-BF16 emulate it far slower than
    # fp32, and then the fp32 op chain on the same bf16-valued inputs is the fair (faster) CPU baseline
    a, b = torch.randn(256, 1024), torch.randn(2048, 1024)
    t = {}
    for dt in (torch.float32, torch.bfloat16):
        x, w = a.to(dt), b.to(dt)
        torch.nn.functional.linear(x, w)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.linear(x, w)
        t[dt] = time.perf_counter() - t0
    dtype = torch.bfloat16 if t[torch.bfloat16] <= 3 * t[torch.float32] else torch.float32
    inp = make_inputs(shape, seed=1234, dtype=torch.bfloat16, device="cpu", plant=PLANT)
    if dtype == torch.float32:
        inp = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in inp.items()}
        inp["W_ref"] = inp["W"]
    return shape, inp, ("bf16 (fp32 softmax maths, as the trainer runs it)" if dtype == torch.bfloat16
                        else "fp32 op chain on the bf16-valued inputs (this host emulates bf16 matmuls)")


def time_cpu_port(steps, warmup, config="cfg2"):
    import torch
    from oracle import dpa_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape, inp, dt = cpu_sample_inputs(config)
    for _ in range(warmup):
        dpa_port.step_cpu(inp)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        dpa_port.step_cpu(inp)
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return dict(value=shape.tokens * steps / total, ms_per_step=1e3 * total / steps, cores=cores, tokens_per_step=shape.tokens, dtype=dt,
                sample=f"1 of the {config} batch's sample pairs + 1 reference sample per step (T={shape.T}, H={shape.H}, V={shape.V}, "
                       f"{shape.tokens} tokens), {dt}, lm_head trainable, {cores} host threads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_port(args.steps, max(args.warmup, 1), args.config)
    cfg = config_dict(args, None)       # the workload this arm samples: identical to the GPU arm's config
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16" if r["dtype"].startswith("bf16") else "f32", data="synthetic", impl="reference",
                config=cfg, tokens_per_step=r["tokens_per_step"], sample=r["sample"],
                where=f"host CPU, {r['cores']} threads, one process (rank 0); a step = a 1/16 slice of the workload's batch",
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def config_dict(args, shape):
    from halva_b200.synth import CONFIGS
    shape = shape or CONFIGS[args.config]
    return dict(workload=f"HALVA-7B DPA loss at training shape ({args.config}): {shape.B} pairs + {shape.B_ref} reference samples per GPU, "
                         f"T={shape.T} incl. {shape.n_img} image tokens, H={shape.H}, V={shape.V}, {'fp32' if getattr(args, 'fp32', False) else 'bf16'}, phrase density {shape.density:.2f}, "
                         f"seeded synthetic hidden states (targets planted, gain {PLANT})",
                parallelism=f"{args.parallelism}{args.gpus}", need_dW=not args.no_dw, dense=bool(args.dense),
                backward="recompute" if args.backward == "recompute" else "fused (gradients formed in the forward pass)",
                l2="inputs (1.3 GB hidden states + 262 MB lm_head) exceed the 126 MB L2; no explicit flush")


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Process-wide handles of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            tt = self.torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
            ms = tt.item()
        return ms

    def time_steps(self, fn, steps, warm=3):
        """ms per step of `fn`: `warm` untimed calls, then `steps` calls bracketed by barrier + synchronize, CUDA events
        on the current stream, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        self.e0.record()
        for _ in range(steps):
            fn()
        self.e1.record()
        self.barrier()
        return self.max_over_ranks(self.e0.elapsed_time(self.e1)) / steps


def make_step(t, shape, W, need_dw, world_avg=False, **kw):
    """One forward + backward of dpa_loss over the tensors of `t` (a halva_b200.synth.make_inputs dict)."""
    from halva_b200 import dpa_loss
    from halva_b200.dist import average_gradients
    hp, hr = t["h_pair"], t["h_ref_pol"]

    def step(src=None):
        s = src or t
        s["h_pair"].grad = None
        s["h_ref_pol"].grad = None
        W.grad = None
        loss, align, div = dpa_loss(s["h_pair"], s["labels_pair"], s["signs_pair"], s["h_ref_pol"], s["labels_ref"], W,
                                    h_ref_frozen=s["h_ref_frz"], alpha=shape.alpha, **kw)
        loss.backward()
        if world_avg and need_dw:          # data parallel: average the lm_head gradient over ranks (NCCL over NVLink)
            average_gradients([W.grad])
        return loss, align, div

    hp.requires_grad_(True)
    hr.requires_grad_(True)
    return step


def run_e2e(cx, args, tensors, step, dense_copy):
    """Same step driven from HOST buffers.  Every step: (1) the span plan is rebuilt from THAT step's host labels and
    (2) the pinned hidden states / labels / phrase ids are copied to the device -- both inside the timed region
    (halva_b200.host.HostStager -> libhalva_dpa.so dpa_stage_host_batch: labels in full, hidden rows span-wise, or every
    row with dense_copy) -- double-buffered against the previous step's kernels on a copy stream; the three losses of
    every step are read back on the host, one step late (while the next step's kernels run)."""
    torch = cx.torch
    from halva_b200.host import HostStager
    keys = ("h_pair", "h_ref_pol", "h_ref_frz", "labels_pair", "signs_pair", "labels_ref")
    host = {k: tensors[k].detach().cpu().pin_memory() for k in keys}
    stager = HostStager(dense=dense_copy)
    sets = []
    for _ in range(2):
        s = {k: torch.zeros_like(tensors[k].detach()) for k in keys}
        s["h_pair"].requires_grad_(True)
        s["h_ref_pol"].requires_grad_(True)
        sets.append(s)
    out_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=cx.dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    h2d, plan_us, issue_us, read = [0], [], [], []

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        for i in range(n + 2):
            if i < n:
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_done[b])           # buffer set b free again (step i-2 finished)
                    h2d[0] = stager.stage(host, {k: v.detach() for k, v in sets[b].items()})   # plan + copies, this batch
                    plan_us.append(stager.info.plan_us)
                    issue_us.append(stager.info.issue_us)
                    ev_copied[b].record(copy_stream)
            if 0 < i <= n:
                b = (i - 1) % 2
                cur.wait_event(ev_copied[b])
                l, al, dv = step(sets[b])
                ev_done[b].record(cur)
                out_host[b][:3].copy_(torch.stack([l.detach(), al, dv]), non_blocking=True)
                ev_out[b].record(cur)
            if i > 1:
                b = (i - 2) % 2
                ev_out[b].synchronize()                           # the result of step i-2 is on the host: read it
                read.append(float(out_host[b][0]))

    for b in range(2):
        ev_done[b].record(torch.cuda.current_stream())
    e2e_loop(2)
    del plan_us[:], issue_us[:]
    cx.barrier()
    cx.e0.record()
    e2e_loop(args.steps)
    cx.e1.record()
    cx.barrier()
    ms = cx.max_over_ranks(cx.e0.elapsed_time(cx.e1))
    assert len(read) == 2 + args.steps, len(read)                 # every step's loss was read on the host
    return dict(ms_per_step=ms / args.steps, h2d_bytes=h2d[0], loss=read[-1], plan_ms=sum(plan_us) / len(plan_us) / 1e3,
                issue_ms=sum(issue_us) / len(issue_us) / 1e3, spans=stager.info.spans_pair + stager.info.spans_ref)


def tc_frac(flops, ms, peaks):
    return round(flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"], 4)


def variant_gpu_reference_chain(cx, t, shape, W, steps):
    """SURVEY 2b bar (i): the reference's ATen op sequence k1-k7 (llava/train/halva_trainer.py:392-592, restated in
    oracle/dpa_port.py) forward + autograd backward on the SAME GPU, same batch, bf16 -- the comparator, not the product."""
    torch = cx.torch
    from oracle import dpa_port
    hp, hr = t["h_pair"], t["h_ref_pol"]
    out = {}

    def ref_step():
        hp.grad = hr.grad = W.grad = None
        loss, _, _ = dpa_port.dpa_loss_torch(hp, t["labels_pair"], t["signs_pair"], hr, t["h_ref_frz"], t["labels_ref"], W, W.detach(),
                                             shape.alpha)
        loss.backward()
        out["loss"] = loss.detach()

    try:
        ms = cx.time_steps(ref_step, max(2, min(steps, 5)), warm=1)
        res = dict(ms_per_step=round(ms, 3), value=round(shape.tokens / (ms * 1e-3)), loss=out["loss"].item())
    except torch.cuda.OutOfMemoryError:
        res = dict(error="OOM")
    hp.grad = hr.grad = W.grad = None
    out.clear()
    torch.cuda.empty_cache()
    return res


def variant_fp32_cfg1(cx, steps, peaks):
    """BASELINE configs[0] on the GPU: fp32 hidden states / weights / gradients (split-bf16 x6 tensor-core passes)."""
    torch = cx.torch
    from halva_b200.synth import CONFIGS, algorithmic_flops, make_inputs
    shape = CONFIGS["cfg1"]
    t = make_inputs(shape, seed=1234, device=cx.dev, plant=PLANT, dtype=torch.float32)
    W = t["W"].requires_grad_(True)
    step = make_step(t, shape, W, True)
    ms = cx.time_steps(lambda: step(), steps)
    loss = step()[0].item()
    return dict(ms_per_step=round(ms, 4), value=round(shape.tokens / (ms * 1e-3)), loss=loss,
                fp32_equiv_tflops_dense=round(algorithmic_flops(shape, True) / (ms * 1e-3) / 1e12, 1))


def variant_vp(cx, cfg, steps, peaks, args):
    """BASELINE configs[2]: the SAME batch on every rank, lm_head rows split over the ranks (vocab-parallel), the library's
    own communicator exchanging inside the two C calls.  tokens/s counts the batch once (strong scaling over the vocabulary)."""
    torch, dist = cx.torch, cx.dist
    from halva_b200 import read_status
    from halva_b200.dist import init_library_comm, vocab_shard_bounds
    from halva_b200.synth import CONFIGS, algorithmic_flops, make_inputs
    shape = CONFIGS[cfg]
    t = make_inputs(shape, seed=1234, device=cx.dev, plant=PLANT)
    v0, v1 = vocab_shard_bounds(shape.V, cx.world, cx.rank)
    W = t["W"][v0:v1].contiguous().requires_grad_(True)
    if not args.vp_torch_exchange:
        init_library_comm(dist.group.WORLD)
    step = make_step(t, shape, W, True, vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V, backward=args.vp_backward)
    loss, _, _ = step()
    st = read_status(loss)
    ms = cx.time_steps(lambda: step(), steps)
    f_act = algorithmic_flops(shape, True, st["n_pair_active"], st["n_ref_active"]) / cx.world
    res = dict(config=cfg, ms_per_step=round(ms, 4), value=round(shape.tokens / (ms * 1e-3)), loss=step()[0].item(),
               tc_frac_active=tc_frac(f_act, ms, peaks),
               rows=[st["n_pair_active"], st["n_ref_active"]], v_local=v1 - v0,
               exchange="torch.distributed around the C calls" if args.vp_torch_exchange else "library communicator inside the C calls")
    del t, W, step
    torch.cuda.empty_cache()
    return res


def variant_parity(cx):
    """Multi-GPU parity self-check on hardware: (a) vocab-parallel == this rank's single-GPU result on the same rows
    (loss, dH, its dW rows); (b) data parallel: the averaged dW == the mean of the per-shard single-GPU dW."""
    torch, dist = cx.torch, cx.dist
    from halva_b200 import dpa_loss
    from halva_b200.dist import average_gradients, dp_shard, vocab_shard_bounds
    from halva_b200.synth import CONFIGS, make_inputs
    shape = dataclasses.replace(CONFIGS["cfg3"], B=cx.world, B_ref=cx.world, T=512)
    shape.T_ref = 512
    t = make_inputs(shape, seed=77, device=cx.dev, plant=PLANT, shared_head=False)      # identical on every rank
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))

    def run(W, W_ref, pairs=None, **kw):
        if pairs is None:
            sp = sr = slice(None)
        else:
            sp = torch.tensor(list(pairs) + [shape.B + i for i in pairs], device=cx.dev)
            sr = torch.tensor(list(pairs), device=cx.dev)
        hp = t["h_pair"][sp].clone().requires_grad_(True)
        hr = t["h_ref_pol"][sr].clone().requires_grad_(True)
        Wl = W.clone().requires_grad_(True)
        loss, _, div = dpa_loss(hp, t["labels_pair"][sp], t["signs_pair"][sp], hr, t["labels_ref"][sr], Wl,
                                h_ref_frozen=t["h_ref_frz"][sr], W_ref=W_ref, alpha=shape.alpha, **kw)
        loss.backward()
        return loss.item(), div.item(), hp.grad, hr.grad, Wl.grad

    single = run(t["W"], t["W_ref"])
    v0, v1 = vocab_shard_bounds(shape.V, cx.world, cx.rank)
    vp = run(t["W"][v0:v1].contiguous(), t["W_ref"][v0:v1].contiguous(), vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V)
    res = dict(shape=f"{shape.B} pairs, T={shape.T}, H={shape.H}, V={shape.V}, W_ref != W",
               vp_loss_rel=abs(vp[0] - single[0]) / abs(single[0]), vp_div_rel=abs(vp[1] - single[1]) / max(abs(single[1]), 1e-30),
               vp_dH_pair_rel=rel(vp[2], single[2]), vp_dH_ref_rel=rel(vp[3], single[3]), vp_dW_rel=rel(vp[4], single[4][v0:v1]))
    mine = list(dp_shard(shape.B, cx.world, cx.rank))
    loc = run(t["W"], t["W_ref"], mine)
    dW = loc[4].clone()
    average_gradients([dW])
    want = sum(run(t["W"], t["W_ref"], list(dp_shard(shape.B, cx.world, r)))[4].float() for r in range(cx.world)) / cx.world
    res["dp_dW_rel"] = rel(dW, want)
    flags = torch.tensor([res["vp_loss_rel"], res["vp_div_rel"], res["vp_dH_pair_rel"], res["vp_dH_ref_rel"], res["vp_dW_rel"],
                          res["dp_dW_rel"]], device=cx.dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)                 # worst rank
    worst = flags.tolist()
    tol = [1e-5, 2e-4, 1e-2, 1e-2, 1e-2, 1e-2]                    # losses / gradients (bf16: north star 1e-3 / 1e-2)
    res = {k: float(f"{w:.3g}") for k, w in zip(("vp_loss_rel", "vp_div_rel", "vp_dH_pair_rel", "vp_dH_ref_rel", "vp_dW_rel", "dp_dW_rel"), worst)}
    res.update(shape=f"{shape.B} pairs, T={shape.T}, H={shape.H}, V={shape.V}, W_ref != W", tol=tol, ok=all(w <= x for w, x in zip(worst, tol)),
               note="max over ranks; vp vs the single-GPU result on the same rows, dp dW vs the mean of per-shard dW")
    del t
    torch.cuda.empty_cache()
    return res


def variant_strong(cx, steps, ms_one_gpu):
    """SURVEY 8(d) secondary row: the FIXED 16-pair cfg-2 batch (the one rank 0 runs alone in the headline) split over the
    N ranks by sample pairs, dW all-reduced; tokens/s counts the batch once."""
    torch = cx.torch
    from halva_b200.dist import dp_shard
    from halva_b200.synth import CONFIGS, make_inputs
    shape = CONFIGS["cfg2"]
    mine = list(dp_shard(shape.B, cx.world, cx.rank))
    if not mine:
        return dict(error=f"{shape.B} pairs cannot be split over {cx.world} ranks")
    t = make_inputs(shape, seed=1234, device=cx.dev, plant=PLANT)
    sp = torch.tensor(mine + [shape.B + i for i in mine], device=cx.dev)
    sr = torch.tensor(mine, device=cx.dev)
    sub = dict(h_pair=t["h_pair"][sp].contiguous(), labels_pair=t["labels_pair"][sp].contiguous(), signs_pair=t["signs_pair"][sp].contiguous(),
               h_ref_pol=t["h_ref_pol"][sr].contiguous(), h_ref_frz=t["h_ref_frz"][sr].contiguous(), labels_ref=t["labels_ref"][sr].contiguous())
    W = t["W"].requires_grad_(True)
    del t
    step = make_step(sub, shape, W, True, world_avg=True)
    ms = cx.time_steps(lambda: step(), steps)
    res = dict(ms_per_step=round(ms, 4), value=round(shape.tokens / (ms * 1e-3)), pairs_per_gpu=len(mine),
               speedup_vs_1gpu_same_batch=round(ms_one_gpu / ms, 3), efficiency=round(ms_one_gpu / ms / cx.world, 3))
    del sub, W, step
    torch.cuda.empty_cache()
    return res


def variant_dp_config(cx, cfg, steps, peaks):
    """Another BASELINE config data-parallel (weak scaling: its full batch on every rank), e.g. configs[3] = cfg4."""
    torch = cx.torch
    from halva_b200 import read_status
    from halva_b200.synth import CONFIGS, algorithmic_flops, make_inputs
    shape = CONFIGS[cfg]
    t = make_inputs(shape, seed=1234 + cx.rank, device=cx.dev, plant=PLANT)
    W = t["W"].requires_grad_(True)
    step = make_step(t, shape, W, True, world_avg=cx.world > 1)
    st = read_status(step()[0])
    ms = cx.time_steps(lambda: step(), steps)
    res = dict(config=cfg, ms_per_step=round(ms, 4), value=round(shape.tokens * cx.world / (ms * 1e-3)), loss=step()[0].item(),
               tc_frac_active=tc_frac(algorithmic_flops(shape, True, st["n_pair_active"], st["n_ref_active"]), ms, peaks),
               rows=[st["n_pair_active"], st["n_ref_active"]])
    del t, W, step
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    from halva_b200 import _lib, dpa_loss, read_status
    from halva_b200.synth import CONFIGS, algorithmic_flops, make_inputs

    cx = Ctx()
    dist, dev, world, rank = cx.dist, cx.dev, cx.world, cx.rank
    need_dw = not args.no_dw
    shape = CONFIGS[args.config]
    if args.density is not None:
        shape = dataclasses.replace(shape, density=args.density)
    if args.pairs is not None:
        shape = dataclasses.replace(shape, B=args.pairs, B_ref=args.pairs)
    vp = args.parallelism == "vp" and world > 1
    # dp: every rank its own batch (weak scaling);  vp: the SAME rows on every rank, lm_head rows split over the ranks
    if args.fp32:
        args.no_e2e, args.no_variants = True, True
    t = make_inputs(shape, seed=1234 + (0 if vp else rank), device=dev, plant=PLANT,
                    dtype=torch.float32 if args.fp32 else torch.bfloat16)
    vp_kw = {}
    if vp:
        from halva_b200.dist import vocab_shard_bounds
        v0, v1 = vocab_shard_bounds(shape.V, world, rank)
        t["W"] = t["W"][v0:v1].contiguous()
        vp_kw = dict(vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V)
        if not args.vp_torch_exchange:      # the library's own NCCL communicator: both exchange steps inside the two C calls
            from halva_b200.dist import init_library_comm
            init_library_comm(dist.group.WORLD)
        args.backward = args.vp_backward
    W = t["W"].requires_grad_(need_dw)
    kw = dict(dense=args.dense, chunk_rows=args.chunk_rows, backward=args.backward, **vp_kw)
    step = make_step(t, shape, W, need_dw, world_avg=world > 1 and not vp, **kw)
    tensors = t

    loss, _, _ = step()
    st = read_status(loss)
    assert st["status"] == 0, st
    rows_pair, rows_ref = st["n_pair_active"], st["n_ref_active"]
    for _ in range(max(args.warmup, 3) - 1):
        step()
    cx.barrier()

    sampler = ClockSampler(cx.local) if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER") else None
    if sampler:
        sampler.start()
    _lib.profile_enable(True)
    _lib.profile_collect()
    cx.barrier()
    cx.e0.record()
    for _ in range(args.steps):
        loss, _, _ = step()
    cx.e1.record()
    cx.barrier()
    ms = cx.max_over_ranks(cx.e0.elapsed_time(cx.e1))
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    loss_val = loss.item()
    ms_eager = ms
    K = args.steps

    # ---- the same K steps replayed from a CUDA graph: the path is stream-ordered with no host sync (row counts live on
    #      the device), so one capture removes the launch gaps of a step.  Falls back to the eager timing on any error.
    graph_info = dict(used=False)
    del loss                      # an alive loss keeps the leaves' AccumulateGrad nodes bound to the default stream,
    import gc                     # which a capture on a side stream must not touch
    gc.collect()
    if args.graph and world == 1:      # opt-in, single process only: NCCL collectives must not be captured here
        try:
            cap_stream = torch.cuda.Stream(device=dev)
            cap_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap_stream):
                step()
            torch.cuda.current_stream().wait_stream(cap_stream)
            cx.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss_g, _, _ = step()
            ms_g = cx.time_steps(g.replay, K, warm=2) * K
            lg = loss_g.item()
            if abs(lg - loss_val) > 1e-6 * abs(loss_val):
                raise RuntimeError(f"graph replay loss {lg} != eager loss {loss_val}")
            graph_info = dict(used=True, ms_per_step_eager=ms_eager / K, ms_per_step_graph=ms_g / K)
            ms = ms_g
        except Exception as e:  # noqa: BLE001
            graph_info = dict(used=False, error=repr(e)[:200])
            if os.environ.get("BENCH_DEBUG"):
                import traceback
                traceback.print_exc()
            torch.cuda.synchronize()

    # ---- end to end: host (pinned) buffers in, losses out, per-step planning + H2D double-buffered against the previous
    #      step's compute.  Measured BEFORE the variants: every further minute under load leaves the GPU in a hotter / lower
    #      power state, and the end-to-end number is the headline -- it should see the device state of the timing above.
    e2e_span = e2e_full = None
    if not args.no_e2e:
        e2e_span = run_e2e(cx, args, tensors, step, dense_copy=bool(args.dense or args.e2e_full_copy))
        if not (args.dense or args.e2e_full_copy) and not args.no_variants:
            e2e_full = run_e2e(cx, args, tensors, step, dense_copy=True)
    clocks = sampler.finish() if sampler else None

    peaks = load_peaks()
    variants = {}
    do_var = not args.no_variants and not vp
    hp, hr = t["h_pair"], t["h_ref_pol"]
    tokens_step = shape.tokens * (1 if vp else world)
    tps = lambda ms_: round(tokens_step / (ms_ * 1e-3))

    diag = do_var and world == 1           # single-GPU diagnostics; at N > 1 the line's tail belongs to the multi-GPU variants
    if diag and need_dw:
        # lm_head frozen (the shipped LoRA recipe, /root/reference/llava/train/train_halva.py:167-168): no dW
        Wf = W.detach()
        step_f = make_step(t, shape, Wf, False, **kw)
        ms_v = cx.time_steps(lambda: step_f(), K)
        variants["lm_head_frozen"] = dict(ms_per_step=round(ms_v, 4), value=tps(ms_v), note="LoRA recipe: no dW GEMM")
    if diag and args.backward != "recompute":
        # the frozen model's logits GIVEN (dpa_loss(ref_logits=...)): 3 instead of 4 lm_head GEMMs per reference-sample row
        with torch.no_grad():
            zr = torch.empty(shape.B_ref, shape.T_ref, shape.V, dtype=torch.bfloat16, device=dev)
            for b0 in range(shape.B_ref):
                torch.matmul(t["h_ref_frz"][b0], t["W"].detach().t(), out=zr[b0])
        out = {}

        def step_given():
            hp.grad = hr.grad = W.grad = None
            l, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, ref_logits=zr, alpha=shape.alpha,
                               dense=args.dense, chunk_rows=args.chunk_rows)
            l.backward()
            out["l"] = l.detach()
        ms_v = cx.time_steps(step_given, K)
        variants["ref_logits_given"] = dict(ms_per_step=round(ms_v, 4), value=tps(ms_v), loss=out["l"].item(),
                                            note="frozen logits handed over (cached): 3 instead of 4 lm_head GEMMs per reference row")
        del zr, out
    if diag and shape.B % 4 == 0 and shape.B_ref == shape.B and shape.B >= 8:
        # the reference recipe's form of the same 16 pairs: 4 micro-batches of 4 pairs, gradients accumulated
        # (per_device_train_batch_size 4 x gradient_accumulation_steps 4, /root/reference/src/hallava_7b.sh:21-22)
        mb = shape.B // 4
        micro = []
        for i in range(4):
            r_ = torch.arange(i * mb, (i + 1) * mb, device=dev)
            pr = torch.cat([r_, shape.B + r_])
            micro.append(dict(h_pair=t["h_pair"].detach()[pr].clone().requires_grad_(True), labels_pair=t["labels_pair"][pr].contiguous(),
                              signs_pair=t["signs_pair"][pr].contiguous(), h_ref_pol=t["h_ref_pol"].detach()[r_].clone().requires_grad_(True),
                              h_ref_frz=t["h_ref_frz"][r_].contiguous(), labels_ref=t["labels_ref"][r_].contiguous()))

        def step_micro():
            W.grad = None
            for m in micro:
                m["h_pair"].grad = None
                m["h_ref_pol"].grad = None
                l, _, _ = dpa_loss(m["h_pair"], m["labels_pair"], m["signs_pair"], m["h_ref_pol"], m["labels_ref"], W,
                                   h_ref_frozen=m["h_ref_frz"], alpha=shape.alpha, dense=args.dense, chunk_rows=args.chunk_rows,
                                   backward=args.backward, grad_scale_hint=0.25)
                (l / 4).backward()                       # W.grad accumulates over the micro-batches, as under grad accumulation
        ms_v = cx.time_steps(step_micro, K)
        variants["micro_4x4"] = dict(ms_per_step=round(ms_v, 4), value=tps(ms_v), note="4 micro-batches of B/4 pairs, dW accumulated (the recipe's 4 x 4); loss / 4 before each backward, announced as grad_scale_hint")
        del micro
    if do_var and not args.dense:
        # dense convention (SURVEY 8d primary figure for cfg-1..4): every fed position through every GEMM, no row compaction
        step_d = make_step(t, shape, W, need_dw, world_avg=world > 1, dense=True, chunk_rows=args.chunk_rows, backward=args.backward)
        ms_v = cx.time_steps(lambda: step_d(), min(K, 10), warm=2)
        variants["dense"] = dict(ms_per_step=round(ms_v, 3), value=tps(ms_v), loss=step_d()[0].item(),
                                 step_tc_frac_dense=tc_frac(algorithmic_flops(shape, need_dw), ms_v, peaks),
                                 alg_tflop=round(algorithmic_flops(shape, need_dw) / 1e12, 2),
                                 note="DPA_FLAG_DENSE: what the reference executes (all 98304 positions)")
    if do_var and world == 1 and need_dw and not args.dense:
        r = variant_gpu_reference_chain(cx, t, shape, W, K)
        if "ms_per_step" in r:
            r["speedup_ours"] = round(r["ms_per_step"] / (ms_eager / K), 2)
            r["loss_rel_diff"] = float(f"{abs(r['loss'] - loss_val) / abs(r['loss']):.2g}")
        r["note"] = "reference ATen op chain k1-k7 (halva_trainer.py:392-592 via oracle/dpa_port.py) fwd+bwd, same GPU, same batch, bf16"
        variants["gpu_reference_chain"] = r
    if do_var and world == 1:
        variants["fp32_cfg1"] = variant_fp32_cfg1(cx, K, peaks)
    if do_var and world > 1:
        def guarded(name, fn):
            # a variant that fails the same way on every rank (e.g. peer memory unavailable) must not cost the whole line
            try:
                variants[name] = fn()
            except Exception as e:  # noqa: BLE001
                variants[name] = dict(error=repr(e)[:300])
                torch.cuda.synchronize()
        # rank 0's headline batch IS the fixed batch (seed 1234): its single-GPU time is this run's ms per step
        guarded("strong", lambda: variant_strong(cx, K, ms_eager / K))
        guarded("vp_cfg3", lambda: variant_vp(cx, "cfg3", min(K, 10), peaks, args))
        if "error" in variants["vp_cfg3"] and not args.vp_torch_exchange:
            # the library-side exchange is unavailable here: the portable form (torch.distributed around the C calls)
            from halva_b200.dist import destroy_library_comm
            try:
                destroy_library_comm()
            except Exception:  # noqa: BLE001
                pass
            args.vp_torch_exchange = True
            err = variants["vp_cfg3"]["error"]
            guarded("vp_cfg3", lambda: variant_vp(cx, "cfg3", min(K, 10), peaks, args))
            variants["vp_cfg3"]["library_exchange_error"] = err
        guarded("multi_gpu_parity", lambda: variant_parity(cx))
        guarded("dp_cfg4", lambda: variant_dp_config(cx, "cfg4", min(K, 10), peaks))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms / K
    value = tokens_step / (ms_step * 1e-3)
    unit_flops = 2.0 * shape.H * W.shape[0]          # per GPU: this rank's lm_head rows
    if args.fp32:
        unit_flops *= 6          # executed bf16 tensor-core FLOPs: six split-operand K segments per GEMM
    # executed tensor-core FLOPs of each GEMM class per step (rows actually processed, tile padding not counted)
    n_p, n_r = (2 * shape.B * (shape.T - 1), shape.B_ref * (shape.T_ref - 1)) if args.dense else (rows_pair, rows_ref)
    fused = args.backward != "recompute"
    # GEMM classes: fwd_pair/dz_pair = one [rows x H].[V x H]^T each; fwd_ref/dz_ref/stash_ref = policy + frozen logits of the
    # reference-sample rows (two GEMMs' worth per launch); dH = dz.W; dW = dz^T.H
    exec_flops = dict(fwd_pair=n_p * unit_flops, dz_pair=n_p * unit_flops, dH=(n_p + n_r) * unit_flops,
                      dW=(n_p + n_r) * unit_flops if need_dw else 0.0)
    if fused:
        exec_flops.update(stash_ref=2 * n_r * unit_flops)
    else:
        exec_flops.update(fwd_ref=2 * n_r * unit_flops, dz_ref=2 * n_r * unit_flops)
    kernels = {}
    for k, (t_ms, launches) in prof.items():
        per_step = t_ms / K
        if launches == 0 and per_step == 0:
            continue
        kernels[k] = dict(ms_per_step=round(per_step, 4), launches_per_step=launches / K)
    if fused and kernels.get("dz_pair", {}).get("ms_per_step", 0) < 0.1:
        exec_flops.pop("dz_pair")          # every pair row's logits were kept: no recompute GEMM had rows
    for k in kernels:
        if k in exec_flops and kernels[k]["ms_per_step"] > 0:
            kernels[k]["tflops"] = round(exec_flops[k] / (kernels[k]["ms_per_step"] * 1e-3) / 1e12, 1)
    exec_flops = {k: v for k, v in exec_flops.items() if kernels.get(k, {}).get("ms_per_step", 0) > 0}
    gemm_kinds = list(exec_flops)
    dom = max(gemm_kinds, key=lambda k: kernels[k]["ms_per_step"])
    # launches of the dominant class that had rows to process: the library launches every row chunk of the (host-known)
    # capacity, the chunks past the (device-known) active row count return at once (~5 us each)
    pad = lambda n: -(-n // 256) * 256
    chunk_rows = args.chunk_rows or _lib.default_chunk_rows(2 * shape.B * (shape.T - 1), shape.B_ref * (shape.T_ref - 1))
    chunks_active = max(1, -(-(pad(n_p) + pad(n_r)) // chunk_rows))
    per_chunk = {"stash_ref": 2}.get(dom, 1)
    n_launch_dom = float(per_chunk * (1 if dom in ("fwd_pair", "fwd_ref") else chunks_active))
    dom_ms_launch = kernels[dom]["ms_per_step"] / n_launch_dom
    achieved = exec_flops[dom] / n_launch_dom / (dom_ms_launch * 1e-3) / 1e12
    share = world if vp else 1                       # vp: the algorithmic FLOPs of the batch are split over the ranks
    f_alg_active = algorithmic_flops(shape, need_dw, rows_pair, rows_ref) / share
    gemm_ms = sum(kernels[k]["ms_per_step"] for k in gemm_kinds)
    traffic, traffic_note = load_traffic(dom)
    roofline = dict(bound="tensor", kernel=dom, achieved=round(achieved, 1), peak=peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=round(achieved / peaks["bf16_tflops"], 4), traffic=traffic, traffic_note=traffic_note,
                    peak_source=peaks["source"], frac_of_sustained=round(achieved / peaks["bf16_tflops_sustained"], 4)
                    if peaks.get("bf16_tflops_sustained") else None,
                    flops_per_launch=exec_flops[dom] / n_launch_dom, ms_per_launch=round(dom_ms_launch, 4),
                    launches_with_rows_per_step=n_launch_dom,
                    share_of_step=round(kernels[dom]["ms_per_step"] / (ms_eager / K), 3),
                    note="FLOPs executed by this kernel for the rows it processed (rows with a non-zero contribution)")
    e2e = None
    if e2e_span:
        e2e = dict(value=tokens_step / (e2e_span["ms_per_step"] * 1e-3), unit=UNIT, h2d_bytes_per_step=e2e_span["h2d_bytes"],
                   d2h_bytes_per_step=12, ms_per_step=e2e_span["ms_per_step"], loss=e2e_span["loss"],
                   plan_ms=round(e2e_span["plan_ms"], 4), issue_ms=round(e2e_span["issue_ms"], 4), copies_per_step=e2e_span["spans"] * 1 + 3,
                   note="per step, inside the timed region: span plan rebuilt from that step's host labels (plan_ms), pinned host "
                        "hidden states/labels -> device (labels in full; hidden rows span-wise: only rows that predict a contributing "
                        "label; issue_ms = host time queuing the copies), 3 losses read back on the host one step late")
        if e2e_full:
            e2e["full_copy"] = dict(value=round(tokens_step / (e2e_full["ms_per_step"] * 1e-3)), ms_per_step=round(e2e_full["ms_per_step"], 4),
                                    h2d_bytes_per_step=e2e_full["h2d_bytes"], note="same loop copying EVERY hidden row")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=max(args.warmup, 3), ms_per_step=ms_step,
                higher_is_better=True, scaling="strong" if vp else "weak", vs_baseline=None,
                dtype="f32 (split-bf16 x6 on the tensor cores, fp32 accumulate promoted per K=256)" if args.fp32 else "bf16", data="synthetic",
                config=config_dict(args, shape), e2e=e2e,
                gpu_launches=int(round(sum(v["launches_per_step"] for v in kernels.values()) * K)),
                launches_per_step=sum(v["launches_per_step"] for v in kernels.values()),
                roofline=roofline, kernels=kernels, clocks=clocks,
                rows=dict(pair_active=rows_pair, ref_active=rows_ref, pair_fed=2 * shape.B * shape.T, ref_fed=shape.B_ref * shape.T_ref),
                step_tc_frac_active_alg=tc_frac(f_alg_active, ms_step, peaks),
                step_tc_frac_executed=tc_frac(sum(exec_flops.values()), ms_step, peaks),
                gemm_share_of_step=round(gemm_ms / (ms_eager / K), 3), loss=loss_val, cuda_graph=graph_info)
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_port(2, 1, args.config)
        line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                    ms_per_step=r["ms_per_step"])
    if variants:
        line["variants"] = variants            # last: the driver keeps the tail of the line
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--pairs", type=int, default=None)
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--parallelism", default="dp", choices=["dp", "vp"],
                    help="dp: data parallel over sample pairs (default); vp: vocab-parallel lm_head (same rows on every rank)")
    ap.add_argument("--vp-torch-exchange", action="store_true",
                    help="vp: exchange through torch.distributed around the C calls instead of the library's own communicator")
    ap.add_argument("--vp-backward", default="auto", choices=["auto", "fused", "recompute"], help="gradient strategy of the vp runs")
    ap.add_argument("--fp32", action="store_true", help="fp32 hidden states / weights (BASELINE configs[0]'s precision): split-bf16 "
                    "tensor-core passes with K-chunk promotion, fp32 gradients; the e2e loop and the variants are skipped")
    ap.add_argument("--no-dw", action="store_true", help="lm_head frozen (the shipped LoRA recipe): no dW")
    ap.add_argument("--dense", action="store_true", help="process every fed position (no row compaction)")
    ap.add_argument("--backward", default="auto", choices=["auto", "fused", "recompute"],
                    help="gradient strategy of dpa_loss (auto = fused)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="headline only: skip every `variants` timing")
    ap.add_argument("--e2e-full-copy", action="store_true", help="e2e: copy every hidden row, not only contributing spans")
    ap.add_argument("--graph", action="store_true", help="(1 GPU) also time the K steps as CUDA-graph replays and report that")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end loop (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
