#!/usr/bin/env python
"""bench.py -- DPA loss forward+backward throughput on synthetic LLaVA-1.5-7B lm_head shapes.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference op chain on the host CPU (oracle port)

One "step" = one dpa_loss forward + backward over one batch of BASELINE.json configs[1]
(16 sample pairs + 16 reference samples, T 2048, H 4096, V 32000, bf16; per GPU under data parallelism).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the definitions used here.
"""
from __future__ import annotations

import argparse
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
    return k["dram_bytes"], f"largest {kernel_class} launch in {d.get('source')} ({k['launch']})"


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
def cpu_sample_inputs():
    """Bounded sample of the workload for the host-CPU arms: 1 pair + 1 reference sample, T 1024, fp32 (the
    reference's CPU-runnable configuration, BASELINE.json configs[0])."""
    import torch
    from halva_b200.synth import CONFIGS, make_inputs
    shape = CONFIGS["cfg1"]
    return shape, make_inputs(shape, seed=1234, dtype=torch.float32, device="cpu", plant=PLANT)


def time_cpu_port(steps, warmup):
    import torch
    from oracle import dpa_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape, inp = cpu_sample_inputs()
    for _ in range(warmup):
        dpa_port.step_cpu(inp)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        dpa_port.step_cpu(inp)
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return dict(value=shape.tokens * steps / total, ms_per_step=1e3 * total / steps, cores=cores, tokens_per_step=shape.tokens,
                sample="1 pair + 1 reference sample, T=1024, H=4096, V=32000, fp32, lm_head trainable (cfg-1 shaped sample of the cfg-2 batch)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_port(args.steps, max(args.warmup, 1))
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=config_dict(args, None),
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


def run_e2e(args, tensors, step, barrier, dev, world):
    """Same step driven from HOST buffers: every step the pinned hidden states / labels / phrase ids are copied to the
    device (halva_b200.host.HostStager: labels in full, hidden rows span-wise -- only rows that predict a contributing
    label are ever read by the kernels), double-buffered against the previous step's compute on a copy stream, and the
    three losses are read back to the host."""
    import torch
    import torch.distributed as dist
    from halva_b200.host import HostStager
    host = {k: v.detach().cpu().pin_memory() for k, v in tensors.items()}
    stager = HostStager(host["labels_pair"], host["signs_pair"], host["labels_ref"], dense=bool(args.dense or args.e2e_full_copy))
    sets = []
    for _ in range(2):
        s = {k: torch.zeros_like(v.detach()) for k, v in tensors.items()}
        s["h_pair"].requires_grad_(True)
        s["h_ref_pol"].requires_grad_(True)
        sets.append(s)
    out_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d = [0]
    read = []

    def e2e_loop(n):
        """Step i: H2D of its inputs on the copy stream (overlaps step i-1's kernels), its kernels, D2H of its three losses.
        The host reads the losses of EVERY step, one step late -- while the next step's kernels run -- the way a trainer
        logs; a blocking read straight after each step would leave the GPU idle while the host queues the next ~65 launches."""
        cur = torch.cuda.current_stream()
        for i in range(n + 2):
            if i < n:
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_done[b])           # buffer set b free again (step i-2 finished)
                    h2d[0] = stager.stage(host, sets[b])
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
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = tt.item()
    assert len(read) == 2 + args.steps, len(read)                 # every step's loss was read on the host
    return ms_e2e, h2d[0], read[-1]


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import dataclasses
    import torch
    import torch.distributed as dist
    from halva_b200 import _lib, dpa_loss, read_status
    from halva_b200.dist import average_gradients
    from halva_b200.synth import CONFIGS, algorithmic_flops, make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    need_dw = not args.no_dw
    shape = CONFIGS[args.config]
    if args.density is not None:
        shape = dataclasses.replace(shape, density=args.density)
    if args.pairs is not None:
        shape = dataclasses.replace(shape, B=args.pairs, B_ref=args.pairs)
    vp = args.parallelism == "vp" and world > 1
    # dp: every rank its own batch (weak scaling);  vp: the SAME rows on every rank, lm_head rows split over the ranks
    if args.fp32:
        args.backward, args.no_e2e, args.no_variants = "recompute", True, True
    t = make_inputs(shape, seed=1234 + (0 if vp else rank), device=dev, plant=PLANT,
                    dtype=torch.float32 if args.fp32 else torch.bfloat16)
    hp, hr = t["h_pair"].requires_grad_(True), t["h_ref_pol"].requires_grad_(True)
    vp_kw = {}
    if vp:
        from halva_b200.dist import vocab_shard_bounds
        v0, v1 = vocab_shard_bounds(shape.V, world, rank)
        t["W"] = t["W"][v0:v1].contiguous()
        vp_kw = dict(vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V)
        if not args.vp_torch_exchange:      # the library's own NCCL communicator: both exchange steps inside the two C calls
            from halva_b200.dist import init_library_comm
            init_library_comm(dist.group.WORLD)
    W = t["W"].requires_grad_(need_dw)
    tensors = dict(h_pair=hp, h_ref_pol=hr, h_ref_frz=t["h_ref_frz"], labels_pair=t["labels_pair"], signs_pair=t["signs_pair"],
                   labels_ref=t["labels_ref"])

    def step(src=tensors):
        src["h_pair"].grad = None
        src["h_ref_pol"].grad = None
        W.grad = None
        loss, align, div = dpa_loss(src["h_pair"], src["labels_pair"], src["signs_pair"], src["h_ref_pol"], src["labels_ref"], W,
                                    h_ref_frozen=src["h_ref_frz"], alpha=shape.alpha, dense=args.dense, chunk_rows=args.chunk_rows,
                                    backward="recompute" if vp else args.backward, **vp_kw)
        loss.backward()
        if world > 1 and need_dw and not vp:   # data parallel: average the lm_head gradient over ranks (NCCL over NVLink)
            average_gradients([W.grad])
        return loss, align, div

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss, _, _ = step()
    st = read_status(loss)
    assert st["status"] == 0, st
    rows_pair, rows_ref = st["n_pair_active"], st["n_ref_active"]
    for _ in range(max(args.warmup, 3) - 1):
        step()
    barrier()

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER") else None
    if sampler:
        sampler.start()
    _lib.profile_enable(True)
    _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss, _, _ = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    loss_val = loss.item()
    ms_eager = ms

    # ---- the same K steps replayed from a CUDA graph: the path is stream-ordered with no host sync (row counts live on
    #      the device), so one capture removes the ~65 launch gaps of a step.  Falls back to the eager timing on any error.
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
            barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss_g, _, _ = step()
            for _ in range(2):
                g.replay()
            barrier()
            e0.record()
            for _ in range(args.steps):
                g.replay()
            e1.record()
            barrier()
            ms_g = e0.elapsed_time(e1)
            if world > 1:
                tt = torch.tensor([ms_g], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms_g = tt.item()
            lg = loss_g.item()
            if abs(lg - loss_val) > 1e-6 * abs(loss_val):
                raise RuntimeError(f"graph replay loss {lg} != eager loss {loss_val}")
            graph_info = dict(used=True, ms_per_step_eager=ms_eager / args.steps, ms_per_step_graph=ms_g / args.steps)
            ms = ms_g
        except Exception as e:  # noqa: BLE001
            graph_info = dict(used=False, error=repr(e)[:200])
            if os.environ.get("BENCH_DEBUG"):
                import traceback
                traceback.print_exc()
            torch.cuda.synchronize()

    # ---- end to end: host (pinned) buffers in, losses out, H2D double-buffered against the previous step's compute
    ms_e2e, h2d_bytes, loss_e2e = None, 0, None
    if not args.no_e2e:
        ms_e2e, h2d_bytes, loss_e2e = run_e2e(args, tensors, step, barrier, dev, world)
    # (measured BEFORE the variants: every further minute under load leaves the GPU in a hotter / lower power state, and the
    #  end-to-end number is the headline -- it should see the same device state as the device-resident timing above)

    # ---- variant: lm_head frozen (the shipped LoRA recipe, /root/reference/llava/train/train_halva.py:167-168): no dW
    ms_frozen = None
    if need_dw and not args.no_variants:
        Wf = W.detach()

        def step_frozen():
            hp.grad = None
            hr.grad = None
            l, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], Wf, h_ref_frozen=t["h_ref_frz"],
                               alpha=shape.alpha, dense=args.dense, chunk_rows=args.chunk_rows, backward=args.backward)
            l.backward()

        for _ in range(3):
            step_frozen()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_frozen()
        e1.record()
        barrier()
        ms_frozen = e0.elapsed_time(e1) / args.steps
        if world > 1:
            tt = torch.tensor([ms_frozen], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_frozen = tt.item()

    # ---- variant: the frozen model's logits GIVEN (dpa_loss(ref_logits=...), e.g. cached from an earlier frozen pass) instead
    #      of its hidden states: one lm_head GEMM less per reference-sample row, an HBM-bound copy + statistics pass instead
    ms_given = None
    if not args.no_variants and not vp and args.backward != "recompute":
        with torch.no_grad():
            zr = torch.empty(shape.B_ref, shape.T_ref, shape.V, dtype=torch.bfloat16, device=dev)
            for b0 in range(shape.B_ref):
                torch.matmul(t["h_ref_frz"][b0], t["W"].detach().t(), out=zr[b0])

        def step_given():
            hp.grad = None
            hr.grad = None
            W.grad = None
            l, _, _ = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, ref_logits=zr,
                               alpha=shape.alpha, dense=args.dense, chunk_rows=args.chunk_rows)
            l.backward()
            return l

        for _ in range(3):
            lg_ = step_given()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_given()
        e1.record()
        barrier()
        ms_given = e0.elapsed_time(e1) / args.steps
        loss_given = lg_.item()
        del zr, lg_
        if world > 1:
            tt = torch.tensor([ms_given], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_given = tt.item()

    # ---- variant: the reference recipe's form of the same 16 pairs -- 4 micro-batches of 4 pairs, gradients accumulated
    #      (per_device_train_batch_size 4 x gradient_accumulation_steps 4, /root/reference/src/hallava_7b.sh:21-22)
    ms_micro = None
    if not args.no_variants and not vp and shape.B % 4 == 0 and shape.B_ref == shape.B and shape.B >= 8:
        mb = shape.B // 4
        micro = []
        for i in range(4):
            rows = torch.arange(i * mb, (i + 1) * mb, device=dev)
            pr = torch.cat([rows, shape.B + rows])
            micro.append(dict(h_pair=t["h_pair"].detach()[pr].clone().requires_grad_(True), labels_pair=t["labels_pair"][pr].contiguous(),
                              signs_pair=t["signs_pair"][pr].contiguous(), h_ref_pol=t["h_ref_pol"].detach()[rows].clone().requires_grad_(True),
                              h_ref_frz=t["h_ref_frz"][rows].contiguous(), labels_ref=t["labels_ref"][rows].contiguous()))

        def step_micro():
            W.grad = None
            for m in micro:
                m["h_pair"].grad = None
                m["h_ref_pol"].grad = None
                l, _, _ = dpa_loss(m["h_pair"], m["labels_pair"], m["signs_pair"], m["h_ref_pol"], m["labels_ref"], W,
                                   h_ref_frozen=m["h_ref_frz"], alpha=shape.alpha, dense=args.dense, chunk_rows=args.chunk_rows,
                                   backward=args.backward)
                (l / 4).backward()                       # W.grad accumulates over the micro-batches, as under grad accumulation

        for _ in range(3):
            step_micro()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_micro()
        e1.record()
        barrier()
        ms_micro = e0.elapsed_time(e1) / args.steps
        del micro
        if world > 1:
            tt = torch.tensor([ms_micro], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_micro = tt.item()

    clocks = sampler.finish() if sampler else None


    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    K = args.steps
    tokens_step = shape.tokens * (1 if vp else world)
    if vp:
        args.backward = "recompute"
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
    if fused and (args.chunk_rows or 16384) >= n_p:
        exec_flops.pop("dz_pair")          # every pair row sits in row chunk 0: its logits were kept, no recompute GEMM ran
    kernels = {}
    for k, (t_ms, launches) in prof.items():
        per_step = t_ms / K
        if launches == 0 and per_step == 0:
            continue
        kernels[k] = dict(ms_per_step=round(per_step, 4), launches_per_step=launches / K,
                          tflops=round(exec_flops[k] / (per_step * 1e-3) / 1e12, 1) if k in exec_flops and per_step > 0 else None)
    exec_flops = {k: v for k, v in exec_flops.items() if kernels.get(k, {}).get("ms_per_step", 0) > 0}
    gemm_kinds = list(exec_flops)
    dom = max(gemm_kinds, key=lambda k: kernels[k]["ms_per_step"])
    # launches of the dominant class that had rows to process: the library launches every row chunk of the (host-known)
    # capacity, the chunks past the (device-known) active row count return at once (~5 us each)
    pad = lambda n: -(-n // 256) * 256
    chunks_active = max(1, -(-(pad(n_p) + pad(n_r)) // (args.chunk_rows or 16384)))
    per_chunk = {"stash_ref": 2}.get(dom, 1)
    n_launch_dom = float(per_chunk * (1 if dom == "fwd_pair" or dom == "fwd_ref" else chunks_active))
    dom_ms_launch = kernels[dom]["ms_per_step"] / n_launch_dom
    achieved = exec_flops[dom] / n_launch_dom / (dom_ms_launch * 1e-3) / 1e12
    share = world if vp else 1                       # vp: the algorithmic FLOPs of the batch are split over the ranks
    f_alg_dense = algorithmic_flops(shape, need_dw) / share
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
                    note="FLOPs executed by this kernel for the rows it processed (row compaction: only rows with a non-zero "
                         "contribution are run). Whole-step figures: step_tc_frac_*")
    e2e_ms_step = ms_e2e / K if ms_e2e else None
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=max(args.warmup, 3), ms_per_step=ms_step,
                higher_is_better=True, scaling="strong" if vp else "weak", vs_baseline=None,
                dtype="f32 (split-bf16 x6 on the tensor cores, fp32 accumulate promoted per K=256)" if args.fp32 else "bf16", data="synthetic",
                config=config_dict(args, shape),
                e2e=dict(value=tokens_step / (e2e_ms_step * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=12,
                         ms_per_step=e2e_ms_step, loss=loss_e2e,
                         note="pinned host hidden states/labels -> device every step (labels in full; hidden rows span-wise: only rows "
                              "that predict a contributing label are transferred, the kernels never read the others), double-buffered "
                              "on a copy stream; the three losses of every step are read back on the host, one step late (while the "
                              "next step's kernels run)") if e2e_ms_step else None,
                gpu_launches=int(round(sum(v["launches_per_step"] for v in kernels.values()) * K)),
                roofline=roofline, kernels=kernels, clocks=clocks,
                rows=dict(pair_active=rows_pair, ref_active=rows_ref, pair_fed=2 * shape.B * shape.T, ref_fed=shape.B_ref * shape.T_ref),
                step_tc_frac_dense_alg=round(f_alg_dense / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
                step_tc_frac_active_alg=round(f_alg_active / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
                step_tc_frac_executed=round(sum(exec_flops.values()) / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
                gemm_share_of_step=round(gemm_ms / (ms_eager / K), 3), loss=loss_val, cuda_graph=graph_info)
    if ms_frozen:
        line["variants"] = dict(lm_head_frozen=dict(ms_per_step=ms_frozen, value=tokens_step / (ms_frozen * 1e-3), unit=UNIT,
                                                    note="same batch, lm_head not trainable (LoRA recipe): no dW GEMM"))
    if ms_given:
        line.setdefault("variants", {})["ref_logits_given"] = dict(
            ms_per_step=ms_given, value=tokens_step / (ms_given * 1e-3), unit=UNIT, loss=loss_given,
            note="same batch, the frozen model's bf16 logits handed over as ref_logits [B_ref,T_ref,V] instead of its hidden "
                 "states: 3 instead of 4 lm_head GEMMs per reference-sample row")
    if ms_micro:
        line.setdefault("variants", {})["micro_4x4"] = dict(
            ms_per_step=ms_micro, value=tokens_step / (ms_micro * 1e-3), unit=UNIT,
            note="the same pairs as 4 micro-batches of a quarter each with gradient accumulation (the reference recipe's "
                 "per-device batch 4 x accumulation 4): 4 forward+backward calls per step, dW accumulated by autograd")
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_port(2, 1)
        line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                    ms_per_step=r["ms_per_step"])
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
    ap.add_argument("--fp32", action="store_true", help="fp32 hidden states / weights (BASELINE configs[0]'s precision): split-bf16 "
                    "tensor-core passes with K-chunk promotion, fp32 gradients; the e2e loop and the variants are skipped")
    ap.add_argument("--no-dw", action="store_true", help="lm_head frozen (the shipped LoRA recipe): no dW")
    ap.add_argument("--dense", action="store_true", help="process every fed position (no row compaction)")
    ap.add_argument("--backward", default="auto", choices=["auto", "fused", "recompute"],
                    help="gradient strategy of dpa_loss (auto = fused)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the lm_head-frozen variant timing")
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
