"""Two-GPU checks (NCCL): vocab-parallel lm_head equals the single-GPU result on the same rows; data-parallel
gradient averaging equals the mean of per-shard gradients.  Skipped on boxes with fewer than 2 GPUs
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))


def _worker(rank, world, port, out):
    import dataclasses
    import torch.distributed as dist
    from halva_b200 import dpa_loss
    from halva_b200.dist import average_gradients, dp_shard, vocab_shard_bounds
    from halva_b200.synth import CONFIGS, make_inputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        shape = dataclasses.replace(CONFIGS["small"], V=1024, B=4, B_ref=4)
        t = make_inputs(shape, seed=5, device=dev, shared_head=False)       # identical on every rank (same seed)

        def run(W, W_ref, sel_pair=None, sel_ref=None, src=None, **kw):
            src = t if src is None else src
            hp = (src["h_pair"] if sel_pair is None else src["h_pair"][sel_pair]).clone().requires_grad_(True)
            hr = (src["h_ref_pol"] if sel_ref is None else src["h_ref_pol"][sel_ref]).clone().requires_grad_(True)
            Wl = W.clone().requires_grad_(True)
            pick = lambda k, s: src[k] if s is None else src[k][s]
            loss, align, div = dpa_loss(hp, pick("labels_pair", sel_pair), pick("signs_pair", sel_pair), hr, pick("labels_ref", sel_ref),
                                        Wl, h_ref_frozen=pick("h_ref_frz", sel_ref), W_ref=W_ref, alpha=t["alpha"], **kw)
            loss.backward()
            return loss.detach(), div.detach(), hp.grad, hr.grad, Wl.grad

        # ---- vocab parallel: each rank holds V/2 rows of W and W_ref, all rows of h
        full = run(t["W"], t["W_ref"], backward="recompute")
        v0, v1 = vocab_shard_bounds(shape.V, world, rank)
        part = run(t["W"][v0:v1].contiguous(), t["W_ref"][v0:v1].contiguous(), vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V)
        assert abs(part[0].item() - full[0].item()) <= 1e-5 * abs(full[0].item()), (part[0].item(), full[0].item())
        assert abs(part[1].item() - full[1].item()) <= 1e-4 * abs(full[1].item()) + 1e-7
        assert _rel(part[2], full[2]) <= 5e-3 and _rel(part[3], full[3]) <= 5e-3        # bf16 partial sums all-reduced
        assert _rel(part[4], full[4][v0:v1]) <= 1e-3                                    # dW stays sharded
        # ---- the same in fp32 (split-bf16 path; caller-side exchange, fp32 partial sums all-reduced)
        t32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in t.items()}
        full32 = run(t32["W"], t32["W_ref"], src=t32)
        part32 = run(t32["W"][v0:v1].contiguous(), t32["W_ref"][v0:v1].contiguous(), src=t32, vocab_group=dist.group.WORLD,
                     v_offset=v0, vocab_size=shape.V)
        assert part32[2].dtype == torch.float32
        assert abs(part32[0].item() - full32[0].item()) <= 2e-6 * abs(full32[0].item()), (part32[0].item(), full32[0].item())
        assert _rel(part32[2], full32[2]) <= 1e-5 and _rel(part32[3], full32[3]) <= 1e-5 and _rel(part32[4], full32[4][v0:v1]) <= 1e-5
        # the same with the library's own communicator: statistics all-gathered inside the C calls, dH partial sums reduced
        # through peer memory by the library's own kernels (row counts from device memory, no host read-back)
        from halva_b200 import _lib
        from halva_b200.dist import destroy_library_comm, init_library_comm
        assert init_library_comm(dist.group.WORLD) == world and _lib.load().dpa_comm_world() == world
        shard = lambda: (t["W"][v0:v1].contiguous(), t["W_ref"][v0:v1].contiguous())
        vkw = dict(vocab_group=dist.group.WORLD, v_offset=v0, vocab_size=shape.V)
        own = run(*shard(), backward="recompute", **vkw)
        assert own[0].item() == part[0].item() and torch.equal(own[4], part[4])
        assert _rel(own[2], part[2]) <= 1e-3 and _rel(own[3], part[3]) <= 1e-3
        # fused-gradient forward under vocab parallelism (the default once the library owns the communicator): one exchange of
        # the pair statistics, one per row chunk of the reference-sample rows, dH through peer memory, dW sharded
        full_f = run(t["W"], t["W_ref"])
        for kw in (dict(), dict(chunk_rows=256), dict(chunk_rows=512, dense=True)):       # several chunks: both peer buffers cycle
            fz = run(*shard(), **vkw, **kw)
            assert abs(fz[0].item() - full_f[0].item()) <= 1e-5 * abs(full_f[0].item()), (kw, fz[0].item(), full_f[0].item())
            assert abs(fz[1].item() - full_f[1].item()) <= 1e-4 * abs(full_f[1].item()) + 1e-7
            assert _rel(fz[2], full_f[2]) <= 5e-3 and _rel(fz[3], full_f[3]) <= 5e-3, kw
            assert _rel(fz[4], full_f[4][v0:v1]) <= 2e-3, kw
            assert torch.equal(fz[2].float().abs().sum(-1) == 0, full_f[2].float().abs().sum(-1) == 0)   # the same ROWS carry a gradient
        other = [torch.empty_like(fz[2]) for _ in range(world)]
        dist.all_gather(other, fz[2].contiguous())
        assert all(torch.equal(o, other[0]) for o in other)                               # every rank holds bit-identical dH
        # uneven shards (V = 1016 over 2 ranks: 504 + 512 rows), offsets derived from the shard sizes of all ranks
        shape_u = dataclasses.replace(shape, V=1016)
        tu = make_inputs(shape_u, seed=6, device=dev)
        u0, u1 = vocab_shard_bounds(shape_u.V, world, rank)
        assert (u1 - u0) != shape_u.V // world or world != 2
        full_u = run(tu["W"], tu["W"], src=tu)
        part_u = run(tu["W"][u0:u1].contiguous(), tu["W"][u0:u1].contiguous(), src=tu, vocab_group=dist.group.WORLD)
        assert abs(part_u[0].item() - full_u[0].item()) <= 1e-5 * abs(full_u[0].item()), (part_u[0].item(), full_u[0].item())
        assert _rel(part_u[2], full_u[2]) <= 5e-3 and _rel(part_u[3], full_u[3]) <= 5e-3 and _rel(part_u[4], full_u[4][u0:u1]) <= 2e-3
        with torch.no_grad():                                                             # evaluation under vocab parallelism
            le = dpa_loss(t["h_pair"], t["labels_pair"], t["signs_pair"], t["h_ref_pol"], t["labels_ref"], shard()[0],
                          h_ref_frozen=t["h_ref_frz"], W_ref=shard()[1], alpha=t["alpha"], **vkw)[0]
        assert abs(le.item() - full_f[0].item()) <= 1e-5 * abs(full_f[0].item())
        destroy_library_comm()
        # ---- data parallel: pairs split over ranks, lm_head gradient averaged
        B = shape.B
        mine = list(dp_shard(B, world, rank))
        sel = torch.tensor(mine + [B + i for i in mine], device=dev)
        selr = torch.tensor(mine, device=dev)
        loc = run(t["W"], t["W_ref"], sel, selr)
        dW = loc[4].clone()
        average_gradients([dW])
        others = []
        for r in range(world):
            m = list(dp_shard(B, world, r))
            others.append(run(t["W"], t["W_ref"], torch.tensor(m + [B + i for i in m], device=dev), torch.tensor(m, device=dev))[4].float())
        want = sum(others) / world
        assert _rel(dW, want) <= 1e-2
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_vocab_parallel_and_data_parallel():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
