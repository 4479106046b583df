"""N > 1 host-side logic on CPU: two gloo ranks exercise the data-parallel gradient averaging and the vocab-parallel
exchange protocol (record layout + max-rescale merge) against the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpa_testutil import load_golden
from halva_b200.dist import (average_gradients, dp_shard, library_comm_matches, merge_row_stats, peer_slice, row_stats_record_floats,
                             shard_offsets, vocab_shard_bounds)
from oracle import dpa_oracle

IGNORE = -100


def test_vocab_shard_bounds_cover_and_align():
    for V, G in ((32000, 8), (32000, 3), (1000, 2), (300, 4), (32003, 8)):
        cuts = [vocab_shard_bounds(V, G, r) for r in range(G)]
        assert cuts[0][0] == 0 and cuts[-1][1] == V
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c and b % 8 == 0
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 8 + (V % 8)
    assert vocab_shard_bounds(32000, 8, 3) == (12000, 16000)          # 4000 rows per rank at the 13B cfg-3 split


def test_dp_shard_partitions():
    for n, G in ((16, 8), (5, 2), (3, 4)):
        got = [i for r in range(G) for i in dp_shard(n, G, r)]
        assert got == list(range(n))


def test_peer_slices_partition_the_rows():
    """Rows a rank reduces in the peer-memory all-reduce of the packed dH rows (csrc/dpa_peer.cuh): every row exactly once."""
    for rows in (0, 1, 7, 256, 24832, 32768):
        for G in (2, 3, 4, 8):
            cuts = [peer_slice(rows, r, G) for r in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == rows
            assert all(b == c for (_, b), (c, _) in zip(cuts, cuts[1:]))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_record(g, v0, v1):
    """What one vocab shard contributes for every (dense) row: the layout of dpa_row_stats_view."""
    f64 = lambda k: np.asarray(g[k], np.float64)
    W, Wr = f64("W")[v0:v1], f64("W_ref")[v0:v1]
    zp_pair = f64("h_pair")[:, :-1].reshape(-1, W.shape[1]) @ W.T
    y = np.asarray(g["labels_pair"])[:, 1:].reshape(-1)
    y0 = np.where(y == IGNORE, 0, y)
    m = zp_pair.max(-1)
    s = np.exp(zp_pair - m[:, None]).sum(-1)
    inside = (y0 >= v0) & (y0 < v1)
    zy = np.where(inside, zp_pair[np.arange(len(y0)), np.clip(y0 - v0, 0, v1 - v0 - 1)], 0.0)
    zp = f64("h_ref_pol")[:, :-1].reshape(-1, W.shape[1]) @ W.T
    zr = f64("h_ref_frz")[:, :-1].reshape(-1, W.shape[1]) @ Wr.T
    mp_, mr_ = zp.max(-1), zr.max(-1)
    sp_, e = np.exp(zp - mp_[:, None]).sum(-1), np.exp(zr - mr_[:, None])
    sr_, u = e.sum(-1), (e * (zr - zp)).sum(-1)
    pair = np.stack([m, s, zy, np.zeros_like(m)], -1)
    ref = np.stack([mp_, sp_, mr_, sr_], -1)
    return torch.from_numpy(np.concatenate([pair.reshape(-1), ref.reshape(-1), u]).astype(np.float32))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("tiny_fp32")
        # ---- data parallel: each rank owns one sample pair + its reference sample; gradients are averaged
        B = g["h_pair"].shape[0] // 2
        mine = list(dp_shard(B, world, rank))
        sel = mine + [B + i for i in mine]
        o = dpa_oracle.dpa_loss(g["h_pair"][sel], g["labels_pair"][sel], g["signs_pair"][sel], g["h_ref_pol"][mine], g["h_ref_frz"][mine],
                                g["labels_ref"][mine], g["W"], g["W_ref"], float(g["alpha"]))
        dW = torch.from_numpy(o["dW"]).clone()
        small = torch.full((3,), float(rank + 1), dtype=torch.float64)
        average_gradients([dW, small], bucket_bytes=1 << 30)          # both in one bucket
        gathered = [torch.zeros_like(dW) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(o["dW"]).clone())
        assert torch.allclose(dW, sum(gathered) / world, rtol=1e-12, atol=0)
        assert torch.allclose(small, torch.full((3,), (1 + world) / 2.0, dtype=torch.float64))
        # ---- vocab parallel: partial per-row statistics of my lm_head rows -> all-gather -> merge == full-vocab oracle
        V = g["W"].shape[0]
        v0, v1 = vocab_shard_bounds(V, world, rank)
        rec = _partial_record(g, v0, v1)
        n_pair = g["h_pair"].shape[0] * (g["h_pair"].shape[1] - 1)
        n_ref = g["h_ref_pol"].shape[0] * (g["h_ref_pol"].shape[1] - 1)
        assert rec.numel() == row_stats_record_floats(n_pair, n_ref)
        recs = [torch.zeros_like(rec) for _ in range(world)]
        dist.all_gather(recs, rec)                                     # the one exchange step of the forward
        pair, ref = merge_row_stats(recs, n_pair, n_ref)
        full = dpa_oracle.dpa_loss(g["h_pair"], g["labels_pair"], g["signs_pair"], g["h_ref_pol"], g["h_ref_frz"], g["labels_ref"],
                                   g["W"], g["W_ref"], float(g["alpha"]))
        lse = (pair[:, 0] + torch.log(pair[:, 1])).numpy()
        assert np.allclose(lse, full["lse_pair"].reshape(-1), rtol=1e-5)
        valid = (np.asarray(g["labels_pair"])[:, 1:] != IGNORE).reshape(-1)
        assert np.allclose((pair[:, 2].numpy() - lse) * valid, full["logp_pair"].reshape(-1), rtol=1e-4, atol=1e-5)
        kl = (ref[:, 4] / ref[:, 3] - ((ref[:, 2] - ref[:, 0]) + torch.log(ref[:, 3] / ref[:, 1]))).numpy()
        act = full["active_ref"].reshape(-1)
        assert np.allclose(kl * act, full["kl_rows"].reshape(-1), rtol=2e-3, atol=2e-6)
        # ---- fused-gradient forward under vocab parallelism: with the GLOBAL log-sum-exps known after the exchange, every shard
        #      sums q (z_r - z_p) over its own columns; the row's KL is the sum of the shards' parts minus (LSE_r - LSE_p) once
        #      (rank 0 subtracts it), and the divergence one scalar all-reduce of the shards' sums
        f64 = lambda k: np.asarray(g[k], np.float64)
        zp = f64("h_ref_pol")[:, :-1].reshape(-1, f64("W").shape[1]) @ f64("W")[v0:v1].T
        zr = f64("h_ref_frz")[:, :-1].reshape(-1, f64("W").shape[1]) @ f64("W_ref")[v0:v1].T
        lse_p = (ref[:, 0] + torch.log(ref[:, 1])).double().numpy()
        lse_r = (ref[:, 2] + torch.log(ref[:, 3])).double().numpy()
        t_local = (np.exp(zr - lse_r[:, None]) * (zr - zp)).sum(-1)
        part = torch.from_numpy((t_local - (lse_r - lse_p if rank == 0 else 0.0)) * act)
        tot = part.clone()
        dist.all_reduce(tot)
        assert np.allclose(tot.numpy(), full["kl_rows"].reshape(-1), rtol=2e-3, atol=2e-6)
        div = torch.tensor([part.sum().item()], dtype=torch.float64)
        dist.all_reduce(div)
        assert abs(div.item() / g["h_ref_pol"].shape[0] - full["div"]) <= 2e-3 * abs(full["div"]) + 1e-7
        # ---- shard offsets from the shard sizes of all ranks (uneven shards allowed), and the communicator / group match
        sizes = [5, 11][:world]
        off, total = shard_offsets(sizes[rank], None, "cpu")
        assert (off, total) == (sum(sizes[:rank]), sum(sizes))
        assert shard_offsets(sizes[rank], None, "cpu") == (off, total)       # cached: no second collective needed
        assert library_comm_matches(None) is False                           # no library communicator was created here
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_protocols():
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
