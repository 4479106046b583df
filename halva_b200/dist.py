"""Host-side helpers for the two ways the DPA-loss path shards across the GPUs of one box.

Data parallel   -- sample pairs (+ their reference samples) are independent units: each rank runs ``dpa_loss`` on its
                   shard with no data-path collective; the only exchange is the all-reduce of trainable gradients
                   (what DeepSpeed ZeRO / DDP do for the reference, /root/reference/src/hallava_7b.sh:30-34).
Vocab parallel  -- lm_head rows are split over a process group; ``dpa_loss(..., vocab_group=g)`` exchanges one
                   all-gather of per-row statistics in the forward and all-reduces dH in the backward.

Everything here is backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def vocab_shard_bounds(vocab: int, world: int, rank: int, multiple: int = 8) -> Tuple[int, int]:
    """[v0, v1) of this rank's lm_head rows: contiguous, sizes differ by at most `multiple`, every boundary except the
    last is a multiple of `multiple` (16-byte rows for the TMA descriptors of the bf16 dz scratch)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the group")
    units = -(-vocab // multiple)
    lo = (units * rank) // world * multiple
    hi = (units * (rank + 1)) // world * multiple
    return min(lo, vocab), min(hi, vocab)


def peer_slice(rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[r0, r1) of the `rows` packed dH rows of a chunk that `rank` sums in the library's peer-memory all-reduce
    (csrc/dpa_peer.cuh: peer_reduce_rows_kernel uses the same split): contiguous, every row owned by exactly one rank."""
    return rows * rank // world, rows * (rank + 1) // world


def dp_shard(n_samples: int, world: int, rank: int) -> range:
    """Sample pairs (and their reference samples) owned by `rank`: contiguous, as even as possible."""
    lo = (n_samples * rank) // world
    hi = (n_samples * (rank + 1)) // world
    return range(lo, hi)


def average_gradients(tensors: Iterable[torch.Tensor], group=None, bucket_bytes: int = 512 << 20) -> None:
    """In-place mean over the group of every tensor (gradients of the trainable parameters the path produced, e.g.
    dW [V,H]).  Tensors are packed into flat buckets so that small ones share a launch; NVSwitch makes the cost
    latency + bytes / (per-GPU bandwidth), so buckets are sized for overlap, not for link count."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        if len(bucket) == 1 and bucket[0].is_contiguous():
            dist.all_reduce(bucket[0], group=group)
            bucket[0].div_(world)
        else:
            flat = torch.cat([t.reshape(-1) for t in bucket])
            dist.all_reduce(flat, group=group)
            flat.div_(world)
            off = 0
            for t in bucket:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        bucket, size = [], 0

    for t in tensors:
        if t is None:
            continue
        nbytes = t.numel() * t.element_size()
        if bucket and (size + nbytes > bucket_bytes or t.dtype != bucket[0].dtype):
            flush()
        bucket.append(t)
        size += nbytes
    flush()


_lib_comm_ranks = None          # global ranks of the group the library's communicator was created over
_shard_offset_cache = {}


def library_comm_matches(group) -> bool:
    """True when `group` is the process group given to init_library_comm (same global ranks in the same order): only then
    may dpa_loss let the library exchange by itself -- its single communicator would otherwise talk to the wrong ranks."""
    if _lib_comm_ranks is None:
        return False
    return _group_ranks(group) == _lib_comm_ranks


def _group_ranks(group):
    g = dist.group.WORLD if group is None else group
    return tuple(dist.get_process_group_ranks(g))


def shard_offsets(v_local: int, group, device) -> Tuple[int, int]:
    """(first global vocabulary id of this rank's lm_head shard, global vocabulary size) from the shard sizes of all ranks
    of `group` (uneven shards allowed).  One small all-gather per (group, v_local), cached afterwards."""
    key = (_group_ranks(group), int(v_local))
    if key not in _shard_offset_cache:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = torch.tensor([int(v_local)], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
        sizes = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sizes, mine, group=group)
        sizes = [int(t.item()) for t in sizes]
        _shard_offset_cache[key] = (sum(sizes[:rank]), sum(sizes))
    return _shard_offset_cache[key]


def init_library_comm(group=None) -> int:
    """Give libhalva_dpa.so its own NCCL communicator over `group` (include/halva_dpa.h: dpa_comm_*): vocab-parallel
    dpa_loss(..., vocab_group=group) then exchanges inside the two C calls instead of through torch.distributed.
    The 128-byte NCCL id is created on the group's rank 0 and broadcast through the group.  Returns the world size."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return 1
    global _lib_comm_ranks
    if lib.dpa_comm_world() == world and _lib_comm_ranks == _group_ranks(group):
        return world
    if lib.dpa_comm_world() > 1:
        _lib.check(lib.dpa_comm_destroy(), "dpa_comm_destroy")
        _lib_comm_ranks = None
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(lib.dpa_comm_unique_id(buf), "dpa_comm_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    on_gpu = dist.get_backend(group) == "nccl"
    if on_gpu:
        t = t.cuda()
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().tolist())
    _lib.check(lib.dpa_comm_init(rank, world, raw), "dpa_comm_init")
    _lib_comm_ranks = _group_ranks(group)
    return world


def destroy_library_comm() -> None:
    from . import _lib
    global _lib_comm_ranks
    _lib.check(_lib.load().dpa_comm_destroy(), "dpa_comm_destroy")
    _lib_comm_ranks = None


# ---- layout of the vocab-parallel exchange record (dpa_row_stats_view): [cap_pair x {m,s,z_y,0}] [cap_ref x {m_p,s_p,m_r,s_r}]
#      [cap_ref x u] fp32.  merge_row_stats is the host restatement of merge_strided_kernel, used by the CPU protocol test.
def row_stats_record_floats(cap_pair: int, cap_ref: int) -> int:
    return 4 * cap_pair + 4 * cap_ref + cap_ref


def merge_row_stats(records: Sequence[torch.Tensor], cap_pair: int, cap_ref: int):
    """records: one flat fp32 tensor per vocab shard.  Returns (pair [cap_pair,3] = m,s,z_y; ref [cap_ref,5] = m_p,s_p,m_r,s_r,u)
    with the shards' partial (max, sum-exp) rescaled to the common maximum and z_y / u summed."""
    pair = torch.stack([r[:4 * cap_pair].view(cap_pair, 4) for r in records])                          # [G,cap,4]
    ref = torch.stack([r[4 * cap_pair:4 * cap_pair + 4 * cap_ref].view(cap_ref, 4) for r in records])
    u = torch.stack([r[4 * cap_pair + 4 * cap_ref:].view(cap_ref) for r in records])
    m = pair[..., 0].max(0).values
    s = (pair[..., 1] * torch.exp(pair[..., 0] - m)).sum(0)
    zy = pair[..., 2].sum(0)
    mp, mr = ref[..., 0].max(0).values, ref[..., 2].max(0).values
    wr = torch.exp(ref[..., 2] - mr)
    sp = (ref[..., 1] * torch.exp(ref[..., 0] - mp)).sum(0)
    sr = (ref[..., 3] * wr).sum(0)
    uu = (u * wr).sum(0)
    return torch.stack([m, s, zy], -1), torch.stack([mp, sp, mr, sr, uu], -1)
