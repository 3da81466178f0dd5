"""Batch sharding across the GPUs of one box: the utterance
batch is split contiguously over ranks, every rank runs the identical single-GPU
pipeline, and ONE collective — an all-gather of the wav shards over
NCCL/NVLink — assembles the result in global utterance order.  The reference has
no equivalent (each accelerate rank writes its own files and mis-indexes the
uids, models/vocoders/vocoder_inference.py:360)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous split; the first ``n_items % world_size`` ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def sharded_vocoder_forward(model, mels, group=None):
    """mels: the GLOBAL batch [B, n_mel, T] (every rank holds it or at least its
    own slice filled in).  Returns wav [B, 1, T*hop] on every rank."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = mels.shape[0]
    lo, hi = shard_bounds(B, ws, rank)
    local = model(mels[lo:hi]) if hi > lo else None
    if ws == 1:
        return local
    return gather_shards(local, B, ws, rank, group, like=mels, model=model)


def gather_shards(local, n_items, world_size, rank, group=None, like=None, model=None):
    """All-gather ragged shards (sizes differ by at most one) into global order."""
    base, extra = divmod(n_items, world_size)
    cap = base + (1 if extra else 0)
    if local is not None:
        tail = local.shape[1:]
        dev, dt = local.device, local.dtype
    else:  # this rank got no items: shape from a peer via the padded gather
        raise ValueError("sharded forward needs at least one item per rank")
    padded = local
    if local.shape[0] < cap:
        padded = torch.zeros((cap,) + tuple(tail), device=dev, dtype=dt)
        padded[: local.shape[0]] = local
    out = torch.empty((world_size * cap,) + tuple(tail), device=dev, dtype=dt)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if extra == 0:
        return out
    pieces = []
    for r in range(world_size):
        lo, hi = shard_bounds(n_items, world_size, r)
        pieces.append(out[r * cap: r * cap + (hi - lo)])
    return torch.cat(pieces, 0)
