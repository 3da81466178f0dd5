"""CPU (gloo, world_size 2/3) coverage of the N>1 host logic: contiguous utterance sharding and the final
all-gather in global utterance order, including ragged shard sizes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from amphion_b200.dist import gather_shards, shard_bounds, sharded_vocoder_forward


def test_shard_bounds_partition():
    for n in (1, 2, 7, 8, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_vocoder(mel):  # stand-in with the generator's contract: [B, n_mel, T] -> [B, 1, T*hop]
    return (mel.sum(1, keepdim=True) * 0.01).repeat_interleave(4, dim=-1)


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        mels = torch.randn(n_items, 5, 6, generator=g)          # every rank holds the global batch
        wav = sharded_vocoder_forward(_fake_vocoder, mels)
        torch.save(wav, os.path.join(out_dir, f"wav{rank}.pt"))
        lo, hi = shard_bounds(n_items, world, rank)
        again = gather_shards(_fake_vocoder(mels[lo:hi]), n_items, world, rank)
        assert torch.equal(again, wav)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 6), (2, 7), (3, 8)])
def test_sharded_forward_gathers_in_global_order(tmp_path, world, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    want = _fake_vocoder(torch.randn(n_items, 5, 6, generator=g))
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"wav{r}.pt"))
        assert got.shape == want.shape
        assert torch.equal(got, want)          # same result on every rank, global utterance order
