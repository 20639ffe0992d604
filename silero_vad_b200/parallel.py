"""Multi-GPU data parallelism over independent streams.

Streams are the unit of parallelism: every row of a batch has private (h, c) and audio context
(src/silero_vad/utils_vad.py:65-76), the ~1 MB of weights is replicated, and nothing is exchanged while the
fused kernel runs.  One process per GPU (torch.distributed, NCCL over NVLink); the only collective is one
all-gather of the per-chunk probabilities (4 B per chunk against 2 KB of audio read), so it is a plain NCCL
call, not a fused kernel.  The time axis is never split: the recurrence is serial per stream.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range [lo, hi) of stream rows owned by `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_probs(local_probs: torch.Tensor, n_streams: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather [B_local, T] probability blocks (shard_bounds order) into [n_streams, T] on every rank.
    Uneven shards are padded to the largest block for the fixed-size collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    T = local_probs.shape[1]
    cap = (n_streams + world - 1) // world
    lo, hi = shard_bounds(n_streams, rank, world)
    assert local_probs.shape[0] == hi - lo, "local block does not match shard_bounds"
    buf = local_probs.new_zeros(cap, T)
    buf[: hi - lo] = local_probs
    out = local_probs.new_empty(world * cap, T)
    dist.all_gather_into_tensor(out, buf, group=group)
    if n_streams == world * cap:
        return out
    rows = [out[r * cap: r * cap + (shard_bounds(n_streams, r, world)[1] - shard_bounds(n_streams, r, world)[0])] for r in range(world)]
    return torch.cat(rows, 0)


def sharded_audio_forward(model, audio: torch.Tensor, sr: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """audio_forward for a batch that every rank holds (or can index): this rank computes rows shard_bounds(...)
    with `model` (its own GPU) and all ranks receive the full [B, T] probability matrix."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(audio.shape[0], rank, world)
    local = model.audio_forward_device(audio[lo:hi], sr)
    return gather_probs(local, audio.shape[0], group)
