"""world_size-2 gloo test of the stream-sharding + all-gather host logic (no GPU: each rank's local
probabilities come from the CPU oracle standing in for its GPU; the collective and index algebra are real)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleBackedModel:
    """Duck-types the one method sharded_audio_forward needs; stands in for the per-rank GPU engine."""

    def __init__(self):
        from oracle.oracle import Oracle
        self.o = Oracle()

    def audio_forward_device(self, x, sr):
        return torch.from_numpy(self.o.audio_forward(x.numpy(), sr))


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from silero_vad_b200.parallel import shard_bounds, sharded_audio_forward
        rng = np.random.default_rng(3)
        audio = torch.from_numpy((rng.standard_normal((B, 512 * 3 + 50)) * 0.2).astype(np.float32))
        model = _OracleBackedModel()
        full = sharded_audio_forward(model, audio, 16000)
        want = model.audio_forward_device(audio, 16000)
        lo, hi = shard_bounds(B, rank, world)
        q.put((rank, bool(torch.equal(full, want)), tuple(full.shape), (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7, 1])
def test_sharded_forward_gloo_world2(B):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == (B, 4)
    assert res[0][3][0] == 0 and res[-1][3][1] == B and res[0][3][1] == res[1][3][0]


def test_shard_bounds_cover():
    from silero_vad_b200.parallel import shard_bounds
    for n in (0, 1, 5, 8, 4096, 65536, 65537):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
