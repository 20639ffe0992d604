"""Worker of tests/test_gpu_round2.py::test_multi_gpu_sharded_forward_matches_oracle (launched by torchrun, one rank per GPU).

Every rank computes its shard of an unevenly divisible batch with its own engine, all-gathers the probabilities over NCCL
(silero_vad_b200.parallel) and compares the FULL matrix with the CPU oracle run row by row.  Writes rank<r>.json."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def main():
    out = Path(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle.oracle import Oracle
    from recipes import r1_audio
    from silero_vad_b200 import load_silero_vad
    from silero_vad_b200.parallel import shard_bounds, sharded_audio_forward
    model = load_silero_vad(device=local)
    res = {"rank": rank, "world": world, "ok": True, "err": 0.0}
    for sr, B, T in ((16000, 613, 9), (8000, 301, 7), (16000, 5, 6)):     # tile kernels (uneven shards) and the cluster kernel
        n = 512 if sr == 16000 else 256
        x = np.stack([r1_audio(sr, b, n * T) * (10.0 if b % 3 == 0 else 1.0) for b in range(B)])
        probs = sharded_audio_forward(model, torch.from_numpy(x), sr)
        want = Oracle().audio_forward(x, sr, nthreads=8)
        got = probs.cpu().numpy()
        lo, hi = shard_bounds(B, rank, world)
        res["ok"] = res["ok"] and got.shape == want.shape and hi - lo in (B // world, B // world + 1)
        res["err"] = max(res["err"], float(np.abs(got - want).max()))
    (out / f"rank{rank}.json").write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
