#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref/silero_vad, staged by baseline/stage_reference.py) on this box:
  B1  load_silero_vad() TorchScript, batch 1, torch.set_num_threads(1), model(chunk, 16000) per chunk (model.py:3,17,34)
  B2  TorchScript audio_forward, batch 4096 (bounded T), all host cores (utils_vad.py:94-110)
  B4  the same module moved to the GPU (tuning/tune.py:38), batch 1 per-chunk latency and batch 4096 throughput
Median of >= 5 warmed trials (examples/onnx_sequence/run.py:172-194).  Prints one JSON object.
"""
import json
import os
import statistics
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "baseline" / "_ref"))

import numpy as np
import torch


def cpu_info():
    model, phys = "?", set()
    try:
        core = pkg = None
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
            elif line.startswith("physical id"):
                pkg = line.split(":")[1].strip()
            elif line.startswith("core id"):
                core = line.split(":")[1].strip()
                phys.add((pkg, core))
    except OSError:
        pass
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    return {"model": model, "logical": os.cpu_count(), "physical": len(phys) or None, "affinity": aff}


def median_time(fn, trials=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(trials):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def main():
    import silero_vad
    sr, n = 16000, 512
    out = {"cpu": cpu_info(), "torch": torch.__version__}
    rng = np.random.default_rng(17 + sr)
    model = silero_vad.load_silero_vad()
    # B1
    torch.set_num_threads(1)
    T1 = 400
    x1 = torch.from_numpy((rng.standard_normal(n * T1) * 0.03).astype(np.float32))

    def b1():
        model.reset_states()
        with torch.no_grad():
            for t in range(T1):
                model(x1[t * n:(t + 1) * n], sr).item()
    dt = median_time(b1, trials=5, warm=1)
    out["B1_jit_cpu_batch1_1thread"] = {"chunks_per_s": T1 / dt, "us_per_chunk": dt / T1 * 1e6, "chunks": T1}
    # B2
    ncores = out["cpu"]["affinity"] or os.cpu_count()
    for threads in sorted({ncores, max(1, (out["cpu"]["physical"] or ncores))}):
        torch.set_num_threads(threads)
        B, T = 4096, 8
        xb = torch.from_numpy((rng.standard_normal((B, n * T)) * 0.03).astype(np.float32))
        dt = median_time(lambda: model.audio_forward(xb, sr), trials=5, warm=1)
        out[f"B2_jit_cpu_batch4096_{threads}threads"] = {"chunks_per_s": B * T / dt, "threads": threads, "sample": f"{B} streams x {T} chunks"}
    torch.set_num_threads(1)
    # B4
    if torch.cuda.is_available():
        dev = torch.device("cuda:0")
        gm = silero_vad.load_silero_vad().to(dev)
        xg = x1.to(dev)

        def g1():
            gm.reset_states()
            with torch.no_grad():
                for t in range(T1):
                    gm(xg[t * n:(t + 1) * n], sr).item()
        dt = median_time(g1, trials=5, warm=2)
        out["B4_jit_b200_batch1"] = {"chunks_per_s": T1 / dt, "us_per_chunk": dt / T1 * 1e6}
        B, T = 4096, 64
        xb = (torch.randn(B, n * T, device=dev) * 0.03)

        def g2():
            gm.reset_states()
            with torch.no_grad():
                for t in range(T):
                    p = gm(xb[:, t * n:(t + 1) * n], sr)
            torch.cuda.synchronize()
        dt = median_time(g2, trials=5, warm=2)
        out["B4_jit_b200_batch4096"] = {"chunks_per_s": B * T / dt, "sample": f"{B} streams x {T} chunks, device-resident audio, per-chunk module calls"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
