#!/usr/bin/env python
"""Bring-up check of svad_fused_h16 on a GPU: every activation the kernel dumps for CTA 0, steps 0 and 1 (mag, e0, e1, e2, e3,
LSTM gate pre-activations, h') against the CPU model of the same arithmetic (tools/h16_numerics.py), then probabilities
against the reference goldens.  Usage: h16_check.py [16000|8000]"""
import ctypes
import os
import sys
from pathlib import Path

os.environ["SVAD_DEBUG_LIB"] = "1"   # the -DSVAD_H16_DEBUG build (silero_vad_b200.build.build_debug()) holds the dump code

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))
from h16_numerics import H16Model, read_container  # noqa: E402

from silero_vad_b200 import _cabi, load_silero_vad  # noqa: E402

K = dict(mag=0, nyq=4 * 128 * 32)
K["e0"] = K["nyq"] + 128
K["e1"] = K["e0"] + 4 * 128 * 32
K["e2"] = K["e1"] + 2 * 64 * 32
K["e3"] = K["e2"] + 64 * 32
K["gates"] = K["e3"] + 128 * 32
K["h"] = K["gates"] + 4 * 128 * 32
K["c"] = K["h"] + 128 * 32
STEP = K["c"] + 128 * 32


def main():
    sr = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
    n = 512 if sr == 16000 else 256
    Kt = 128 if sr == 16000 else 64
    name = "test16k" if sr == 16000 else "aepyx8k"
    z = np.load(REPO / f"tests/golden/{name}.npz")
    wav = z["pcm"].astype(np.float32) / 32768.0
    m = load_silero_vad(device=0)
    m.engine.set_kernel("h16")
    m.engine.set_small_batch_max(0)
    m.engine.set_tile_rows(7)
    B, T = 28, 4
    x = np.stack([wav[30000 + 9000 * b: 30000 + 9000 * b + n * T] for b in range(B)]).copy()
    dbg = torch.zeros(2 * STEP, device="cuda")
    L = _cabi.lib()
    L.svad_engine_set_debug_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.svad_engine_set_debug_buffer(m.engine._h, dbg.data_ptr())
    p = m.audio_forward(torch.from_numpy(x), sr).numpy()
    torch.cuda.synchronize()
    L.svad_engine_set_debug_buffer(m.engine._h, None)
    d = dbg.cpu().numpy()
    tm = read_container(REPO / "silero_vad_b200/data/silero_vad_v6.weights")
    worst = {}
    pm = np.zeros((B, T), np.float32)
    for b in range(B):
        mod = H16Model(tm, sr)
        mod.trace = []
        pm[b] = mod.run(x[b], T)
        for s in range(2):
            tr = mod.trace[s]
            blk = d[s * STEP:(s + 1) * STEP]
            got = {
                "mag": blk[K["mag"]: K["mag"] + 4 * Kt * 32].reshape(4, Kt, 32)[:, :, b].T,          # [Kt, 4]
                "e0": blk[K["e0"]: K["e0"] + 4 * 128 * 32].reshape(4, 128, 32)[:, :, b].T,
                "e1": blk[K["e1"]: K["e1"] + 2 * 64 * 32].reshape(2, 64, 32)[:, :, b].T,
                "e2": blk[K["e2"]: K["e2"] + 64 * 32].reshape(64, 32)[:, b],
                "e3": blk[K["e3"]: K["e3"] + 128 * 32].reshape(128, 32)[:, b],
                "gates": blk[K["gates"]: K["gates"] + 4 * 128 * 32].reshape(512, 32)[:, b],
                "h": blk[K["h"]: K["h"] + 128 * 32].reshape(128, 32)[:, b],
            }
            want = {"mag": tr["mag"][:Kt], "e0": tr["e0"], "e1": tr["e1"], "e2": tr["e2"], "e3": tr["e3"],
                    "gates": tr["gates"] - mod.bl, "h": tr["h"]}
            for k in got:
                e = float(np.abs(got[k] - want[k]).max())
                key = (s, k)
                if e > worst.get(key, (0, 0, 0))[0]:
                    worst[key] = (e, b, float(np.abs(want[k]).max()))
    for (s, k), (e, b, mx) in sorted(worst.items()):
        print(f"step {s} {k:6s}: max|gpu - model| = {e:.3e} (stream {b}, max|value| {mx:.3g})")
    print("probabilities vs CPU model:", float(np.abs(p - pm).max()))
    # golden: whole fixture through the h16 kernel as ONE stream copied into 300 rows (tile kernels need B > small_max = 0 here)
    m.engine.set_tile_rows(0)
    Lw = (len(wav) // n) * n
    got = m.audio_forward(torch.from_numpy(wav[:Lw])[None], sr).numpy()[0]
    err = float(np.abs(got - z["probs"][: len(got)]).max())
    print(f"{name}: max|p - p_ref| = {err:.3e} over {len(got)} chunks")
    xb = torch.from_numpy(np.stack([wav[(2500 * b) % 900000: (2500 * b) % 900000 + 40 * n] for b in range(4096)]))
    pa = m.audio_forward(xb, sr)
    m.engine.set_kernel("tc")
    pb = m.audio_forward(xb, sr)
    print("B=4096 x 40 chunks, h16 vs tc kernel:", float((pa - pb).abs().max()))


if __name__ == "__main__":
    main()
