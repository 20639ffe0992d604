#!/usr/bin/env python
"""Per-role phase times of svad_fused_h16 (CTA 0, middle step of a B=4096 x T=64 launch), from the clock64 stamps the kernel
records when a debug buffer is set.  Usage: h16_phase_times.py [16000|8000] [batch]"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from silero_vad_b200 import _cabi, load_silero_vad  # noqa: E402

STEP = 4 * 128 * 32 + 128 + 4 * 128 * 32 + 2 * 64 * 32 + 64 * 32 + 128 * 32 + 4 * 128 * 32 + 128 * 32 + 128 * 32
NAMES = {
    0: ["loop top", "front region free (f_done)", "window staged", "STFT acc ready", "mag written", "enc0 acc ready", "e0 written",
        "  (window rows stored)", "  (group sync)", "  (reflect rows copied)"],
    1: ["loop top", "xp ready", "enc1 acc drained", "STFT issued", "mag ready", "enc0 issued", "e0 ready", "enc1 issued"],
    2: ["loop top", "f_done", "e1 written", "enc2 acc", "e2 written", "enc3 acc", "e3 written", "LSTM acc", "h written+sync", "head done"],
    3: ["loop top", "e1 ready", "enc2 issued", "e2 ready", "enc3 issued", "e3 ready", "LSTM issued"],
}
ROLE = ["EF (front epilogue)", "MF (front MMA issue)", "EB (back epilogue)", "MB (back MMA issue)"]


def main():
    sr = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    n = 512 if sr == 16000 else 256
    T = 64
    m = load_silero_vad(device=0)
    m.engine.set_kernel("h16")
    m.engine.set_small_batch_max(0)
    x = torch.randn(B, n * T, device="cuda") * 0.03
    nfl = 2 * STEP + 2 + (64 + 160) * 2 + 16
    dbg = torch.zeros(nfl, device="cuda")
    L = _cabi.lib()
    L.svad_engine_set_debug_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    m.audio_forward_device(x, sr)
    L.svad_engine_set_debug_buffer(m.engine._h, dbg.data_ptr())
    m.audio_forward_device(x, sr)
    torch.cuda.synchronize()
    L.svad_engine_set_debug_buffer(m.engine._h, None)
    raw = dbg.cpu().numpy().view(np.uint8)
    off = ((2 * STEP * 4 + 7) // 8) * 8
    st = raw[off: off + 64 * 8].view(np.int64).reshape(4, 16)
    t0 = min(int(st[r][0]) for r in range(4) if st[r][0])
    for r in range(4):
        print(ROLE[r])
        prev = None
        for v, name in sorted((int(st[r][k]), name) for k, name in enumerate(NAMES[r])):
            if not v:
                continue
            print(f"   {name:32s} t = {v - t0:8d}   (+{0 if prev is None else v - prev})")
            prev = v
    spans = raw[off + 64 * 8: off + (64 + 148) * 8].view(np.int64)
    spans = spans[spans > 0]
    span = int(spans.max())
    print(f"CTA spans (own clocks): min {spans.min()} median {int(np.median(spans))} max {span} cycles over {len(spans)} CTAs; "
          f"slowest = {span / T:.0f} cycles per step")
    print("kernel time / steps:")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.audio_forward_device(x, sr)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"   {ms:.3f} ms per launch = {ms * 1e-3 / T * 1.965e9:.0f} cycles per step at 1965 MHz; {B * T / ms * 1e3:.4e} chunks/s; "
          f"implied SM clock of CTA 0 = {span / (ms * 1e-3) / 1e6:.0f} MHz")


if __name__ == "__main__":
    main()
