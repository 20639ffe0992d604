#!/usr/bin/env python
"""Does the row stride of the audio matrix matter?  bench.py's [4096, 32768] fp32 rows are exactly 128 KB apart: the 32 streams of a
tile read the same offset of 32 rows, i.e. addresses that differ only in bits >= 17.  Times svad_fused_h16 on the same audio with
rows padded by `pad` samples (the C ABI takes the row stride).  Usage: h16_stride_probe.py [sr] [pads...]"""
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from silero_vad_b200 import load_silero_vad  # noqa: E402
from silero_vad_b200.model import _ptr  # noqa: E402


def main():
    sr = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
    pads = [int(a) for a in sys.argv[2:]] or [0, 32, 64, 256, 1024, 4096 + 64]
    B, T = 4096, 64
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    L = n * T
    m = load_silero_vad(device=0)
    m.engine.set_kernel("h16")
    m.engine.set_small_batch_max(0)
    if os.environ.get("SVAD_TILE_ROWS"):   # streams per tile = 4 x rows (default: the batch spread over every SM)
        m.engine.set_tile_rows(int(os.environ["SVAD_TILE_ROWS"]))
    g = torch.Generator(device="cuda").manual_seed(1)
    ref = None
    for pad in pads:
        buf = torch.zeros(B, L + pad, device="cuda")
        buf[:, :L] = torch.randn(B, L, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 0.03
        state = torch.zeros(2, B, 128, device="cuda")
        cx = torch.zeros(B, ctx, device="cuda")
        probs = torch.empty(B, T, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def run():
            state.zero_(); cx.zero_()
            m.engine.forward_device_ex(sr, B, L, L + pad, _ptr(buf), 0, 1, _ptr(state), _ptr(cx), _ptr(state), _ptr(cx), _ptr(probs), T, st)

        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if ref is None:
            ref = probs.clone()
        print(f"sr {sr} row stride {L + pad:6d} samples (pad {pad:5d}): {ms:.3f} ms per launch, {B * T / ms * 1e3:.4e} chunks/s, "
              f"max |p - p(pad 0)| = {(probs - ref).abs().max().item():.1e}")


if __name__ == "__main__":
    main()
