"""Deterministic synthetic-audio recipes shared by the golden generators (oracle/gen_golden*.py) and the tests.

R1 = /root/reference/examples/onnx_sequence/run.py:159-162 (N(0, 0.03^2) noise, default_rng(17 + sr + stream)).
R2 = /root/reference/examples/openvino/verify.py:31-51 (restated): 3 s silence, 4 s mains hum, 3 s 0.05-noise,
     4 s gated formant-like tones, 3 s chirp, 3 s 0.3-noise, 2 s silence; default_rng(42).
`r2_segments` is the reference harness's thresholder (verify.py:116-127): runs of >= min_chunks chunks with p >= thr.
"""
import numpy as np


def r1_audio(sr, stream, nsamples):
    return (np.random.default_rng(17 + sr + stream).standard_normal(nsamples) * 0.03).astype(np.float32)


def synthetic_r2(sr, rng=None):
    rng = rng if rng is not None else np.random.default_rng(42)

    def t(sec):
        return np.arange(int(sec * sr)) / sr
    parts = [np.zeros(int(3 * sr), np.float32)]
    tt = t(4)
    parts.append((0.02 * np.sin(2 * np.pi * 60 * tt) + 0.01 * np.sin(2 * np.pi * 120 * tt)
                  + 0.005 * np.sin(2 * np.pi * 180 * tt)).astype(np.float32))
    parts.append((0.05 * rng.standard_normal(int(3 * sr))).astype(np.float32))
    tt = t(4)
    env = 0.5 * (1 + np.sign(np.sin(2 * np.pi * 4 * tt)))
    car = np.sin(2 * np.pi * 220 * tt) + 0.6 * np.sin(2 * np.pi * 710 * tt) + 0.3 * np.sin(2 * np.pi * 2400 * tt)
    parts.append((0.15 * env * car + 0.02 * rng.standard_normal(len(tt))).astype(np.float32))
    tt = t(3)
    parts.append((0.1 * np.sin(2 * np.pi * (100 + 900 * tt) * tt)).astype(np.float32))
    parts.append((0.3 * rng.standard_normal(int(3 * sr))).astype(np.float32))
    parts.append(np.zeros(int(2 * sr), np.float32))
    return np.concatenate(parts)


def r2_segments(probs, thr=0.5, min_chunks=8):
    segs, start = [], None
    for i, p in enumerate(probs):
        if p >= thr and start is None:
            start = i
        elif p < thr and start is not None:
            if i - start >= min_chunks:
                segs.append((start, i))
            start = None
    if start is not None and len(probs) - start >= min_chunks:
        segs.append((start, len(probs)))
    return segs
