#!/usr/bin/env python
"""Generate tests/golden/* from the REFERENCE itself (build container only; needs /root/reference).

The reference's own tests pin no numbers (tests/test_basic.py asserts `is not None`), so the golden
vectors are outputs of the reference TorchScript model (`load_silero_vad()`, the default loader,
src/silero_vad/model.py:17,34) and of the reference's own `get_speech_timestamps` / `VADIterator`
(src/silero_vad/utils_vad.py:212-549), run here on torch CPU with one thread.  onnxruntime is not
installed in this image, so the ONNX twin cannot be executed (its weights are bit-identical).

Audio fixtures are the reference's test clips re-packed as int16 npz (WAV decoding convention
int16/32768: examples/onnx_sequence/run.py:104-119, examples/cpp/wav.h:95-136).

Usage:  python oracle/gen_golden.py        (rewrites tests/golden/)
"""
import hashlib
import json
import sys
import wave
import warnings
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = REPO / "tests" / "golden"
sys.path.insert(0, str(REF / "src"))

torch.set_num_threads(1)
from silero_vad import VADIterator, get_speech_timestamps, load_silero_vad  # noqa: E402


def load_wav(path):
    w = wave.open(str(path))
    assert w.getnchannels() == 1 and w.getsampwidth() == 2
    return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy(), w.getframerate()


def synthetic_r2(sr, rng):
    """Structured synthetic signal, recipe of examples/openvino/verify.py:31-51 (restated)."""
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


def segs(ts):
    return [[d["start"], d["end"]] for d in ts]


class FakeModel:
    """Feeds a fixed probability sequence through the reference's post-processing."""

    def __init__(self, probs):
        self.probs, self.i = probs, 0

    def reset_states(self):
        self.i = 0

    def __call__(self, chunk, sr):
        p = self.probs[self.i]
        self.i += 1
        return torch.tensor([[p]], dtype=torch.float32)


def prob_sequence(rng, n):
    """Piecewise speech / silence / hovering-near-threshold probability track."""
    out = []
    while len(out) < n:
        kind = rng.integers(0, 4)
        ln = int(rng.integers(1, 60))
        if kind == 0:
            seg = rng.uniform(0.0, 0.3, ln)
        elif kind == 1:
            seg = rng.uniform(0.6, 1.0, ln)
        elif kind == 2:
            seg = rng.uniform(0.3, 0.6, ln)
        else:
            seg = rng.uniform(0.0, 1.0, ln)
        out.extend(seg.tolist())
    return np.asarray(out[:n], np.float32)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    model = load_silero_vad()
    meta = {}

    # ------------------------------------------------------------------ WAV fixtures
    fixtures = {
        "test16k": REF / "tests/data/test.wav",
        "aepyx16k": REF / "examples/c++/aepyx.wav",
        "aepyx8k": REF / "examples/c++/aepyx_8k.wav",
    }
    for name, path in fixtures.items():
        pcm, sr = load_wav(path)
        audio = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
        with torch.no_grad():
            probs = model.audio_forward(audio[None], sr).numpy()[0]
            ts = get_speech_timestamps(audio, model, sampling_rate=sr)
        rec = {"sr": sr, "samples": int(pcm.size), "chunks": int(probs.size), "sum_p": float(probs.astype(np.float64).sum()),
               "n_ge_05": int((probs >= 0.5).sum()), "segments": segs(ts),
               "segments_md5": hashlib.md5(str([(a, b) for a, b in segs(ts)]).encode()).hexdigest()}
        if name == "test16k":
            with torch.no_grad(), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rec["variants"] = {
                    "max_speech_5": segs(get_speech_timestamps(audio, model, max_speech_duration_s=5)),
                    "max_speech_5_legacy": segs(get_speech_timestamps(audio, model, max_speech_duration_s=5,
                                                                      use_max_poss_sil_at_max_speech=False)),
                    "threshold_03": segs(get_speech_timestamps(audio, model, threshold=0.3)),
                    "seconds": segs(get_speech_timestamps(audio, model, return_seconds=True)),
                    "seconds_res3": segs(get_speech_timestamps(audio, model, return_seconds=True, time_resolution=3)),
                    "sr8000_decimated": segs(get_speech_timestamps(audio[::2], model, sampling_rate=8000)),
                    "sr32000_interleaved": segs(get_speech_timestamps(audio.repeat_interleave(2), model, sampling_rate=32000)),
                    "ragged_tail": segs(get_speech_timestamps(audio[:200_123], model)),
                }
                it = VADIterator(model)
                ev = []
                for i in range(0, len(audio) - 511, 512):
                    e = it(audio[i:i + 512])
                    if e:
                        ev.append(e)
                it.reset_states()
                rec["vad_iterator_events"] = ev
                it = VADIterator(model)
                ev = []
                for i in range(0, len(audio) - 511, 512):
                    e = it(audio[i:i + 512], return_seconds=True)
                    if e:
                        ev.append(e)
                rec["vad_iterator_events_seconds"] = ev
                probs8 = model.audio_forward(audio[None, ::2], 8000).numpy()[0]
            np.savez_compressed(OUT / "test16k_as8k_probs.npz", probs=probs8)
        meta[name] = rec
        np.savez_compressed(OUT / f"{name}.npz", pcm=pcm, sr=np.int32(sr), probs=probs)
        print(name, sr, pcm.size, probs.size, len(ts), rec["sum_p"])

    # ------------------------------------------------------------------ synthetic, chained, stateless contract
    # R1 (examples/onnx_sequence/run.py:159-169): N(0, 0.03^2) audio, random initial state * 0.01.
    # R2 (examples/openvino/verify.py:31-51): structured 22 s signal, default_rng(42).
    syn = {}
    for sr in (16000, 8000):
        n, ctx = (512, 64) if sr == 16000 else (256, 32)
        net = model._model if sr == 16000 else model._model_8k
        B, T = 5, 48
        audio = np.stack([(np.random.default_rng(17 + sr + b).standard_normal(n * T) * 0.03).astype(np.float32) for b in range(B)])
        audio[3] *= 10.0      # louder rows so the probabilities leave the floor
        audio[4] = synthetic_r2(sr, np.random.default_rng(42))[7 * sr: 7 * sr + n * T]
        state0 = (np.random.default_rng(29 + sr).standard_normal((2, B, 128)) * 0.01).astype(np.float32)
        ctx0 = (np.random.default_rng(31 + sr).standard_normal((B, ctx)) * 0.03).astype(np.float32)
        st, cx = torch.from_numpy(state0), torch.from_numpy(ctx0)
        probs = []
        with torch.no_grad():
            for t in range(T):
                x1 = torch.cat([cx, torch.from_numpy(audio[:, t * n:(t + 1) * n])], 1)
                out, st = net(x1, st)
                cx = x1[:, -ctx:]
                probs.append(out.numpy()[:, 0])
        syn[f"r1_{sr}_audio"] = audio
        syn[f"r1_{sr}_state0"] = state0
        syn[f"r1_{sr}_ctx0"] = ctx0
        syn[f"r1_{sr}_probs"] = np.stack(probs, 1)
        syn[f"r1_{sr}_stateN"] = st.numpy()
        syn[f"r1_{sr}_ctxN"] = cx.numpy()
        r2 = synthetic_r2(sr, np.random.default_rng(42))
        with torch.no_grad():
            syn[f"r2_{sr}_probs"] = model.audio_forward(torch.from_numpy(r2)[None], sr).numpy()[0]
        # ragged bulk call, B=3, L not a multiple of n (zero-padded tail, utils_vad.py:100-102)
        Lr = n * 7 + 101
        rag = np.stack([r2[5 * sr + b * 1000: 5 * sr + b * 1000 + Lr] for b in range(3)]) * np.float32(2.0)
        with torch.no_grad():
            syn[f"ragged_{sr}_probs"] = model.audio_forward(torch.from_numpy(rag), sr).numpy()
        syn[f"ragged_{sr}_audio"] = rag
    # wrapper protocol: implicit resets on sr / batch change, 1-D input, sr = 2*16000 decimation
    rng = np.random.default_rng(7)
    calls, outs = [], []
    seq = [(2, 16000)] * 3 + [(2, 8000)] * 2 + [(3, 8000)] * 2 + [(1, 16000)] * 2 + [(1, 32000)] * 2 + [(0, 16000)] * 2
    model.reset_states()
    with torch.no_grad():
        for i, (B, sr) in enumerate(seq):
            n = {16000: 512, 8000: 256, 32000: 1024}[sr]
            shape = (n,) if B == 0 else (B, n)         # B == 0 encodes a 1-D chunk
            x = (rng.standard_normal(shape) * 0.2).astype(np.float32)
            y = model(torch.from_numpy(x), sr).numpy()
            syn[f"proto_x{i}"] = x
            syn[f"proto_y{i}"] = y
            calls.append([B, sr])
    meta["protocol_calls"] = calls
    np.savez_compressed(OUT / "synthetic.npz", **syn)

    # ------------------------------------------------------------------ state machine on scripted probabilities
    rng = np.random.default_rng(2024)
    cases = []
    for k in range(64):
        sr = [16000, 8000, 32000, 48000][k % 4] if k % 3 == 0 else 16000
        step = sr // 16000 if sr > 16000 else 1
        msr = 16000 if sr >= 16000 else 8000
        w = 512 if msr == 16000 else 256
        nchunks = int(rng.integers(1, 400))
        probs = prob_sequence(rng, nchunks)
        tail = int(rng.integers(0, w))                       # ragged last chunk
        alen_model = (nchunks - 1) * w + (tail if tail else w)
        alen = alen_model * step
        kw = {}
        if k % 2:
            kw["max_speech_duration_s"] = float(rng.choice([1.0, 2.5, 4.0]))
        if k % 4 == 3:
            kw["use_max_poss_sil_at_max_speech"] = False
        if k % 5 == 0:
            kw["threshold"] = float(rng.choice([0.3, 0.7, 0.1]))
        if k % 7 == 0:
            kw["neg_threshold"] = float(rng.choice([0.2, 0.45]))
        if k % 6 == 0:
            kw["min_silence_duration_ms"] = int(rng.choice([0, 50, 300]))
            kw["min_speech_duration_ms"] = int(rng.choice([0, 100, 500]))
        if k % 8 == 0:
            kw["speech_pad_ms"] = int(rng.choice([0, 100, 200]))
        if k % 9 == 0:
            kw["return_seconds"] = True
            kw["time_resolution"] = int(rng.choice([1, 2, 3]))
        if k % 11 == 0:
            kw["min_silence_at_max_speech"] = int(rng.choice([0, 30, 200]))
        audio = torch.zeros(alen)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ts = get_speech_timestamps(audio, FakeModel(probs.tolist()), sampling_rate=sr, **kw)
        cases.append({"sampling_rate": sr, "audio_len": alen, "probs": [float(p) for p in probs], "kwargs": kw, "segments": segs(ts)})
    meta["state_machine_cases"] = len(cases)
    (OUT / "state_machine_cases.json").write_text(json.dumps(cases))
    (OUT / "meta.json").write_text(json.dumps(meta, indent=1))
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
