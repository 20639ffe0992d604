#!/usr/bin/env python
"""Round-2 additions to tests/golden/, generated from the REFERENCE itself (build container only; needs /root/reference).
Kept separate from gen_golden.py so the round-1 fixtures are not rewritten.

  round2.npz / round2.json:
    * R2 audio check: md5 of the regenerated structured signal (tests/recipes.synthetic_r2) -- the r2_{sr}_probs of
      synthetic.npz were computed from exactly this signal -- plus its segmentation under the reference harness's
      thresholder (examples/openvino/verify.py:116-127).
    * short clips (< one window) through the reference's get_speech_timestamps (utils_vad.py:323-327 pads each chunk).
    * collect_chunks / drop_chunks (utils_vad.py:552-646) on tests/data/test.wav with its own segments, samples and seconds.
    * sr = 32000 / 48000 audio_forward and model() (utils_vad.py:39-42 `x[:, ::step]`).
    * R1 bench-workload rows: 64 streams of BASELINE configs[2] (rows spread over the first / last tiles) x 64 chunks at
      both rates, so the GPU test of the bench workload is pinned to the reference and not only to the C oracle.
"""
import hashlib
import json
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = REPO / "tests" / "golden"
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REPO / "tests"))

torch.set_num_threads(1)
from silero_vad import collect_chunks, drop_chunks, get_speech_timestamps, load_silero_vad  # noqa: E402
from recipes import r1_audio, r2_segments, synthetic_r2  # noqa: E402

BENCH_ROWS = [0, 1, 15, 16, 17, 27, 28, 29, 31, 55, 56, 57] + [2048 + i for i in range(8)] + list(range(4096 - 44, 4096))   # 64 rows


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    model = load_silero_vad()
    meta, arr = {}, {}
    syn = np.load(OUT / "synthetic.npz")
    for sr in (16000, 8000):
        r2 = synthetic_r2(sr)
        with torch.no_grad():
            probs = model.audio_forward(torch.from_numpy(r2)[None], sr).numpy()[0]
        assert np.array_equal(probs, syn[f"r2_{sr}_probs"]), "regenerated R2 signal does not reproduce the round-1 golden"
        meta[f"r2_{sr}"] = {"audio_md5": md5(r2), "samples": int(r2.size), "segments": r2_segments(probs),
                           "segments_thr005_min2": r2_segments(probs, thr=0.05, min_chunks=2)}
    # short clips
    z = np.load(OUT / "test16k.npz")
    wav = torch.from_numpy(z["pcm"].astype(np.float32) / 32768.0)
    short = {}
    for sr, lens in ((16000, (1, 400, 511, 512, 513)), (8000, (1, 100, 255, 256, 300))):
        for ln in lens:
            a = wav[44000: 44000 + ln] if sr == 16000 else wav[44000: 44000 + 2 * ln: 2]
            with torch.no_grad():
                ts = get_speech_timestamps(a, model, sampling_rate=sr, min_speech_duration_ms=0, speech_pad_ms=0, threshold=0.02)
            short[f"{sr}_{ln}"] = [[d["start"], d["end"]] for d in ts]
    meta["short_clips"] = short
    # collect / drop
    with torch.no_grad():
        ts = get_speech_timestamps(wav, model)
        ts_s = get_speech_timestamps(wav, model, return_seconds=True)
    c, d = collect_chunks(ts, wav), drop_chunks(ts, wav)
    cs, ds = collect_chunks(ts_s, wav, seconds=True, sampling_rate=16000), drop_chunks(ts_s, wav, seconds=True, sampling_rate=16000)
    meta["chunks"] = {"segments": [[x["start"], x["end"]] for x in ts], "segments_seconds": [[x["start"], x["end"]] for x in ts_s],
                      "collect_len": int(c.numel()), "collect_md5": md5(c.numpy()), "drop_len": int(d.numel()), "drop_md5": md5(d.numpy()),
                      "collect_s_len": int(cs.numel()), "collect_s_md5": md5(cs.numpy()), "drop_s_len": int(ds.numel()), "drop_s_md5": md5(ds.numpy())}
    # decimation
    for sr in (32000, 48000):
        k = sr // 16000
        a = wav[: 16000 * 6].repeat_interleave(k)
        a = a + 0.001 * torch.sin(torch.arange(a.numel()) * 0.37)   # make the dropped samples differ from the kept ones
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            arr[f"decim_{sr}_audio"] = a.numpy()
            arr[f"decim_{sr}_probs"] = model.audio_forward(a[None], sr).numpy()
            model.reset_states()
            y = [model(a[i * 512 * k:(i + 1) * 512 * k], sr).numpy()[0, 0] for i in range(6)]
            arr[f"decim_{sr}_call_probs"] = np.asarray(y, np.float32)
            ts = get_speech_timestamps(a, model, sampling_rate=sr)
            meta[f"decim_{sr}_segments"] = [[x["start"], x["end"]] for x in ts]
    # bench workload rows
    torch.set_num_threads(8)
    for sr in (16000, 8000):
        n = 512 if sr == 16000 else 256
        x = np.stack([r1_audio(sr, b, n * 64) for b in BENCH_ROWS])
        with torch.no_grad():
            arr[f"bench_{sr}_probs"] = model.audio_forward(torch.from_numpy(x), sr).numpy()
    meta["bench_rows"] = BENCH_ROWS
    np.savez_compressed(OUT / "round2.npz", **arr)
    (OUT / "round2.json").write_text(json.dumps(meta, indent=1))
    print("wrote round2.npz / round2.json", {k: v.shape for k, v in arr.items()})
    print(json.dumps(meta["short_clips"]), meta["r2_16000"]["segments"], meta["r2_8000"]["segments"])


if __name__ == "__main__":
    main()
