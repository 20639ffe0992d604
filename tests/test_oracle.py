"""Pins the CPU oracle (oracle/) against golden outputs of the reference TorchScript model and of the
reference's own get_speech_timestamps / VADIterator (tests/golden/, made by oracle/gen_golden.py)."""
import numpy as np
import pytest

from oracle import oracle as O

TOL = 2e-5  # fp32 summation-order noise between ATen and the C restatement (observed ~2e-6)


@pytest.mark.parametrize("name", ["test16k", "aepyx16k", "aepyx8k"])
def test_bulk_probs_match_reference(oracle, fixtures, name):
    fx = fixtures[name]
    p = oracle.audio_forward(fx["audio"], fx["sr"])[0]
    assert p.shape == fx["probs"].shape
    assert np.abs(p - fx["probs"]).max() < TOL


@pytest.mark.parametrize("name", ["test16k", "aepyx16k", "aepyx8k"])
def test_segments_match_reference(oracle, fixtures, meta, name):
    fx = fixtures[name]
    p = oracle.audio_forward(fx["audio"], fx["sr"])[0]
    ts = O.get_speech_timestamps(p.tolist(), len(fx["audio"]), sampling_rate=fx["sr"])
    assert [[d["start"], d["end"]] for d in ts] == meta[name]["segments"]


def test_8k_decimated(oracle, fixtures):
    gold = np.load(O._REPO / "tests/golden/test16k_as8k_probs.npz")["probs"]
    p = oracle.audio_forward(fixtures["test16k"]["audio"][::2].copy(), 8000)[0]
    assert np.abs(p - gold).max() < TOL


@pytest.mark.parametrize("sr", [16000, 8000])
def test_chained_stateless_contract(oracle, synthetic, sr):
    s = synthetic
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    audio, st, cx = s[f"r1_{sr}_audio"], s[f"r1_{sr}_state0"].copy(), s[f"r1_{sr}_ctx0"].copy()
    T = audio.shape[1] // n
    # (a) per-chunk stateless step
    st_a, cx_a, probs = st.copy(), cx.copy(), []
    for t in range(T):
        x1 = np.concatenate([cx_a, audio[:, t * n:(t + 1) * n]], 1)
        out, st_a = oracle.step(x1, st_a, sr)
        cx_a = x1[:, -ctx:]
        probs.append(out)
    probs = np.stack(probs, 1)
    assert np.abs(probs - s[f"r1_{sr}_probs"]).max() < TOL
    assert np.abs(st_a - s[f"r1_{sr}_stateN"]).max() < TOL
    # (b) bulk path with carried state/context
    pb = oracle.audio_forward(audio, sr, state=st, context=cx)
    assert np.abs(pb - s[f"r1_{sr}_probs"]).max() < TOL
    assert np.abs(st - s[f"r1_{sr}_stateN"]).max() < TOL
    assert np.array_equal(cx, s[f"r1_{sr}_ctxN"])


@pytest.mark.parametrize("sr", [16000, 8000])
def test_ragged_bulk(oracle, synthetic, sr):
    p = oracle.audio_forward(synthetic[f"ragged_{sr}_audio"], sr)
    assert p.shape == synthetic[f"ragged_{sr}_probs"].shape
    assert np.abs(p - synthetic[f"ragged_{sr}_probs"]).max() < TOL


def test_wrapper_protocol(oracle, synthetic, meta):
    m = O.OracleModel(oracle)
    for i, (B, sr) in enumerate(meta["protocol_calls"]):
        y = m(synthetic[f"proto_x{i}"], sr)
        assert y.shape == synthetic[f"proto_y{i}"].shape
        assert np.abs(y - synthetic[f"proto_y{i}"]).max() < TOL, (i, B, sr)


def test_state_machine_cases(sm_cases):
    for c in sm_cases:
        sr = c["sampling_rate"]
        step = sr // 16000 if sr > 16000 else 1
        msr = 16000 if sr >= 16000 else sr
        ts = O.get_speech_timestamps(c["probs"], c["audio_len"] // step if step > 1 else c["audio_len"],
                                     sampling_rate=msr, step=step, **c["kwargs"])
        assert [[d["start"], d["end"]] for d in ts] == c["segments"], c["kwargs"]


def test_variants_and_iterator(oracle, fixtures, meta):
    fx = fixtures["test16k"]
    v = meta["test16k"]["variants"]
    p = oracle.audio_forward(fx["audio"], 16000)[0].tolist()
    L = len(fx["audio"])
    seg = lambda ts: [[d["start"], d["end"]] for d in ts]
    assert seg(O.get_speech_timestamps(p, L, max_speech_duration_s=5)) == v["max_speech_5"]
    assert seg(O.get_speech_timestamps(p, L, max_speech_duration_s=5, use_max_poss_sil_at_max_speech=False)) == v["max_speech_5_legacy"]
    assert seg(O.get_speech_timestamps(p, L, threshold=0.3)) == v["threshold_03"]
    assert seg(O.get_speech_timestamps(p, L, return_seconds=True)) == v["seconds"]
    assert seg(O.get_speech_timestamps(p, L, return_seconds=True, time_resolution=3)) == v["seconds_res3"]
    ev = [e for e in O.vad_iterator_events(p, 512) if e]
    assert ev == meta["test16k"]["vad_iterator_events"]
    ev = [e for e in O.vad_iterator_events(p, 512, return_seconds=True) if e]
    assert ev == meta["test16k"]["vad_iterator_events_seconds"]
    pr = oracle.audio_forward(fx["audio"][:200_123], 16000)[0].tolist()
    assert seg(O.get_speech_timestamps(pr, 200_123)) == v["ragged_tail"]


@pytest.mark.parametrize("sr", [16000, 8000])
def test_r2_structured_signal(oracle, synthetic, sr):
    """Reference harness examples/openvino/verify.py:31-51,157-182: the structured 22 s synthetic signal, chained state,
    max-abs below tolerance AND identical segmentation under the harness's thresholder (verify.py:116-127)."""
    import hashlib
    import json
    from recipes import r2_segments, synthetic_r2
    meta = json.loads((O._REPO / "tests/golden/round2.json").read_text())[f"r2_{sr}"]
    audio = synthetic_r2(sr)
    assert hashlib.md5(audio.tobytes()).hexdigest() == meta["audio_md5"]
    p = oracle.audio_forward(audio, sr)[0]
    assert p.shape == synthetic[f"r2_{sr}_probs"].shape
    assert np.abs(p - synthetic[f"r2_{sr}_probs"]).max() < TOL
    assert [list(s) for s in r2_segments(p)] == meta["segments"]
    assert [list(s) for s in r2_segments(p, thr=0.05, min_chunks=2)] == meta["segments_thr005_min2"]


@pytest.mark.parametrize("sr", [16000, 8000])
def test_bench_workload_rows(oracle, sr):
    """The 64 rows of bench.py's workload (R1 noise, 64 chunks) that the GPU test checks: oracle vs the reference model."""
    import json
    from recipes import r1_audio
    rows = json.loads((O._REPO / "tests/golden/round2.json").read_text())["bench_rows"]
    gold = np.load(O._REPO / "tests/golden/round2.npz")[f"bench_{sr}_probs"]
    n = 512 if sr == 16000 else 256
    x = np.stack([r1_audio(sr, b, n * 64) for b in rows])
    assert np.abs(oracle.audio_forward(x, sr, nthreads=8) - gold).max() < TOL


@pytest.mark.parametrize("sr", [32000, 48000])
def test_decimated_rates(oracle, sr):
    """sr = k * 16000 (utils_vad.py:39-42): the oracle on audio[::k] against the reference called with the full-rate audio."""
    z = np.load(O._REPO / "tests/golden/round2.npz")
    p = oracle.audio_forward(z[f"decim_{sr}_audio"][:: sr // 16000].copy(), 16000)
    assert np.abs(p - z[f"decim_{sr}_probs"]).max() < TOL
