"""CPU-only tests of the product's host side: the C ABI library loads and exports every symbol the header
declares, the C++ segment automaton equals the reference on scripted probabilities, the barrier-level CPU
emulation of the fused kernel (same source as the CUDA kernel) matches the oracle, argument validation."""
import ctypes
import re
import subprocess

import numpy as np
import pytest

from conftest import REPO


def test_cabi_exports_every_declared_symbol():
    from silero_vad_b200 import _cabi
    L = _cabi.lib()
    header = (REPO / "include" / "silero_vad_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(svad_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_cabi.EXPORTS) == declared
    assert L.svad_abi_version() == 1


def test_no_gpu_fails_loudly():
    """Without a CUDA device the engine must refuse to construct (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from silero_vad_b200 import _cabi, load_silero_vad
    from silero_vad_b200.model import WEIGHTS
    with pytest.raises(RuntimeError):
        load_silero_vad()
    with pytest.raises(_cabi.SvadError):
        _cabi.Engine(WEIGHTS, 0)
    with pytest.raises(Exception):
        load_silero_vad(onnx=True, opset_version=14)


def test_product_does_not_touch_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    for p in (REPO / "silero_vad_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".h", ".cpp"}:
            assert "oracle" not in p.read_text(), p


def _params(kw, sr):
    from silero_vad_b200.utils_vad import _segment_params
    d = dict(threshold=0.5, neg_threshold=None, min_speech_duration_ms=250, max_speech_duration_s=float("inf"),
             min_silence_duration_ms=100, speech_pad_ms=30, min_silence_at_max_speech=98, use_max_poss_sil_at_max_speech=True)
    d.update(kw)
    return _segment_params(sr, **d)


def test_segments_cpp_equals_reference_cases(sm_cases):
    from silero_vad_b200 import _cabi
    from silero_vad_b200.utils_vad import _finish
    for c in sm_cases:
        sr = c["sampling_rate"]
        step = sr // 16000 if sr > 16000 else 1
        msr = 16000 if sr >= 16000 else sr
        kw = dict(c["kwargs"])
        rs, tr = kw.pop("return_seconds", False), kw.pop("time_resolution", 1)
        alen = c["audio_len"] // step
        segs = _cabi.speech_segments(np.asarray(c["probs"], np.float32)[None], np.asarray([alen]), _params(kw, msr))[0]
        out = _finish(segs, msr, alen, rs, tr, step)
        assert [[d["start"], d["end"]] for d in out] == c["segments"], c["kwargs"]


def test_segments_cpp_equals_oracle_random_batch():
    """Batched call on random probability tracks vs the pure-Python restatement, many parameter sets."""
    from oracle import oracle as O
    from silero_vad_b200 import _cabi
    rng = np.random.default_rng(11)
    for trial in range(40):
        sr = int(rng.choice([16000, 8000]))
        w = 512 if sr == 16000 else 256
        B, T = int(rng.integers(1, 9)), int(rng.integers(1, 300))
        probs = np.clip(np.cumsum(rng.normal(0, 0.25, (B, T)), axis=1) % 2.0, 0, 1).astype(np.float32)
        lens = np.array([int(rng.integers(1, T * w + 1)) for _ in range(B)])
        kw = dict(threshold=float(rng.choice([0.5, 0.35, 0.8])), min_speech_duration_ms=int(rng.choice([0, 250, 700])),
                  min_silence_duration_ms=int(rng.choice([0, 100, 400])), speech_pad_ms=int(rng.choice([0, 30, 150])),
                  max_speech_duration_s=float(rng.choice([float("inf"), 1.5, 3.0])),
                  use_max_poss_sil_at_max_speech=bool(rng.integers(0, 2)), min_silence_at_max_speech=int(rng.choice([98, 10])))
        got = _cabi.speech_segments(probs, lens, _params(kw, sr))
        for b in range(B):
            Tb = (lens[b] + w - 1) // w
            want = O.get_speech_timestamps(probs[b, :Tb].tolist(), int(lens[b]), sampling_rate=sr, **kw)
            assert got[b] == [(d["start"], d["end"]) for d in want], (trial, b, kw)


@pytest.fixture(scope="module")
def emu():
    so = REPO / "tests" / "emu" / "libsvad_emu.so"
    src = REPO / "tests" / "emu" / "svad_emu.cpp"
    deps = [src] + list((REPO / "silero_vad_b200" / "csrc").glob("*.h"))
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    fp = ctypes.c_void_p
    lib.svad_emu_forward.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long, fp, ctypes.c_int, fp, fp, fp, fp, fp]
    lib.svad_emu_forward_tc.argtypes = lib.svad_emu_forward.argtypes
    return lib


@pytest.mark.parametrize("sr,rm,B", [(16000, 8, 5), (16000, 7, 37), (8000, 8, 5), (8000, 5, 23), (16000, 4, 1)])
def test_emulated_kernel_matches_oracle(emu, oracle, fixtures, sr, rm, B):
    """The CUDA kernel's per-thread code and barrier schedule, executed on CPU threads, vs the oracle:
    chained state over 5 chunks, ragged tail, random carried-in state and context."""
    from silero_vad_b200.model import WEIGHTS
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    a = fixtures["test16k"]["audio"]
    dec = 1 if sr == 16000 else 2
    x = np.stack([a[(9000 * b)::dec][: n * 5 - 37] for b in range(B)]).copy()
    st = (np.random.default_rng(1).standard_normal((2, B, 128)) * 0.1).astype(np.float32)
    cx = (np.random.default_rng(2).standard_normal((B, ctx)) * 0.1).astype(np.float32)
    st_o, cx_o = st.copy(), cx.copy()
    want = oracle.audio_forward(x, sr, state=st_o, context=cx_o, nthreads=2)
    probs = np.zeros_like(want)
    st_e, cx_e = np.zeros_like(st), np.zeros_like(cx)
    rc = emu.svad_emu_forward(str(WEIGHTS).encode(), sr, rm, B, x.shape[1], x.ctypes.data, 0, st.ctypes.data, cx.ctypes.data,
                              st_e.ctypes.data, cx_e.ctypes.data, probs.ctypes.data)
    assert rc == 0
    assert np.abs(probs - want).max() < 2e-5
    assert np.abs(st_e - st_o).max() < 2e-5
    assert np.array_equal(cx_e, cx_o)


@pytest.mark.parametrize("sr", [16000, 8000])
def test_tensor_core_weight_stream_tables(emu, sr):
    """TapeTC (svad_pack.h): the slabs tile the tape, every buffer lies in its region and is 1 KB aligned, the e0-region buffers
    are not refilled before enc1 (their last reader) is done, and no copy can land on a slab that may still be unconsumed
    when it is issued (dep_delta vs the buffer map, simulated over three steps)."""
    msg = ctypes.create_string_buffer(256)
    rc = emu.svad_emu_tape_check(sr, msg, 256)
    assert rc == 0, msg.value.decode()


@pytest.mark.parametrize("sr,rm,B", [(16000, 8, 5), (16000, 7, 37), (8000, 8, 5)])
def test_emulated_tensor_core_kernel_matches_oracle(emu, oracle, fixtures, sr, rm, B):
    """The tensor-core schedule (svad_tc.h) on CPU threads: tcgen05.mma modelled as truncated-TF32 products read through
    the declared swizzled layouts, split precision hi/lo, TMEM as an array.  Checks layouts, ring hand-off, epilogues."""
    from silero_vad_b200.model import WEIGHTS
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    a = fixtures["test16k"]["audio"]
    dec = 1 if sr == 16000 else 2
    x = np.stack([a[(9000 * b)::dec][: n * 4 - 37] for b in range(B)]).copy()
    st = (np.random.default_rng(1).standard_normal((2, B, 128)) * 0.1).astype(np.float32)
    cx = (np.random.default_rng(2).standard_normal((B, ctx)) * 0.1).astype(np.float32)
    st_o, cx_o = st.copy(), cx.copy()
    want = oracle.audio_forward(x, sr, state=st_o, context=cx_o, nthreads=2)
    probs = np.zeros_like(want)
    st_e, cx_e = np.zeros_like(st), np.zeros_like(cx)
    rc = emu.svad_emu_forward_tc(str(WEIGHTS).encode(), sr, rm, B, x.shape[1], x.ctypes.data, 0, st.ctypes.data, cx.ctypes.data,
                                 st_e.ctypes.data, cx_e.ctypes.data, probs.ctypes.data)
    assert rc == 0
    assert np.abs(probs - want).max() < 2e-5
    assert np.abs(st_e - st_o).max() < 5e-5
    assert np.array_equal(cx_e, cx_o)


@pytest.mark.parametrize("sr,L", [(16000, 300), (16000, 700), (8000, 256)])
def test_emulated_tensor_core_kernel_short_streams(emu, oracle, fixtures, sr, L):
    """One- and two-chunk streams through the tensor-core schedule: no next chunk to prefetch, zero-padded tail, zero initial
    state (null pointers), the weight stream wound up after a single step."""
    from silero_vad_b200.model import WEIGHTS
    a = fixtures["test16k"]["audio"]
    x = np.stack([a[40000 + 7000 * b:][:L] for b in range(3)]).copy()
    want = oracle.audio_forward(x, sr, nthreads=2)
    probs = np.zeros_like(want)
    st_e = np.zeros((2, 3, 128), np.float32)
    rc = emu.svad_emu_forward_tc(str(WEIGHTS).encode(), sr, 8, 3, L, x.ctypes.data, 0, None, None, st_e.ctypes.data, None, probs.ctypes.data)
    assert rc == 0
    assert np.abs(probs - want).max() < 2e-5


def test_emulated_tensor_core_kernel_pcm16_equals_f32(emu, fixtures):
    """int16 PCM ingest of the tensor-core schedule: bit-identical to feeding int16/32768 as fp32 (rows of 7-row tiles)."""
    from silero_vad_b200.model import WEIGHTS
    pcm = np.stack([fixtures["test16k"]["pcm"][20000 * b: 20000 * b + 512 * 3 + 100] for b in range(3)]).copy()
    f32 = pcm.astype(np.float32) / 32768.0
    out = []
    for arr, flag in ((f32, 0), (pcm, 1)):
        p = np.zeros((3, 4), np.float32)
        assert emu.svad_emu_forward_tc(str(WEIGHTS).encode(), 16000, 7, 3, arr.shape[1], arr.ctypes.data, flag, None, None, None, None,
                                       p.ctypes.data) == 0
        out.append(p)
    assert np.array_equal(out[0], out[1]) and out[0].max() > 0.5


def test_emulated_kernel_pcm16_equals_f32(emu, fixtures):
    """int16 PCM ingest (scaled by 2^-15 on load) must be bit-identical to feeding int16/32768 as fp32."""
    from silero_vad_b200.model import WEIGHTS
    pcm = np.stack([fixtures["test16k"]["pcm"][20000 * b: 20000 * b + 512 * 3 + 100] for b in range(3)]).copy()
    f32 = pcm.astype(np.float32) / 32768.0
    out = []
    for arr, flag in ((f32, 0), (pcm, 1)):
        p = np.zeros((3, 4), np.float32)
        assert emu.svad_emu_forward(str(WEIGHTS).encode(), 16000, 7, 3, arr.shape[1], arr.ctypes.data, flag, None, None, None, None,
                                    p.ctypes.data) == 0
        out.append(p)
    assert np.array_equal(out[0], out[1]) and out[0].max() > 0.5


def test_validate_input_messages():
    from silero_vad_b200.model import SileroVADB200
    import torch
    m = SileroVADB200.__new__(SileroVADB200)     # host-side validation needs no device
    m.sample_rates = [8000, 16000]
    x, sr = m._validate_input(torch.zeros(1024), 32000)
    assert tuple(x.shape) == (1, 512) and sr == 16000
    with pytest.raises(ValueError, match="Too many dimensions"):
        m._validate_input(torch.zeros(1, 1, 512), 16000)
    with pytest.raises(ValueError, match="Supported sampling rates"):
        m._validate_input(torch.zeros(512), 22050)
    with pytest.raises(ValueError, match="too short"):
        m._validate_input(torch.zeros(100), 16000)


def test_header_is_valid_c():
    """include/silero_vad_b200.h must compile as plain C (it is what cgo / JNI / Rust bindgen consume)."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(REPO / "include" / "silero_vad_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


class _ScriptedModel:
    """Duck-typed model that replays scripted probabilities [T][B] (host logic tests need no GPU)."""

    def __init__(self, probs):
        import torch
        self.p, self.i, self.torch = probs, 0, torch

    def reset_states(self):
        self.i = 0

    def __call__(self, x, sr):
        row = self.p[self.i]
        self.i += 1
        return self.torch.tensor(row, dtype=self.torch.float32).reshape(-1, 1)


def test_vad_iterator_batch_equals_independent_iterators():
    import torch
    from oracle import oracle as O
    from silero_vad_b200 import VADIterator, VADIteratorBatch
    rng = np.random.default_rng(5)
    B, T = 7, 400
    probs = np.clip(np.cumsum(rng.normal(0, 0.2, (T, B)), axis=0) % 2.0, 0, 1).astype(np.float32)
    it = VADIteratorBatch(_ScriptedModel(probs), B)
    got = [[] for _ in range(B)]
    for t in range(T):
        for b, e in enumerate(it(torch.zeros(B, 512))):
            got[b].append(e)
    for b in range(B):
        single = VADIterator(_ScriptedModel(probs[:, b:b + 1]))
        want = [single(torch.zeros(512)) for _ in range(T)]
        assert got[b] == want
        assert want == O.vad_iterator_events(probs[:, b].tolist(), 512)


def test_hub_entry_point_signature_and_argument_checks():
    """hubconf.silero_vad mirrors the reference hub entry (hubconf.py:26-56): same parameters, same utils tuple, the
    reference's exception for an unknown ONNX opset; constructing the engine needs a GPU and must fail loudly without one."""
    import inspect
    import hubconf
    assert list(inspect.signature(hubconf.silero_vad).parameters) == ["onnx", "force_onnx_cpu", "opset_version"]
    with pytest.raises(Exception, match="Available ONNX opset_version"):
        hubconf.silero_vad(onnx=True, opset_version=14)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            hubconf.silero_vad()


def test_reference_model_constructors_exist():
    """utils_vad.init_jit_model / OnnxWrapper (reference utils_vad.py:10-31, 194-199) exist with the reference's parameters and
    refuse a CPU device instead of falling back."""
    import inspect
    import torch
    from silero_vad_b200 import utils_vad as U
    assert list(inspect.signature(U.init_jit_model).parameters) == ["model_path", "device"]
    assert list(inspect.signature(U.OnnxWrapper.__init__).parameters) == ["self", "path", "force_onnx_cpu"]
    with pytest.raises(RuntimeError, match="CUDA device only"):
        U.init_jit_model("silero_vad.jit", torch.device("cpu"))


def test_segments_threaded_batch_equals_per_row():
    """svad_speech_segments cuts large batches into row ranges, one host thread each: same lists as one row at a time."""
    from silero_vad_b200 import _cabi
    rng = np.random.default_rng(11)
    B, T = 700, 90
    probs = rng.uniform(0, 1, (B, T)).astype(np.float32)
    probs[rng.uniform(size=(B, T)) < 0.5] *= 0.2
    lens = rng.integers(1, T * 512 + 1, B).astype(np.int64)
    p = _params({}, 16000)
    whole = _cabi.speech_segments(probs, lens, p)
    for b in range(0, B, 37):
        assert _cabi.speech_segments(probs[b:b + 1], lens[b:b + 1], p)[0] == whole[b], b


def test_collect_chunks_plan_sizes_like_python_slices():
    """Sizing pass of svad_collect_chunks_device (host only, no engine): per-row output lengths equal what torch slicing gives."""
    import ctypes
    from silero_vad_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(3)
    B, ld = 5, 1000
    row_len = np.asarray([1000, 0, 700, 1000, 10], np.int64)
    rows, bounds = [], []
    for b in range(B):
        for _ in range(int(rng.integers(0, 6))):
            a, e = sorted(rng.integers(0, 1300, 2).tolist())
            rows.append(b); bounds.append([a, e])
    rows, bounds = np.asarray(rows, np.int64), np.asarray(bounds, np.int64).reshape(-1, 2)
    for drop in (0, 1):
        offs = np.zeros(B + 1, np.int64)
        rc = L.svad_collect_chunks_device(None, None, 4, B, ld, row_len.ctypes.data, rows.ctypes.data, bounds.ctypes.data, len(rows), drop,
                                          None, 0, offs.ctypes.data, None)
        assert rc == 0, L.svad_last_error()
        for b in range(B):
            w = np.arange(row_len[b])
            tab = bounds[rows == b]
            if drop:
                parts, cur = [], 0
                for a, e in tab:
                    parts.append(w[cur:a]); cur = e
                parts.append(w[cur:])
                want = sum(len(x) for x in parts)
            else:
                want = sum(len(w[a:e]) for a, e in tab)
            assert offs[b + 1] - offs[b] == want, (drop, b)
    bad = np.asarray([[-1, 5]], np.int64)
    offs = np.zeros(2, np.int64)
    assert L.svad_collect_chunks_device(None, None, 4, 1, ld, row_len.ctypes.data, np.zeros(1, np.int64).ctypes.data, bad.ctypes.data, 1, 0,
                                        None, 0, offs.ctypes.data, None) == -1


def test_corrupt_weight_container_is_an_error_not_a_crash(tmp_path):
    """svad_engine_create: malformed containers return SVAD_EWEIGHTS (no exception across the C boundary, no huge allocation)."""
    import ctypes
    import struct
    from silero_vad_b200 import _cabi
    from silero_vad_b200.model import WEIGHTS
    L = _cabi.lib()
    good = WEIGHTS.read_bytes()
    cases = {"trunc": good[: len(good) // 2], "magic": b"XXXXXXXX" + good[8:],
             "hugedim": good[:12] + struct.pack("<I", 5) + b"hello" + struct.pack("<I", 2) + struct.pack("<II", 0x7fffffff, 0x7fffffff),
             "count": good[:8] + struct.pack("<I", 0xffffffff) + good[12:]}
    for name, blob in cases.items():
        f = tmp_path / f"{name}.weights"
        f.write_bytes(blob)
        h = ctypes.c_void_p()
        rc = L.svad_engine_create(str(f).encode(), 0, ctypes.byref(h))
        assert rc == -2 and not h.value, (name, rc, L.svad_last_error())
    h = ctypes.c_void_p()
    assert L.svad_engine_create(str(tmp_path / "missing").encode(), 0, ctypes.byref(h)) == -2


def test_build_hash_tracks_sources():
    """_cabi.lib() rebuilds when the sources change: the staleness check is a content hash, not file times."""
    from silero_vad_b200 import build as B
    assert not B.stale()
    assert B.HASH.read_text().strip() == B.source_hash()
