"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the golden
outputs of the reference TorchScript model (tests/golden/) and against the CPU oracle on seeded inputs.

Tolerance: BASELINE.json north_star asks for per-chunk probabilities within 1e-4 max-abs of the reference
and bit-exact segment indices; the fp32 kernel is expected near 1e-6 and the tests print what they see."""
import ctypes
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4
TIGHT = 2e-5   # what an all-fp32 implementation should reach; failing this but not TOL is worth a look


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module", params=["auto", "h16", "tc", "fp32"])
def model(torch_cuda, request):
    """All kernels behind the same surface.  auto = the engine's defaults (small batches on the cluster kernel, large
    ones on the tensor-core tile kernel); tc / fp32 force the respective tile kernel for every batch size."""
    from silero_vad_b200 import load_silero_vad
    m = load_silero_vad(device=0)
    if request.param != "auto":
        m.engine.set_kernel(request.param)
        m.engine.set_small_batch_max(0)
    return m


def seg(ts):
    return [[d["start"], d["end"]] for d in ts]


@pytest.mark.parametrize("name", ["test16k", "aepyx16k", "aepyx8k"])
def test_fixture_probs_and_segments(torch_cuda, model, fixtures, meta, name):
    torch = torch_cuda
    from silero_vad_b200 import get_speech_timestamps
    fx = fixtures[name]
    wav = torch.from_numpy(fx["audio"])
    p = model.audio_forward(wav[None], fx["sr"]).numpy()[0]
    err = float(np.abs(p - fx["probs"]).max())
    print(f"{name}: max|p - p_ref| = {err:.3e} over {p.size} chunks")
    assert p.shape == fx["probs"].shape and err < TOL
    assert err < TIGHT
    ts = get_speech_timestamps(wav, model, sampling_rate=fx["sr"])
    assert seg(ts) == meta[name]["segments"]


def test_variants(torch_cuda, model, fixtures, meta):
    torch = torch_cuda
    import warnings
    from silero_vad_b200 import VADIterator, get_speech_timestamps
    wav = torch.from_numpy(fixtures["test16k"]["audio"])
    v = meta["test16k"]["variants"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert seg(get_speech_timestamps(wav, model, max_speech_duration_s=5)) == v["max_speech_5"]
        assert seg(get_speech_timestamps(wav, model, max_speech_duration_s=5, use_max_poss_sil_at_max_speech=False)) == v["max_speech_5_legacy"]
        assert seg(get_speech_timestamps(wav, model, threshold=0.3)) == v["threshold_03"]
        assert seg(get_speech_timestamps(wav, model, return_seconds=True)) == v["seconds"]
        assert seg(get_speech_timestamps(wav, model, return_seconds=True, time_resolution=3)) == v["seconds_res3"]
        assert seg(get_speech_timestamps(wav[::2], model, sampling_rate=8000)) == v["sr8000_decimated"]
        assert seg(get_speech_timestamps(wav.repeat_interleave(2), model, sampling_rate=32000)) == v["sr32000_interleaved"]
        assert seg(get_speech_timestamps(wav[:200_123], model)) == v["ragged_tail"]
    it = VADIterator(model)
    ev = [e for e in (it(wav[i:i + 512]) for i in range(0, len(wav) - 511, 512)) if e]
    assert ev == meta["test16k"]["vad_iterator_events"]
    it = VADIterator(model)
    ev = [e for e in (it(wav[i:i + 512], return_seconds=True) for i in range(0, 512 * 400, 512)) if e]
    assert ev == meta["test16k"]["vad_iterator_events_seconds"][:len(ev)] and len(ev) > 4


@pytest.mark.parametrize("sr", [16000, 8000])
def test_stateless_step_contract(torch_cuda, model, synthetic, sr):
    """svad_step_device chained over T chunks with a random initial state (ONNX contract)."""
    torch = torch_cuda
    s = synthetic
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    audio = torch.from_numpy(s[f"r1_{sr}_audio"]).cuda()
    st = torch.from_numpy(s[f"r1_{sr}_state0"]).cuda()
    cx = torch.from_numpy(s[f"r1_{sr}_ctx0"]).cuda()
    B, T = audio.shape[0], audio.shape[1] // n
    probs = torch.empty(T, B, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for t in range(T):
        x1 = torch.cat([cx, audio[:, t * n:(t + 1) * n]], 1).contiguous()
        model.engine.step_device(sr, B, x1.data_ptr(), st.data_ptr(), probs[t].data_ptr(), st.data_ptr(), stream)
        cx = x1[:, -ctx:]
    torch.cuda.synchronize()
    e_p = float(np.abs(probs.t().cpu().numpy() - s[f"r1_{sr}_probs"]).max())
    e_s = float(np.abs(st.cpu().numpy() - s[f"r1_{sr}_stateN"]).max())
    s_tol = TIGHT * max(1.0, float(np.abs(s[f"r1_{sr}_stateN"]).max()))   # the cell state reaches |c| ~ 40 on the loud rows
    print(f"step sr={sr}: prob err {e_p:.3e}, state err {e_s:.3e} (tol {s_tol:.1e})")
    assert e_p < TIGHT and e_s < s_tol
    # bulk entry with the same carried-in state/context must agree with the chained steps
    st2 = torch.from_numpy(s[f"r1_{sr}_state0"]).cuda()
    cx2 = torch.from_numpy(s[f"r1_{sr}_ctx0"]).cuda()
    pb = torch.empty(B, T, device="cuda")
    model.engine.forward_device(sr, B, T * n, audio.stride(0), audio.data_ptr(), st2.data_ptr(), cx2.data_ptr(), st2.data_ptr(),
                                cx2.data_ptr(), pb.data_ptr(), T, stream)
    torch.cuda.synchronize()
    assert float(np.abs(pb.cpu().numpy() - s[f"r1_{sr}_probs"]).max()) < TIGHT
    assert float(np.abs(st2.cpu().numpy() - s[f"r1_{sr}_stateN"]).max()) < s_tol
    assert np.array_equal(cx2.cpu().numpy(), s[f"r1_{sr}_ctxN"])


@pytest.mark.parametrize("sr", [16000, 8000])
def test_ragged_and_r2(torch_cuda, model, synthetic, oracle, sr):
    torch = torch_cuda
    p = model.audio_forward(torch.from_numpy(synthetic[f"ragged_{sr}_audio"]), sr).numpy()
    assert p.shape == synthetic[f"ragged_{sr}_probs"].shape
    assert float(np.abs(p - synthetic[f"ragged_{sr}_probs"]).max()) < TIGHT
    # R2 structured synthetic signal: regenerate deterministically via the oracle's golden probs only
    assert synthetic[f"r2_{sr}_probs"].ndim == 1


def test_wrapper_protocol(torch_cuda, model, synthetic, meta):
    torch = torch_cuda
    model.reset_states()
    for i, (B, sr) in enumerate(meta["protocol_calls"]):
        x = torch.from_numpy(synthetic[f"proto_x{i}"])
        y = model(x, sr)
        assert y.device == x.device and tuple(y.shape) == synthetic[f"proto_y{i}"].shape
        assert float(np.abs(y.numpy() - synthetic[f"proto_y{i}"]).max()) < TIGHT, (i, B, sr)
    with pytest.raises(ValueError):
        model(torch.zeros(100), 16000)
    with pytest.raises(ValueError):
        model(torch.zeros(512), 44100)
    with pytest.raises(ValueError):
        model(torch.zeros(1, 1, 512), 16000)
    with pytest.raises(ValueError):
        model(torch.zeros(640), 16000)


@pytest.mark.parametrize("rows", [0, 4, 5, 6, 7, 8])
def test_batch_rows_independent_of_tiling(torch_cuda, model, fixtures, oracle, rows):
    """Many streams over many tiles: every copy of a row must give bit-identical probabilities whatever tile /
    slot it lands in, and match the oracle (rows are independent in the reference: utils_vad.py:65-76)."""
    torch = torch_cuda
    a = fixtures["test16k"]["audio"]
    base = np.stack([a[40000 * k: 40000 * k + 512 * 12 + 77] for k in range(16)])
    B = 333
    idx = np.arange(B) % 16
    x = torch.from_numpy(base[idx])
    model.engine.set_tile_rows(rows)
    try:
        p = model.audio_forward(x, 16000).numpy()
    finally:
        model.engine.set_tile_rows(0)
    want = oracle.audio_forward(base, 16000, nthreads=4)
    for k in range(16):
        grp = p[idx == k]
        assert (grp == grp[0]).all(), f"row {k}: copies differ across tiles"
    err = float(np.abs(p[:16] - want).max())
    print(f"rows={rows}: err vs oracle {err:.3e}")
    assert err < TIGHT


def test_full_size_batch_property(torch_cuda, model, fixtures, request):
    """BASELINE config size (B=4096 streams): duplicate-row invariance + agreement with the single-stream run."""
    torch = torch_cuda
    a = torch.from_numpy(fixtures["aepyx16k"]["audio"][: 512 * 200 * 8]).view(8, -1)
    x = a.repeat(512, 1).cuda()                      # 4096 x 102400 samples (200 chunks each)
    p = model.audio_forward_device(x, 16000)
    p = p.view(512, 8, -1)
    assert bool((p == p[0:1]).all())
    single = torch.stack([model.audio_forward(a[i:i + 1], 16000)[0] for i in range(8)])
    if request.node.callspec.params["model"] == "auto":   # B=1 runs on the fp32 cluster kernel, B=4096 on the default tile kernel:
        d = float((p[0].cpu() - single).abs().max())      # two implementations, each within TIGHT of the reference
        print(f"tile kernel vs cluster kernel: {d:.3e}")
        assert d < 2 * TIGHT
    else:
        assert bool((p[0].cpu() == single).all())


def test_host_entry_points(torch_cuda, model, synthetic, oracle):
    sr = 16000
    x = np.ascontiguousarray(synthetic["r1_16000_audio"])
    B, L = x.shape
    T = L // 512
    st = synthetic["r1_16000_state0"].copy()
    cx = synthetic["r1_16000_ctx0"].copy()
    probs = np.zeros((B, T), np.float32)
    model.engine.forward_host(sr, B, L, L, x.ctypes.data, st.ctypes.data, cx.ctypes.data, st.ctypes.data, cx.ctypes.data,
                              probs.ctypes.data, T)
    assert float(np.abs(probs - synthetic["r1_16000_probs"]).max()) < TIGHT
    assert float(np.abs(st - synthetic["r1_16000_stateN"]).max()) < TIGHT * max(1.0, float(np.abs(synthetic["r1_16000_stateN"]).max()))
    assert np.array_equal(cx, synthetic["r1_16000_ctxN"])
    x1 = np.concatenate([synthetic["r1_16000_ctx0"], x[:, :512]], 1)
    st = synthetic["r1_16000_state0"].copy()
    pr = np.zeros(B, np.float32)
    model.engine.step_host(sr, B, x1.ctypes.data, st.ctypes.data, pr.ctypes.data, st.ctypes.data)
    assert float(np.abs(pr - synthetic["r1_16000_probs"][:, 0]).max()) < TIGHT


def test_batch_timestamps_equal_single(torch_cuda, model, fixtures):
    torch = torch_cuda
    from silero_vad_b200 import get_speech_timestamps, get_speech_timestamps_batch
    a = torch.from_numpy(fixtures["test16k"]["audio"])
    lens = [200_000, 123_457, 960_000, 51_200]
    L = max(lens)
    x = torch.zeros(len(lens), L)
    for i, n in enumerate(lens):
        x[i, :n] = a[i * 1000: i * 1000 + n] if i != 2 else a
    got = get_speech_timestamps_batch(x, model, lengths=lens)
    for i, n in enumerate(lens):
        assert got[i] == get_speech_timestamps(x[i, :n], model), i


def test_pcm16_paths_bit_identical(torch_cuda, model, fixtures):
    """int16 PCM through the device and the (time-sliced, pipelined) host entry points == the fp32 path."""
    torch = torch_cuda
    pcm = torch.from_numpy(np.stack([fixtures["aepyx16k"]["pcm"][100000 * b: 100000 * b + 512 * 37 + 300] for b in range(9)]).copy())
    f32 = pcm.to(torch.float32) / 32768.0
    want = model.audio_forward(f32, 16000)
    got = model.audio_forward(pcm, 16000)
    assert torch.equal(want, got) and float(want.max()) > 0.9
    B, L = pcm.shape
    T = (L + 511) // 512
    for arr, fn in ((pcm.numpy(), model.engine.forward_host_pcm16), (f32.numpy(), model.engine.forward_host)):
        for pinned in (False, True):
            a = torch.from_numpy(np.ascontiguousarray(arr))
            a = a.pin_memory() if pinned else a
            out = torch.zeros(B, T).pin_memory() if pinned else torch.zeros(B, T)
            st = np.zeros((2, B, 128), np.float32)
            fn(16000, B, L, L, a.data_ptr(), 0, 0, st.ctypes.data, 0, out.data_ptr(), T)
            assert torch.equal(out.cpu(), want), (arr.dtype, pinned)
            assert np.abs(st).max() > 0


def test_vad_iterator_batch_on_device(torch_cuda, model, fixtures, meta):
    """8 phone lines in parallel, each fed a different offset of the fixture: every row must emit what a private
    VADIterator emits; row 0 (offset 0) must reproduce the reference's golden event list."""
    torch = torch_cuda
    from silero_vad_b200 import VADIterator, VADIteratorBatch
    wav = torch.from_numpy(fixtures["test16k"]["audio"])
    B, T = 8, 500
    offs = [0] + [7000 * b + 123 for b in range(1, B)]
    it = VADIteratorBatch(model, B)
    got = [[] for _ in range(B)]
    for t in range(T):
        x = torch.stack([wav[o + 512 * t: o + 512 * (t + 1)] for o in offs])
        for b, e in enumerate(it(x)):
            if e:
                got[b].append(e)
    assert got[0] == meta["test16k"]["vad_iterator_events"][:len(got[0])] and len(got[0]) > 6
    for b in (3, 7):
        single = VADIterator(model)
        want = [e for e in (single(wav[offs[b] + 512 * t: offs[b] + 512 * (t + 1)]) for t in range(T)) if e]
        assert got[b] == want
