"""GPU parity of the tensor-core kernel (tcgen05, split-precision TF32 for all dense layers) against the reference goldens,
the oracle and the fp32 CUDA-core kernel.  Same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL, TIGHT = 1e-4, 2e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def model_tc(torch_cuda):
    from silero_vad_b200 import load_silero_vad
    m = load_silero_vad(device=0)
    m.engine.set_kernel("tc")
    m.engine.set_small_batch_max(0)
    return m


def seg(ts):
    return [[d["start"], d["end"]] for d in ts]


@pytest.mark.parametrize("name", ["test16k", "aepyx16k", "aepyx8k"])
def test_tc_fixture_probs_and_segments(torch_cuda, model_tc, fixtures, meta, name):
    torch = torch_cuda
    from silero_vad_b200 import get_speech_timestamps
    fx = fixtures[name]
    wav = torch.from_numpy(fx["audio"])
    p = model_tc.audio_forward(wav[None], fx["sr"]).numpy()[0]
    err = float(np.abs(p - fx["probs"]).max())
    print(f"tc {name}: max|p - p_ref| = {err:.3e} over {p.size} chunks")
    assert err < TOL
    assert err < TIGHT
    assert seg(get_speech_timestamps(wav, model_tc, sampling_rate=fx["sr"])) == meta[name]["segments"]


@pytest.mark.parametrize("sr", [16000, 8000])
def test_tc_state_contract(torch_cuda, model_tc, synthetic, sr):
    torch = torch_cuda
    s = synthetic
    n, ctx = (512, 64) if sr == 16000 else (256, 32)
    audio = torch.from_numpy(s[f"r1_{sr}_audio"]).cuda()
    st = torch.from_numpy(s[f"r1_{sr}_state0"]).cuda()
    cx = torch.from_numpy(s[f"r1_{sr}_ctx0"]).cuda()
    B, T = audio.shape[0], audio.shape[1] // n
    pb = torch.empty(B, T, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    # two pieces with carried state: T//2 chunks, then the rest
    h = (T // 2) * n
    model_tc.engine.forward_device(sr, B, h, audio.stride(0), audio.data_ptr(), st.data_ptr(), cx.data_ptr(), st.data_ptr(), cx.data_ptr(),
                                   pb.data_ptr(), T, stream)
    model_tc.engine.forward_device(sr, B, T * n - h, audio.stride(0), audio[:, h:].data_ptr(), st.data_ptr(), cx.data_ptr(), st.data_ptr(),
                                   cx.data_ptr(), pb[:, T // 2:].data_ptr(), T, stream)
    torch.cuda.synchronize()
    e_p = float(np.abs(pb.cpu().numpy() - s[f"r1_{sr}_probs"]).max())
    e_s = float(np.abs(st.cpu().numpy() - s[f"r1_{sr}_stateN"]).max())
    smax = float(np.abs(s[f"r1_{sr}_stateN"]).max())
    print(f"tc sr={sr}: prob err {e_p:.3e} state err {e_s:.3e} (max |state| {smax:.2f})")
    assert e_p < TIGHT and e_s < 2e-5 * max(1.0, smax)
    assert np.array_equal(cx.cpu().numpy(), s[f"r1_{sr}_ctxN"])


@pytest.mark.parametrize("rows", [0, 7, 8])
def test_tc_tiling_invariance(torch_cuda, model_tc, fixtures, oracle, rows):
    torch = torch_cuda
    a = fixtures["test16k"]["audio"]
    base = np.stack([a[40000 * k: 40000 * k + 512 * 12 + 77] for k in range(16)])
    idx = np.arange(333) % 16
    model_tc.engine.set_tile_rows(rows)
    try:
        p = model_tc.audio_forward(torch.from_numpy(base[idx]), 16000).numpy()
    finally:
        model_tc.engine.set_tile_rows(0)
    for k in range(16):
        grp = p[idx == k]
        assert (grp == grp[0]).all(), f"row {k}: copies differ across tiles"
    err = float(np.abs(p[:16] - oracle.audio_forward(base, 16000, nthreads=4)).max())
    print(f"tc rows={rows}: err vs oracle {err:.3e}")
    assert err < TIGHT


def test_tc_matches_fp32_kernel_full_batch(torch_cuda, model_tc, fixtures):
    torch = torch_cuda
    from silero_vad_b200 import load_silero_vad
    ref = load_silero_vad(device=0)
    ref.engine.set_kernel("fp32")
    a = torch.from_numpy(fixtures["aepyx16k"]["audio"][: 512 * 100 * 8]).view(8, -1)
    x = a.repeat(512, 1).cuda()                      # 4096 streams x 100 chunks
    p_tc = model_tc.audio_forward_device(x, 16000)
    p_32 = ref.audio_forward_device(x, 16000)
    d = float((p_tc - p_32).abs().max())
    print(f"tc vs fp32 kernel, 4096 x 100 chunks: max diff {d:.3e}")
    assert d < TIGHT
    assert bool((p_tc.view(512, 8, -1) == p_tc.view(512, 8, -1)[0:1]).all())
