#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Silero-VAD engine.

Metric (BASELINE.json): chunks/sec, one chunk = 512 samples @ 16 kHz (256 @ 8 kHz with --sr 8000).
Workload at N=1 (BASELINE.json configs[2]): batch=4096 independent 16 kHz streams, T chunks each (default 64:
537 MB of fp32 audio per step, larger than the 126 MB L2, so nothing is served from cache between steps).
A "step" is one pass of the fused kernel over that batch.  N>1: every rank owns its own 4096 streams
(weak scaling, streams are independent units) and the per-chunk probabilities are all-gathered over NCCL.

  python bench.py --gpus 1 --steps 20 --warmup 3            own arm  (GPU, device-resident inputs + e2e)
  python bench.py --impl reference ...                      CPU arm  (oracle port of the reference, all host threads)

Prints ONE JSON line (rank 0).  Keys: see the task contract; `roofline` is the HBM roofline of the fused
kernel (algorithmic bytes 2052 B/chunk = 2048 B audio + 4 B probability) against MEASURED_PEAKS.json, with
the fp32-FMA fraction (the pipe that actually binds) reported beside it as `fp32_frac`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

ALGO_BYTES = {16000: 2052, 8000: 1028}          # audio fp32 in + prob fp32 out, per chunk
ALGO_FLOP = {16000: 0.730e6, 8000: 0.553e6}     # algorithmic minimum FLOP per chunk (rFFT, dead taps skipped)
HBM_FALLBACK_GBS = 6650.0                       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# MACs of the four convolutions + LSTM that the tensor-core kernel runs on tcgen05 (everything but the Nyquist bin and the head)
TENSOR_MAC = {16000: 352256, 8000: 270336}   # 353 664 - 1 280 (Nyquist bin) - 128 (head); 271 744 - 1 280 - 128
BF16_FALLBACK_TFLOPS = 2250.0                   # nominal dense bf16 peak when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--chunks", type=int, default=64, help="chunks per stream per step")
    ap.add_argument("--sr", type=int, default=16000, choices=[16000, 8000])
    ap.add_argument("--kernel", default="tc", choices=["h16", "tc", "fp32"], help="h16: tcgen05 split-fp16 two-loop kernel; tc: tcgen05 split-TF32; fp32: all CUDA cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)", 1965.0


def bf16_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return BF16_FALLBACK_TFLOPS, "nominal (2.25 PFLOP/s dense bf16)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(args, threads=None, budget_s=12.0):
    """The reference's algorithm on the host cores: the C oracle port (oracle/svad_oracle.c, OpenMP over streams),
    bounded sample of the same workload (same sr, same synthetic recipe, fewer streams/chunks)."""
    from oracle.oracle import Oracle
    o = Oracle()
    threads = threads or o.best_threads(args.sr)
    n = 512 if args.sr == 16000 else 256
    rng = np.random.default_rng(17 + args.sr)
    # calibrate: 32 streams/thread x 4 chunks (8 per thread left the OpenMP fork/join and the 4-row GEMM blocks badly amortised)
    Bs = min(args.batch, 32 * threads)
    x = (rng.standard_normal((Bs, n * 4)) * 0.03).astype(np.float32)
    o.audio_forward(x, args.sr, nthreads=threads)   # spin the thread pool up
    t0 = time.perf_counter(); o.audio_forward(x, args.sr, nthreads=threads); dt = time.perf_counter() - t0
    rate = Bs * 4 / max(dt, 1e-6)
    T = int(max(4, min(args.chunks, budget_s * rate / Bs)))
    x = (rng.standard_normal((Bs, n * T)) * 0.03).astype(np.float32)
    return o, x, T, Bs, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o, x, T, Bs, threads = cpu_arm(args, budget_s=4.0)
    for _ in range(max(1, min(args.warmup, 3))):
        o.audio_forward(x, args.sr, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.audio_forward(x, args.sr, nthreads=threads)
    dt = time.perf_counter() - t0
    val = Bs * T * args.steps / dt
    sample = f"{Bs} streams x {T} chunks per step (of {args.batch} x {args.chunks}), N(0,0.03^2) audio"
    print(json.dumps({
        "impl": "reference", "metric": "chunks/sec", "value": val, "unit": "chunks/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch={args.batch} independent {args.sr} Hz streams x {args.chunks} chunks (bounded CPU sample)",
                   "sr": args.sr, "impl": "C port of the reference graph (oracle/svad_oracle.c), OpenMP over streams"},
        "cpu_baseline": {"value": val, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from silero_vad_b200 import load_silero_vad

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    model = load_silero_vad(device=local)
    eng = model.engine
    eng.set_kernel(args.kernel)
    sr, B, T = args.sr, args.batch, args.chunks
    n = 512 if sr == 16000 else 256
    L = n * T
    g = torch.Generator(device=dev); g.manual_seed(17 + sr + rank)
    x = torch.randn(B, L, device=dev, generator=g) * 0.03          # recipe R1 (examples/onnx_sequence/run.py:159-162)
    probs = torch.empty(B, T, device=dev)
    gathered = torch.empty(world * B, T, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        eng.forward_device(sr, B, L, L, x.data_ptr(), 0, 0, 0, 0, probs.data_ptr(), T, stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, probs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for i in range(args.steps):
            kev[i][0].record()
            eng.forward_device(sr, B, L, L, x.data_ptr(), 0, 0, 0, 0, probs.data_ptr(), T, stream)
            kev[i][1].record()
            if world > 1:
                dist.all_gather_into_tensor(gathered, probs)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    launches = eng.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    chunks_per_step = B * T * world
    value = chunks_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers (pinned), H2D + kernel + D2H every step, through the C ABI host entry point
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory()
        ph = torch.empty(B, T).pin_memory()
        esteps = args.steps
        for _ in range(2):
            eng.forward_host(sr, B, L, L, xh.data_ptr(), 0, 0, 0, 0, ph.data_ptr(), T)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            eng.forward_host(sr, B, L, L, xh.data_ptr(), 0, 0, 0, 0, ph.data_ptr(), T)
            if world > 1:
                dist.all_gather_into_tensor(gathered, probs)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert torch.equal(ph, probs.cpu()), "host entry point disagrees with device entry point"
        e2e = {"value": chunks_per_step * esteps / dt, "unit": "chunks/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * T * 4,
               "ms_per_step": dt / esteps * 1e3,
               "api": "svad_forward_host (C ABI): fp32 host audio in pinned memory, time-sliced H2D overlapped with the kernel, D2H of probabilities"}
        # same call with int16 PCM host audio (what WAV files hold): half the PCIe bytes, bit-identical probabilities
        xi = (xh * 32768.0).round().clamp(-32768, 32767).to(torch.int16).pin_memory()
        pi = torch.empty(B, T).pin_memory()
        for _ in range(2):
            eng.forward_host_pcm16(sr, B, L, L, xi.data_ptr(), 0, 0, 0, 0, pi.data_ptr(), T)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            eng.forward_host_pcm16(sr, B, L, L, xi.data_ptr(), 0, 0, 0, 0, pi.data_ptr(), T)
        barrier()
        dti = time.perf_counter() - t0
        e2e["pcm16"] = {"value": B * T * world * esteps / dti, "unit": "chunks/s", "h2d_bytes_per_step": B * L * 2, "ms_per_step": dti / esteps * 1e3}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src, sm_max = peaks()
    clocks = clk.summary()
    achieved_gbs = ALGO_BYTES[sr] * B * T / (kernel_ms * 1e-3) / 1e9
    sm_mhz = clocks["sm_mhz"] or sm_max
    fp32_peak_now = eng.sm_count * 128 * 2 * sm_mhz * 1e6          # FLOP/s at the SM clock seen under load
    fp32_peak_max = eng.sm_count * 128 * 2 * sm_max * 1e6
    achieved_flops = ALGO_FLOP[sr] * B * T / (kernel_ms * 1e-3)
    traffic = None
    tp = REPO / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(f"bytes_per_launch_{sr}_{B}x{T}")
        except Exception:
            traffic = None
    out = {
        "metric": "chunks/sec", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32x3" if args.kernel == "tc" else "f32",   # tf32x3 = 3-product hi/lo split on tcgen05, fp32 accumulate (fp32-class results)
        "data": "synthetic",
        "config": {"workload": f"batch={B} independent {sr} Hz streams per GPU x {T} chunks per step (BASELINE configs[2])",
                   "sr": sr, "batch_per_gpu": B, "chunks_per_stream": T, "global_streams": B * world,
                   "l2": "inputs (%.0f MB/step/GPU) larger than L2, no flush needed" % (B * L * 4 / 1e6),
                   "parallelism": f"dp{world}: streams sharded, weights replicated, NCCL all-gather of probabilities" if world > 1 else "single GPU",
                   "kernel": ("svad_fused_tc (tcgen05 split-TF32 for enc0-3 + LSTM; STFT, gate math and head on the CUDA cores)" if args.kernel == "tc" else "svad_fused_fp32 (fp32 FFMA2, CUDA cores only)"),
                   "precision": ("x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in tf32 with fp32 accumulate; probabilities within 7e-6 of the fp32 reference" if args.kernel == "tc" else "IEEE fp32")},
        "gpu_launches": int(launches),
        "kernel_ms": kernel_ms,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_chunk": ALGO_BYTES[sr],
                     "fp32_frac": achieved_flops / fp32_peak_now, "fp32_frac_of_max_clock": achieved_flops / fp32_peak_max,
                     "fp32_note": "the path is compute-bound (356 FLOP/B vs ridge ~11); fp32_frac = algorithmic 0.73 MFLOP/chunk over "
                                  "SMs*128*2*clock, the CUDA-core FMA roofline an all-fp32 implementation is held to"},
    }
    if args.kernel == "tc":
        # second view for the tensor-core kernel: what it executes on tcgen05 vs the tensor peak.  tf32 runs at half the bf16
        # rate and split precision issues 3 products per MAC, so 1/6 of the bf16 peak is the ceiling for algorithmic FLOP.
        bf16, bf16_src = bf16_peak()
        algo_t = 2.0 * TENSOR_MAC[sr] * B * T / (kernel_ms * 1e-3) / 1e12
        out["roofline"]["tensor"] = {"algorithmic_tflops": algo_t, "executed_tf32_tflops": 3.0 * algo_t, "peak_bf16_tflops": bf16,
                                     "peak_source": bf16_src, "frac_of_tf32_peak": 3.0 * algo_t / (0.5 * bf16),
                                     "note": "dense-layer MACs on tcgen05 (x3 products of the hi/lo split) over half the bf16 peak (tf32 rate)"}
    if e2e:
        out["e2e"] = e2e
    if world == 1:
        # BASELINE configs[1]: batch = 1 streaming, one 512-sample chunk per call through the stateless C ABI step
        # (host buffers, H2D + cluster kernel + D2H + sync per call)
        n1 = 512 if sr == 16000 else 256
        x1 = np.zeros((1, n1 + n1 // 8), np.float32)
        st1 = np.zeros((2, 1, 128), np.float32)
        pr1 = np.zeros(1, np.float32)
        lat = []
        for i in range(300):
            x1[0, n1 // 8:] = np.random.default_rng(i).standard_normal(n1).astype(np.float32) * 0.03
            t0 = time.perf_counter()
            eng.step_host(sr, 1, x1.ctypes.data, st1.ctypes.data, pr1.ctypes.data, st1.ctypes.data)
            lat.append((time.perf_counter() - t0) * 1e6)
        lat = np.sort(np.asarray(lat[50:]))
        xs = torch.randn(1, n1 * 256, device=dev) * 0.03
        ps = torch.empty(1, 256, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.forward_device(sr, 1, n1 * 256, n1 * 256, xs.data_ptr(), 0, 0, 0, 0, ps.data_ptr(), 256, stream)
        e0.record()
        eng.forward_device(sr, 1, n1 * 256, n1 * 256, xs.data_ptr(), 0, 0, 0, 0, ps.data_ptr(), 256, stream)
        e1.record(); torch.cuda.synchronize()
        out["latency_b1"] = {"step_host_us_median": float(np.median(lat)), "step_host_us_p99": float(lat[int(0.99 * len(lat))]),
                             "kernel_us_per_chunk": e0.elapsed_time(e1) * 1e3 / 256,
                             "note": "batch=1 (BASELINE configs[1]): svad_step_host per chunk incl. copies and sync; kernel = 8-CTA cluster kernel, 256 chunks in one launch"}
    if world == 1 and not args.no_cpu_baseline:
        o, xs, Ts, Bs, threads = cpu_arm(args)
        t0 = time.perf_counter(); o.audio_forward(xs, sr, nthreads=threads); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": Bs * Ts / dt, "unit": "chunks/s", "cores": threads, "kind": "port",
                               "sample": f"{Bs} streams x {Ts} chunks, same recipe; C port of the reference graph (oracle/svad_oracle.c)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
