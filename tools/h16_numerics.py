#!/usr/bin/env python
"""Numerics study for the fp16 split-precision tensor-core kernel (svad_fused_h16), CPU only.

Every dense layer (and the STFT as a dense DFT-basis product) is evaluated as
    x.w ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo
with hi = fp16(round-to-nearest) of the scaled operand and lo = fp16(operand - hi): 22 mantissa bits per operand, the
same as the tf32 hi/lo split, at twice the tensor-core rate and half the bytes.  fp16 has a 5-bit exponent, so operands
are pre-scaled by exact powers of two (per layer for activations, per tensor for weights) and the accumulator is
descaled in the epilogue.  Products of fp16 values are exact in fp32; the accumulation is modelled in fp32 (one add per
K=16 block, the instruction granularity).

Compares chained probabilities with the reference goldens (tests/golden/*.npz).  Usage: h16_numerics.py [nchunks]
"""
import struct
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]


def read_container(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"SVADW001"
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack("<%dI" % nd, f.read(4 * nd))
            out[name] = np.frombuffer(f.read(4 * int(np.prod(dims))), dtype="<f4").reshape(dims).copy()
    return out


def split16(x, scale):
    """x (fp32) * scale -> (hi, lo) fp16 pair as float32 arrays (values exactly representable in fp16)."""
    xs = x * scale if x.dtype == np.float64 else (x.astype(np.float32) * np.float32(scale)).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(xs.dtype)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def mm3(wh, wl, xh, xl, kblk=16):
    """D[m, n] = sum_k (wh xh + wh xl + wl xh), fp32 accumulation in K blocks of `kblk` (products exact in fp64 -> fp32)."""
    M, K = wh.shape
    acc = np.zeros((M, xh.shape[1]), np.float32)
    for k0 in range(0, K, kblk):
        s = slice(k0, k0 + kblk)
        blk = (wh[:, s].astype(np.float64) @ xh[s].astype(np.float64) + wh[:, s].astype(np.float64) @ xl[s].astype(np.float64)
               + wl[:, s].astype(np.float64) @ xh[s].astype(np.float64))
        acc = (acc + blk.astype(np.float32)).astype(np.float32)
    return acc


def pow2_scale(maxabs, target):
    return 2.0 ** np.floor(np.log2(target / maxabs))


class H16Model:
    def __init__(self, tm, sr, act_scale=None):
        p = "_model." if sr == 16000 else "_model_8k."
        self.sr = sr
        self.N = 256 if sr == 16000 else 128
        self.n = 2 * self.N
        self.ctx = self.N // 4
        self.F = self.N // 2 + 1
        N, F = self.N, self.F
        m = np.arange(N)
        win = 0.5 - 0.5 * np.cos(2 * np.pi * m / N)
        k = np.arange(F)[:, None]
        re = win * np.cos(2 * np.pi * k * m / N)
        im = -win * np.sin(2 * np.pi * k * m / N)
        # 2 x 128-row tiles: tile0 = re[0..N/2-1]; tile1 = [re[N/2], im[1..N/2-1]]
        self.basis = np.concatenate([re[: N // 2], re[N // 2: N // 2 + 1], im[1: N // 2]])   # [N, N] float64: split directly, like the packer
        g = lambda s: tm[p + s]
        self.w0, self.b0 = g("encoder.0.reparam_conv.weight"), g("encoder.0.reparam_conv.bias")
        self.w1, self.b1 = g("encoder.1.reparam_conv.weight"), g("encoder.1.reparam_conv.bias")
        self.w2, self.b2 = g("encoder.2.reparam_conv.weight"), g("encoder.2.reparam_conv.bias")
        self.w3, self.b3 = g("encoder.3.reparam_conv.weight"), g("encoder.3.reparam_conv.bias")
        self.wl = np.concatenate([g("decoder.rnn.weight_ih"), g("decoder.rnn.weight_hh")], axis=1)   # [512, 256]
        self.bl = g("decoder.rnn.bias_ih") + g("decoder.rnn.bias_hh")
        self.wo, self.bo = g("decoder.decoder.2.weight")[0, :, 0], g("decoder.decoder.2.bias")[0]
        # weight scales: per tensor, max |w| -> ~2^14
        self.S = {}
        for name, w in (("basis", self.basis), ("w0", self.w0), ("w1", self.w1), ("w2", self.w2), ("w3", self.w3), ("wl", self.wl)):
            self.S[name] = pow2_scale(np.abs(w).max(), 2.0 ** 14)
        # activation scales (fixed powers of two; worst-case bounds are checked by the caller)
        # the kernel's activation scales (svad_h16_pack.h: kSx, kSmag, kSe0..kSe3, kSh)
        self.A = dict(x=2.0 ** 11, mag=2.0 ** 5, e0=2.0 ** 3, e1=2.0 ** 3, e2=2.0 ** 3, e3=2.0 ** 8, h=2.0 ** 8)
        if act_scale:
            self.A.update(act_scale)
        self.wsplit = {k_: split16(getattr(self, k_), self.S[k_]) for k_ in ("basis", "w0", "w1", "w2", "w3", "wl")}
        self.trace = None
        self.maxes = {k_: 0.0 for k_ in self.A}

    def gemm(self, wname, wsel, x, aname):
        """W[wsel] . x with the split; x [K, n] fp32 unscaled."""
        wh, wl = self.wsplit[wname]
        wh, wl = wsel(wh), wsel(wl)
        self.maxes[aname] = max(self.maxes[aname], float(np.abs(x).max()))
        xh, xl = split16(x, self.A[aname])
        d = mm3(wh, wl, xh, xl)
        return d * np.float32(1.0 / (self.S[wname] * self.A[aname]))

    def step(self, x1, h, c):
        """x1 [ctx+n] -> prob, h', c'."""
        N, F = self.N, self.F
        P = N // 4
        L = len(x1)
        xp = np.concatenate([x1, x1[L - 2: L - 2 - P: -1]]).astype(np.float32)
        frames = np.stack([xp[(N // 2) * f: (N // 2) * f + N] for f in range(4)], axis=1)   # [N, 4]
        d = self.gemm("basis", lambda w: w, frames, "x")     # [N, 4]: re[0..N/2-1], re[N/2], im[1..N/2-1]
        re = np.concatenate([d[: N // 2], d[N // 2: N // 2 + 1]])
        im = np.concatenate([np.zeros((1, 4), np.float32), d[N // 2 + 1:], np.zeros((1, 4), np.float32)])
        mag = np.sqrt(re * re + im * im).astype(np.float32)   # [F, 4]
        Kt = F - 1
        # enc0: taps via shifted frames; the Nyquist bin (row Kt) as an fp32 rank-1 update
        e0 = np.zeros((128, 4), np.float32)
        for j in range(3):
            xs = np.zeros((Kt, 4), np.float32)
            for t in range(4):
                f = t + j - 1
                if 0 <= f < 4:
                    xs[:, t] = mag[:Kt, f]
            e0 += self.gemm("w0", lambda w: w[:, :Kt, j], xs, "mag")
            for t in range(4):
                f = t + j - 1
                if 0 <= f < 4:
                    e0[:, t] += self.w0[:, Kt, j] * mag[Kt, f]
        e0 = np.maximum(e0 + self.b0[:, None], 0).astype(np.float32)
        e1 = np.zeros((64, 2), np.float32)
        for j in range(3):
            xs = np.zeros((128, 2), np.float32)
            for t in range(2):
                f = 2 * t + j - 1
                if 0 <= f < 4:
                    xs[:, t] = e0[:, f]
            e1 += self.gemm("w1", lambda w: w[:, :, j], xs, "e0")
        e1 = np.maximum(e1 + self.b1[:, None], 0).astype(np.float32)
        e2 = np.zeros((64, 1), np.float32)
        for j in (1, 2):
            e2 += self.gemm("w2", lambda w: w[:, :, j], e1[:, j - 1: j], "e1")
        e2 = np.maximum(e2 + self.b2[:, None], 0).astype(np.float32)
        e3 = np.maximum(self.gemm("w3", lambda w: w[:, :, 1], e2, "e2") + self.b3[:, None], 0).astype(np.float32)
        g = self.gemm("wl", lambda w: w[:, :128], e3, "e3") + self.gemm("wl", lambda w: w[:, 128:], h[:, None], "h")
        g = (g[:, 0] + self.bl).astype(np.float32)
        sig = lambda v: (1.0 / (1.0 + np.exp(-v.astype(np.float64)))).astype(np.float32)
        i, f_, gg, o = sig(g[:128]), sig(g[128:256]), np.tanh(g[256:384]), sig(g[384:])
        c2 = (f_ * c + i * gg).astype(np.float32)
        h2 = (o * np.tanh(c2)).astype(np.float32)
        p = sig(np.float32(np.dot(self.wo, np.maximum(h2, 0)) + self.bo))
        if self.trace is not None:
            self.trace.append(dict(mag=mag, e0=e0, e1=e1, e2=e2[:, 0], e3=e3[:, 0], gates=g, h=h2, c=c2, p=float(p)))
        return float(p), h2, c2

    def run(self, audio, nchunks=None):
        T = len(audio) // self.n if nchunks is None else min(nchunks, len(audio) // self.n)
        h = np.zeros(128, np.float32)
        c = np.zeros(128, np.float32)
        ctx = np.zeros(self.ctx, np.float32)
        out = []
        for t in range(T):
            chunk = audio[t * self.n: (t + 1) * self.n]
            x1 = np.concatenate([ctx, chunk])
            p, h, c = self.step(x1, h, c)
            ctx = x1[-self.ctx:]
            out.append(p)
        return np.array(out, np.float32)


def bounds(m):
    """Worst-case |activation| bounds for |audio| <= 1 from L1 row norms (what the fixed scales must cover)."""
    b = {}
    b["x"] = 1.0
    b["mag"] = float(np.abs(m.basis).sum(1).max()) * np.sqrt(2)
    b["e0"] = float((np.abs(m.w0).sum((1, 2)) * b["mag"] + np.abs(m.b0)).max())
    b["e1"] = float((np.abs(m.w1).sum((1, 2)) * b["e0"] + np.abs(m.b1)).max())
    b["e2"] = float((np.abs(m.w2).sum((1, 2)) * b["e1"] + np.abs(m.b2)).max())
    b["e3"] = float((np.abs(m.w3).sum((1, 2)) * b["e2"] + np.abs(m.b3)).max())
    b["h"] = 1.0
    return b


def main():
    nch = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    tm = read_container(REPO / "silero_vad_b200/data/silero_vad_v6.weights")
    for name, sr in (("test16k", 16000), ("aepyx8k", 8000), ("aepyx16k", 16000)):
        z = np.load(REPO / f"tests/golden/{name}.npz")
        audio = z["pcm"].astype(np.float32) / 32768.0
        m = H16Model(tm, sr)
        got = m.run(audio, nch)
        want = z["probs"][: len(got)]
        print(f"{name}: {len(got)} chunks  max|p - p_ref| = {np.abs(got - want).max():.3e}   observed max |act|: "
              + " ".join(f"{k}={v:.3g}" for k, v in m.maxes.items()))
        bb = bounds(m)
        print("   worst-case bounds: " + " ".join(f"{k}={v:.3g}" for k, v in bb.items()),
              "  scaled max (must be < 65504): " + " ".join(f"{k}={bb[k] * m.A[k]:.3g}" for k in bb))
        print("   weight scales: " + " ".join(f"{k}=2^{int(np.log2(v))}" for k, v in m.S.items()))


if __name__ == "__main__":
    main()
