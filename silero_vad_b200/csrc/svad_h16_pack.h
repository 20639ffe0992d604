// svad_h16_pack.h -- geometry, shared-memory map and host-side weight packing of the fp16 split-precision kernel
// `svad_fused_h16` (svad_h16.cuh).
//
// Every contraction of the path -- the STFT as a dense windowed-DFT basis product (the reference's own formulation:
// silero_vad.jit::_model.stft.transform_ is conv1d with forward_basis_buffer), the four encoder convolutions and the LSTM
// cell -- runs on tcgen05 with kind::f16 operands and fp32 accumulation in TMEM:
//     x . w  ~=  x_hi . w_hi  +  x_lo . w_hi  +  x_hi . w_lo,      hi = fp16(s v),  lo = fp16(s v - hi)
// 11 + 11 significand bits per operand (the same 22 bits as the tf32 hi/lo split of svad_tc.h) at twice the tensor-core
// rate (K = 16 per instruction) and half the bytes per element, both in the weight stream and in shared memory.  fp16 has
// a 5-bit exponent, so operands are pre-scaled by exact powers of two (`s`: per tensor for the weights, per layer for the
// activations) and the accumulator is descaled in the epilogue; conversions saturate (cvt.satfinite) instead of
// overflowing.  tools/h16_numerics.py models exactly this arithmetic on the CPU: <= 4e-6 against the reference on the three
// WAV fixtures, the same as the tf32 kernel.
//
// Weights: A operand, K-major SWIZZLE_128B tiles of [M x 64] fp16 (M = 128: 16 KB, M = 64: 8 KB), row r at
// (r/8)*1024 + (r%8)*128 B, 16-byte chunk (k/8) ^ (r%8).  Two tapes per branch, one per pipeline loop of the kernel:
//   F tape: STFT basis, enc0, enc1        B tape: enc2, enc3, LSTM        (16 KB slabs: one M = 128 tile, or the {hi | lo} pair of an M = 64 tile)
// A bulk copy L2 -> shared memory takes ~1000 cycles to land whatever its size, so what a ring sustains is (bytes in flight) /
// latency: three 16 KB stages per loop is what the 227 KB of shared memory leave room for next to the activations.
#pragma once
#include <cuda_fp16.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "svad_pack.h"

namespace svad {

template <bool SR16>
struct H16Geo {
    static constexpr int n = SR16 ? 512 : 256;       // chunk samples
    static constexpr int ctx = SR16 ? 64 : 32;
    static constexpr int N = SR16 ? 256 : 128;       // filter length
    static constexpr int hop = N / 2;
    static constexpr int Kt = N / 2;                 // bins 0 .. Kt-1 on the tensor core; bin Kt (Nyquist) is a rank-1 fp32 update
    static constexpr int L1 = ctx + n;               // 576 / 288
    static constexpr int XR = L1 + N / 4;            // rows of the padded window [context | chunk | reflect]: 640 / 320
    static constexpr int kcs = N / 64;               // 64-wide K chunks of the STFT: 4 / 2
    static constexpr int stft_M = SR16 ? 128 : 64;   // 16k: two M=128 tiles (re[0..127] | re[128], im[1..127]); 8k: two M=64 tiles (re[0..63] | re[64], im[1..63])
    static constexpr int nslab_stft = SR16 ? 16 : 4; // 16k: (tile, kc) x {hi, lo}; 8k: (kc, tile) with {hi | lo} in one slab
    static constexpr int e0_chunks = Kt / 64;        // 2 / 1
    static constexpr int nslab_e0 = 3 * e0_chunks * 2;   // (tap, chunk) x {hi, lo}
    static constexpr int nslab_e1 = 6;               // (tap, chunk): {hi | lo} of a [64 x 64] tile
    static constexpr int nslabF = nslab_stft + nslab_e0 + nslab_e1;   // 34 / 16 slabs per chunk step
    static constexpr int nslabB = 1 + 1 + 16;                         // enc2 (both taps), enc3 {hi | lo}, LSTM 16 x {hi | lo}: 32 KB each
};
// One thread issues a bulk copy about every 310 cycles whatever its size (tools/ubench_feed.cu: 16 KB copies 52 B/clk/SM, 32 KB copies
// 78-96 B/clk/SM, ceiling ~105-110): the back tape, dominated by the LSTM's 512 KB per step, moves in 32 KB slabs.
constexpr int kH16SlabF = 16384, kH16SlabB = 32768;
constexpr int kH16StagesF = 3, kH16StagesB = 2;
// warps 0-3 front epilogue, 4-7 back epilogue, 8 / 9 MMA issue (front / back), 10 / 11 weight streams; with SVAD_H16_EF_WARPS == 8 a second
// front-epilogue group (warps 12-15) takes frames 2, 3 of the |X| / enc0 epilogues and every other block of the window staging
#ifndef SVAD_H16_EF_WARPS
#define SVAD_H16_EF_WARPS 4
#endif
constexpr int kH16EfWarps = SVAD_H16_EF_WARPS;
static_assert(kH16EfWarps == 4 || kH16EfWarps == 8, "front epilogue: one or two groups of four warps");
constexpr int kH16Threads = 384 + 32 * (kH16EfWarps - 4);

// activation scales (exact powers of two).  |x| <= 1 for normalised audio; mag <= 181 |x|; e0..e3 were observed <= 72 on speech
// at amplitude 0.55 (tools/h16_numerics.py) -- the conversions saturate at 65504, far above anything normalised audio produces.
constexpr float kSx = 2048.0f, kSmag = 32.0f, kSe0 = 8.0f, kSe1 = 8.0f, kSe2 = 8.0f, kSe3 = 256.0f, kSh = 256.0f;

// shared-memory map (byte offsets).  Each layer's output overwrites its input: the MMAs that read the input have
// completed (accumulators in TMEM) before the epilogue writes.
struct H16Map {
    static constexpr int R = 0;                       // front region: window xp (hi | lo) -> mag (hi | lo) -> e0 (hi | lo)
    static constexpr int R_bytes = 81920;             // 2 x 640 rows x 64 B
    static constexpr int P = R + R_bytes;             // back region: e1 -> e2 -> e3 (hi at +0, lo at +8192 / +4096 for e2)
    static constexpr int P_bytes = 16384;
    static constexpr int H = P + P_bytes;             // LSTM hidden state h (hi at +0, lo at +8192)
    static constexpr int H_bytes = 16384;
    static constexpr int FR = H + H_bytes;            // front weight ring
    static constexpr int BR = FR + kH16StagesF * kH16SlabF;
    static constexpr int C = BR + kH16StagesB * kH16SlabB;   // shared-memory scratch (floats): |X| of the Nyquist bin [4 frames][32], head weights
    static constexpr int s_nyq = 0, s_wout = 128, s_floats = 256;
    // constants block in GLOBAL memory (biases and scales are read once per thread into registers)
    static constexpr int c_b0 = 0, c_b1 = 128, c_b2 = 192, c_b3 = 256, c_bl = 384, c_wout = 896, c_wnyq = 1024,
                         c_bout = 1536, c_scale = 1540;      // c_scale: d_stft, d_e0, d_e1, d_e2, d_e3, d_lstm
    static constexpr int c_floats = 1552;
    static constexpr int BAR = C + s_floats * 4;      // mbarriers + TMEM slot
    static constexpr int total = BAR + 512;
};
static_assert(H16Map::FR % 1024 == 0 && H16Map::BR % 1024 == 0 && H16Map::P % 1024 == 0 && H16Map::H % 1024 == 0, "tile alignment");
static_assert(H16Map::total <= 232448, "shared memory budget");

struct PackedH16 {
    std::vector<unsigned char> tapeF, tapeB;
    std::vector<float> consts;
};

inline float h16_pow2_scale(double maxabs, double target) { return (float)std::exp2(std::floor(std::log2(target / maxabs))); }

// (hi, lo) of the scaled value; v in double so that analytically computed operands (the DFT basis) keep 22 bits
inline void h16_split(double v, __half& hi, __half& lo) {
    hi = __float2half_rn((float)v);
    lo = __float2half_rn((float)(v - (double)__half2float(hi)));
}

// A(r, k) for r < M, k < 64 -> K-major SWIZZLE_128B tile pair: hi tile at dst_hi, lo tile at dst_lo (M * 128 bytes each)
template <class F>
inline void h16_pack_tile(int M, F&& A, double scale, unsigned char* dst_hi, unsigned char* dst_lo) {
    __half* hi = reinterpret_cast<__half*>(dst_hi);
    __half* lo = reinterpret_cast<__half*>(dst_lo);
    for (int r = 0; r < M; r++)
        for (int k = 0; k < 64; k++) {
            const int pos = (r / 8) * 512 + (r % 8) * 64 + (((k / 8) ^ (r % 8)) * 8) + (k % 8);
            h16_split(scale * A(r, k), hi[pos], lo[pos]);
        }
}

template <bool SR16>
inline bool pack_branch_h16(const TensorMap& tm, PackedH16& out, std::string& err) {
    using G = H16Geo<SR16>;
    const std::string p = SR16 ? "_model." : "_model_8k.";
    auto get = [&](const char* s) -> const float* {
        auto it = tm.find(p + s);
        if (it == tm.end()) { err = "missing tensor " + p + s; return nullptr; }
        return it->second.data.data();
    };
    const int F = G::Kt + 1;
    const float* w0 = get("encoder.0.reparam_conv.weight");   // [128][F][3]
    const float* b0 = get("encoder.0.reparam_conv.bias");
    const float* w1 = get("encoder.1.reparam_conv.weight");   // [64][128][3]
    const float* b1 = get("encoder.1.reparam_conv.bias");
    const float* w2 = get("encoder.2.reparam_conv.weight");   // [64][64][3]
    const float* b2 = get("encoder.2.reparam_conv.bias");
    const float* w3 = get("encoder.3.reparam_conv.weight");   // [128][64][3]
    const float* b3 = get("encoder.3.reparam_conv.bias");
    const float* wih = get("decoder.rnn.weight_ih");          // [512][128]
    const float* whh = get("decoder.rnn.weight_hh");
    const float* bih = get("decoder.rnn.bias_ih");
    const float* bhh = get("decoder.rnn.bias_hh");
    const float* wo = get("decoder.decoder.2.weight");
    const float* bo = get("decoder.decoder.2.bias");
    if (!w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !wih || !whh || !bih || !bhh || !wo || !bo) return false;
    auto maxabs = [](const float* w, size_t n) { double m = 0; for (size_t i = 0; i < n; i++) m = std::fmax(m, std::fabs((double)w[i])); return m; };
    const double kTarget = 16384.0;
    const float Sb = 16384.0f;                                               // |basis| <= 1
    const float S0 = h16_pow2_scale(maxabs(w0, (size_t)128 * F * 3), kTarget);
    const float S1 = h16_pow2_scale(maxabs(w1, (size_t)64 * 128 * 3), kTarget);
    const float S2 = h16_pow2_scale(maxabs(w2, (size_t)64 * 64 * 3), kTarget);
    const float S3 = h16_pow2_scale(maxabs(w3, (size_t)128 * 64 * 3), kTarget);
    const float Sl = h16_pow2_scale(std::fmax(maxabs(wih, 512 * 128), maxabs(whh, 512 * 128)), kTarget);   // one scale: both halves of K share the accumulator

    out.tapeF.assign((size_t)G::nslabF * kH16SlabF, 0);
    out.tapeB.assign((size_t)G::nslabB * kH16SlabB, 0);
    unsigned char* tf = out.tapeF.data();
    unsigned char* tb = out.tapeB.data();
    // ---- STFT basis: rows of the reference's forward_basis_buffer, rebuilt analytically (periodic Hann x [cos ; -sin])
    const int N = G::N;
    auto win = [&](int m) { return 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)m / (double)N); };
    auto re_row = [&](int bin, int m) { return win(m) * std::cos(2.0 * M_PI * (double)((bin * m) % N) / (double)N); };
    auto im_row = [&](int bin, int m) { return -win(m) * std::sin(2.0 * M_PI * (double)((bin * m) % N) / (double)N); };
    // second tile: row 0 = the Nyquist bin's real part (its imaginary part and bin 0's are identically zero), rows r >= 1 = im[r]
    auto tile1 = [&](int r, int m) { return r == 0 ? re_row(N / 2, m) : im_row(r, m); };
    if (SR16) {
        for (int mt = 0; mt < 2; mt++)
            for (int kc = 0; kc < 4; kc++) {
                unsigned char* s = tf + (size_t)((mt * 4 + kc) * 2) * kH16SlabF;
                h16_pack_tile(128, [&](int r, int k) { return mt == 0 ? re_row(r, 64 * kc + k) : tile1(r, 64 * kc + k); }, Sb, s, s + kH16SlabF);
            }
    } else {
        for (int kc = 0; kc < 2; kc++) {
            unsigned char* s = tf + (size_t)(kc * 2) * kH16SlabF;
            h16_pack_tile(64, [&](int r, int k) { return re_row(r, 64 * kc + k); }, Sb, s, s + 8192);
            h16_pack_tile(64, [&](int r, int k) { return tile1(r, 64 * kc + k); }, Sb, s + kH16SlabF, s + kH16SlabF + 8192);
        }
    }
    // ---- enc0: taps in the order 1, 0, 2 (tap 1 reaches all four output frames: its first instruction overwrites the accumulator)
    const int tap_order0[3] = {1, 0, 2};
    for (int jo = 0; jo < 3; jo++)
        for (int c = 0; c < G::e0_chunks; c++) {
            const int j = tap_order0[jo];
            unsigned char* s = tf + (size_t)(G::nslab_stft + (jo * G::e0_chunks + c) * 2) * kH16SlabF;
            h16_pack_tile(128, [&](int o, int k) { return (double)w0[((size_t)o * F + 64 * c + k) * 3 + j]; }, S0, s, s + kH16SlabF);
        }
    // ---- enc1: taps 1, 2, 0; per slab two 64-channel chunks of [64 x 64] tiles {hi, lo}
    const int tap_order1[3] = {1, 2, 0};
    for (int jo = 0; jo < 3; jo++) {
        const int j = tap_order1[jo];
        for (int c = 0; c < 2; c++) {
            unsigned char* s = tf + (size_t)(G::nslab_stft + G::nslab_e0 + jo * 2 + c) * kH16SlabF;
            h16_pack_tile(64, [&](int o, int k) { return (double)w1[((size_t)o * 128 + 64 * c + k) * 3 + j]; }, S1, s, s + 8192);
        }
    }
    // ---- back tape: enc2 (taps 1, 2: tap 0 only ever multiplies zero padding), enc3 (tap 1), LSTM
    for (int q = 0; q < 2; q++) {
        unsigned char* s = tb + (size_t)q * 16384;   // slab 0: tap 1 {hi | lo}, tap 2 {hi | lo}
        h16_pack_tile(64, [&](int o, int k) { return (double)w2[((size_t)o * 64 + k) * 3 + q + 1]; }, S2, s, s + 8192);
    }
    h16_pack_tile(128, [&](int o, int k) { return (double)w3[((size_t)o * 64 + k) * 3 + 1]; }, S3, tb + kH16SlabB, tb + kH16SlabB + 16384);
    for (int kc = 0; kc < 4; kc++)
        for (int m = 0; m < 4; m++) {
            unsigned char* s = tb + (size_t)(2 + kc * 4 + m) * kH16SlabB;
            const float* src = kc < 2 ? wih : whh;
            h16_pack_tile(128, [&](int j, int k) { return (double)src[(size_t)(m * 128 + j) * 128 + 64 * (kc & 1) + k]; }, Sl, s, s + 16384);
        }
    // ---- constants
    out.consts.assign(H16Map::c_floats, 0.0f);
    float* cs = out.consts.data();
    memcpy(cs + H16Map::c_b0, b0, 128 * 4);
    memcpy(cs + H16Map::c_b1, b1, 64 * 4);
    memcpy(cs + H16Map::c_b2, b2, 64 * 4);
    memcpy(cs + H16Map::c_b3, b3, 128 * 4);
    for (int g = 0; g < 512; g++) cs[H16Map::c_bl + g] = bih[g] + bhh[g];
    memcpy(cs + H16Map::c_wout, wo, 128 * 4);
    for (int j = 0; j < 3; j++)
        for (int o = 0; o < 128; o++) cs[H16Map::c_wnyq + j * 128 + o] = w0[((size_t)o * F + (F - 1)) * 3 + j];
    cs[H16Map::c_bout] = bo[0];
    cs[H16Map::c_scale + 0] = 1.0f / (kSx * Sb);
    cs[H16Map::c_scale + 1] = 1.0f / (kSmag * S0);
    cs[H16Map::c_scale + 2] = 1.0f / (kSe0 * S1);
    cs[H16Map::c_scale + 3] = 1.0f / (kSe1 * S2);
    cs[H16Map::c_scale + 4] = 1.0f / (kSe2 * S3);
    cs[H16Map::c_scale + 5] = 1.0f / (kSe3 * Sl);
    return true;
}

}  // namespace svad
