// svad_pack.h -- host side: read the SVADW001 weight container and lay the parameters out as the
// "weight tape" the fused kernel streams through shared memory once per chunk step.
//
// Tape (fp32, consumption order; every slab is a contiguous, 16-byte aligned byte range):
//   enc0  W0p[c][j][o]    c < F, j < 3, o < 128          (= encoder.0.reparam_conv.weight[o][c][j])
//   enc1  W1p[c][j][o]    c < 128, j < 3, o < 64
//   enc2  W2p[c][jj][o]   c < 64, jj < 2 (taps 1,2: tap 0 only ever multiplies zero padding), o < 64
//   enc3  W3p[c][o]       c < 64, tap 1 only, o < 128
//   lstm  Wl[k][n']       k < 256 ([W_ih ; W_hh] along k), n' = 64*w + 32*u + 4*l + g  <->  row g*128 + j of
//                         the PyTorch LSTMCell weights with j = 16*w + 2*l + u (gate order i,f,g,o)
// Constants block (loaded into shared memory once per CTA), see svad::SmemMap::c_*:
//   b0[128] b1[64] b2[64] b3[128] bl[512] (= b_ih + b_hh, permuted like n') wout[128] bout pad[3] win[256]
//   win[m] = 0.5 * (0.5 - 0.5 cos(2 pi m / N))   (periodic Hann; the 1/2 belongs to the two-for-one FFT split)
// Reference for the parameter names: silero_vad.jit state_dict.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "svad_core.h"

namespace svad {

struct HostTensor { std::vector<uint32_t> dims; std::vector<float> data; };
using TensorMap = std::map<std::string, HostTensor>;

inline bool read_container(const char* path, TensorMap& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    char magic[8]; uint32_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SVADW001", 8) || fread(&n, 4, 1, f) != 1 || n > 4096) { fclose(f); err = "bad magic"; return false; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t nl = 0, nd = 0;
        if (fread(&nl, 4, 1, f) != 1 || nl > 255) { fclose(f); err = "bad name"; return false; }
        std::string name(nl, 0);
        if (fread(&name[0], 1, nl, f) != nl || fread(&nd, 4, 1, f) != 1 || nd > 4) { fclose(f); err = "bad tensor header"; return false; }
        HostTensor t; t.dims.resize(nd); size_t numel = 1;
        for (uint32_t d = 0; d < nd; d++) {
            if (fread(&t.dims[d], 4, 1, f) != 1) { fclose(f); err = "bad dims"; return false; }
            numel *= t.dims[d];
            if (numel > (size_t)1 << 26) { fclose(f); err = "tensor too large (corrupt container?)"; return false; }   // 64 Mi floats: 250x the largest real tensor
        }
        t.data.resize(numel);
        if (fread(t.data.data(), 4, numel, f) != numel) { fclose(f); err = "truncated"; return false; }
        out[name] = std::move(t);
    }
    fclose(f);
    return true;
}

// ---- static slab schedule (shared by host packer, kernel and emulator)
template <bool SR16>
struct Tape {
    using G = Geo<SR16>;
    static constexpr int e0_off = 0;
    static constexpr int e1_off = e0_off + G::F * 384;
    static constexpr int e2_off = e1_off + 128 * 192;
    static constexpr int e3_off = e2_off + 64 * 128;
    static constexpr int l_off = e3_off + 64 * 128;
    static constexpr int total = l_off + 256 * 512;
    // enc0 slab s covers channels [e0_c0(s), e0_c0(s+1))
    SVAD_HD static constexpr int e0_c0(int s) {
        return s * (G::F / G::e0_nslab) + (s < G::F % G::e0_nslab ? s : G::F % G::e0_nslab);
    }
    // slab index -> (float offset, float count)
    SVAD_HD static constexpr int slab_off(int i) {
        if (i < G::e0_nslab) return e0_off + e0_c0(i) * 384;
        i -= G::e0_nslab;
        if (i < 4) return e1_off + i * 32 * 192;
        i -= 4;
        if (i == 0) return e2_off;
        if (i == 1) return e3_off;
        i -= 2;
        return l_off + i * 16 * 512;
    }
    SVAD_HD static constexpr int slab_len(int i) {
        if (i < G::e0_nslab) return (e0_c0(i + 1) - e0_c0(i)) * 384;
        i -= G::e0_nslab;
        if (i < 4) return 32 * 192;
        i -= 4;
        if (i < 2) return 64 * 128;
        return 16 * 512;
    }
};

struct PackedBranch {
    std::vector<float> tape;
    std::vector<float> consts;
};

template <bool SR16>
inline bool pack_branch(const TensorMap& tm, PackedBranch& out, std::string& err) {
    using G = Geo<SR16>;
    using T = Tape<SR16>;
    const std::string p = SR16 ? "_model." : "_model_8k.";
    auto get = [&](const char* s, std::vector<uint32_t> dims) -> const float* {
        auto it = tm.find(p + s);
        if (it == tm.end()) { err = "missing tensor " + p + s; return nullptr; }
        if (it->second.dims != dims) { err = "unexpected shape for " + p + s; return nullptr; }
        return it->second.data.data();
    };
    const uint32_t F = G::F;
    const float* w0 = get("encoder.0.reparam_conv.weight", {128, F, 3});
    const float* b0 = get("encoder.0.reparam_conv.bias", {128});
    const float* w1 = get("encoder.1.reparam_conv.weight", {64, 128, 3});
    const float* b1 = get("encoder.1.reparam_conv.bias", {64});
    const float* w2 = get("encoder.2.reparam_conv.weight", {64, 64, 3});
    const float* b2 = get("encoder.2.reparam_conv.bias", {64});
    const float* w3 = get("encoder.3.reparam_conv.weight", {128, 64, 3});
    const float* b3 = get("encoder.3.reparam_conv.bias", {128});
    const float* wih = get("decoder.rnn.weight_ih", {512, 128});
    const float* whh = get("decoder.rnn.weight_hh", {512, 128});
    const float* bih = get("decoder.rnn.bias_ih", {512});
    const float* bhh = get("decoder.rnn.bias_hh", {512});
    const float* wo = get("decoder.decoder.2.weight", {1, 128, 1});
    const float* bo = get("decoder.decoder.2.bias", {1});
    if (!w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !wih || !whh || !bih || !bhh || !wo || !bo) return false;

    out.tape.assign(T::total, 0.0f);
    float* t = out.tape.data();
    for (int c = 0; c < G::F; c++)
        for (int j = 0; j < 3; j++)
            for (int o = 0; o < 128; o++) t[T::e0_off + (c * 3 + j) * 128 + o] = w0[(o * G::F + c) * 3 + j];
    for (int c = 0; c < 128; c++)
        for (int j = 0; j < 3; j++)
            for (int o = 0; o < 64; o++) t[T::e1_off + (c * 3 + j) * 64 + o] = w1[(o * 128 + c) * 3 + j];
    for (int c = 0; c < 64; c++)
        for (int jj = 0; jj < 2; jj++)
            for (int o = 0; o < 64; o++) t[T::e2_off + (c * 2 + jj) * 64 + o] = w2[(o * 64 + c) * 3 + (jj + 1)];
    for (int c = 0; c < 64; c++)
        for (int o = 0; o < 128; o++) t[T::e3_off + c * 128 + o] = w3[(o * 64 + c) * 3 + 1];
    out.consts.assign(SmemMap::consts_floats, 0.0f);
    float* cs = out.consts.data();
    for (int w = 0; w < 8; w++)
        for (int u = 0; u < 2; u++)
            for (int l = 0; l < 8; l++)
                for (int g = 0; g < 4; g++) {
                    const int np = 64 * w + 32 * u + 4 * l + g, j = 16 * w + 2 * l + u, row = g * 128 + j;
                    for (int k = 0; k < 128; k++) {
                        t[T::l_off + k * 512 + np] = wih[row * 128 + k];
                        t[T::l_off + (128 + k) * 512 + np] = whh[row * 128 + k];
                    }
                    cs[SmemMap::c_bl + np] = bih[row] + bhh[row];
                }
    memcpy(cs + SmemMap::c_b0, b0, 128 * 4);
    memcpy(cs + SmemMap::c_b1, b1, 64 * 4);
    memcpy(cs + SmemMap::c_b2, b2, 64 * 4);
    memcpy(cs + SmemMap::c_b3, b3, 128 * 4);
    memcpy(cs + SmemMap::c_wout, wo, 128 * 4);
    cs[SmemMap::c_bout] = bo[0];
    for (int m = 0; m < G::N; m++)
        cs[SmemMap::c_win + m] = (float)(0.5 * (0.5 - 0.5 * cos(2.0 * M_PI * (double)m / (double)G::N)));
    return true;
}

// ---------------------------------------------------------------- tensor-core tape (svad_tc.h)
// Every dense layer runs on tcgen05 (kind::tf32, weights = M operand, stream slots = N, K = 8), split precision
//   x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo,   hi = fp32 value as is (the tensor core TRUNCATES fp32 containers to
// tf32: tools/umma_unit.cu), lo = v - trunc_tf32(v) (exact in fp32).  A weight slab holds K-major SWIZZLE_128B tiles
// (row r at (r/8)*1024 + (r%8)*128 B, 16-byte chunk (k/4)^(r%8)); slabs in consumption order:
//   enc0  24 / 12 slabs of 16 KB: [128 x 32] tiles, {hi, lo} alternating; pairs ordered tap 1 of every k-chunk first, then
//         taps 0, 2 chunk by chunk (e0_pair_kc / e0_pair_tap); rows o < 128, k = bins kc*32..+32 of W0[o][bin][tap]
//         (Kt = 128 / 64 bins; the Nyquist bin F-1 is a rank-1 update on the CUDA cores from consts c_wnyq)
//   enc1  12 slabs of 16 KB: [64 x 32] tiles {hi | lo}; taps in order 1, 2, 0, k-chunk kc < 4 of W1[o][c][tap]
//   enc2  4 slabs: [64 x 32] {hi | lo}; q = (tap - 1) * 2 + channel half of W2[o][c][tap], taps 1, 2
//   enc3  4 slabs: [128 x 32] tiles, (kc, hi | lo) of W3[o][c][tap 1]
//   lstm  32 slabs of 32 KB: [128 x 32] tile pairs {hi | lo}; for kc < 8, gate block m < 4: rows j < 128 (hidden unit),
//         k = kc*32..+32 of [W_ih ; W_hh][m*128 + j][k]
template <bool SR16>
struct TapeTC {
    using G = Geo<SR16>;
    static constexpr int Kt = G::F - 1;                  // 128 / 64 bins on the tensor core
    static constexpr int tile = 128 * 32;                // floats per [128 x 32] tile = one 16 KB slab
    static constexpr int e0_nslab = (Kt / 32) * 3 * 2;   // (kc, tap) x {hi, lo}: 24 / 12
    SVAD_HD static constexpr int e0_pair_kc(int pr) { return pr < Kt / 32 ? pr : (pr - Kt / 32) >> 1; }
    SVAD_HD static constexpr int e0_pair_tap(int pr) { return pr < Kt / 32 ? 1 : (((pr - Kt / 32) & 1) ? 2 : 0); }
    static constexpr int e1_nslab = 12;                  // (tap, kc): [64 x 32] tiles {hi | lo} = 16 KB, read by MMA warp kc
    static constexpr int e2_nslab = 4;                   // (tap, channel half): [64 x 32] tiles {hi | lo}, read by MMA warp q
    static constexpr int e3_nslab = 4;                   // (kc, hi | lo): [128 x 32] tiles, read by MMA warp 2 kc + lo
    static constexpr int l_nslab = 32;                   // (kc, gate block): one 32 KB slab = tile pair {hi | lo}, read by ONE MMA warp
    static constexpr int e0_off = 0;
    static constexpr int e1_off = e0_off + e0_nslab * tile;
    static constexpr int e2_off = e1_off + e1_nslab * tile;
    static constexpr int e3_off = e2_off + e2_nslab * tile;
    static constexpr int l_off = e3_off + e3_nslab * tile;
    static constexpr int total = l_off + l_nslab * 2 * tile;
    static constexpr int nslab = e0_nslab + e1_nslab + e2_nslab + e3_nslab + l_nslab;
    // ---- slab -> shared-memory buffer.  Buffers 0-3 are the 4 x 16 KB ring (enc0, enc1, enc2 tiles); buffers 4-7 are the e0
    // region, dead once enc1 has run: enc3's four slabs land there while enc2 still computes.  Every bulk copy costs the same
    // ~450 cycles up to 32 KB (tools/ubench_ingest.cu), so the LSTM streams 32 KB tile pairs through four double-size buffers
    // (ids 0, 2 = ring halves, 4, 6 = e0 region halves).  dep_delta(idx) = how many slabs back the event lies that frees the
    // buffer(s) of slab idx (slabs are issued, and their consumption observed, strictly in order).
    static constexpr int E3 = e0_nslab + e1_nslab + e2_nslab;   // first enc3 slab
    static constexpr int NA = E3 + e3_nslab;                    // slabs before the LSTM (a multiple of 4)
    static constexpr int kBufs = 8;
    SVAD_HD static constexpr int buf(int idx) { return idx < E3 ? (idx & 3) : (idx < NA ? 4 + (idx - E3) : (((idx - NA) & 3) << 1)); }
    SVAD_HD static constexpr int dep_delta(int idx) {
        if (idx < E3) return idx >= 4 ? 4 : (idx == 0 ? 4 : (idx == 3 ? 6 : 5));   // first slabs of a step wait for the last LSTM pairs
        if (idx < NA) return idx - (e0_nslab + e1_nslab - 1);                       // enc3: the last enc1 slab (reader of the e0 region)
        const int l = idx - NA;
        return l >= 4 ? 4 : 7 - l;   // pairs 0, 1 follow enc2 slabs (0,1), (2,3) in the ring; pairs 2, 3 follow enc3 slabs (0,1), (2,3)
    }
    template <class M>
    SVAD_HD static constexpr int buf_off(int b) { return (b & 4) ? M::e0 + (b & 3) * M::stage_floats : M::stage + b * M::stage_floats; }
    // parity toggles a warp misses when it sits out an MMA phase: XOR of (1 << buf) over the phase's slabs
    SVAD_HD static constexpr uint32_t phase_mask(int first, int count) {
        uint32_t m = 0;
        for (int i = 0; i < count; i++) m ^= 1u << buf(first + i);
        return m;
    }
    SVAD_HD static constexpr int slab_off(int i) {
        if (i < e0_nslab) return e0_off + i * tile;
        i -= e0_nslab;
        if (i < e1_nslab) return e1_off + i * tile;
        i -= e1_nslab;
        if (i < e2_nslab) return e2_off + i * tile;
        i -= e2_nslab;
        if (i < e3_nslab) return e3_off + i * tile;
        i -= e3_nslab;
        return l_off + i * 2 * tile;
    }
    SVAD_HD static constexpr int slab_len(int i) {
        if (i < e0_nslab) return tile;
        i -= e0_nslab;
        return i < e1_nslab + e2_nslab + e3_nslab ? tile : 2 * tile;   // enc1-3: 16 KB; LSTM: a [128 x 32] tile pair
    }
};

inline float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

// A[r][k] (r < 128, k < 32, row stride lda floats, k stride ldk) -> K-major SWIZZLE_128B tile pair {hi | lo}
inline void pack_umma_a(const float* A, long lda, long ldk, float* dst) {
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < 32; k++) {
            const float v = A[r * lda + k * ldk];
            const int pos = (r / 8) * 256 + (r % 8) * 32 + (((k / 4) ^ (r % 8)) * 4) + (k % 4);
            dst[pos] = v;
            dst[128 * 32 + pos] = v - trunc_tf32(v);
        }
}

// A[r][k] (r < 64, k < 32) -> K-major SWIZZLE_128B half tile pair {hi | lo}, 2048 floats each (M = 64 instructions)
inline void pack_umma_a64(const float* A, long lda, long ldk, float* dst) {
    for (int r = 0; r < 64; r++)
        for (int k = 0; k < 32; k++) {
            const float v = A[r * lda + k * ldk];
            const int pos = (r / 8) * 256 + (r % 8) * 32 + (((k / 4) ^ (r % 8)) * 4) + (k % 4);
            dst[pos] = v;
            dst[64 * 32 + pos] = v - trunc_tf32(v);
        }
}

template <bool SR16>
inline bool pack_branch_tc(const TensorMap& tm, PackedBranch& out, std::string& err) {
    using G = Geo<SR16>;
    using T = TapeTC<SR16>;
    using T1 = Tape<SR16>;
    PackedBranch v1;
    if (!pack_branch<SR16>(tm, v1, err)) return false;
    const std::string p = SR16 ? "_model." : "_model_8k.";
    const float* w0 = tm.at(p + "encoder.0.reparam_conv.weight").data.data();   // [128][F][3]
    const float* wih = tm.at(p + "decoder.rnn.weight_ih").data.data();          // [512][128]
    const float* whh = tm.at(p + "decoder.rnn.weight_hh").data.data();
    const float* bih = tm.at(p + "decoder.rnn.bias_ih").data.data();
    const float* bhh = tm.at(p + "decoder.rnn.bias_hh").data.data();
    out.tape.assign(T::total, 0.0f);
    float* t = out.tape.data();
    // enc0 tile pairs {hi | lo} in issue order: first tap 1 of every k-chunk (the only tap that reaches all four frames, so the
    // first instruction into each of the four accumulators overwrites every column), then taps 0 and 2 chunk by chunk
    for (int pr = 0; pr < T::e0_nslab / 2; pr++) {
        const int kc = T::e0_pair_kc(pr), j = T::e0_pair_tap(pr);
        pack_umma_a(w0 + (size_t)(kc * 32) * 3 + j, (long)G::F * 3, 3, t + T::e0_off + pr * 2 * T::tile);   // {hi | lo} = 2 slabs
    }
    const float* w1 = tm.at(p + "encoder.1.reparam_conv.weight").data.data();   // [64][128][3]
    for (int jo = 0; jo < 3; jo++) {   // tap order 1, 2, 0: the first slab of every MMA warp covers both output frames
        const int j = jo == 2 ? 0 : jo + 1;
        for (int kc = 0; kc < 4; kc++) pack_umma_a64(w1 + (size_t)(kc * 32) * 3 + j, 128 * 3, 3, t + T::e1_off + (jo * 4 + kc) * T::tile);
    }
    const float* w2 = tm.at(p + "encoder.2.reparam_conv.weight").data.data();   // [64][64][3], taps 1, 2 live
    for (int q = 0; q < 4; q++) pack_umma_a64(w2 + (size_t)((q & 1) * 32) * 3 + (q >> 1) + 1, 64 * 3, 3, t + T::e2_off + q * T::tile);
    const float* w3 = tm.at(p + "encoder.3.reparam_conv.weight").data.data();   // [128][64][3], tap 1 live
    for (int kc = 0; kc < 2; kc++) pack_umma_a(w3 + (size_t)(kc * 32) * 3 + 1, 64 * 3, 3, t + T::e3_off + kc * 2 * T::tile);   // {hi | lo} = 2 slabs
    for (int kc = 0; kc < 8; kc++)
        for (int m = 0; m < 4; m++) {
            const float* src = (kc < 4 ? wih : whh) + (size_t)(m * 128) * 128 + (kc & 3) * 32;
            pack_umma_a(src, 128, 1, t + T::l_off + (kc * 4 + m) * 2 * T::tile);
        }
    out.consts.assign(SmemMapTC::consts_floats, 0.0f);
    float* cs = out.consts.data();
    memcpy(cs, v1.consts.data(), sizeof(float) * SmemMapTC::c_twr);             // b0..b3, (bl), wout, bout, window
    for (int g = 0; g < 512; g++) cs[SmemMapTC::c_bl + g] = bih[g] + bhh[g];    // natural gate order here
    for (int j = 0; j < 3; j++)
        for (int o = 0; o < 128; o++) cs[SmemMapTC::c_wnyq + j * 128 + o] = w0[((size_t)o * G::F + (G::F - 1)) * 3 + j];
    return true;
}

}  // namespace svad
