// svad_small.h -- small-batch / low-latency kernel: one thread-block CLUSTER of 8 CTAs per group of 4 streams.
//
// The tile kernels (svad_tile.h, svad_tc.h) need >= 4096 streams to fill the GPU: one SM computes a whole 32-slot tile and
// streams all 0.9 MB of weights through its shared memory every step, so a single stream costs ~60 us per chunk.  Here the
// NETWORK is split instead: CTA r of the cluster keeps 1/8 of every layer's output channels (and of the DFT basis) resident
// in its shared memory for the whole launch (146 KB), computes that slice of each layer as warp-per-row dot products, and
// all-gathers the slice into the 8 CTAs' activation buffers through distributed shared memory (st.shared::cluster), one
// cluster barrier per layer.  No weight traffic per step at all; 6 cluster barriers + ~80 k MAC per CTA per step.
// Used for B <= 4 * (SMs / 8) streams (config 2 of BASELINE.json: batch = 1 streaming, and any small batch).
//
// The STFT here is the dense DFT-basis product (the slice is only 17 bins per CTA), i.e. the reference's own formulation
// (silero_vad.jit::_model.stft.transform_) with the basis rebuilt analytically: w[m] cos / -w[m] sin (2 pi k m / N),
// periodic Hann w (equal to forward_basis_buffer to 7.7e-8).
#pragma once
#include "svad_pack.h"

namespace svad {

constexpr int kSmallCtas = 8;      // cluster size
constexpr int kSmallNS = 4;        // streams per cluster (interleaved innermost in every activation buffer)
constexpr int kSmallThreads = 256;

template <bool SR16>
struct SmallMap {
    using G = Geo<SR16>;
    static constexpr int BPC = (G::F + kSmallCtas - 1) / kSmallCtas;   // bins per CTA: 17 / 9
    // ---- weight blob of one CTA (float offsets), every matrix K-major: [k][row] with the rows of the slice contiguous,
    // so that a warp whose lanes own consecutive rows reads it conflict-free and the activation x[k] is a broadcast
    static constexpr int RB = 2 * BPC;                                  // basis rows of the slice: (bin, re|im)
    static constexpr int w_basis = 0;                                   // [N][RB]
    static constexpr int w_e0 = w_basis + G::N * RB;                    // [3][F][16]
    static constexpr int w_b0 = w_e0 + 16 * 3 * G::F;                   // [16]
    static constexpr int w_e1 = w_b0 + 16;                              // [3][128][8]
    static constexpr int w_b1 = w_e1 + 8 * 384;                         // [8]
    static constexpr int w_e2 = w_b1 + 8;                               // [2][64][8]
    static constexpr int w_b2 = w_e2 + 8 * 128;                         // [8]
    static constexpr int w_e3 = w_b2 + 8;                               // [64][16]
    static constexpr int w_b3 = w_e3 + 16 * 64;                         // [16]
    static constexpr int w_l = w_b3 + 16;                               // [256][64], row = unit * 4 + gate
    static constexpr int w_bl = w_l + 16 * 4 * 256;                     // [64]
    static constexpr int w_out = w_bl + 64;                             // [128] + bout
    static constexpr int blob_floats = (w_out + 129 + 3) / 4 * 4;
    // ---- activations (every CTA holds full copies), [..][kSmallNS]
    static constexpr int a_xp = blob_floats;                            // [L1 + N/4][4]
    static constexpr int a_mag = a_xp + (G::L1 + G::N / 4) * 4;         // [4][F][4]
    static constexpr int a_e0 = a_mag + 4 * G::F * 4;                   // [4][128][4]
    static constexpr int a_e1 = a_e0 + 4 * 128 * 4;                     // [2][64][4]
    static constexpr int a_e2 = a_e1 + 2 * 64 * 4;                      // [64][4]
    static constexpr int a_xh = a_e2 + 64 * 4;                          // [256][4]: e3 (0..127) then h (128..255) = the LSTM input
    static constexpr int a_gates = a_xh + 256 * 4;                      // [16][4][4]  local
    static constexpr int a_c = a_gates + 256;                           // [16][4]     local cell state
    static constexpr int a_h2 = a_c + 64;                               // [128][4]    second copy of h (ping-pong by step parity)
    static constexpr int a_red = a_h2 + 128 * 4;                        // [256] float4 partial sums of the split-K dot products
    static constexpr int total_floats = a_red + 256 * 4;
};

template <bool SR16>
inline void pack_small(const TensorMap& tm, std::vector<float>& blobs /* [8][blob_floats] */) {
    using G = Geo<SR16>;
    using M = SmallMap<SR16>;
    const std::string p = SR16 ? "_model." : "_model_8k.";
    const float* w0 = tm.at(p + "encoder.0.reparam_conv.weight").data.data();
    const float* b0 = tm.at(p + "encoder.0.reparam_conv.bias").data.data();
    const float* w1 = tm.at(p + "encoder.1.reparam_conv.weight").data.data();
    const float* b1 = tm.at(p + "encoder.1.reparam_conv.bias").data.data();
    const float* w2 = tm.at(p + "encoder.2.reparam_conv.weight").data.data();
    const float* b2 = tm.at(p + "encoder.2.reparam_conv.bias").data.data();
    const float* w3 = tm.at(p + "encoder.3.reparam_conv.weight").data.data();
    const float* b3 = tm.at(p + "encoder.3.reparam_conv.bias").data.data();
    const float* wih = tm.at(p + "decoder.rnn.weight_ih").data.data();
    const float* whh = tm.at(p + "decoder.rnn.weight_hh").data.data();
    const float* bih = tm.at(p + "decoder.rnn.bias_ih").data.data();
    const float* bhh = tm.at(p + "decoder.rnn.bias_hh").data.data();
    const float* wo = tm.at(p + "decoder.decoder.2.weight").data.data();
    const float* bo = tm.at(p + "decoder.decoder.2.bias").data.data();
    blobs.assign((size_t)kSmallCtas * M::blob_floats, 0.0f);
    for (int r = 0; r < kSmallCtas; r++) {
        float* d = blobs.data() + (size_t)r * M::blob_floats;
        for (int lb = 0; lb < M::BPC; lb++) {
            const int k = r * M::BPC + lb;
            if (k >= G::F) continue;
            for (int m = 0; m < G::N; m++) {
                const double w = 0.5 - 0.5 * cos(2.0 * M_PI * m / G::N), ang = 2.0 * M_PI * (double)((long)k * m % G::N) / G::N;
                d[M::w_basis + m * M::RB + lb * 2 + 0] = (float)(w * cos(ang));
                d[M::w_basis + m * M::RB + lb * 2 + 1] = (float)(-w * sin(ang));
            }
        }
        for (int o = 0; o < 16; o++) {
            for (int j = 0; j < 3; j++)
                for (int c = 0; c < G::F; c++) d[M::w_e0 + (j * G::F + c) * 16 + o] = w0[((16 * r + o) * G::F + c) * 3 + j];
            d[M::w_b0 + o] = b0[16 * r + o];
            for (int c = 0; c < 64; c++) d[M::w_e3 + c * 16 + o] = w3[((16 * r + o) * 64 + c) * 3 + 1];
            d[M::w_b3 + o] = b3[16 * r + o];
            for (int g = 0; g < 4; g++) {
                const int row = g * 128 + 16 * r + o;
                for (int k = 0; k < 128; k++) {
                    d[M::w_l + k * 64 + o * 4 + g] = wih[row * 128 + k];
                    d[M::w_l + (128 + k) * 64 + o * 4 + g] = whh[row * 128 + k];
                }
                d[M::w_bl + o * 4 + g] = bih[row] + bhh[row];
            }
        }
        for (int o = 0; o < 8; o++) {
            for (int j = 0; j < 3; j++)
                for (int c = 0; c < 128; c++) d[M::w_e1 + (j * 128 + c) * 8 + o] = w1[((8 * r + o) * 128 + c) * 3 + j];
            d[M::w_b1 + o] = b1[8 * r + o];
            for (int jj = 0; jj < 2; jj++)
                for (int c = 0; c < 64; c++) d[M::w_e2 + (jj * 64 + c) * 8 + o] = w2[((8 * r + o) * 64 + c) * 3 + jj + 1];
            d[M::w_b2 + o] = b2[8 * r + o];
        }
        for (int j = 0; j < 128; j++) d[M::w_out + j] = wo[j];
        d[M::w_out + 128] = bo[0];
    }
}

#if defined(__CUDACC__)
// thread-per-row partial dot product over k in [k0, k1): w is K-major with `rows` rows per k (this thread owns `row`),
// x4[k] holds the 4 interleaved streams (a warp-wide broadcast read)
__device__ __forceinline__ void dotT(const float* __restrict__ w, int rows, int row, const float4* __restrict__ x4, int k0, int k1, float4& acc) {
    const float* wp = w + (size_t)k0 * rows + row;
    int k = k0;
#pragma unroll 4
    for (; k < k1; k++, wp += rows) {
        const float wv = *wp;
        const float4 x = x4[k];
        acc.x = fmaf(wv, x.x, acc.x); acc.y = fmaf(wv, x.y, acc.y); acc.z = fmaf(wv, x.z, acc.z); acc.w = fmaf(wv, x.w, acc.w);
    }
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
#endif

}  // namespace svad
