// svad_tc.h -- tensor-core variant of the fused kernel: enc0 (47 % of the MACs) and the LSTM cell (37 %) run on
// tcgen05 (5th-gen tensor cores, accumulators in TMEM); the STFT, enc1-3, gate math and head stay on the CUDA cores.
//
// Orientation: the WEIGHTS are the M = 128 operand (A, K-major SWIZZLE_128B tiles streamed from the tape), the
// stream slots are N = 32 (B, MN-major SWIZZLE_128B_BASE32B = the activation rows as they already sit in shared
// memory), K = 8 per instruction, fp32 accumulate.  tcgen05 cost is proportional to N, so a 28-32 stream tile per
// CTA keeps the tensor pipe efficient while the activations of a tile still fit one SM.
//
// Precision: plain TF32 fails parity (2e-3).  Split precision x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// with hi = the fp32 container as is (the tensor core truncates it to tf32; measured in tools/umma_unit.cu) and
// lo = v - trunc_tf32(v), exact in fp32: three MMAs per k-step, products carry ~21 mantissa bits.
// The activation buffer itself is the "hi" operand; only the "lo" rows are staged (one pass of the CUDA cores).
//
// Env (GPU: svad_api.cu, CPU emulator: tests/emu) adds to the fp32 kernel's interface:
//   mma(col, a_tile, ks, b_rows, acc)   D[128 x 32 @ TMEM column col] (+)= A_tile[:, ks*8..+8] * B_rows[8 x 32]
//   mma_slab_done(it) / acc_commit()    tcgen05.commit to the stage's / the layer's mbarrier      (issuer thread)
//   acc_wait()                          all threads: the layer's accumulators are complete
//   tmem_ld16(lane_quarter, col, v)     16 consecutive columns of this thread's TMEM lane
//   fence_async()                       make generic-proxy smem writes visible to the MMA (async proxy)
//   tc_fence_before() / tc_fence_after()  order tcgen05.ld against later MMAs across a CTA barrier
//   slab_wait(it)                       weight ring: wait until slab it has landed in its stage
//   thread 0 only (it keeps `issued` / `freed` counters, so the calls are idempotent):
//     mark_free(it)     slab it was consumed by the CUDA cores (a CTA barrier has passed)
//     free_upto(it)     wait until the MMAs reading every slab <= it have completed
//     refill_upto(it)   issue the TMA copies of all not yet issued slabs <= it (slab i reuses the stage of i-2)
#pragma once
#include "svad_tile.h"

namespace svad {

SVAD_HD float lo_part(float v) {   // v - trunc_tf32(v), exact
#if defined(__CUDA_ARCH__)
    return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
    uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float t; memcpy(&t, &u, 4); return v - t;
#endif
}

// rows [0, nrows) of `src` (32 floats each) -> lo parts into `dst`, same row layout; all threads
SVAD_HD void stage_lo(int tid, const float* src, float* dst, int nrows) {
    for (int i = tid; i < nrows * 8; i += kThreads) {
        const f4 v = *reinterpret_cast<const f4*>(src + i * 4);
        *reinterpret_cast<f4*>(dst + i * 4) = f4{lo_part(v.x), lo_part(v.y), lo_part(v.z), lo_part(v.w)};
    }
}

template <bool SR16, int RM, typename S, class Env>
SVAD_HD void run_cta_tc(Env& env, const TileArgs& a, int first_tile, int tile_stride, int ntiles) {
    using G = Geo<SR16>;
    using TP = TapeTC<SR16>;
    using M = SmemMapTC;
    constexpr int Kt = TP::Kt, KC0 = Kt / 32;
    const Tc tc(env.tid());
    float* sm = env.smem();
    const S* audio = static_cast<const S*>(a.audio);
    Regs rg;
    constexpr int BT = 4 * RM;
    // epilogue coordinates: TMEM lane = weight row; warps w and w+4 share lane quarter w%4 and split the 32 slot columns
    const int lq = tc.warp & 3, row = 32 * lq + tc.lane, half = tc.warp >> 2;

    for (int i = tc.tid; i < M::c_twr; i += kThreads) sm[M::consts + i] = a.consts[i];
    for (int i = tc.tid; i < 384; i += kThreads) sm[M::consts + M::c_wnyq + i] = a.consts[M::c_wnyq + i];
    {
        const int k = tc.tid >> 4, r = tc.tid & 15;
        float s = 0.0f, c = 1.0f;
        if (k < G::NQ) {
            const float x = -2.0f * (float)((k * r) % G::N) / (float)G::N;
#if defined(__CUDA_ARCH__)
            sincospif(x, &s, &c);
#else
            s = (float)sin(M_PI * (double)x); c = (float)cos(M_PI * (double)x);
#endif
        }
        sm[M::consts + M::c_twr + tc.tid] = c;
        sm[M::consts + M::c_twi + tc.tid] = s;
    }
    int my_tiles = 0;
    for (int tile = first_tile; tile < ntiles; tile += tile_stride) my_tiles++;
    const long total_slabs = (long)my_tiles * a.T * TP::nslab;
    long it = 0;
    float cst[16];   // LSTM cell state of hidden unit `row` for slots 16*half .. +16

    for (int tile = first_tile; tile < ntiles; tile += tile_stride) {
        const int g0 = tile * BT;
        env.sync();
        for (int i = tc.tid; i < kHid * kSlots; i += kThreads) {
            const int j = i >> 5, s = i & 31;
            const int g = g0 + slot_to_local<RM>(s);
            float v = 0.0f;
            if (a.state_in && slot_valid<RM>(s) && g < a.B) v = a.state_in[(long)g * kHid + j];
            sm[M::h + j * kSlots + tc_slot(s, j)] = v;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int s = 16 * half + i, g = g0 + slot_to_local<RM>(s);
            cst[i] = (a.state_in && slot_valid<RM>(s) && g < a.B) ? a.state_in[((long)a.B + g) * kHid + row] : 0.0f;
        }
        const S* aud[2];
        const float* cxp[2];
#pragma unroll
        for (int hs = 0; hs < 2; hs++) {
            const int sl = 16 * hs + (tc.tid >> 4);
            const int gs = g0 + slot_to_local<RM>(sl);
            const bool v = slot_valid<RM>(sl) && gs < a.B;
            aud[hs] = v ? audio + (long)gs * a.ld : nullptr;
            cxp[hs] = (v && a.ctx_in) ? a.ctx_in + (long)gs * a.ctx_ld : nullptr;
        }
        env.sync();

        float xa[G::NQ], xb[G::NQ];
        for (long t = 0; t < a.T; t++) {
            // ---------------- STFT (CUDA cores) -> mag rows in tcgen05 atom layout
            const bool fast = (t > 0) && ((t + 1) * G::n <= a.L);
            if (t + 1 < a.T) {
                constexpr int kPerLine = 128 / (int)sizeof(S), kLines = G::n / kPerLine;
                for (int i = tc.tid; i < BT * kLines; i += kThreads) {
                    const int loc = i / kLines, line = i % kLines, g = g0 + loc;
                    const long off = (t + 1) * G::n + line * kPerLine;
                    if (g < a.B && off < a.L) env.prefetch_l2(audio + (long)g * a.ld + off);
                }
            }
            if (t == 0) stft_load<SR16, S>(tc.tid, 0, aud[0], cxp[0], a.L, t, fast, xa, xb);
#pragma unroll 1
            for (int rnd = 0; rnd < 4; rnd++) {
                const int hs = rnd >> 1, fp = rnd & 1;
                float na[G::NQ], nb[G::NQ];
                if (rnd < 3) stft_load<SR16, S>(tc.tid, (rnd + 1) & 1, rnd >= 1 ? aud[1] : aud[0], rnd >= 1 ? cxp[1] : cxp[0], a.L, t, fast, na, nb);
                stft_pass_a<SR16, M>(tc, sm, xa, xb);
                env.sync();
#pragma unroll
                for (int kk = 0; kk < G::NQ / 8; kk++) stft_pass_c<SR16, M>(tc, sm, hs, fp, tc.warp * (G::NQ / 8) + kk);
                env.sync();
                if (rnd < 3) {
#pragma unroll
                    for (int q = 0; q < G::NQ; q++) { xa[q] = na[q]; xb[q] = nb[q]; }
                }
            }
            // ---------------- enc0 on the tensor core
            // lo rows of mag[f][0..Kt) -> lo0[f][0..Kt)  (the Z planes there are dead now)
#pragma unroll
            for (int f = 0; f < 4; f++) stage_lo(tc.tid, sm + M::mag + f * M::mag_pitch * kSlots, sm + M::lo0 + f * Kt * kSlots, Kt);
            env.fence_async();
            env.tc_fence_before();     // the previous step's tcgen05.ld of these TMEM columns are done
            env.sync();
            if (tc.warp == 0) {
                if (tc.lane == 0) {
                    env.tc_fence_after();
#pragma unroll 1
                    for (int s = 0; s < TP::e0_nslab; s++) {
                        const long is = it + s;
                        const int kc = s / 3, j = s % 3;
                        const float* slab = env.slab_wait(is);
#pragma unroll 1
                        for (int tt = 0; tt < 4; tt++) {
                            const int f = tt + j - 1;
                            if (f < 0 || f > 3) continue;
                            const float* bh = sm + M::mag + (f * M::mag_pitch + kc * 32) * kSlots;
                            const float* bl = sm + M::lo0 + (f * Kt + kc * 32) * kSlots;
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                const bool first = (kc == 0) && (ks == 0) && (j == (tt == 0 ? 1 : 0));
                                env.mma(tt * 32, slab, ks, bh + ks * 8 * kSlots, !first);
                                env.mma(tt * 32, slab, ks, bl + ks * 8 * kSlots, true);
                                env.mma(tt * 32, slab + TP::tile, ks, bh + ks * 8 * kSlots, true);
                            }
                        }
                        env.mma_slab_done(is);
                        env.free_upto(is - 1);
                        env.refill_upto(is + 1, total_slabs);
                    }
                    env.acc_commit();
                }
                env.warp_sync();
            }
            it += TP::e0_nslab;
            env.acc_wait();
            // epilogue: this thread owns channel `row`, frames 2*half, 2*half+1: + bias + Nyquist-bin rank-1 term, ReLU -> e0
            {
                const float b0 = sm[M::consts + M::c_b0 + row];
                const float wn0 = sm[M::consts + M::c_wnyq + row], wn1 = sm[M::consts + M::c_wnyq + 128 + row],
                            wn2 = sm[M::consts + M::c_wnyq + 256 + row];
#pragma unroll
                for (int ff = 0; ff < 2; ff++) {
                    const int tt = 2 * half + ff;
                    float v[32];
                    env.tmem_ld16(lq, tt * 32, *reinterpret_cast<float(*)[16]>(v));
                    env.tmem_ld16(lq, tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
                    const float* nyq = sm + M::mag + (M::mag_pitch * tt + Kt) * kSlots;   // row Kt of frame tt (identity swizzle)
#pragma unroll
                    for (int s = 0; s < 32; s++) {
                        float acc = v[s] + b0;
                        if (tt > 0) acc = fmaf(wn0, nyq[s - M::mag_pitch * kSlots], acc);
                        acc = fmaf(wn1, nyq[s], acc);
                        if (tt < 3) acc = fmaf(wn2, nyq[s + M::mag_pitch * kSlots], acc);
                        v[s] = relu(acc);
                    }
                    float* dst = sm + M::e0 + (tt * 128 + row) * kSlots;
#pragma unroll
                    for (int gq = 0; gq < 8; gq++)
                        *reinterpret_cast<f4*>(dst + ((gq ^ key_hi(row)) << 2)) = f4{v[4 * gq], v[4 * gq + 1], v[4 * gq + 2], v[4 * gq + 3]};
                }
            }
            env.tc_fence_before();
            env.sync();
            if (tc.tid == 0) {   // the stage of the last enc0 slab is free now
                env.free_upto(it - 1);
                env.refill_upto(it + 1, total_slabs);
            }
            // ---------------- enc1..enc3 on the CUDA cores (as in the fp32 kernel)
            enc1_init<RM, M>(tc, sm, rg);
#pragma unroll 1
            for (int s = 0; s < 4; s++, it++) {
                const float* slab = env.slab_wait(it);
                enc1_slab<RM, M>(tc, sm, slab, rg, s * 32, s * 32 + 32);
                if (s == 3) enc1_store<RM, M>(tc, sm, rg);
                env.sync();
                if (tc.tid == 0) { env.mark_free(it); env.refill_upto(it + 2, total_slabs); }
            }
            {
                const float* slab = env.slab_wait(it);
                enc2_all<RM, M>(tc, sm, slab, rg);
                env.sync();
                if (tc.tid == 0) { env.mark_free(it); env.refill_upto(it + 2, total_slabs); }
                it++;
                slab = env.slab_wait(it);
                enc3_all<RM, M>(tc, sm, slab, rg);
                // lo rows of h (carried from the previous step) can be staged before the barrier
                stage_lo(tc.tid, sm + M::h, sm + M::lol + kHid * kSlots, kHid);
                env.sync();
                if (tc.tid == 0) { env.mark_free(it); env.refill_upto(it + 2, total_slabs); }
                it++;
            }
            // ---------------- LSTM on the tensor core: gates[m*128 + j][slot] = sum_k W[.][k] * [e3 ; h][k][slot]
            stage_lo(tc.tid, sm + M::e3, sm + M::lol, kHid);
            env.fence_async();
            env.sync();
            if (tc.warp == 0) {
                if (tc.lane == 0) {
                    env.tc_fence_after();
#pragma unroll 1
                    for (int s = 0; s < 32; s++) {
                        const long is = it + s;
                        const int kc = s >> 2, m = s & 3;
                        const float* slab = env.slab_wait(is);
                        const float* bh = (kc < 4) ? sm + M::e3 + kc * 32 * kSlots : sm + M::h + (kc - 4) * 32 * kSlots;
                        const float* bl = sm + M::lol + kc * 32 * kSlots;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const bool first = (kc == 0) && (ks == 0);
                            env.mma(128 + m * 32, slab, ks, bh + ks * 8 * kSlots, !first);
                            env.mma(128 + m * 32, slab, ks, bl + ks * 8 * kSlots, true);
                            env.mma(128 + m * 32, slab + TP::tile, ks, bh + ks * 8 * kSlots, true);
                        }
                        env.mma_slab_done(is);
                        env.free_upto(is - 1);
                        env.refill_upto(is + 1, total_slabs);
                    }
                    env.acc_commit();
                }
                env.warp_sync();
            }
            it += 32;
            env.acc_wait();
            // epilogue: hidden unit `row`, slots 16*half..+16
            {
                float gi[16], gf[16], gg[16], go[16];
                env.tmem_ld16(lq, 128 + 0 * 32 + 16 * half, gi);
                env.tmem_ld16(lq, 128 + 1 * 32 + 16 * half, gf);
                env.tmem_ld16(lq, 128 + 2 * 32 + 16 * half, gg);
                env.tmem_ld16(lq, 128 + 3 * 32 + 16 * half, go);
                const float* bl = sm + M::consts + M::c_bl;
                const float bi = bl[row], bf = bl[128 + row], bg = bl[256 + row], bo = bl[384 + row];
                float hv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float ig = sigmoid_acc(gi[i] + bi), fg = sigmoid_acc(gf[i] + bf), g2 = tanhf(gg[i] + bg), og = sigmoid_acc(go[i] + bo);
                    const float cn = fmaf(fg, cst[i], ig * g2);
                    cst[i] = cn;
                    hv[i] = og * tanhf(cn);
                }
                float* hrow = sm + M::h + row * kSlots;
#pragma unroll
                for (int gq = 0; gq < 4; gq++)   // logical float4 groups 4*half + gq
                    *reinterpret_cast<f4*>(hrow + (tc_f4(4 * half + gq, row) << 2)) = f4{hv[4 * gq], hv[4 * gq + 1], hv[4 * gq + 2], hv[4 * gq + 3]};
            }
            env.tc_fence_before();
            env.sync();
            if (tc.tid == 0) {
                env.free_upto(it - 1);
                env.refill_upto(it + 1, total_slabs);
            }
            if (tc.tid < kSlots) {
                const int g = g0 + slot_to_local<RM>(tc.tid);
                if (slot_valid<RM>(tc.tid) && g < a.B) {
                    const float* wout = sm + M::consts + M::c_wout;
                    float a0 = sm[M::consts + M::c_bout], a1 = 0.f;
#pragma unroll 8
                    for (int j = 0; j < kHid; j += 2) {
                        a0 = fmaf(wout[j], relu(sm[M::h + j * kSlots + tc_slot(tc.tid, j)]), a0);
                        a1 = fmaf(wout[j + 1], relu(sm[M::h + (j + 1) * kSlots + tc_slot(tc.tid, j + 1)]), a1);
                    }
                    a.probs[(long)g * a.ldp + t] = sigmoid_acc(a0 + a1);
                }
            }
            if (t + 1 < a.T) stft_load<SR16, S>(tc.tid, 0, aud[0], cxp[0], a.L, t + 1, ((t + 2) * G::n <= a.L), xa, xb);
        }
        // ---- tile exit
        env.sync();
        if (a.state_out) {
            for (int i = tc.tid; i < kHid * kSlots; i += kThreads) {
                const int s = i & 31, j = i >> 5;
                const int g = g0 + slot_to_local<RM>(s);
                if (slot_valid<RM>(s) && g < a.B) a.state_out[(long)g * kHid + j] = sm[M::h + j * kSlots + tc_slot(s, j)];
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int s = 16 * half + i, g = g0 + slot_to_local<RM>(s);
                if (slot_valid<RM>(s) && g < a.B) a.state_out[((long)a.B + g) * kHid + row] = cst[i];
            }
        }
        if (a.ctx_out) {
            for (int i = tc.tid; i < BT * G::ctx; i += kThreads) {
                const int loc = i / G::ctx, k = i % G::ctx, g = g0 + loc;
                if (g < a.B) {
                    const float* cx = a.ctx_in ? a.ctx_in + (long)g * a.ctx_ld : nullptr;
                    a.ctx_out[(long)g * G::ctx + k] = (a.T > 0) ? window_sample<SR16, S>(audio + (long)g * a.ld, a.L, cx, a.T - 1, G::n + k)
                                                                : (cx ? cx[k] : 0.0f);
                }
            }
        }
    }
}

}  // namespace svad
