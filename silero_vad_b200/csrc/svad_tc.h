// svad_tc.h -- tensor-core variant of the fused kernel: the four encoder convolutions and the LSTM cell (all dense-layer
// MACs) run on tcgen05 (5th-gen tensor cores, accumulators in TMEM); the STFT, the epilogues (bias, ReLU, lo split, gate
// math) and the head stay on the CUDA cores.
//
// Orientation: the WEIGHTS are the M operand (A, K-major SWIZZLE_128B tiles streamed from the tape: 128 rows for enc0 /
// enc3 / a gate block, 64 rows for enc1 / enc2), the stream slots are N (B, MN-major SWIZZLE_128B_BASE32B = the activation
// rows as they already sit in shared memory; several 32-slot atoms -- frames, or hi | lo rows -- at the descriptor's LBO
// stride), K = 8 per instruction, fp32 accumulate.  One tcgen05.mma costs 41 / 49 / 65 cycles of the tensor pipe at
// N = 32 / 64 / 128 (tools/ubench_umma.cu), so instructions are made as wide in N as the layer allows.
//
// Precision: plain TF32 fails parity (2e-3).  Split precision x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// with hi = the fp32 container as is (the tensor core truncates it to tf32; measured in tools/umma_unit.cu) and
// lo = v - trunc_tf32(v), exact in fp32; products carry ~21 mantissa bits.  The activation buffer itself is the "hi"
// operand; the producing epilogue writes the "lo" rows next to it (mag and h are split by one CUDA-core pass).
//
// Env (GPU: svad_api.cu, CPU emulator: tests/emu) adds to the fp32 kernel's interface:
//   mma_ks4<M, NP>(col, a0, b0, a1, b1, a2, b2, acc, n, n12)   the 4 k-steps of one 32-wide k-chunk, per k-step NP descriptor
//                                       pairs (A_i, B_i) into D[M x n @ TMEM column col], issued from ONE elected region
//   mma_slab_done(it) / acc_commit()    tcgen05.commit to the stage's / the layer's mbarrier      (issuer thread)
//   acc_wait()                          all threads: the layer's accumulators are complete
//   tmem_ld16(lane_quarter, col, v)     16 consecutive columns of this thread's TMEM lane
//   fence_async()                       make generic-proxy smem writes visible to the MMA (async proxy)
//   tc_fence_before() / tc_fence_after()  order tcgen05.ld against later MMAs across a CTA barrier
//   slab_wait(idx) / slab_pass(idx)     weight ring: wait until slab idx (per-step index) has landed, returns its buffer / a warp
//                                       that does not read the slab only keeps that buffer's parity bit in step
//   skip_phase(mask, n)                 a warp that sat out an MMA phase fixes up its buffer parities
//   ring warp only (all lanes, warp-uniform bookkeeping; the TMA issue is elected inside):
//     wait_consumed_group(idx, 1)  the MMAs reading slab idx have completed (each slab has ONE consuming MMA warp, and the
//                                  warps drift apart, so every slab is waited for individually, in order)
//     ring_freed(total)   the next slab (in order) is consumed: advance and issue every slab whose buffer is now free
//   MMA warps: mma_a(tile) / mma_b(rows, lbo) descriptors; mma<M>(col, adesc, bdesc, ks, acc, ncols) issues one instruction
#pragma once
#include "svad_tile.h"

namespace svad {

constexpr int kRingWarp = 4;   // owns the weight ring (TMA issue, stage-free bookkeeping); warps 0-3 are the MMA warps

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

#if defined(SVAD_EXACT_GATES)
#define SVAD_SIG sigmoid_acc
#define SVAD_TANH tanhf
#else
#define SVAD_SIG sigmoid_fast
#define SVAD_TANH tanh_fast
#endif
#if defined(__CUDA_ARCH__)
#define SVAD_STAMP(k) do { if (a.dbg && first_tile == 0 && t == a.T / 2 && tc.tid == 0) a.dbg[k] = clock64(); } while (0)
#define SVAD_CLK(v) const long long v = clock64()
#define SVAD_ACC(k, d) do { if (a.dbg && first_tile == 0 && t == a.T / 2 && tc.tid == 0) a.dbg[k] += (d); } while (0)
#else
#define SVAD_STAMP(k) do { } while (0)
#define SVAD_CLK(v) do { } while (0)
#define SVAD_ACC(k, d) do { } while (0)
#endif

template <bool SR16, int RM, typename S, class Env>
SVAD_HD void run_cta_tc(Env& env, const TileArgs& a, int first_tile, int tile_stride, int ntiles) {
    using G = Geo<SR16>;
    using TP = TapeTC<SR16>;
    using M = SmemMapTC;
    constexpr int Kt = TP::Kt, KC0 = Kt / 32;
    const Tc tc(env.tid());
    float* sm = env.smem();
    const S* audio = static_cast<const S*>(a.audio);
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
    const int total_slabs = (int)((long)my_tiles * a.T * TP::nslab);   // < 2^31 (checked on the host)
    int it = 0;
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
            SVAD_STAMP(0);
            // ---------------- STFT (CUDA cores) -> mag rows in tcgen05 atom layout
            const bool fast = (t > 0) && ((t + 1) * G::n <= a.L) && a.dec == 1;
            if (t + 1 < a.T && a.dec == 1) {
                constexpr int kPerLine = 128 / (int)sizeof(S), kLines = G::n / kPerLine;
                for (int i = tc.tid; i < BT * kLines; i += kThreads) {
                    const int loc = i / kLines, line = i % kLines, g = g0 + loc;
                    const long off = (t + 1) * G::n + line * kPerLine;
                    if (g < a.B && off < a.L) env.prefetch_l2(audio + (long)g * a.ld + off);
                }
            }
            if (t == 0) stft_load<SR16, S>(tc.tid, 0, aud[0], cxp[0], a.L, t, fast, xa, xb, a.dec);
#pragma unroll 1
            for (int rnd = 0; rnd < 4; rnd++) {
                const int hs = rnd >> 1, fp = rnd & 1;
                float na[G::NQ], nb[G::NQ];
                if (rnd < 3) stft_load<SR16, S>(tc.tid, (rnd + 1) & 1, rnd >= 1 ? aud[1] : aud[0], rnd >= 1 ? cxp[1] : cxp[0], a.L, t, fast, na, nb, a.dec);
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
            SVAD_STAMP(1);
            // ---------------- enc0 on the tensor core
            // lo rows of mag[f][0..Kt) -> lo0[f][0..Kt)  (the Z planes there are dead now)
#pragma unroll
            for (int f = 0; f < 4; f++) stage_lo(tc.tid, sm + M::mag + f * M::mag_pitch * kSlots, sm + M::lo0 + f * Kt * kSlots, Kt);
            if (tc.tid < 4 * kSlots)   // row Kt of every frame (identity swizzle): the Nyquist bin, used by the enc0 epilogue
                sm[M::consts + M::c_nyq + tc.tid] = sm[M::mag + (M::mag_pitch * (tc.tid >> 5) + Kt) * kSlots + (tc.tid & 31)];
            env.fence_async();
            env.tc_fence_before();     // the previous step's tcgen05.ld of these TMEM columns are done
            env.sync();
            SVAD_STAMP(2);
            // enc0: one instruction covers all frames a tap contributes to (N = 96 or 128: the frames are consecutive 32-column
            // atoms of the B operand, LBO = frame pitch).  The slabs alternate {w_hi, w_lo} tile by tile; four MMA warps take
            // them round robin -- warp 2 (pair & 1) + lo -- each into its own 128-column accumulator (all 512 TMEM columns; the
            // epilogue adds the four), so the tensor pipe always has the next warp's instructions queued while one warp
            // commits and waits for its next slab.  The tape starts with the tap-1 pairs: the first instruction into every
            // accumulator overwrites all four frame blocks.
            if (tc.warp < 4) {
                env.tc_fence_after();
#pragma unroll 1
                for (int s = 0; s < TP::e0_nslab; s++) {
                    const int pr = s >> 1, lo = s & 1;
                    if (tc.warp != (((pr & 1) << 1) | lo)) { env.slab_pass(s); continue; }   // not this warp's slab: only keep the buffer parity in step
                    const int kc = TP::e0_pair_kc(pr), j = TP::e0_pair_tap(pr);
                    SVAD_CLK(c0);
                    const float* tile = env.slab_wait(s);
                    SVAD_CLK(c1); SVAD_ACC(11, c1 - c0);
                    {
                        const int f0 = (j == 2) ? 1 : 0, t0 = f0 + 1 - j, nf = (j == 1) ? 4 : 3;
                        const int dcol = lo * 256 + (pr & 1) * 128 + t0 * 32;
                        const auto ad = env.mma_a(tile);
                        const auto bh = env.mma_b(sm + M::mag + (f0 * M::mag_pitch + kc * 32) * kSlots, M::mag_pitch * kSlots * 4);
                        const auto bl = env.mma_b(sm + M::lo0 + (f0 * Kt + kc * 32) * kSlots, Kt * kSlots * 4);
                        const bool first = pr < 2;
                        if (!lo) env.template mma_ks4<128, 2>(dcol, ad, bh, ad, bl, ad, bl, !first, 32 * nf);   // w_hi * x_hi, w_hi * x_lo
                        else env.template mma_ks4<128, 1>(dcol, ad, bh, ad, bh, ad, bh, !first, 32 * nf);        // w_lo * x_hi
                        env.mma_slab_done(s);
                    }
                }
                env.acc_commit();
                SVAD_STAMP(3);
            } else {
                env.skip_phase(TP::phase_mask(0, TP::e0_nslab), TP::e0_nslab);
                if (tc.warp == kRingWarp) {   // ring manager: as each slab is released, issue whatever its buffer unblocks
#pragma unroll 1
                    for (int s = 0; s < TP::e0_nslab; s++) { env.wait_consumed_group(s, 1); env.ring_freed(total_slabs); }
                }
            }
            it += TP::e0_nslab;
            env.acc_wait();
            SVAD_STAMP(4);
            // epilogue: this thread owns channel `row`, frames 2*half, 2*half+1: + bias + Nyquist-bin rank-1 term, ReLU -> e0
            {
                const float b0 = sm[M::consts + M::c_b0 + row];
                const float wn0 = sm[M::consts + M::c_wnyq + row], wn1 = sm[M::consts + M::c_wnyq + 128 + row],
                            wn2 = sm[M::consts + M::c_wnyq + 256 + row];
#pragma unroll
                for (int ff = 0; ff < 2; ff++) {
                    const int tt = 2 * half + ff;
                    float v[32];
                    float v2[32];
                    env.tmem_ld16(lq, tt * 32, *reinterpret_cast<float(*)[16]>(v));
                    env.tmem_ld16(lq, tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
                    env.tmem_ld16(lq, 256 + tt * 32, *reinterpret_cast<float(*)[16]>(v2));
                    env.tmem_ld16(lq, 256 + tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v2 + 16));
#pragma unroll
                    for (int s = 0; s < 32; s++) v[s] += v2[s];
                    env.tmem_ld16(lq, 128 + tt * 32, *reinterpret_cast<float(*)[16]>(v2));
                    env.tmem_ld16(lq, 128 + tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v2 + 16));
#pragma unroll
                    for (int s = 0; s < 32; s++) v[s] += v2[s];
                    env.tmem_ld16(lq, 384 + tt * 32, *reinterpret_cast<float(*)[16]>(v2));
                    env.tmem_ld16(lq, 384 + tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v2 + 16));
                    const float* nyq = sm + M::consts + M::c_nyq + tt * kSlots;
                    float lo[32];
#pragma unroll
                    for (int s = 0; s < 32; s++) {
                        float acc = (v[s] + v2[s]) + b0;
                        if (tt > 0) acc = fmaf(wn0, nyq[s - kSlots], acc);
                        acc = fmaf(wn1, nyq[s], acc);
                        if (tt < 3) acc = fmaf(wn2, nyq[s + kSlots], acc);
                        v[s] = relu(acc);
                        lo[s] = lo_part(v[s]);
                    }
                    // e0 (hi = as is) and its lo parts, both as tcgen05 B rows: enc1 runs on the tensor core too
                    float* dst = sm + M::e0 + (tt * 128 + row) * kSlots;
                    float* dlo = sm + M::e0lo + (tt * 128 + row) * kSlots;
#pragma unroll
                    for (int gq = 0; gq < 8; gq++) {
                        const int pq = tc_f4(gq, row) << 2;
                        *reinterpret_cast<f4*>(dst + pq) = f4{v[4 * gq], v[4 * gq + 1], v[4 * gq + 2], v[4 * gq + 3]};
                        *reinterpret_cast<f4*>(dlo + pq) = f4{lo[4 * gq], lo[4 * gq + 1], lo[4 * gq + 2], lo[4 * gq + 3]};
                    }
                }
            }
            env.fence_async();
            env.tc_fence_before();
            env.sync();
            SVAD_STAMP(5);
            // ---------------- enc1 on the tensor core: M = 64 output channels, N = 2 output frames x 32 slots, K = 3 taps x 128.
            // out[tt] = sum_j W_j e0[2tt - 1 + j]: tap 1 reads frames (0, 2), tap 2 frames (1, 3) (N = 64, LBO = 2 frames), tap 0
            // frame 1 for tt = 1 only (N = 32; frame -1 is the zero padding).  MMA warp w owns k-chunk w of every tap and its own
            // accumulator (64 TMEM columns; enc0's are free again), so the four issue streams never serialise on one accumulator.
            if (tc.warp < 4) {
                env.tc_fence_after();
                const int abase = (tc.warp & 1) * 64 + (tc.warp >> 1) * 256;
#pragma unroll 1
                for (int s = 0; s < TP::e1_nslab; s++) {
                    const int jo = s >> 2, kc = s & 3;
                    if (kc != tc.warp) { env.slab_pass(TP::e0_nslab + s); continue; }
                    const float* tile = env.slab_wait(TP::e0_nslab + s);
                    const int f0 = (jo == 0) ? 0 : 1, ncols = (jo == 2) ? 32 : 64, col = abase + ((jo == 2) ? 32 : 0);
                    const auto ah = env.mma_a(tile), al = env.mma_a(tile + 64 * 32);
                    const auto bh = env.mma_b(sm + M::e0 + (f0 * 128 + kc * 32) * kSlots, 2 * 128 * kSlots * 4);
                    const auto bl = env.mma_b(sm + M::e0lo + (f0 * 128 + kc * 32) * kSlots, 2 * 128 * kSlots * 4);
                    env.template mma_ks4<64, 3>(col, ah, bh, ah, bl, al, bh, jo != 0, ncols);
                    env.mma_slab_done(TP::e0_nslab + s);
                }
                env.acc_commit();
            } else {
                env.skip_phase(TP::phase_mask(TP::e0_nslab, TP::e1_nslab), TP::e1_nslab);
                if (tc.warp == kRingWarp) {
#pragma unroll 1
                    for (int s = 0; s < TP::e1_nslab; s++) { env.wait_consumed_group(TP::e0_nslab + s, 1); env.ring_freed(total_slabs); }
                }
            }
            it += TP::e1_nslab;
            env.acc_wait();
            SVAD_STAMP(20);
            // epilogue: an M = 64 accumulator keeps row r in TMEM lane 32 * (r / 16) + r % 16, so lanes 0-15 of every warp hold
            // channel 16 * lq + lane; warps w and w + 4 take output frame 0 / 1.  Sum the four accumulators, + bias, ReLU -> e1
            {
                float v[32], p[32];
                const int o = 16 * lq + (tc.lane & 15), tt = half;
                env.tmem_ld16(lq, tt * 32, *reinterpret_cast<float(*)[16]>(v));
                env.tmem_ld16(lq, tt * 32 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
#pragma unroll
                for (int w = 1; w < 4; w++) {
                    const int ab = (w & 1) * 64 + (w >> 1) * 256;
                    env.tmem_ld16(lq, ab + tt * 32, *reinterpret_cast<float(*)[16]>(p));
                    env.tmem_ld16(lq, ab + tt * 32 + 16, *reinterpret_cast<float(*)[16]>(p + 16));
#pragma unroll
                    for (int s = 0; s < 32; s++) v[s] += p[s];
                }
                if (tc.lane < 16) {
                    const float b1 = sm[M::consts + M::c_b1 + o];
                    float* dst = sm + M::e1 + (tt * 64 + o) * kSlots;
                    float* dlo = sm + M::e1lo + (tt * 64 + o) * kSlots;
#pragma unroll
                    for (int gq = 0; gq < 8; gq++) {
                        const f4 x = f4{relu(v[4 * gq] + b1), relu(v[4 * gq + 1] + b1), relu(v[4 * gq + 2] + b1), relu(v[4 * gq + 3] + b1)};
                        const int pq = tc_f4(gq, o) << 2;
                        *reinterpret_cast<f4*>(dst + pq) = x;
                        *reinterpret_cast<f4*>(dlo + pq) = f4{lo_part(x.x), lo_part(x.y), lo_part(x.z), lo_part(x.w)};
                    }
                }
            }
            env.fence_async();
            env.tc_fence_before();
            env.sync();
            SVAD_STAMP(21);
            // ---------------- enc2: M = 64, N = 32, K = taps (1, 2) x 64 channels; tap jj reads e1 frame jj.  MMA warp q owns
            // k-chunk q = (jj, channel half) and accumulator columns 32 q.
            if (tc.warp < 4) {
                env.tc_fence_after();
                const int q = tc.warp, idx = TP::e0_nslab + TP::e1_nslab + q;
                for (int s = 0; s < q; s++) env.slab_pass(idx - q + s);
                const float* tile = env.slab_wait(idx);
                SVAD_STAMP(23);
                const auto ah = env.mma_a(tile), al = env.mma_a(tile + 64 * 32);
                const auto bh = env.mma_b(sm + M::e1 + q * 32 * kSlots, 4096), bl = env.mma_b(sm + M::e1lo + q * 32 * kSlots, 4096);
                env.template mma_ks4<64, 3>(32 * q, ah, bh, ah, bl, al, bh, false, 32);
                env.mma_slab_done(idx);
                for (int s = q + 1; s < 4; s++) env.slab_pass(idx - q + s);
                env.acc_commit();
                SVAD_STAMP(24);
            } else {
                env.skip_phase(TP::phase_mask(TP::e0_nslab + TP::e1_nslab, TP::e2_nslab), TP::e2_nslab);
                if (tc.warp == kRingWarp) {
#pragma unroll 1
                    for (int s = 0; s < TP::e2_nslab; s++) { env.wait_consumed_group(TP::e0_nslab + TP::e1_nslab + s, 1); env.ring_freed(total_slabs); }
                }
            }
            it += TP::e2_nslab;
            env.acc_wait();
            SVAD_STAMP(25);
            {   // epilogue: channel 16 lq + lane (lanes 0-15), slots 16 half .. +16
                float v[16], p[16];
                const int o = 16 * lq + (tc.lane & 15);
                env.tmem_ld16(lq, 16 * half, v);
#pragma unroll
                for (int w = 1; w < 4; w++) {
                    env.tmem_ld16(lq, 32 * w + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] += p[i];
                }
                if (tc.lane < 16) {
                    const float b2 = sm[M::consts + M::c_b2 + o];
#pragma unroll
                    for (int gq = 0; gq < 4; gq++) {
                        const f4 x = f4{relu(v[4 * gq] + b2), relu(v[4 * gq + 1] + b2), relu(v[4 * gq + 2] + b2), relu(v[4 * gq + 3] + b2)};
                        const int pq = tc_f4(4 * half + gq, o) << 2;
                        *reinterpret_cast<f4*>(sm + M::e2 + o * kSlots + pq) = x;
                        *reinterpret_cast<f4*>(sm + M::e2lo + o * kSlots + pq) = f4{lo_part(x.x), lo_part(x.y), lo_part(x.z), lo_part(x.w)};
                    }
                }
            }
            env.fence_async();
            env.tc_fence_before();
            env.sync();
            SVAD_STAMP(22);
            // ---------------- enc3: M = 128, N = 32, K = 64 (tap 1).  MMA warp 2 kc + lo owns slab (kc, hi | lo) and columns 32 w
            if (tc.warp < 4) {
                env.tc_fence_after();
                const int w = tc.warp, kc = w >> 1, idx = TP::E3 + w;
                for (int s = 0; s < w; s++) env.slab_pass(idx - w + s);
                const float* tile = env.slab_wait(idx);
                SVAD_STAMP(26);
                const auto ad = env.mma_a(tile);
                const auto bh = env.mma_b(sm + M::e2 + kc * 32 * kSlots, 4096), bl = env.mma_b(sm + M::e2lo + kc * 32 * kSlots, 4096);
                if (!(w & 1)) env.template mma_ks4<128, 2>(32 * w, ad, bh, ad, bl, ad, bl, false, 32);   // w_hi * x_hi, w_hi * x_lo
                else env.template mma_ks4<128, 1>(32 * w, ad, bh, ad, bh, ad, bh, false, 32);            // w_lo * x_hi
                env.mma_slab_done(idx);
                for (int s = w + 1; s < 4; s++) env.slab_pass(idx - w + s);
                env.acc_commit();
                SVAD_STAMP(27);
            } else {
                env.skip_phase(TP::phase_mask(TP::E3, TP::e3_nslab), TP::e3_nslab);
                if (tc.warp == kRingWarp) {
#pragma unroll 1
                    for (int s = 0; s < TP::e3_nslab; s++) { env.wait_consumed_group(TP::E3 + s, 1); env.ring_freed(total_slabs); }
                }
            }
            it += TP::e3_nslab;
            stage_lo(tc.tid, sm + M::h, sm + M::lol_h, kHid);   // h is from the previous step; e1lo (underneath) is dead
            SVAD_STAMP(28);
            env.acc_wait();
            SVAD_STAMP(29);
            {   // epilogue: channel `row`, slots 16 half .. +16 -> e3 and its lo rows, the first half of the LSTM's B operand
                float v[16], p[16];
                env.tmem_ld16(lq, 16 * half, v);
#pragma unroll
                for (int w = 1; w < 4; w++) {
                    env.tmem_ld16(lq, 32 * w + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] += p[i];
                }
                const float b3 = sm[M::consts + M::c_b3 + row];
#pragma unroll
                for (int gq = 0; gq < 4; gq++) {
                    const f4 x = f4{relu(v[4 * gq] + b3), relu(v[4 * gq + 1] + b3), relu(v[4 * gq + 2] + b3), relu(v[4 * gq + 3] + b3)};
                    const int pq = tc_f4(4 * half + gq, row) << 2;
                    *reinterpret_cast<f4*>(sm + M::e3 + row * kSlots + pq) = x;
                    *reinterpret_cast<f4*>(sm + M::lol_x + row * kSlots + pq) = f4{lo_part(x.x), lo_part(x.y), lo_part(x.z), lo_part(x.w)};
                }
            }
            SVAD_STAMP(6);
            // ---------------- LSTM on the tensor core: gates[m*128 + j][slot] = sum_k W[.][k] * [e3 ; h][k][slot]
            env.fence_async();
            env.tc_fence_before();
            env.sync();
            SVAD_STAMP(7);
            if (tc.warp < 4) {   // MMA warp m issues gate block m
                env.tc_fence_after();
#pragma unroll 1
                for (int s = 0; s < TP::l_nslab; s++) {
                    const int kc = s >> 2, m = s & 3;
                    if (m != tc.warp) { env.slab_pass(TP::NA + s); continue; }
                    SVAD_CLK(c0);
                    const float* tile = env.slab_wait(TP::NA + s);
                    SVAD_CLK(c1); SVAD_ACC(13, c1 - c0);
                    // one instruction covers w_hi * [x_hi | x_lo]: the lo rows sit one N atom (LBO) above the hi rows and land in
                    // the block's second 32 columns; w_lo * x_hi (N = 32) accumulates into the first 32
                    const auto ah = env.mma_a(tile), al = env.mma_a(tile + TP::tile);
                    const float* xh = (kc < 4) ? sm + M::e3 + kc * 32 * kSlots : sm + M::h + (kc - 4) * 32 * kSlots;
                    const auto bhl = env.mma_b(xh, ((kc < 4) ? (M::lol_x - M::e3) : (M::lol_h - M::h)) * 4);
                    env.template mma_ks4<128, 2>(128 + m * 64, ah, bhl, al, bhl, al, bhl, kc != 0, 64, 32);
                    env.mma_slab_done(TP::NA + s);
                }
                env.acc_commit();
                SVAD_STAMP(8);
            } else {
                env.skip_phase(TP::phase_mask(TP::NA, TP::l_nslab), TP::l_nslab);
                if (tc.warp == kRingWarp) {
#pragma unroll 1
                    for (int s = 0; s < TP::l_nslab; s++) { env.wait_consumed_group(TP::NA + s, 1); env.ring_freed(total_slabs); }
                }
            }
            it += TP::l_nslab;
            env.acc_wait();
            SVAD_STAMP(9);
            // epilogue: hidden unit `row`, slots 16*half..+16
            {
                float gi[16], gf[16], gg[16], go[16];
                {
                    float p[16];
                    env.tmem_ld16(lq, 128 + 0 * 64 + 16 * half, gi); env.tmem_ld16(lq, 128 + 0 * 64 + 32 + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) gi[i] += p[i];
                    env.tmem_ld16(lq, 128 + 1 * 64 + 16 * half, gf); env.tmem_ld16(lq, 128 + 1 * 64 + 32 + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) gf[i] += p[i];
                    env.tmem_ld16(lq, 128 + 2 * 64 + 16 * half, gg); env.tmem_ld16(lq, 128 + 2 * 64 + 32 + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) gg[i] += p[i];
                    env.tmem_ld16(lq, 128 + 3 * 64 + 16 * half, go); env.tmem_ld16(lq, 128 + 3 * 64 + 32 + 16 * half, p);
#pragma unroll
                    for (int i = 0; i < 16; i++) go[i] += p[i];
                }
                const float* bl = sm + M::consts + M::c_bl;
                const float bi = bl[row], bf = bl[128 + row], bg = bl[256 + row], bo = bl[384 + row];
                float hv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float ig = SVAD_SIG(gi[i] + bi), fg = SVAD_SIG(gf[i] + bf), g2 = SVAD_TANH(gg[i] + bg), og = SVAD_SIG(go[i] + bo);
                    const float cn = fmaf(fg, cst[i], ig * g2);
                    cst[i] = cn;
                    hv[i] = og * SVAD_TANH(cn);
                }
                float* hrow = sm + M::h + row * kSlots;
#pragma unroll
                for (int gq = 0; gq < 4; gq++)   // logical float4 groups 4*half + gq
                    *reinterpret_cast<f4*>(hrow + (tc_f4(4 * half + gq, row) << 2)) = f4{hv[4 * gq], hv[4 * gq + 1], hv[4 * gq + 2], hv[4 * gq + 3]};
            }
            env.tc_fence_before();
            env.sync();
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
            if (t + 1 < a.T) stft_load<SR16, S>(tc.tid, 0, aud[0], cxp[0], a.L, t + 1, ((t + 2) * G::n <= a.L) && a.dec == 1, xa, xb, a.dec);
            SVAD_STAMP(10);
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
                    a.ctx_out[(long)g * G::ctx + k] = (a.T > 0) ? window_sample<SR16, S>(audio + (long)g * a.ld, a.L, cx, a.T - 1, G::n + k, a.dec)
                                                                : (cx ? cx[k] : 0.0f);
                }
            }
        }
    }
}

}  // namespace svad
