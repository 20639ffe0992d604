// svad_h16.cuh -- `svad_fused_h16`: the whole per-chunk forward pass on tcgen05 with fp16 split-precision operands, as
// TWO software-pipelined loops per CTA that overlap the CUDA-core and the tensor-core work of consecutive chunk steps.
//
// One persistent CTA per SM owns a tile of up to 32 streams (N of every MMA) and walks it through all T chunks with
// (h, c) and the audio context resident on the SM.  Per chunk step:
//
//   FRONT loop (step t+1)                                      BACK loop (step t)
//   EF  stage window  -> xp (fp16 hi | lo, [sample][slot])
//   MF  STFT = basis . xp            (tcgen05, N = 128)
//   EF  |re, im| -> mag (hi | lo)                               EB  enc1 accumulators -> e1      <- hand-over F -> B
//   MF  enc0 (3 taps x Kt bins)      (N = 96 / 128)             MB  enc2        EB -> e2
//   EF  + bias + Nyquist bin, ReLU -> e0                        MB  enc3        EB -> e3
//   MF  enc1 (stride 2)              (M = 64)   ----------->    MB  LSTM gates = W . [e3 ; h]    (4 x M = 128, N = 64 | 32)
//                                                               EB  gate math, c, h' -> h ; head -> probability
//
// The front half of step t+1 depends only on the audio, so it runs while the back half of step t (which carries the
// recurrence) is still in flight: the tensor pipe works on one loop's MMAs while the CUDA cores run the other loop's
// epilogue.  Roles are fixed per warp (EF = warps 0-3, EB = warps 4-7: one warp per TMEM lane quarter each; MF / MB = one
// thread each issuing tcgen05.mma; RF / RB = one thread each streaming the loop's weight tape L2 -> shared memory with
// cp.async.bulk through its own ring).  All hand-overs are mbarriers that complete exactly once per step.
//
// Shared memory (H16Map): every layer's output overwrites its input -- the MMAs reading the input have completed into TMEM
// before the epilogue writes -- so one 80 KB region serves xp -> mag -> e0 and one 16 KB region e1 -> e2 -> e3.  Activation
// buffers are the B operands as they sit: rows of 32 slots x fp16 = 64 B per channel, MN-major SWIZZLE_64B atoms (8 rows,
// 16-byte chunk ^= (row >> 1) & 3); frames / hi | lo blocks are further N atoms at the descriptor's LBO stride.
// TMEM (512 columns): front [0,256): STFT tiles -> enc0 [0,128), enc1 [192,256); back [256,512): enc2 [256,288),
// enc3 [288,320), then the four LSTM gate blocks of 64 columns (w_hi . [x_hi | x_lo] is ONE N = 64 instruction).
#pragma once
#include <cuda_fp16.h>

#include "svad_h16_pack.h"
#include "svad_tile.h"

namespace svad {
namespace h16 {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarriers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits until the phase with this parity has completed.  The spin is bounded: a protocol error traps (the launch fails loudly)
// instead of hanging the GPU.
// Between polls the thread sleeps ~32 ns: twelve warps share four schedulers, and a waiter that polls flat out takes issue
// slots (and the pipe SYNCS runs on) from the warps doing the epilogue math -- measured: 58 % of all executed instructions were
// spin-loop instructions and every phase ran 3x slower than alone.
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        __nanosleep(64);
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) mbar_wait_slow(bar, parity);
}
// The four single-thread roles (MMA issue, weight streams) poll without sleeping: their waits sit on the round trip of the weight
// rings (slab landed -> MMAs issued -> committed -> next copy issued), where a 64 ns nap per hand-over is a tenth of the trip.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 28); spin++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
#if defined(SVAD_H16_SPIN_NAP)
        __nanosleep(SVAD_H16_SPIN_NAP);
#endif
    }
    __trap();
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The MMA-issue and weight-stream roles run as WHOLE, converged warps whose 32 lanes compute identical values; the one
// instruction that must be issued once (tcgen05.mma / commit / the bulk copy) is predicated on elect.sync.  Run by a single lane of a
// diverged warp instead, the descriptors live in per-thread registers and every tcgen05.mma is wrapped in a lane-broadcast loop
// (ELECT / R2UR.BROADCAST / BRA.U.ANY): ~50 cycles of issue per instruction, as long as the instruction takes to execute.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    __syncwarp();
}
// CTA-pair mode: the two CTAs of a cluster share every weight slab (each fetches half of it and multicasts it into both shared
// memories), so a ring stage may be refilled only when BOTH issue warps have released it: the release is committed to both CTAs.
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    if (elect_one()) {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }   // one epilogue group
__device__ __forceinline__ void ef_sync() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kH16EfWarps) : "memory"); }   // all front-epilogue warps

// ---------------------------------------------------------------- descriptors / MMA issue (one thread)
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t desc_a(uint32_t saddr) { return desc(saddr, 16, 1024, 2); }              // K-major SWIZZLE_128B weight tile
__device__ __forceinline__ uint64_t desc_b(uint32_t saddr, uint32_t lbo) { return desc(saddr, lbo, 512, 4); }   // MN-major SWIZZLE_64B activation rows
// D fp32, A / B fp16, A K-major, B MN-major
__device__ __forceinline__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// The four K = 16 steps of one 64-wide K chunk, NP (A, B) descriptor pairs per step, all into accumulator `d`.  One asm block:
// the descriptors advance in registers (A by 32 B, B by 16 rows = 1024 B per step) -- issuing instruction by instruction from
// C++ costs ~300 cycles each (tools/umma_f16_unit.cu), far above the 41-65 cycles the tensor pipe needs.
// Pair 0 uses instruction descriptor i0, pairs 1 and 2 use i12 (the LSTM's w_lo . x_hi has another N).
#define SVAD_H16_STEP(N) "add.s64 a0, %1, " #N "*2; add.s64 b0, %2, " #N "*64; add.s64 a1, %3, " #N "*2; add.s64 b1, %4, " #N "*64; add.s64 a2, %5, " #N "*2; add.s64 b2, %6, " #N "*64;\n"
#define SVAD_H16_M0(P) "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a0, b0, %7, " P ";\n"
#define SVAD_H16_M1 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %9, t;\n"
#define SVAD_H16_M2 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %9, t;\n"
template <int NP>
__device__ __forceinline__ void mma_chunk(uint32_t d, uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1, uint64_t a2, uint64_t b2, uint32_t i0, uint32_t i12,
                                          bool acc_first) {
    const uint32_t accf = acc_first ? 1u : 0u;
    if constexpr (NP == 3) {
        asm volatile("{\n.reg .pred q, p, t;\n.reg .b64 a0, b0, a1, b1, a2, b2;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %8, 0;\nsetp.eq.u32 t, %8, %8;\n"
                     SVAD_H16_STEP(0) SVAD_H16_M0("p") SVAD_H16_M1 SVAD_H16_M2 SVAD_H16_STEP(1) SVAD_H16_M0("t") SVAD_H16_M1 SVAD_H16_M2
                     SVAD_H16_STEP(2) SVAD_H16_M0("t") SVAD_H16_M1 SVAD_H16_M2 SVAD_H16_STEP(3) SVAD_H16_M0("t") SVAD_H16_M1 SVAD_H16_M2 "}\n"
                     ::"r"(d), "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "r"(i0), "r"(accf), "r"(i12) : "memory");
    } else if constexpr (NP == 2) {
        asm volatile("{\n.reg .pred q, p, t;\n.reg .b64 a0, b0, a1, b1, a2, b2;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %8, 0;\nsetp.eq.u32 t, %8, %8;\n"
                     SVAD_H16_STEP(0) SVAD_H16_M0("p") SVAD_H16_M1 SVAD_H16_STEP(1) SVAD_H16_M0("t") SVAD_H16_M1
                     SVAD_H16_STEP(2) SVAD_H16_M0("t") SVAD_H16_M1 SVAD_H16_STEP(3) SVAD_H16_M0("t") SVAD_H16_M1 "}\n"
                     ::"r"(d), "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "r"(i0), "r"(accf), "r"(i12) : "memory");
    } else {
        asm volatile("{\n.reg .pred q, p, t;\n.reg .b64 a0, b0, a1, b1, a2, b2;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %8, 0;\nsetp.eq.u32 t, %8, %8;\n"
                     SVAD_H16_STEP(0) SVAD_H16_M0("p") SVAD_H16_STEP(1) SVAD_H16_M0("t") SVAD_H16_STEP(2) SVAD_H16_M0("t") SVAD_H16_STEP(3) SVAD_H16_M0("t") "}\n"
                     ::"r"(d), "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "r"(i0), "r"(accf), "r"(i12) : "memory");
    }
}
#undef SVAD_H16_STEP
#undef SVAD_H16_M0
#undef SVAD_H16_M1
#undef SVAD_H16_M2

// ---------------------------------------------------------------- TMEM loads, fp16 split stores
// 16 consecutive accumulator columns of this thread's TMEM lane.  The load is asynchronous: several are issued back to back and
// tmem_wait() is called once before the first use (a TMEM round trip costs hundreds of cycles while the other loop's MMAs run).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                   "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float sqrt_fast(float v) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }   // rel. error 2^-23
// (hi, lo) halves of two scaled values packed as {v0, v1}; saturating conversions
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    uint32_t h;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v1), "f"(v0));
    const __half2 hh = *reinterpret_cast<const __half2*>(&h);
    const float2 hf = __half22float2(hh);
    uint32_t l;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(v1 - hf.y), "f"(v0 - hf.x));
    hi = h; lo = l;
}
// 16 consecutive slots [16*half, +16) of activation row `row` (already scaled) -> the row's hi and lo images (row pitch 64 B)
#if defined(SVAD_H16_NOINLINE_STORE)   // measured: 1.86e8 -> 1.38e8 (the 16 values go through local memory)
__device__ __noinline__
#else
__device__ __forceinline__
#endif
void store_row16(unsigned char* hi_base, unsigned char* lo_base, int row, int half, const float (&v)[16]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const int sw = (row >> 1) & 3;
    const int c0 = ((2 * half) ^ sw) * 16, c1 = ((2 * half + 1) ^ sw) * 16;
    *reinterpret_cast<uint4*>(hi_base + row * 64 + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(hi_base + row * 64 + c1) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(lo_base + row * 64 + c0) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(lo_base + row * 64 + c1) = make_uint4(l[4], l[5], l[6], l[7]);
}
__device__ __forceinline__ float relu_f(float v) { return v > 0.0f ? v : 0.0f; }
// Gate activations from one MUFU.EX2 and one MUFU.RCP each (no Newton step: rcp.approx is good to 1 ulp, the results enter products
// that are split to 22 bits anyway); exact saturation: 2^x = inf -> rcp = 0.
__device__ __forceinline__ float rcp_approx(float y) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_mufu(float v) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v)); }
__device__ __forceinline__ float tanh_mufu(float v) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * v)), 1.0f); }

// Debug dump (svad_engine_set_debug_buffer): CTA 0 writes the fp32 value of every activation of its steps 0 and 1 before the
// fp16 split, [step][region][row][slot]; tools/h16_check.py compares them with the CPU model of tools/h16_numerics.py.
constexpr int kDumpMag = 0, kDumpNyq = kDumpMag + 4 * 128 * 32, kDumpE0 = kDumpNyq + 128, kDumpE1 = kDumpE0 + 4 * 128 * 32, kDumpE2 = kDumpE1 + 2 * 64 * 32,
              kDumpE3 = kDumpE2 + 64 * 32, kDumpGates = kDumpE3 + 128 * 32, kDumpH = kDumpGates + 4 * 128 * 32, kDumpC = kDumpH + 128 * 32,
              kDumpStep = kDumpC + 128 * 32;
__device__ __forceinline__ void dump16(const TileArgs& a, long s, int region, int row, int half, const float (&v)[16], float inv_scale) {
#if defined(SVAD_H16_DEBUG)
    if (a.dbg && blockIdx.x == 0 && s < 2) {
        float* d = reinterpret_cast<float*>(a.dbg) + s * kDumpStep + region + row * 32 + 16 * half;
#pragma unroll
        for (int i = 0; i < 16; i++) d[i] = v[i] * inv_scale;
    }
#endif
}

// barrier indices
enum : int { kXpFull = 0, kStftAcc, kMagFull, kE0Acc, kE0Full, kFDone, kE1Ready, kE2Acc, kE2Full, kE3Acc, kE3Full, kLAcc,
             kFFull0, kFEmpty0 = kFFull0 + kH16StagesF, kBFull0 = kFEmpty0 + kH16StagesF, kBEmpty0 = kBFull0 + kH16StagesB,
             kNumBars = kBEmpty0 + kH16StagesB };

// profiling stamps (debug buffer set): clock64 of CTA 0 at its middle step, [role 0..3 = EF, MF, EB, MB][16], after the dumps
#define SVAD_H16_STAMP(role, k) do { if (c.stamps && s == c.stamp_step) c.stamps[(role) * 16 + (k)] = clock64(); } while (0)

struct Ctx {
    long long* stamps;   // null unless this thread records (one thread per role of CTA 0)
    long stamp_step;
    unsigned char* sm;
    uint32_t sm32;       // shared-space address of sm
    uint32_t bars;       // shared-space address of the barrier array
    uint32_t tmem;
    __device__ __forceinline__ uint32_t bar(int i) const { return bars + 8u * (uint32_t)i; }
    __device__ __forceinline__ float* scratch() const { return reinterpret_cast<float*>(sm + H16Map::C); }
};

// ================================================================ weight streams (one thread each)
template <bool CL>
__device__ __forceinline__ void run_stream(const Ctx& c, const unsigned char* tape, long nsteps, int nslab, int slab_bytes, int nstages, int ring_off, int full0, int empty0) {
    const long total = nsteps * nslab;
    int idx = 0, stage = 0;
    uint32_t round = 0;   // how often the ring has wrapped
    const uint32_t half = CL ? cluster_rank() * (uint32_t)(slab_bytes / 2) : 0u;   // pair mode: this CTA fetches its half of every slab for both
#pragma unroll 1
    for (long i = 0; i < total; i++) {
        if (round > 0) mbar_wait_spin(c.bar(empty0 + stage), (round - 1) & 1u);
        const uint32_t bar = c.bar(full0 + stage), dst = c.sm32 + (uint32_t)(ring_off + stage * slab_bytes);
        if (elect_one()) {
            mbar_expect_tx(bar, (uint32_t)slab_bytes);   // the whole slab: the other half arrives from the peer's copy
            if constexpr (CL) {
                const uint16_t mask = 3;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"(dst + half), "l"(tape + (size_t)idx * slab_bytes + half), "r"(slab_bytes / 2), "r"(bar), "h"(mask) : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(tape + (size_t)idx * slab_bytes), "r"(slab_bytes), "r"(bar)
                             : "memory");
            }
        }
        __syncwarp();
        if (++idx == nslab) idx = 0;
        if (++stage == nstages) { stage = 0; round++; }
    }
}

// ================================================================ MF: front MMA issue (one thread)
template <bool SR16, bool CL>
__device__ __forceinline__ void run_mf(const Ctx& c, long nsteps) {
    using G = H16Geo<SR16>;
    using M = H16Map;
    int stage = 0;        // front ring position and wrap parity
    uint32_t round = 0;
    auto slab = [&]() -> uint32_t {
        mbar_wait_spin(c.bar(kFFull0 + stage), round & 1u);
        return c.sm32 + (uint32_t)(M::FR + stage * kH16SlabF);
    };
    auto slab_done = [&]() {
        if constexpr (CL) tc_commit_pair(c.bar(kFEmpty0 + stage)); else tc_commit(c.bar(kFEmpty0 + stage));
        if (++stage == kH16StagesF) { stage = 0; round++; }
    };
    constexpr uint32_t xp_hi = M::R, xp_lo = M::R + G::XR * 64;
    constexpr uint32_t mag_hi = M::R, mag_lo = M::R + 4 * G::Kt * 64;
    constexpr uint32_t e0_hi = M::R, e0_lo = M::R + 32768;
    for (long s = 0; s < nsteps; s++) {
        const uint32_t par = (uint32_t)(s & 1);
        // ---- STFT: D[row][frame*32 + slot] = sum_m basis[row][m] * xp[hop*frame + m][slot]; the four frames are N atoms `hop` rows apart
        SVAD_H16_STAMP(1, 0);
        mbar_wait(c.bar(kXpFull), par);
        SVAD_H16_STAMP(1, 1);
        if (s > 0) mbar_wait(c.bar(kE1Ready), par ^ 1u);   // enc1's accumulator of the previous step (columns 192..255) has been drained
        SVAD_H16_STAMP(1, 2);
        tc_after();
        if constexpr (SR16) {
#pragma unroll 1
            for (int mt = 0; mt < 2; mt++)
#pragma unroll 1
                for (int kc = 0; kc < 4; kc++) {
                    const uint64_t bh = desc_b(c.sm32 + xp_hi + kc * 4096, G::hop * 64), bl = desc_b(c.sm32 + xp_lo + kc * 4096, G::hop * 64);
                    uint64_t a = desc_a(slab());                                     // w_hi . (x_hi, x_lo)
                    mma_chunk<2>(c.tmem + 128 * mt, a, bh, a, bl, a, bl, idesc(128, 128), idesc(128, 128), kc != 0);
                    slab_done();
                    a = desc_a(slab());                                              // w_lo . x_hi
                    mma_chunk<1>(c.tmem + 128 * mt, a, bh, a, bh, a, bh, idesc(128, 128), idesc(128, 128), true);
                    slab_done();
                }
        } else {
#pragma unroll 1
            for (int kc = 0; kc < 2; kc++) {
                const uint64_t bh = desc_b(c.sm32 + xp_hi + kc * 4096, G::hop * 64), bl = desc_b(c.sm32 + xp_lo + kc * 4096, G::hop * 64);
#pragma unroll 1
                for (int mt = 0; mt < 2; mt++) {
                    const uint32_t w = slab();
                    const uint64_t ah = desc_a(w), al = desc_a(w + 8192);
                    mma_chunk<3>(c.tmem + 128 * mt, ah, bh, ah, bl, al, bh, idesc(64, 128), idesc(64, 128), kc != 0);
                    slab_done();
                }
            }
        }
        tc_commit(c.bar(kStftAcc));
        SVAD_H16_STAMP(1, 3);
        // ---- enc0: out frame t reads in frame t + j - 1; tap order 1, 0, 2; one instruction covers every frame a tap reaches
        mbar_wait(c.bar(kMagFull), par);
        SVAD_H16_STAMP(1, 4);
        tc_after();
#pragma unroll 1
        for (int jo = 0; jo < 3; jo++)
            for (int ch = 0; ch < G::e0_chunks; ch++) {
                const int j = jo == 0 ? 1 : (jo == 1 ? 0 : 2);
                const int f0 = (j == 2) ? 1 : 0, nf = (j == 1) ? 4 : 3, t0 = f0 + 1 - j;
                const uint32_t boff = (uint32_t)((f0 * G::Kt + 64 * ch) * 64);
                const uint64_t bh = desc_b(c.sm32 + mag_hi + boff, G::Kt * 64), bl = desc_b(c.sm32 + mag_lo + boff, G::Kt * 64);
                const uint32_t id = idesc(128, 32 * nf);
                uint64_t a = desc_a(slab());
                mma_chunk<2>(c.tmem + 32 * t0, a, bh, a, bl, a, bl, id, id, !(jo == 0 && ch == 0));
                slab_done();
                a = desc_a(slab());
                mma_chunk<1>(c.tmem + 32 * t0, a, bh, a, bh, a, bh, id, id, true);
                slab_done();
            }
        tc_commit(c.bar(kE0Acc));
        SVAD_H16_STAMP(1, 5);
        // ---- enc1 (M = 64, stride 2): out frame tt reads in frames 2 tt - 1 + j; taps 1 and 2 reach both output frames (N = 64, the
        // two input frames are N atoms two frames apart), tap 0 only tt = 1 (frame -1 is the zero padding)
        mbar_wait(c.bar(kE0Full), par);
        SVAD_H16_STAMP(1, 6);
        tc_after();
#pragma unroll 1
        for (int jo = 0; jo < 3; jo++) {
            const int f0 = (jo == 0) ? 0 : 1, ncols = (jo == 2) ? 32 : 64, col = 192 + ((jo == 2) ? 32 : 0);
            const uint32_t id = idesc(64, ncols);
#pragma unroll 1
            for (int ch = 0; ch < 2; ch++) {
                const uint32_t w = slab();
                const uint64_t ah = desc_a(w), al = desc_a(w + 8192);
                const uint32_t boff = (uint32_t)((f0 * 128 + 64 * ch) * 64);
                const uint64_t bh = desc_b(c.sm32 + e0_hi + boff, 2 * 128 * 64), bl = desc_b(c.sm32 + e0_lo + boff, 2 * 128 * 64);
                mma_chunk<3>(c.tmem + col, ah, bh, ah, bl, al, bh, id, id, !(jo == 0 && ch == 0));
                slab_done();
            }
        }
        tc_commit(c.bar(kFDone));   // enc1's accumulators are complete AND the front region is free for the next window
        SVAD_H16_STAMP(1, 7);
    }
}

// ================================================================ MB: back MMA issue (one thread)
template <bool CL>
__device__ __forceinline__ void run_mb(const Ctx& c, long nsteps) {
    using M = H16Map;
    int stage = 0;
    uint32_t round = 0;
    auto slab = [&]() -> uint32_t {
        mbar_wait_spin(c.bar(kBFull0 + stage), round & 1u);
        return c.sm32 + (uint32_t)(M::BR + stage * kH16SlabB);
    };
    auto slab_done = [&]() {
        if constexpr (CL) tc_commit_pair(c.bar(kBEmpty0 + stage)); else tc_commit(c.bar(kBEmpty0 + stage));
        if (++stage == kH16StagesB) { stage = 0; round++; }
    };
    constexpr uint32_t e1_hi = M::P, e1_lo = M::P + 8192, e2_hi = M::P, e2_lo = M::P + 4096, e3_hi = M::P, h_hi = M::H;
    for (long s = 0; s < nsteps; s++) {
        const uint32_t par = (uint32_t)(s & 1);
        // ---- enc2 (M = 64, N = 32): taps 1, 2 read e1 frames 0, 1; one slab = both taps {hi | lo}
        SVAD_H16_STAMP(3, 0);
        mbar_wait(c.bar(kE1Ready), par);
        SVAD_H16_STAMP(3, 1);
        tc_after();
        {
            const uint32_t w = slab();
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint64_t ah = desc_a(w + q * 16384), al = desc_a(w + q * 16384 + 8192);
                const uint64_t bh = desc_b(c.sm32 + e1_hi + q * 4096, 4096), bl = desc_b(c.sm32 + e1_lo + q * 4096, 4096);
                mma_chunk<3>(c.tmem + 256, ah, bh, ah, bl, al, bh, idesc(64, 32), idesc(64, 32), q != 0);
            }
            slab_done();
        }
        tc_commit(c.bar(kE2Acc));
        SVAD_H16_STAMP(3, 2);
        // ---- enc3 (M = 128, N = 32, K = 64): one slab {w_hi | w_lo}
        mbar_wait(c.bar(kE2Full), par);
        SVAD_H16_STAMP(3, 3);
        tc_after();
        {
            const uint32_t w = slab();
            const uint64_t bh = desc_b(c.sm32 + e2_hi, 4096), bl = desc_b(c.sm32 + e2_lo, 4096);
            const uint64_t ah = desc_a(w), al = desc_a(w + 16384);
            mma_chunk<3>(c.tmem + 288, ah, bh, ah, bl, al, bh, idesc(128, 32), idesc(128, 32), false);
            slab_done();
        }
        tc_commit(c.bar(kE3Acc));
        SVAD_H16_STAMP(3, 4);
        // ---- LSTM: gates[m*128 + j][slot] = sum_k W[.][k] * [e3 ; h][k][slot]; w_hi . [x_hi | x_lo] is one N = 64 instruction (the lo rows
        // sit one N atom = 8 KB above the hi rows and land in the block's second 32 columns), w_lo . x_hi (N = 32) adds into the first 32
        mbar_wait(c.bar(kE3Full), par);
        SVAD_H16_STAMP(3, 5);
        tc_after();
#pragma unroll 1
        for (int kc = 0; kc < 4; kc++) {
#if !defined(SVAD_H16_NO_LSTM_HOLD)
            // The front loop is the critical path and its STFT (6 k cycles of tensor time) would otherwise share the pipe with this LSTM.
            // The x half (W_ih . e3) runs right away, under the front loop's window staging, when the pipe is idle; the h half (W_hh . h)
            // waits until the STFT of the NEXT step has completed and streams under the front loop's |X| epilogue.
            if (kc == 2 && s + 1 < nsteps) { mbar_wait(c.bar(kStftAcc), par ^ 1u); tc_after(); }
#endif
            const uint32_t xrows = (kc < 2 ? e3_hi : h_hi) + (uint32_t)((kc & 1) * 4096);
            const uint64_t bhl = desc_b(c.sm32 + xrows, 8192);
#pragma unroll 1
            for (int m = 0; m < 4; m++) {
                const uint32_t w = slab();
                const uint64_t ah = desc_a(w), al = desc_a(w + 16384);
                mma_chunk<2>(c.tmem + 256 + 64 * m, ah, bhl, al, bhl, al, bhl, idesc(128, 64), idesc(128, 32), kc != 0);
                slab_done();
            }
        }
        tc_commit(c.bar(kLAcc));
        SVAD_H16_STAMP(3, 6);
    }
}

// ================================================================ EF: front epilogue group (warps 0-3)
// Staging of the padded window [context | chunk | reflect pad] of chunk t into xp: a warp takes 32-row blocks, lane = row; 32
// coalesced loads (one per slot) put row (m0 + lane) of all 32 slots into registers -- exactly one activation row -- which is
// split and stored.  The loads of the warp's next block are issued before the current block is converted.
template <bool SR16, typename S>
__device__ __forceinline__ void stage_load(const S* p0, long ld, int nvalid, int blk, int lane, float (&v)[32]) {
    using G = H16Geo<SR16>;
    const int m = 32 * blk + lane;
    const int src = (m >= G::L1) ? 2 * G::L1 - 2 - m : m;            // xp[L1 + j] = x1[L1 - 2 - j]
    const S* p = p0 + src;
#pragma unroll
    for (int s = 0; s < 32; s++) v[s] = (s < nvalid) ? ld_sample(p + (long)s * ld) : 0.0f;
}
template <bool SR16>
__device__ __forceinline__ void stage_store(const Ctx& c, int blk, int lane, const float (&v)[32]) {
    using G = H16Geo<SR16>;
    unsigned char* hi = c.sm + H16Map::R;
    unsigned char* lo = c.sm + H16Map::R + G::XR * 64;
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = v[16 * hf + i] * kSx;
        store_row16(hi, lo, 32 * blk + lane, hf, w);
    }
}
// Vector variant (rows 16-byte aligned: ld % 4 == 0): a warp block is still 32 rows x 32 slots, but lane (rg = lane & 7, cg = lane >> 3)
// loads 4 consecutive samples (rows 4 rg .. 4 rg + 3) of the 8 slots 8 cg .. 8 cg + 7 with eight 16-byte loads -- a 4 x 8 patch whose
// rows are whole 16-byte chunks of activation rows.  A quarter of the load instructions of the scalar variant, which was
// instruction-bound (address arithmetic), not latency-bound.  The four row stores are issued in an order that depends on the
// parity of rg so that one store instruction covers both halves of the 128-byte bank lines (4 wavefronts instead of 8).
// Every audio byte is read exactly once: no L1 allocation (SVAD_H16_LDMODE 1; 0 = plain ld.global.nc -- no measurable difference).
// The window is the one thing this kernel reads through the LSU, and it arrives slowly behind the weight streams (TMA, ~80 % of the
// L2's throughput): 16 kHz runs at 2.23e8 chunks/s with these loads and 2.52e8 with them stubbed out (SVAD_H16_NOLOAD).  Issuing a
// step's loads in one early burst (more registers: SVAD_H16_REGSPLIT, SVAD_H16_EF_WARPS=8) made it worse, not better -- 1.97e8 / 1.66e8:
// the back loop's weight stream slows down when many loads are in flight -- so two blocks per warp are kept in flight, round-robin.
#ifndef SVAD_H16_LDMODE
#define SVAD_H16_LDMODE 1
#endif
__device__ __forceinline__ void ld4(const float* p, float (&x)[4]) {
#if SVAD_H16_LDMODE == 1
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "l"(p));
#else
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
#endif
}
__device__ __forceinline__ void ld4(const int16_t* p, float (&x)[4]) {
#if SVAD_H16_LDMODE == 1
    uint2 v;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#else
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
#endif
    x[0] = (float)(int16_t)(v.x & 0xffffu) * (1.0f / 32768.0f); x[1] = (float)(int16_t)(v.x >> 16) * (1.0f / 32768.0f);
    x[2] = (float)(int16_t)(v.y & 0xffffu) * (1.0f / 32768.0f); x[3] = (float)(int16_t)(v.y >> 16) * (1.0f / 32768.0f);
}
template <typename S>
__device__ __forceinline__ void stagev_load(const S* p0, long ld, int nvalid, int blk, int lane, float (&v)[8][4]) {
    const int rg = ((lane >> 2) & 1) | ((lane >> 3) << 1), cg = lane & 3;   // see stagev_store for the lane map
    const S* p = p0 + 32 * blk + 4 * rg + (long)(8 * cg) * ld;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (8 * cg + j < nvalid) ld4(p, v[j]);
        else { v[j][0] = 0.0f; v[j][1] = 0.0f; v[j][2] = 0.0f; v[j][3] = 0.0f; }
        p += ld;
    }
}
template <bool SR16>
__device__ __forceinline__ void stagev_store(const Ctx& c, int blk, int lane, const float (&v)[8][4]) {
    using G = H16Geo<SR16>;
    // A 16-byte shared store is served a quarter warp (8 consecutive lanes) at a time, so those 8 lanes must hit the 8 different 16-byte
    // columns of a 128-byte bank line: 4 slot groups (cg) x 2 row parities.  With lane = (rg, cg) in the natural order every quarter
    // landed on two columns (16 wavefronts per store instead of 4, ncu round 2).
    const int rg = ((lane >> 2) & 1) | ((lane >> 3) << 1), cg = lane & 3;
    unsigned char* hi = c.sm + H16Map::R;
    const bool odd = rg & 1;
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {   // row pairs (0, 1), (2, 3): odd row groups store the pair in swapped order
        uint32_t h[2][4], l[2][4];
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int w = 0; w < 4; w++) split2(v[2 * w][2 * pr + k] * kSx, v[2 * w + 1][2 * pr + k] * kSx, h[k][w], l[k][w]);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int row = 32 * blk + 4 * rg + 2 * pr + (i ^ (int)odd);
            const int off = row * 64 + ((cg ^ ((row >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(hi + off) = odd ? make_uint4(h[i ^ 1][0], h[i ^ 1][1], h[i ^ 1][2], h[i ^ 1][3]) : make_uint4(h[i][0], h[i][1], h[i][2], h[i][3]);
            *reinterpret_cast<uint4*>(hi + G::XR * 64 + off) = odd ? make_uint4(l[i ^ 1][0], l[i ^ 1][1], l[i ^ 1][2], l[i ^ 1][3]) : make_uint4(l[i][0], l[i][1], l[i][2], l[i][3]);
        }
    }
}
// reflect pad rows xp[L1 + i] = xp[L1 - 2 - i], i < N / 4, copied inside shared memory (hi and lo images, re-swizzled per row)
template <bool SR16>
__device__ __forceinline__ void stage_reflect(const Ctx& c, int tid) {
    using G = H16Geo<SR16>;
#pragma unroll 1
    for (int idx = tid; idx < (G::N / 4) * 8; idx += 32 * kH16EfWarps) {
        const int ch = idx & 3, arr = (idx >> 3) & 1, i = ((idx >> 4) << 1) | ((idx >> 2) & 1);   // a quarter warp = 4 chunks x 2 adjacent rows
        const int src = G::L1 - 2 - i, dst = G::L1 + i;
        unsigned char* base = c.sm + H16Map::R + arr * (G::XR * 64);
        *reinterpret_cast<uint4*>(base + dst * 64 + ((ch ^ ((dst >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(base + src * 64 + ((ch ^ ((src >> 1) & 3)) << 4));
    }
}

// first / last chunk of a row, or decimated input: context, zero tail and sample stride resolved per sample (cold path)
template <bool SR16, typename S>
__device__ __noinline__ void stage_generic(unsigned char* sm, const S* audio, long a_ld, long a_L, const float* a_ctx_in, long a_ctx_ld, int a_B, int a_dec,
                                           int g0, int bt, long t, int warp, int lane) {
    // (everything by value: a reference to the kernel's TileArgs or Ctx here would force those structs into local memory for the
    // whole kernel, and every `a.T`, `c.bar(i)` of the hot loops would become a local load)
    using G = H16Geo<SR16>;
    unsigned char* hi = sm + H16Map::R;
    unsigned char* lo = sm + H16Map::R + G::XR * 64;
#pragma unroll 1
    for (int blk = warp; blk < G::XR / 32; blk += kH16EfWarps) {
        const int m = 32 * blk + lane;
#pragma unroll 1
        for (int hf = 0; hf < 2; hf++) {
            float w[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int sl = 16 * hf + i, g = g0 + sl;
                w[i] = (sl < bt && g < a_B) ? kSx * window_sample<SR16, S>(audio + (long)g * a_ld, a_L, a_ctx_in ? a_ctx_in + (long)g * a_ctx_ld : nullptr, t, m, a_dec) : 0.0f;
            }
            store_row16(hi, lo, m, hf, w);
        }
    }
}

template <bool SR16, typename S>
__device__ __forceinline__ void run_ef(const Ctx& c, const TileArgs& a, int first_tile, int tile_stride, int ntiles, int bt) {
    using G = H16Geo<SR16>;
    using M = H16Map;
    constexpr int EW = kH16EfWarps;
    const int wid = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int warp = wid & 3;                   // TMEM lane quarter (warps 0-3 and, with two groups, 12-15)
    const int half = wid >= 12 ? 1 : 0;         // second group: frames 2, 3
    const int w8 = warp + 4 * half;             // index among the front-epilogue warps (staging)
    const int etid = 32 * w8 + lane;
    const int f_lo = EW == 8 ? 2 * half : 0, f_hi = EW == 8 ? f_lo + 2 : 4;
    const S* audio = static_cast<const S*>(a.audio);
    const float* cs = a.consts;   // global; everything needed is read into registers here
    float* nyq = c.scratch() + M::s_nyq;
    const float d_stft = cs[M::c_scale + 0], d_e0 = cs[M::c_scale + 1];
    const uint32_t tq = c.tmem + ((uint32_t)(warp * 32) << 16);
    // channel / bin owned by this thread
    const int row = SR16 ? 32 * warp + lane : 16 * warp + (lane & 15);   // STFT bin (8 kHz: M = 64 accumulators keep row r in lane 32 (r / 16) + r % 16)
    const bool stft_active = SR16 || lane < 16;
    const int o = 32 * warp + lane;                                      // enc0 output channel
    const float b0 = cs[M::c_b0 + o], wn0 = cs[M::c_wnyq + o], wn1 = cs[M::c_wnyq + 128 + o], wn2 = cs[M::c_wnyq + 256 + o];
    const bool vec_ok = (a.ld % 4 == 0) && (reinterpret_cast<uintptr_t>(audio) % (4 * sizeof(S)) == 0);   // rows start on 4-sample boundaries
    const bool ctx_vec_ok = !a.ctx_in || ((a.ctx_ld % 4 == 0) && (reinterpret_cast<uintptr_t>(a.ctx_in) % 16 == 0));
    long s = 0;
    for (int tile = first_tile; tile < ntiles; tile += tile_stride) {
        const int g0 = tile * bt;
        for (long t = 0; t < a.T; t++, s++) {
            const uint32_t par = (uint32_t)(s & 1);
            // ---- stage the window of chunk t (the front region is free once enc1 of the previous step has completed)
            SVAD_H16_STAMP(0, 0);
            // the first chunk of a row takes its context blocks from the carried-in context (or zeros) instead of the row
            const bool fast = ((t + 1) * G::n <= a.L) && a.dec == 1 && (t > 0 || (vec_ok && ctx_vec_ok));
            constexpr int NBV = G::L1 / 32;        // 18 / 9 blocks of [context | chunk]; the reflect rows are copied afterwards
            constexpr int KBV = (NBV + EW - 1) / EW;   // blocks per warp (warp w takes w, w + EW, ...)
#ifndef SVAD_H16_EARLY
#if SVAD_H16_EF_WARPS == 8 || defined(SVAD_H16_REGSPLIT)
#define SVAD_H16_EARLY KBV
#else
#define SVAD_H16_EARLY 2   // 168 registers per thread: two blocks in flight across the wait, the rest loaded after it
                           // (measured 16 kHz / 8 kHz: 1 block 2.25e8 / 2.87e8, 2 blocks 2.27e8 / 3.02e8, 3 blocks -- spills -- 2.04e8 / 2.66e8)
#endif
#endif
            float v[SVAD_H16_EARLY][8][4];
            const S* p0 = audio + (long)g0 * a.ld + (t * G::n - G::ctx);
            const int nvalid = (a.B - g0 < bt) ? a.B - g0 : bt;
            // every block of this warp is loaded (into registers) while enc1 of the previous step still owns the front region
            auto load_block = [&](int blk, float (&vk)[8][4]) {
#ifdef SVAD_H16_NOLOAD   // experiment: no audio loads at all (what do the window stores cost alone?)
#pragma unroll
                for (int j = 0; j < 8; j++) { vk[j][0] = 0.01f * (float)blk; vk[j][1] = 0.02f; vk[j][2] = -0.01f * (float)lane; vk[j][3] = 0.005f; }
                return;
#endif
                if (t == 0 && blk < G::ctx / 32) {
                    if (a.ctx_in) {
                        stagev_load<float>(a.ctx_in + (long)g0 * a.ctx_ld, a.ctx_ld, nvalid, blk, lane, vk);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) { vk[j][0] = 0.0f; vk[j][1] = 0.0f; vk[j][2] = 0.0f; vk[j][3] = 0.0f; }
                    }
                } else {
                    stagev_load<S>(p0, a.ld, nvalid, blk, lane, vk);
                }
            };
            if (fast && vec_ok) {
#pragma unroll
                for (int k = 0; k < KBV; k++)
                    if (k < SVAD_H16_EARLY && w8 + EW * k < NBV) load_block(w8 + EW * k, v[k % SVAD_H16_EARLY]);
            }
            if (s > 0) mbar_wait(c.bar(kFDone), par ^ 1u);
            SVAD_H16_STAMP(0, 1);
            {
                if (t + 1 < a.T && a.dec == 1) {   // pull the next chunk of every stream of the tile into L2
#ifndef SVAD_H16_PREFETCH   // 0: one prefetch.global.L2 per 128-byte line (default), 1: per 64 bytes, 2: one cp.async.bulk.prefetch.L2 per stream
#define SVAD_H16_PREFETCH 0   // measured 16 kHz / 8 kHz: 2.232e8 / 2.881e8, 2.209e8 / 2.926e8, 2.18e8 / 2.83e8 chunks/s
#endif
                    const long off = (t + 1) * G::n;
#if SVAD_H16_PREFETCH == 2
                    // one bulk prefetch per stream
                    if (vec_ok && (a.ld * sizeof(S)) % 16 == 0) {
                        if (etid < bt && g0 + etid < a.B && off < a.L) {
                            const long cnt = (a.L - off < G::n) ? a.L - off : (long)G::n;
                            const uint32_t bytes = (uint32_t)(cnt * sizeof(S)) & ~15u;
                            if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(audio + (long)(g0 + etid) * a.ld + off), "r"(bytes) : "memory");
                        }
                    } else
#endif
                    {
                        constexpr int kPerLine = (SVAD_H16_PREFETCH == 1 ? 64 : 128) / (int)sizeof(S), kLines = G::n / kPerLine;
#pragma unroll 1
                        for (int i = etid; i < bt * kLines; i += 32 * EW) {
                            const int loc = i / kLines, line = i % kLines, g = g0 + loc;
                            const long o2 = off + line * kPerLine;
                            if (g < a.B && o2 < a.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(audio + (long)g * a.ld + o2));
                        }
                    }
                }
                if (fast && vec_ok) {
                    // blocks beyond the early ones reuse the register buffers round-robin: store block k, then load block k + EARLY into it
#pragma unroll
                    for (int k = 0; k < KBV; k++) {
                        if (w8 + EW * k < NBV) stagev_store<SR16>(c, w8 + EW * k, lane, v[k % SVAD_H16_EARLY]);
                        if (k + SVAD_H16_EARLY < KBV && w8 + EW * (k + SVAD_H16_EARLY) < NBV) load_block(w8 + EW * (k + SVAD_H16_EARLY), v[k % SVAD_H16_EARLY]);
                    }
                    SVAD_H16_STAMP(0, 7);
                    ef_sync();
                    SVAD_H16_STAMP(0, 8);
                    stage_reflect<SR16>(c, etid);
                    SVAD_H16_STAMP(0, 9);
                } else if (fast) {
                    constexpr int NB = G::XR / 32;   // 20 / 10 blocks, warp w takes w, w + 4, ...
                    const S* p0 = audio + (long)g0 * a.ld + (t * G::n - G::ctx);
                    const int nvalid = (a.B - g0 < bt) ? a.B - g0 : bt;
                    float va[32], vb[32];
                    stage_load<SR16, S>(p0, a.ld, nvalid, w8, lane, va);
#pragma unroll 1
                    for (int blk = w8; blk < NB; blk += 2 * EW) {
                        if (blk + EW < NB) stage_load<SR16, S>(p0, a.ld, nvalid, blk + EW, lane, vb);
                        stage_store<SR16>(c, blk, lane, va);
                        if (blk + EW < NB) {
                            if (blk + 2 * EW < NB) stage_load<SR16, S>(p0, a.ld, nvalid, blk + 2 * EW, lane, va);
                            stage_store<SR16>(c, blk + EW, lane, vb);
                        }
                    }
                } else {
                    stage_generic<SR16, S>(c.sm, audio, a.ld, a.L, a.ctx_in, a.ctx_ld, a.B, a.dec, g0, bt, t, w8, lane);
                }
            }
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kXpFull));
            SVAD_H16_STAMP(0, 2);
            // ---- STFT epilogue: |re + i im| -> mag rows (hi | lo); bin 0 and the Nyquist bin are real
            mbar_wait(c.bar(kStftAcc), par);
            SVAD_H16_STAMP(0, 3);
            tc_after();
            {
                unsigned char* hi = c.sm + M::R;
                unsigned char* lo = c.sm + M::R + 4 * G::Kt * 64;
#pragma unroll 1
                for (int f = f_lo; f < f_hi; f++) {
                    float re[2][16], im[2][16];
                    tmem_ld16(tq + (uint32_t)(f * 32), re[0]);
                    tmem_ld16(tq + (uint32_t)(f * 32 + 16), re[1]);
                    tmem_ld16(tq + (uint32_t)(128 + f * 32), im[0]);
                    tmem_ld16(tq + (uint32_t)(128 + f * 32 + 16), im[1]);
                    tmem_wait();
                    if (stft_active) {
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) {
                            float mg[16];
                            if (row == 0) {
#pragma unroll
                                for (int i = 0; i < 16; i++) { nyq[f * 32 + 16 * hf + i] = fabsf(im[hf][i]) * d_stft; mg[i] = fabsf(re[hf][i]) * (d_stft * kSmag); }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; i++) mg[i] = sqrt_fast(re[hf][i] * re[hf][i] + im[hf][i] * im[hf][i]) * (d_stft * kSmag);
                            }
                            store_row16(hi, lo, f * G::Kt + row, hf, mg);
                            dump16(a, s, kDumpMag, f * G::Kt + row, hf, mg, 1.0f / kSmag);
                        }
                    }
                }
            }
            tc_before();
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kMagFull));
            SVAD_H16_STAMP(0, 4);
            // ---- enc0 epilogue: + bias + Nyquist-bin rank-1 term (fp32), ReLU -> e0 rows (hi | lo)
            mbar_wait(c.bar(kE0Acc), par);
            SVAD_H16_STAMP(0, 5);
            tc_after();
            {
                unsigned char* hi = c.sm + M::R;
                unsigned char* lo = c.sm + M::R + 32768;
#pragma unroll 1
                for (int tt = f_lo; tt < f_hi; tt++) {
                    float vv[2][16];
                    tmem_ld16(tq + (uint32_t)(tt * 32), vv[0]);
                    tmem_ld16(tq + (uint32_t)(tt * 32 + 16), vv[1]);
                    tmem_wait();
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        float (&v)[16] = vv[hf];
                        const float* ny = nyq + tt * 32 + 16 * hf;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float acc = fmaf(v[i], d_e0, b0);
                            if (tt > 0) acc = fmaf(wn0, ny[i - 32], acc);
                            acc = fmaf(wn1, ny[i], acc);
                            if (tt < 3) acc = fmaf(wn2, ny[i + 32], acc);
                            v[i] = relu_f(acc) * kSe0;
                        }
                        store_row16(hi, lo, tt * 128 + o, hf, v);
                        dump16(a, s, kDumpE0, tt * 128 + o, hf, v, 1.0f / kSe0);
                    }
                }
            }
            tc_before();
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kE0Full));
            SVAD_H16_STAMP(0, 6);
        }
    }
}

// ================================================================ EB: back epilogue group (warps 4-7)
template <bool SR16, typename S>
__device__ __forceinline__ void run_eb(const Ctx& c, const TileArgs& a, int first_tile, int tile_stride, int ntiles, int bt) {
    using G = H16Geo<SR16>;
    using M = H16Map;
    const int warp = (int)(threadIdx.x >> 5) - 4, lane = (int)(threadIdx.x & 31);   // warp 0..3 = TMEM lane quarter
    const int tid = warp * 32 + lane;
    const float* cs = a.consts;
    const float* wout = c.scratch() + M::s_wout;
    const float bout = cs[M::c_bout];
    const float d_e1 = cs[M::c_scale + 2], d_e2 = cs[M::c_scale + 3], d_e3 = cs[M::c_scale + 4], d_l = cs[M::c_scale + 5];
    const uint32_t tq = c.tmem + ((uint32_t)(warp * 32) << 16);
    const int o64 = 16 * warp + (lane & 15);   // channel of the M = 64 layers (lanes 0-15 of each quarter)
    const bool act64 = lane < 16;
    const int j = 32 * warp + lane;            // enc3 channel / LSTM hidden unit
    const float b1 = cs[M::c_b1 + o64], b2 = cs[M::c_b2 + o64], b3 = cs[M::c_b3 + j];
    const float bi = cs[M::c_bl + j], bf = cs[M::c_bl + 128 + j], bg = cs[M::c_bl + 256 + j], bo = cs[M::c_bl + 384 + j];
    unsigned char* p_hi = c.sm + M::P;
    unsigned char* h_hi = c.sm + M::H;
    float creg[32], hreg[32];   // cell and hidden state of unit j for the 32 slots (fp32, never leave the registers between steps)
    const S* audio = static_cast<const S*>(a.audio);
    long s = 0;
    for (int tile = first_tile; tile < ntiles; tile += tile_stride) {
        const int g0 = tile * bt;
        // ---- tile start: carried-in state.  The previous tile's last LSTM MMAs have completed (this thread waited for them).
#pragma unroll
        for (int sl = 0; sl < 32; sl++) {
            const int g = g0 + sl;
            const bool v = a.state_in && sl < bt && g < a.B;
            hreg[sl] = v ? a.state_in[(long)g * kHid + j] : 0.0f;
            creg[sl] = v ? a.state_in[((long)a.B + g) * kHid + j] : 0.0f;
        }
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = hreg[16 * hf + i] * kSh;
            store_row16(h_hi, h_hi + 8192, j, hf, v);
        }
        for (long t = 0; t < a.T; t++, s++) {
            const uint32_t par = (uint32_t)(s & 1);
            // ---- enc1 epilogue: accumulator columns 192..255 (out frame tt at 32 tt) -> e1 [tt][64][32]
            SVAD_H16_STAMP(2, 0);
            mbar_wait(c.bar(kFDone), par);
            SVAD_H16_STAMP(2, 1);
            tc_after();
#pragma unroll 1
            for (int tt = 0; tt < 2; tt++) {
                float vv[2][16];
                tmem_ld16(tq + (uint32_t)(192 + tt * 32), vv[0]);
                tmem_ld16(tq + (uint32_t)(192 + tt * 32 + 16), vv[1]);
                tmem_wait();
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    float (&v)[16] = vv[hf];
                    if (act64) {
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = relu_f(fmaf(v[i], d_e1, b1)) * kSe1;
                        store_row16(p_hi, p_hi + 8192, tt * 64 + o64, hf, v);
                        dump16(a, s, kDumpE1, tt * 64 + o64, hf, v, 1.0f / kSe1);
                    }
                }
            }
            tc_before();
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kE1Ready));
            SVAD_H16_STAMP(2, 2);
            // ---- enc2 epilogue: columns 256..287 -> e2 [64][32]
            mbar_wait(c.bar(kE2Acc), par);
            SVAD_H16_STAMP(2, 3);
            tc_after();
            {
                float vv[2][16];
                tmem_ld16(tq + 256u, vv[0]);
                tmem_ld16(tq + 272u, vv[1]);
                tmem_wait();
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                float (&v)[16] = vv[hf];
                if (act64) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = relu_f(fmaf(v[i], d_e2, b2)) * kSe2;
                    store_row16(p_hi, p_hi + 4096, o64, hf, v);
                    dump16(a, s, kDumpE2, o64, hf, v, 1.0f / kSe2);
                }
              }
            }
            tc_before();
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kE2Full));
            SVAD_H16_STAMP(2, 4);
            // ---- enc3 epilogue: columns 288..319 -> e3 [128][32], the first half of the LSTM's K
            mbar_wait(c.bar(kE3Acc), par);
            SVAD_H16_STAMP(2, 5);
            tc_after();
            {
                float vv[2][16];
                tmem_ld16(tq + 288u, vv[0]);
                tmem_ld16(tq + 304u, vv[1]);
                tmem_wait();
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    float (&v)[16] = vv[hf];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = relu_f(fmaf(v[i], d_e3, b3)) * kSe3;
                    store_row16(p_hi, p_hi + 8192, j, hf, v);
                    dump16(a, s, kDumpE3, j, hf, v, 1.0f / kSe3);
                }
            }
            tc_before();
            fence_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.bar(kE3Full));
            SVAD_H16_STAMP(2, 6);
            // ---- LSTM epilogue: gate math for hidden unit j, all 32 slots; c and h' stay in registers, h' also goes to the B rows
            mbar_wait(c.bar(kLAcc), par);
            SVAD_H16_STAMP(2, 7);
            tc_after();
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {   // unrolled: creg / hreg are indexed statically and stay in registers
                float gi[16], gf[16], gg[16], go[16], q[16];
                {
                    float q0[16];
                    tmem_ld16(tq + (uint32_t)(256 + 0 * 64 + 16 * hf), gi); tmem_ld16(tq + (uint32_t)(256 + 0 * 64 + 32 + 16 * hf), q0);
                    tmem_ld16(tq + (uint32_t)(256 + 1 * 64 + 16 * hf), gf); tmem_ld16(tq + (uint32_t)(256 + 1 * 64 + 32 + 16 * hf), q);
                    tmem_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) { gi[i] += q0[i]; gf[i] += q[i]; }
                    tmem_ld16(tq + (uint32_t)(256 + 2 * 64 + 16 * hf), gg); tmem_ld16(tq + (uint32_t)(256 + 2 * 64 + 32 + 16 * hf), q0);
                    tmem_ld16(tq + (uint32_t)(256 + 3 * 64 + 16 * hf), go); tmem_ld16(tq + (uint32_t)(256 + 3 * 64 + 32 + 16 * hf), q);
                    tmem_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) { gg[i] += q0[i]; go[i] += q[i]; }
                }
                dump16(a, s, kDumpGates + 0 * 4096, j, hf, gi, d_l); dump16(a, s, kDumpGates + 1 * 4096, j, hf, gf, d_l);
                dump16(a, s, kDumpGates + 2 * 4096, j, hf, gg, d_l); dump16(a, s, kDumpGates + 3 * 4096, j, hf, go, d_l);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float ig = sigmoid_mufu(fmaf(gi[i], d_l, bi)), fg = sigmoid_mufu(fmaf(gf[i], d_l, bf));
                    const float g2 = tanh_mufu(fmaf(gg[i], d_l, bg)), og = sigmoid_mufu(fmaf(go[i], d_l, bo));
                    const float cn = fmaf(fg, creg[16 * hf + i], ig * g2);
                    creg[16 * hf + i] = cn;
                    const float hn = og * tanh_mufu(cn);
                    hreg[16 * hf + i] = hn;
                    q[i] = hn * kSh;
                }
                store_row16(h_hi, h_hi + 8192, j, hf, q);
                dump16(a, s, kDumpH, j, hf, q, 1.0f / kSh);
            }
            tc_before();
            fence_async();
            group_sync(1);   // every unit's h' row is in shared memory
            SVAD_H16_STAMP(2, 8);
            // ---- head: p = sigmoid(sum_j wout[j] relu(h'[j]) + bout); warp w takes slots 8w..8w+7, lane = (slot, quarter of the units)
            {
                const int sl = 8 * warp + (lane & 7), part = lane >> 3;
                const __half* hh = reinterpret_cast<const __half*>(h_hi);
                const __half* hl = reinterpret_cast<const __half*>(h_hi + 8192);
                float acc = 0.0f;
#pragma unroll 8
                for (int u = 0; u < 32; u++) {
                    const int r = 4 * u + part;   // the four parts read adjacent rows: different bank groups (rows 32 apart share banks)
                    const int pos = r * 32 + ((((sl >> 3) ^ ((r >> 1) & 3)) << 3) | (sl & 7));
                    const float hv = (__half2float(hh[pos]) + __half2float(hl[pos])) * (1.0f / kSh);
                    acc = fmaf(wout[r], relu_f(hv), acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                const int g = g0 + sl;
                if (part == 0 && sl < bt && g < a.B) a.probs[(long)g * a.ldp + t] = sigmoid_acc(acc + bout);
            }
            SVAD_H16_STAMP(2, 9);
        }
        // ---- tile end: carry state / context out
        group_sync(1);   // the head of the last step has read every h row before the next tile's state overwrites them
        if (a.state_out) {
#pragma unroll
            for (int sl = 0; sl < 32; sl++) {
                const int g = g0 + sl;
                if (sl < bt && g < a.B) {
                    a.state_out[(long)g * kHid + j] = hreg[sl];
                    a.state_out[((long)a.B + g) * kHid + j] = creg[sl];
                }
            }
        }
        if (a.ctx_out) {
            for (int i = tid; i < bt * G::ctx; i += 128) {
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

}  // namespace h16

// ================================================================ kernel
template <bool SR16, typename S, bool CL = false>
__global__ void __launch_bounds__(kH16Threads, 1) svad_fused_h16(TileArgs a, const unsigned char* tapeF, const unsigned char* tapeB, int ntiles, int bt) {
    using namespace h16;
    using G = H16Geo<SR16>;
    using M = H16Map;
    extern __shared__ __align__(1024) unsigned char smem[];
    Ctx c;
    c.stamps = nullptr;
    c.stamp_step = -1;
    c.sm = smem;
    c.sm32 = s32(smem);
    c.bars = c.sm32 + (uint32_t)M::BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + M::BAR + 8 * kNumBars);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = (int)(threadIdx.x & 31);   // warp-uniform by construction
    if (threadIdx.x < 128) c.scratch()[M::s_wout + threadIdx.x] = a.consts[M::c_wout + threadIdx.x];
    if (threadIdx.x == 0) {
        for (int b = 0; b < kNumBars; b++) {
            // group barriers: one arrival per warp of the 4-warp epilogue group; commit / TMA barriers: one arrival
            const bool grp_f = (b == kXpFull || b == kMagFull || b == kE0Full), grp_b = (b == kE1Ready || b == kE2Full || b == kE3Full);
            const bool ring_empty = (b >= kFEmpty0 && b < kFEmpty0 + kH16StagesF) || (b >= kBEmpty0 && b < kBEmpty0 + kH16StagesB);
            mbar_init(c.bar(b), grp_f ? (uint32_t)kH16EfWarps : grp_b ? 4u : (CL && ring_empty) ? 2u : 1u);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_before();
    __syncthreads();
    if constexpr (CL) cluster_sync_all();   // the peer's barriers exist before anything is multicast at them
    tc_after();
    c.tmem = *tmem_slot;
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) my_tiles++;
    const long nsteps = (long)my_tiles * a.T;
    if (a.dbg && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128 || (warp >= 8 && warp < 10 && lane == 0))) {
        c.stamps = a.dbg + (2 * h16::kDumpStep * 4 + 7) / 8;   // after the two activation dumps (floats), 8-byte aligned
        c.stamp_step = nsteps / 2;
    }
    const long long cta_t0 = clock64();   // whole-CTA spans (every CTA): implied SM clock = span / kernel time, spread = slowest / fastest
    // Register file split by warpgroup (setmaxnreg works on 4 aligned warps).  Two front groups (512 threads, compiled for 128 registers):
    // the issue / stream warps give back 88 each, the back epilogue (cell + hidden state of 32 slots in registers) takes 72, the front
    // groups 8: 136 + 200 + 40 + 136 = 4 * 128.  One front group (384 threads x 168): optional 232 + 232 + 40 (SVAD_H16_REGSPLIT).
    // (each setmaxnreg sits at the head of its role's branch: ptxas allocates the code it dominates against that limit)
#if SVAD_H16_EF_WARPS == 8
    constexpr int kRegEf = 136, kRegEb = 200, kRegRole = 40;
    constexpr bool kRegSplit = true;
#elif defined(SVAD_H16_REGSPLIT)
    constexpr int kRegEf = 232, kRegEb = 232, kRegRole = 40;
    constexpr bool kRegSplit = true;
#else
    constexpr int kRegEf = 0, kRegEb = 0, kRegRole = 0;
    constexpr bool kRegSplit = false;
#endif
    if (warp < 4 || warp >= 12) {
        if (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegSplit ? kRegEf : 24));
        run_ef<SR16, S>(c, a, (int)blockIdx.x, (int)gridDim.x, ntiles, bt);
    } else if (warp < 8) {
        if (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegSplit ? kRegEb : 24));
        run_eb<SR16, S>(c, a, (int)blockIdx.x, (int)gridDim.x, ntiles, bt);
    } else {
        if (kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegSplit ? kRegRole : 24));
        if (warp == 8) run_mf<SR16, CL>(c, nsteps);
        else if (warp == 9) run_mb<CL>(c, nsteps);
        else if (warp == 10) run_stream<CL>(c, tapeF, nsteps, G::nslabF, kH16SlabF, kH16StagesF, M::FR, kFFull0, kFEmpty0);
        else run_stream<CL>(c, tapeB, nsteps, G::nslabB, kH16SlabB, kH16StagesB, M::BR, kBFull0, kBEmpty0);
    }
    tc_before();
    __syncthreads();
    if constexpr (CL) cluster_sync_all();   // neither CTA leaves while the other may still signal its barriers
    if (a.dbg && threadIdx.x == 0) (a.dbg + (2 * h16::kDumpStep * 4 + 7) / 8)[64 + blockIdx.x] = clock64() - cta_t0;
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(512u) : "memory");
}

}  // namespace svad
