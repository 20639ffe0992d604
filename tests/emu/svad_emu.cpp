// svad_emu.cpp -- CPU emulator of one CTA of the fused fp32 kernel (TEST INFRASTRUCTURE).
// Runs svad::run_cta<> (the exact schedule + per-thread arithmetic the CUDA kernel compiles) on 256
// OS threads with a pthread barrier standing in for __syncthreads and memcpy standing in for the
// TMA bulk copies of weight slabs.  Lets the layout / index algebra be checked against the oracle
// in the build container, which has no GPU.  Not part of the product; never timed.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <memory>

#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../silero_vad_b200/csrc/svad_tc.h"

using namespace svad;

namespace {
struct Shared {
    std::vector<float> smem;
    pthread_barrier_t bar;
    const float* tape;
    std::atomic<long> issued{0};                 // slabs copied into the ring so far (the "full" side)
    std::atomic<long> released[kStages];         // per stage: thread arrivals so far (the "empty" side)
};
template <bool SR16>
struct EmuEnv {
    Shared* sh;
    int tid_;
    int tid() const { return tid_; }
    float* smem() { return sh->smem.data(); }
    void sync() { pthread_barrier_wait(&sh->bar); }
    void prefetch_l2(const void*) {}
    void issue(long it) {
        const int idx = (int)(it % Geo<SR16>::nslab), stage = (int)(it % kStages);
        memcpy(sh->smem.data() + SmemMap::stage + stage * SmemMap::stage_floats, sh->tape + Tape<SR16>::slab_off(idx),
               sizeof(float) * Tape<SR16>::slab_len(idx));
    }
    const float* slab_acquire(long it, long total) {
        if (tid_ == 0 && it >= 1 && it + 1 < total) {
            const long prev = it - 1;   // its stage is the one slab it+1 goes into
            while (sh->released[prev % kStages].load(std::memory_order_acquire) < (long)kThreads * (prev / kStages + 1)) sched_yield();
            issue(it + 1);
            sh->issued.store(it + 2, std::memory_order_release);
        }
        while (sh->issued.load(std::memory_order_acquire) <= it) sched_yield();
        return sh->smem.data() + SmemMap::stage + (it % kStages) * SmemMap::stage_floats;
    }
    void slab_done(long it) { sh->released[it % kStages].fetch_add(1, std::memory_order_acq_rel); }
};

template <bool SR16, int RM, typename S>
void run(const TileArgs& a, int ntiles) {
    Shared sh;
    sh.smem.assign(SmemMap::total_floats, 0.0f);
    sh.tape = a.tape;
    pthread_barrier_init(&sh.bar, nullptr, kThreads);
    {
        EmuEnv<SR16> e0{&sh, 0};
        const long total = (long)ntiles * a.T * Geo<SR16>::nslab;
        for (int i = 0; i < kStages; i++) sh.released[i] = 0;
        long pre = 0;
        for (long i = 0; i < kStages && i < total; i++) { e0.issue(i); pre = i + 1; }
        sh.issued = pre;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < kThreads; t++)
        th.emplace_back([&, t] {
            EmuEnv<SR16> env{&sh, t};
            run_cta<SR16, RM, S>(env, a, 0, 1, ntiles);
        });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&sh.bar);
}
}  // namespace

// ---------------------------------------------------------------- tensor-core schedule (svad_tc.h) on the CPU
// tcgen05.mma is modelled as: D[128][32] (+)= trunc_tf32(A[128][8]) * trunc_tf32(B[8][32]) with A / B fetched through
// the very swizzled shared-memory layouts the descriptors declare, fp32 accumulate.  Executed synchronously by the
// issuing thread; TMEM is a plain array.
namespace {
struct SharedTC {
    std::vector<float> smem;
    std::vector<float> tmem;   // [128 lanes][512 columns]
    pthread_barrier_t bar;
    const float* tape;
    // mbarriers, modelled as completed-phase counters: try_wait.parity(P) succeeds once (count & 1) != P, exactly the
    // hardware rule, and every thread keeps its own parity bits (fpar / mpar) as on the GPU -- so a bookkeeping slip
    // shows up here as a hang or as a read of a buffer whose copy has not been made yet
    std::atomic<int> full_c[8];    // "landed": one phase per bulk copy into the buffer
    std::atomic<int> mdone_c[8];   // "consumed": one phase per tcgen05.commit of the MMA warp that read the buffer
};
inline float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <bool SR16>
struct EmuEnvTC {
    SharedTC* sh;
    int tid_;
    int issued = 0, freed = -1;
    int tid() const { return tid_; }
    float* smem() { return sh->smem.data(); }
    void sync() { pthread_barrier_wait(&sh->bar); }
    void prefetch_l2(const void*) {}
    void fence_async() {}
    void tc_fence_before() {}
    void tc_fence_after() {}
    bool lane0() const { return (tid_ & 31) == 0; }
    using TP = TapeTC<SR16>;
    int issued_idx = 0;   // ring warp: per-step index of the next slab to issue
    uint32_t fpar = 0, mpar = 0;
    void wait_parity(std::atomic<int>& c, uint32_t parity, const char* what, int idx) {
        long spins = 0;
        while ((uint32_t)(c.load(std::memory_order_acquire) & 1) == parity) {
            sched_yield();
            if (++spins == 200000000L) {   // a hang is a bookkeeping bug: say where, then die so the test fails fast
                fprintf(stderr, "svad_emu: thread %d stuck waiting for %s of slab %d (count %d, parity %u)\n", tid_, what, idx, c.load(), parity);
                abort();
            }
        }
    }
    void issue(int idx) {
        const int b = TP::buf(idx);
        memcpy(sh->smem.data() + TP::template buf_off<SmemMapTC>(b), sh->tape + TP::slab_off(idx), sizeof(float) * TP::slab_len(idx));
        sh->full_c[b].fetch_add(1, std::memory_order_release);
    }
    const float* slab_wait(int idx) {
        const int b = TP::buf(idx);
        // the lanes of a warp run in lockstep on the GPU; here they are free-running OS threads, and a lane that lags by two
        // phases would misread the parity.  Only lane 0 consumes the slab (it issues the MMAs), so only lane 0 waits.
        if (lane0()) wait_parity(sh->full_c[b], (fpar >> b) & 1u, "landed", idx);
        fpar ^= 1u << b;
        return sh->smem.data() + TP::template buf_off<SmemMapTC>(b);
    }
    void skip_phase(uint32_t mask, int) { fpar ^= mask; }
    void slab_pass(int idx) { fpar ^= 1u << TP::buf(idx); }
    void wait_consumed_group(int idx, int n) {
        if (!lane0()) return;   // as above: lane 0 stands for the (lockstep) ring warp
        for (int i = 0; i < n - 1; i++) mpar ^= 1u << TP::buf(idx + i);
        const int b = TP::buf(idx + n - 1);
        wait_parity(sh->mdone_c[b], (mpar >> b) & 1u, "consumed", idx + n - 1);
        mpar ^= 1u << b;
    }
    void ring_freed(int total) {
        if (!lane0()) return;
        freed++;
        while (issued < total && issued - TP::dep_delta(issued_idx) <= freed) {
            issue(issued_idx);
            issued++;
            if (++issued_idx == TP::nslab) issued_idx = 0;
        }
    }
    struct BDesc { const float* rows; int lbo_floats; };
    const float* mma_a(const float* tile) { return tile; }
    BDesc mma_b(const float* rows, int lbo_bytes) { return BDesc{rows, lbo_bytes / 4}; }
    template <int MM = 128>
    void mma(int col, const float* a_tile, BDesc b, int ks, bool acc, int ncols) {
        if (!lane0()) return;
        for (int r = 0; r < MM; r++) {
            const int dlane = (MM == 64) ? 32 * (r / 16) + r % 16 : r;   // M = 64: lanes 0-15 of every subpartition
            for (int n = 0; n < ncols; n++) {
                const float* b_rows = b.rows + (n >> 5) * b.lbo_floats + ks * 8 * 32;   // N atom n/32 at stride LBO
                const int nn = n & 31;
                double s = 0.0;
                for (int kk = 0; kk < 8; kk++) {
                    const int k = ks * 8 + kk;
                    const float av = a_tile[(r / 8) * 256 + (r % 8) * 32 + (((k / 4) ^ (r % 8)) * 4) + (k % 4)];
                    const float bv = b_rows[kk * 32 + ((((nn >> 3) ^ (kk & 3)) << 3) | (nn & 7))];   // rows start at a multiple of 8: (row & 3) == (kk & 3)
                    s += (double)tf32_trunc(av) * (double)tf32_trunc(bv);
                }
                float& d = sh->tmem[dlane * 512 + col + n];
                d = (acc ? d : 0.0f) + (float)s;
            }
        }
    }
    template <int MM, int NP>
    void mma_ks4(int col, const float* a0, BDesc b0, const float* a1, BDesc b1, const float* a2, BDesc b2, bool acc_first, int ncols, int ncols12 = 0) {
        if (!ncols12) ncols12 = ncols;
        for (int ks = 0; ks < 4; ks++) {
            mma<MM>(col, a0, b0, ks, ks != 0 || acc_first, ncols);
            if (NP >= 2) mma<MM>(col, a1, b1, ks, true, ncols12);
            if (NP >= 3) mma<MM>(col, a2, b2, ks, true, ncols12);
        }
    }
    void mma_slab_done(int idx) { if (lane0()) sh->mdone_c[TP::buf(idx)].fetch_add(1, std::memory_order_acq_rel); }
    void acc_commit() {}
    void acc_wait() { pthread_barrier_wait(&sh->bar); }
    void tmem_ld16(int lq, int col, float (&v)[16]) {
        const int lane = 32 * lq + (tid_ & 31);
        for (int i = 0; i < 16; i++) v[i] = sh->tmem[lane * 512 + col + i];
    }
};

template <bool SR16, int RM, typename S>
void run_tc(const TileArgs& a, int ntiles) {
    SharedTC sh;
    sh.smem.assign(SmemMapTC::total_floats, 0.0f);
    sh.tmem.assign(128 * 512, 0.0f);
    sh.tape = a.tape;
    pthread_barrier_init(&sh.bar, nullptr, kThreads);
    const int total = (int)((long)ntiles * a.T * TapeTC<SR16>::nslab);
    for (int i = 0; i < 8; i++) { sh.full_c[i] = 0; sh.mdone_c[i] = 0; }
    EmuEnvTC<SR16> boot{&sh, 0};
    boot.freed = -2;
    boot.ring_freed(total);   // freed = -1: primes every buffer whose first slab has no predecessor
    std::vector<std::thread> th;
    for (int t = 0; t < kThreads; t++)
        th.emplace_back([&, t] {
            EmuEnvTC<SR16> env{&sh, t};
            env.issued = boot.issued; env.issued_idx = boot.issued_idx; env.freed = -1;
            run_cta_tc<SR16, RM, S>(env, a, 0, 1, ntiles);
        });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&sh.bar);
}
}  // namespace

extern "C" int svad_emu_forward_tc(const char* weights, int sr, int rm, int B, long L, const void* audio, int pcm16,
                                   const float* state_in, const float* ctx_in, float* state_out, float* ctx_out, float* probs) {
    TensorMap tm;
    std::string err;
    if (!read_container(weights, tm, err)) return -1;
    PackedBranch pb;
    const bool sr16 = sr == 16000;
    if (!(sr16 ? pack_branch_tc<true>(tm, pb, err) : pack_branch_tc<false>(tm, pb, err))) return -2;
    const int n = sr16 ? 512 : 256;
    TileArgs a{};
    a.audio = audio; a.ld = L; a.L = L; a.dec = 1; a.B = B; a.T = (L + n - 1) / n;
    a.state_in = state_in; a.ctx_in = ctx_in; a.ctx_ld = sr16 ? 64 : 32; a.state_out = state_out; a.ctx_out = ctx_out;
    a.probs = probs; a.ldp = a.T; a.tape = pb.tape.data(); a.consts = pb.consts.data();
    const int bt = 4 * rm, ntiles = (B + bt - 1) / bt;
    if (sr16 && rm == 8) pcm16 ? run_tc<true, 8, int16_t>(a, ntiles) : run_tc<true, 8, float>(a, ntiles);
    else if (sr16 && rm == 7) pcm16 ? run_tc<true, 7, int16_t>(a, ntiles) : run_tc<true, 7, float>(a, ntiles);
    else if (!sr16 && rm == 8) pcm16 ? run_tc<false, 8, int16_t>(a, ntiles) : run_tc<false, 8, float>(a, ntiles);
    else if (!sr16 && rm == 7) pcm16 ? run_tc<false, 7, int16_t>(a, ntiles) : run_tc<false, 7, float>(a, ntiles);
    else return -3;
    return 0;
}

extern "C" int svad_emu_forward(const char* weights, int sr, int rm, int B, long L, const void* audio, int pcm16, const float* state_in,
                                const float* ctx_in, float* state_out, float* ctx_out, float* probs) {
    TensorMap tm;
    std::string err;
    if (!read_container(weights, tm, err)) return -1;
    PackedBranch pb;
    const bool sr16 = sr == 16000;
    if (!(sr16 ? pack_branch<true>(tm, pb, err) : pack_branch<false>(tm, pb, err))) return -2;
    const int n = sr16 ? 512 : 256;
    TileArgs a{};
    a.audio = audio; a.ld = L; a.L = L; a.dec = 1; a.B = B; a.T = (L + n - 1) / n;
    a.state_in = state_in; a.ctx_in = ctx_in; a.ctx_ld = sr16 ? 64 : 32; a.state_out = state_out; a.ctx_out = ctx_out;
    a.probs = probs; a.ldp = a.T; a.tape = pb.tape.data(); a.consts = pb.consts.data();
    const int bt = 4 * rm, ntiles = (B + bt - 1) / bt;
    switch (rm * 2 + (sr16 ? 1 : 0)) {
        case 9: pcm16 ? run<true, 4, int16_t>(a, ntiles) : run<true, 4, float>(a, ntiles); break;   case 8: pcm16 ? run<false, 4, int16_t>(a, ntiles) : run<false, 4, float>(a, ntiles); break;
        case 11: pcm16 ? run<true, 5, int16_t>(a, ntiles) : run<true, 5, float>(a, ntiles); break;  case 10: pcm16 ? run<false, 5, int16_t>(a, ntiles) : run<false, 5, float>(a, ntiles); break;
        case 13: pcm16 ? run<true, 6, int16_t>(a, ntiles) : run<true, 6, float>(a, ntiles); break;  case 12: pcm16 ? run<false, 6, int16_t>(a, ntiles) : run<false, 6, float>(a, ntiles); break;
        case 15: pcm16 ? run<true, 7, int16_t>(a, ntiles) : run<true, 7, float>(a, ntiles); break;  case 14: pcm16 ? run<false, 7, int16_t>(a, ntiles) : run<false, 7, float>(a, ntiles); break;
        case 17: pcm16 ? run<true, 8, int16_t>(a, ntiles) : run<true, 8, float>(a, ntiles); break;  case 16: pcm16 ? run<false, 8, int16_t>(a, ntiles) : run<false, 8, float>(a, ntiles); break;
        default: return -3;
    }
    return 0;
}

// ---- static checks of the tensor-core weight stream tables (TapeTC): tape tiling, buffer bounds, and that no slab is ever
// copied into shared memory that an earlier, possibly unconsumed slab still occupies
template <bool SR16>
static int tape_check(char* msg, int n) {
    using TP = TapeTC<SR16>;
    using M = SmemMapTC;
    auto fail = [&](const char* what, int a, int b) { snprintf(msg, n, "%s (%d, %d)", what, a, b); return 1; };
    int off = 0;
    for (int i = 0; i < TP::nslab; i++) {
        if (TP::slab_off(i) != off) return fail("tape is not tiled by the slabs at slab", i, TP::slab_off(i));
        off += TP::slab_len(i);
        const int b = TP::buf(i), lo = TP::template buf_off<M>(b), hi = lo + TP::slab_len(i);
        if (b < 0 || b >= TP::kBufs) return fail("buffer id out of range", i, b);
        const bool ring = b < 4;
        const int r0 = ring ? M::stage : M::e0, r1 = ring ? M::stage + kTcStages * M::stage_floats : M::e0 + M::e0_floats;
        if (lo < r0 || hi > r1) return fail("slab leaves its buffer region", i, b);
        if ((lo * 4) % 1024) return fail("tile base not 1 KB aligned", i, lo * 4);
        // the e0 region holds Z planes / lo0 / e0 until the last enc1 slab has been consumed
        if (!ring && i - TP::dep_delta(i) < TP::e0_nslab + TP::e1_nslab - 1) return fail("e0-region buffer filled before enc1 is done", i, TP::dep_delta(i));
    }
    if (off != TP::total) return fail("tape length", off, TP::total);
    if (TP::NA % 4 || TP::l_nslab % 4) return fail("phase lengths must be multiples of the ring depth", TP::NA, TP::l_nslab);
    for (int g = TP::nslab; g < 3 * TP::nslab; g++) {   // steady state: steps 1 and 2 of 3
        const int i = g % TP::nslab, d = TP::dep_delta(i);
        if (d < 1) return fail("dep_delta < 1", i, d);
        const int lo = TP::template buf_off<M>(TP::buf(i)), hi = lo + TP::slab_len(i);
        for (int gp = g - d + 1; gp < g; gp++) {   // issued earlier, not known to be consumed when slab g goes out
            const int j = gp % TP::nslab, lj = TP::template buf_off<M>(TP::buf(j)), hj = lj + TP::slab_len(j);
            if (lo < hj && lj < hi) return fail("slab overwrites an unconsumed slab", i, j);
        }
        // and the copy must not wait for a slab that comes AFTER something it blocks: d <= slabs in flight capacity
        if (d > TP::nslab) return fail("dep_delta larger than a step", i, d);
    }
    for (int first = 0; first < TP::nslab; first += 4) {
        uint32_t m = 0;
        for (int k = 0; k < 4; k++) m ^= 1u << TP::buf(first + k);
        if (m != TP::phase_mask(first, 4)) return fail("phase_mask", first, (int)m);
    }
    snprintf(msg, n, "ok: %d slabs, %d floats", TP::nslab, TP::total);
    return 0;
}
extern "C" int svad_emu_tape_check(int sr, char* msg, int n) { return sr == 16000 ? tape_check<true>(msg, n) : tape_check<false>(msg, n); }
