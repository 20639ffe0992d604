// svad_emu.cpp -- CPU emulator of one CTA of the fused fp32 kernel (TEST INFRASTRUCTURE).
// Runs svad::run_cta<> (the exact schedule + per-thread arithmetic the CUDA kernel compiles) on 256
// OS threads with a pthread barrier standing in for __syncthreads and memcpy standing in for the
// TMA bulk copies of weight slabs.  Lets the layout / index algebra be checked against the oracle
// in the build container, which has no GPU.  Not part of the product; never timed.
#include <pthread.h>
#include <sched.h>

#include <atomic>

#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../silero_vad_b200/csrc/svad_tile.h"

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
    a.audio = audio; a.ld = L; a.L = L; a.B = B; a.T = (L + n - 1) / n;
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
