// ubench_ingest.cu -- how fast can every SM pull the same L2-resident weight tape into shared memory with
// cp.async.bulk (TMA bulk copy)?  Decides whether a tensor-core path (2x weight bytes per step) can be fed.
//   mode 0: every CTA copies the whole tape itself;  mode C (2,4,8): cluster of C CTAs, each copies 1/C of every
//   slab and multicasts it to all C.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

#ifndef KSTAGE
#define KSTAGE 32768
#endif
#ifndef KSTAGES
#define KSTAGES 4
#endif
#ifndef KISSUERS
#define KISSUERS 1   // bulk copies are issued round-robin by this many different warps
#endif
constexpr int kStage = KSTAGE, kStages = KSTAGES;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <int C>
__global__ void __launch_bounds__(128, 1) ingest(const char* tape, int tape_bytes, int reps, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStage * kStages);
    uint32_t rank = 0;
    if (C > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(full + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (C > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    const int nslab = tape_bytes / kStage;
    const long total = (long)nslab * reps;
    float acc = 0.f;
    long long t0 = clock64();
    // issue-ahead by kStages-1; consumers just touch one word per slab (we measure the copy engine, not LDS)
    for (long it = 0; it < total + kStages - 1; it++) {
        if (it < total && threadIdx.x == 32 * (int)(it % KISSUERS)) {
            const int stage = it % kStages;
            const uint32_t bar = su32(full + stage);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kStage) : "memory");
            const char* src = tape + (size_t)(it % nslab) * kStage;
            if (C == 1) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(smem + stage * kStage)), "l"(src), "r"(kStage), "r"(bar) : "memory");
            } else {
                const int part = kStage / C;
                const uint16_t mask = (uint16_t)((1u << C) - 1);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(su32(smem + stage * kStage + rank * part)), "l"(src + rank * part), "r"(part), "r"(bar), "h"(mask) : "memory");
            }
        }
        const long c = it - (kStages - 1);
        if (c >= 0) {
            mbar_wait(su32(full + (c % kStages)), (uint32_t)((c / kStages) & 1));
            acc += reinterpret_cast<const float*>(smem + (c % kStages) * kStage)[threadIdx.x];
            if (C > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
            else __syncthreads();
        }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int C>
void run(const char* tape, int bytes, int sms) {
    const int reps = 40;
    const size_t smem = kStage * kStages + 64;
    cudaFuncSetAttribute(ingest<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (C > 8) cudaFuncSetAttribute(ingest<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    long long* cyc; float* sink;
    const int grid = sms / C * C;
    cudaMalloc(&cyc, 8 * grid); cudaMalloc(&sink, 4 * grid * 128);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, ingest<C>, tape, bytes, 2, cyc, sink);
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, ingest<C>, tape, bytes, reps, cyc, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, 8 * grid, cudaMemcpyDeviceToHost);
    double per_sm_bytes = (double)bytes * reps;
    printf("cluster=%d grid=%d: %s  %.1f B/clk/SM landed, %.2f TB/s aggregate landed, L2 read %.2f TB/s (%.3f ms)\n", C, grid, cudaGetErrorString(err),
           per_sm_bytes / (double)h[0], per_sm_bytes * grid / (ms * 1e-3) / 1e12, per_sm_bytes * grid / C / (ms * 1e-3) / 1e12, ms);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int bytes = 1792 * 1024;   // ~ hi+lo weight tape of the 16 kHz branch
    char* tape; cudaMalloc(&tape, bytes); cudaMemset(tape, 0, bytes);
    printf("%s, %d SMs, tape %d KB\n", p.name, p.multiProcessorCount, bytes / 1024);
    printf("stage %d B x %d stages, %d issuing warps\n", kStage, kStages, KISSUERS);
    run<1>(tape, bytes, p.multiProcessorCount);
    run<2>(tape, bytes, p.multiProcessorCount);   // every CTA of the cluster fetches 1/C of each stage and multicasts it
    run<4>(tape, bytes, p.multiProcessorCount);
    return 0;
}
