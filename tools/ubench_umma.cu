// ubench_umma.cu -- issue/throughput cost of one tcgen05.mma.kind::tf32 (K = 8) as a function of M, N, the B-operand
// layout (MN-major SWIZZLE_128B_BASE32B = activation rows as the fused kernel keeps them, or K-major SWIZZLE_128B) and
// the number of warps issuing concurrently (each into its own TMEM columns).  Operands are zeros; only time matters.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_umma tools/ubench_umma.cu && ./ubench_umma
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ bool elect() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}

// smem: A tile [128 x 32] K-major (16 KB) | B up to [256 n x 32 k] (32 KB)
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int b_kmajor, int nwarps, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(nwarps));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const int warp = threadIdx.x >> 5;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((b_kmajor ? 0u : 1u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
    if (warp < nwarps) {
        const uint64_t ad = make_desc(su32(smem), 16, 1024, 2);
        const uint64_t bd = b_kmajor ? make_desc(su32(smem + 16384), 16, 1024, 2) : make_desc(su32(smem + 16384), 4096, 512, 1);
        const uint32_t col = tm + (uint32_t)(warp * N);
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t a2 = ad + (uint64_t)(ks * 2), b2 = bd + (uint64_t)(b_kmajor ? ks * 2 : ks * 64);
                if (elect())
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(col), "l"(a2),
                                 "l"(b2), "r"(idesc), "r"(1u)
                                 : "memory");
            }
        }
        if (elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t smem = 16384 + 32768;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d; cudaMalloc(&d, 16);
    const int reps = 64;   // x 4 k-steps = 256 MMAs per warp
    printf("%s: cycles per tcgen05.mma (kind::tf32, K = 8), %d MMAs per issuing warp; issue = issue loop only, done = until commit lands\n", p.name, reps * 4);
    for (int bk = 0; bk < 2; bk++)
        for (int M = 64; M <= 128; M += 64)
            for (int nw = 1; nw <= 4; nw *= 4)
                for (int N = 32; N <= 256; N *= 2) {
                    if (nw * N > 512) continue;
                    bench<<<p.multiProcessorCount, 128, smem>>>(M, N, bk, nw, 4, d);
                    bench<<<p.multiProcessorCount, 128, smem>>>(M, N, bk, nw, reps, d);
                    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    cudaError_t e = cudaDeviceSynchronize();
                    const double n = (double)reps * 4 * nw;
                    printf("B %s  M=%3d N=%3d warps=%d: issue %.1f cyc/MMA/warp, done %.1f cyc/MMA overall -> %.0f MAC/clk/SM %s\n", bk ? "K-major " : "MN-major", M, N, nw,
                           (double)h[0] / (reps * 4), (double)h[1] / n, (double)M * N * 8 * n / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    // start-up cost: a short burst of MMAs (M=128, N=32) issued cold / after an idle gap
    for (int reps = 1; reps <= 16; reps *= 2) {
        bench<<<p.multiProcessorCount, 128, smem>>>(128, 32, 0, 1, reps, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("burst of %3d MMAs (M=128 N=32, 1 warp): issue %lld cycles, done %lld cycles\n", reps * 4, h[0], h[1]);
    }
    for (int reps = 1; reps <= 16; reps *= 4) {
        bench<<<p.multiProcessorCount, 128, smem>>>(64, 32, 0, 4, reps, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("burst of %3d MMAs per warp (M=64 N=32, 4 warps): issue %lld cycles, done %lld cycles\n", reps * 4, h[0], h[1]);
    }
    return 0;
}
