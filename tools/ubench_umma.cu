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

// The fused kernel's enc2 phase verbatim: 4 warps, each 4 k-steps x {hi*hi, hi*lo, lo*hi} (M = 64, N = 32) on its own A tile pair,
// B rows and accumulator, descriptors rebuilt per instruction, one elect per instruction.  uniform != 0: the warp index comes
// from a shuffle broadcast, which lets the compiler keep the descriptor arithmetic in uniform registers.
__device__ __forceinline__ void mma_one(uint32_t tm, int col, uint64_t ad, uint64_t bd, int ks, bool acc, int ncols, int M) {
    const uint64_t a2 = ad + (uint64_t)(ks * 2), b2 = bd + (uint64_t)(ks * 64);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (elect())
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm + (uint32_t)col), "l"(a2), "l"(b2),
                     "r"(idesc), "r"(acc ? 1u : 0u)
                     : "memory");
}
__global__ void __launch_bounds__(256, 1) bench_phase(int uniform, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (4 * 16384 + 2 * 16384) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(4));
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
    int warp = threadIdx.x >> 5;
    if (uniform) warp = __shfl_sync(0xffffffffu, warp, 0);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; rep++) {   // rep 0 warms the instruction cache
        __syncthreads();
        t0 = clock64();
        if (warp < 4) {
            const int q = warp;
            const float* tile = reinterpret_cast<const float*>(smem) + q * 4096;
            const float* bh_rows = reinterpret_cast<const float*>(smem) + 4 * 4096 + q * 1024;
            const float* bl_rows = bh_rows + 4096;
            const uint64_t ah = make_desc(su32(tile), 16, 1024, 2), al = make_desc(su32(tile + 2048), 16, 1024, 2);
            const uint64_t bh = make_desc(su32(bh_rows), 4096, 512, 1), bl = make_desc(su32(bl_rows), 4096, 512, 1);
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                mma_one(tm, 32 * q, ah, bh, ks, ks != 0, 32, 64);
                mma_one(tm, 32 * q, ah, bl, ks, true, 32, 64);
                mma_one(tm, 32 * q, al, bh, ks, true, 32, 64);
            }
            if (elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
        }
        t1 = clock64();
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
        t2 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// Do bulk copies (cp.async.bulk global -> shared, the fused kernel's weight stream) in flight slow the MMA issue?  Warps 0..nw-1
// issue MMAs as in bench(); warp 7 streams `ncopy` copies of copy_bytes through two staging buffers at the same time.
__global__ void __launch_bounds__(256, 1) bench_mix(int M, int N, int nwarps, int reps, int copy_bytes, int ncopy, const char* src, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, cbar[2];
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(nwarps));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&cbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&cbar[1])));
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
    long long t1 = t0;
    if (warp < nwarps) {
        const uint64_t ad = make_desc(su32(smem), 16, 1024, 2);
        const uint64_t bd = make_desc(su32(smem + 16384), 4096, 512, 1);
        const uint32_t col = tm + (uint32_t)(warp * N);
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t a2 = ad + (uint64_t)(ks * 2), b2 = bd + (uint64_t)(ks * 64);
                if (elect())
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(col), "l"(a2),
                                 "l"(b2), "r"(idesc), "r"(1u)
                                 : "memory");
            }
        }
        if (elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
        t1 = clock64();
    } else if (warp == 7 && copy_bytes > 0) {
        unsigned char* stage = smem + 16384 + 32768;
        for (int c = 0; c < ncopy + 1; c++) {
            if (c < ncopy && (threadIdx.x & 31) == 0) {
                const uint32_t b = su32(&cbar[c & 1]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(copy_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(stage + (c & 1) * copy_bytes)),
                             "l"(src + (size_t)(c % 16) * copy_bytes), "r"(copy_bytes), "r"(b)
                             : "memory");
            }
            if (c >= 1) {
                const int w = c - 1;
                asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&cbar[w & 1])), "r"((uint32_t)((w >> 1) & 1)) : "memory");
            }
        }
        t1 = clock64();
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        if (threadIdx.x == 224) out[2] = t1 - t0;
    }
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
    {
        char* src; cudaMalloc(&src, 16 * 65536); cudaMemset(src, 0, 16 * 65536);
        long long* d3; cudaMalloc(&d3, 24);
        const size_t sm2 = 16384 + 32768 + 2 * 65536;
        cudaFuncSetAttribute(bench_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        const int cfg[][2] = {{0, 0}, {16384, 40}, {32768, 40}, {65536, 30}, {4096, 80}};
        for (int nw = 1; nw <= 4; nw *= 4)
            for (auto& c : cfg) {
                cudaMemset(d3, 0, 24);
                bench_mix<<<p.multiProcessorCount, 256, sm2>>>(128, 32, nw, 64, c[0], c[1], src, d3);
                long long h[3]; cudaMemcpy(h, d3, 24, cudaMemcpyDeviceToHost);
                printf("MMA (M=128 N=32, %d warps x 256) with %2d concurrent bulk copies of %5d B: MMA issue %.1f cyc/MMA/warp, all done %lld cycles; copy warp busy %lld cycles  %s\n", nw,
                       c[1], c[0], (double)h[0] / 256, h[1], h[2], cudaGetErrorString(cudaDeviceSynchronize()));
            }
    }
    cudaFuncSetAttribute(bench_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384);
    for (int u = 0; u < 2; u++) {
        bench_phase<<<p.multiProcessorCount, 256, 6 * 16384>>>(u, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("enc2-like phase (4 warps x 12 MMAs, M=64 N=32, per-instruction descriptors, %s warp index): issue %lld cycles, done %lld cycles  %s\n",
               u ? "shuffle-broadcast" : "plain", h[0], h[1], cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
