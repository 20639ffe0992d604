// ubench_feed.cu -- how fast does a ring of S stages x 16 KB pull an L2-resident weight tape into shared memory, every SM at once?
//   mode 0: one thread issues cp.async.bulk (TMA bulk copy), completion on an mbarrier (what svad_fused_h16 does)
//   mode 1: one warp issues cp.async.cg 16-byte copies (LDGSTS), completion through cp.async.mbarrier.arrive.noinc
// The consumer (another thread) waits for a stage, "uses" it for `hold` cycles, and releases it: the steady-state cycles per slab
// give latency = S * cycles_per_slab when the ring is latency-bound.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench_feed tools/ubench_feed.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}


// `rings` independent rings (each: producer warp 2r, consumer warp 2r + 1), each of `stages` x kSlab bytes
__global__ void __launch_bounds__(256, 1) feed(const unsigned char* tape, int nslab_tape, int stages, int mode, int hold, long total, long long* cycles, int kSlab, int rings) {
    extern __shared__ __align__(1024) unsigned char smem_all[];
    const int ring = threadIdx.x >> 6;
    unsigned char* smem = smem_all + (size_t)ring * stages * kSlab;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_all + (size_t)rings * stages * kSlab) + ring * 32;
    uint64_t* empty = full + 16;
    const int warp = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
    if (ring >= rings) return;
    if ((threadIdx.x & 63) == 0) {
        for (int s = 0; s < stages; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(full + s)), "r"(mode == 0 ? 1 : 32));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(empty + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("bar.sync %0, 64;" ::"r"(ring + 1) : "memory");
    const unsigned char* base = tape + (size_t)(blockIdx.x % 7) * 4096 + (size_t)ring * 3 * 16384;   // CTAs do not all start on the same line
    if (warp == 0) {   // producer
        int stage = 0; uint32_t round = 0; int idx = 0;
        for (long i = 0; i < total; i++) {
            if (round > 0 && (mode == 1 || lane == 0)) mbar_wait(su32(empty + stage), (round - 1) & 1u);
            const unsigned char* src = base + (size_t)idx * kSlab;
            const uint32_t dst = su32(smem + stage * kSlab), bar = su32(full + stage);
            if (mode == 0) {
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kSlab) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(kSlab), "r"(bar) : "memory");
                }
            } else {
#pragma unroll 8
                for (int k = 0; k < kSlab / 512; k++)   // kSlab is a multiple of 512
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 512 + lane * 16), "l"(src + k * 512 + lane * 16) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
            }
            if (++idx == nslab_tape) idx = 0;
            if (++stage == stages) { stage = 0; round++; }
        }
    } else if (warp == 1 && lane == 0) {   // consumer
        int stage = 0; uint32_t round = 0;
        long long t0 = 0;
        for (long i = 0; i < total; i++) {
            mbar_wait(su32(full + stage), round & 1u);
            if (i == 16) t0 = clock64();
            const long long t = clock64();
            while (clock64() - t < hold) {}
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(empty + stage)) : "memory");
            if (++stage == stages) { stage = 0; round++; }
        }
        if (blockIdx.x == 0 && ring == 0) cycles[0] = clock64() - t0;
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int tape_bytes = 68 * 16384;
    unsigned char* tape;
    cudaMalloc(&tape, (size_t)tape_bytes + 8 * 16384);
    cudaMemset(tape, 1, (size_t)tape_bytes + 8 * 16384);
    long long* dcy;
    cudaMalloc(&dcy, 8);
    struct Cfg { int mode, slab, stages, rings; };
    const Cfg cfgs[] = {{0, 16384, 1, 1}, {0, 16384, 2, 1}, {0, 16384, 3, 1}, {0, 16384, 6, 1}, {0, 8192, 4, 1}, {0, 8192, 8, 1}, {0, 32768, 2, 1}, {0, 32768, 3, 1},
                        {0, 49152, 2, 1}, {0, 65536, 2, 1}, {0, 65536, 3, 1}, {0, 16384, 3, 2}, {0, 16384, 2, 4}, {0, 32768, 2, 2}, {0, 32768, 1, 2}, {0, 32768, 1, 4},
                        {1, 16384, 3, 1}, {1, 16384, 3, 2}, {1, 16384, 2, 4}};
    for (const Cfg& c : cfgs) {
        const int nslab_tape = tape_bytes / c.slab;
        const long total = (long)nslab_tape * 16;
        const size_t smem = (size_t)c.rings * c.stages * c.slab + 1024;
        cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        feed<<<prop.multiProcessorCount, 256, smem>>>(tape, nslab_tape, c.stages, c.mode, 0, total, dcy, c.slab, c.rings);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long cy = 0;
        cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cy / (double)(total - 16);
        printf("%s slab %2d KB x %d stages x %d rings (%3d KB in flight): %7.1f cycles per slab per ring = %6.1f B/clk/SM total\n", c.mode == 0 ? "TMA bulk" : "LDGSTS  ",
               c.slab / 1024, c.stages, c.rings, c.slab * c.stages * c.rings / 1024, per, (double)c.slab * c.rings / per);
    }
    return 0;
}
