// ubench_tma2d.cu -- the audio box of svad_fused_h16 in isolation: a 2-D tensor map over a [rows][L] fp32 matrix, box 32 rows x 128 samples,
// one cp.async.bulk.tensor.2d per box into shared memory, completion on an mbarrier.  Checks values (incl. negative / out-of-range
// coordinates = zero fill) and times a ring of boxes per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench_tma2d tools/ubench_tma2d.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int x, int y, float* out, long long* cyc, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(su32(smem)), "l"(&tm), "r"(x + 128 * (r % 4)), "r"(y + 32 * (int)blockIdx.x), "r"(su32(bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(su32(bar)), "r"(r & 1) : "memory");
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
    if (blockIdx.x == 0 && reps == 1)
        for (int i = threadIdx.x; i < 4096; i += 128) out[i] = reinterpret_cast<const float*>(smem)[i];
}

int main() {
    const int rows = 4096, L = 2048;
    std::vector<float> h((size_t)rows * L);
    for (int r = 0; r < rows; r++)
        for (int i = 0; i < L; i++) h[(size_t)r * L + i] = (float)r + 1e-4f * (float)i;
    float *d, *out;
    long long* cyc;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (!p) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)L, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)L * 4};
    const cuuint32_t box[2] = {128, 32}, es[2] = {1, 1};
    CUresult r = ((fn_t)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64);
    const int xs[] = {0, 448, -64, L - 64};
    for (int x : xs) {
        probe<<<1, 128, 16384 + 64>>>(tm, x, 4, out, cyc, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("x=%d: CUDA error %s\n", x, cudaGetErrorString(e)); return 1; }
        std::vector<float> o(4096);
        cudaMemcpy(o.data(), out, 4096 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < 32; rr++)
            for (int i = 0; i < 128; i++) {
                const int xi = x + i;
                const float want = (xi < 0 || xi >= L) ? 0.0f : (float)(4 + rr) + 1e-4f * (float)xi;
                if (o[rr * 128 + i] != want) bad++;
            }
        printf("x = %5d: box[0][0..2] = %.4f %.4f %.4f  box[1][0] = %.4f  mismatches %d\n", x, o[0], o[1], o[2], o[128], bad);
    }
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    probe<<<prop.multiProcessorCount, 128, 16384 + 64>>>(tm, 0, 0, out, cyc, 64);
    cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%d CTAs, 64 serial boxes of 16 KB each: %.0f cycles per box (issue -> landed)\n", prop.multiProcessorCount, (double)c / 64);
    return 0;
}
