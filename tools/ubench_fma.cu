// ubench_fma.cu -- measures what the fp32 FMA pipe of one B200 SM delivers for the register-tile patterns the
// fused kernel uses (roofline denominator for `fp32_frac`; MEASURED_PEAKS.json has no fp32 figure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma tools/ubench_fma.cu && ./ubench_fma
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <int MODE>  // 0: operands in registers, 1: operands re-read from shared memory every k (LDS.128 x4 / 64 FFMA)
__global__ void __launch_bounds__(512, 1) fma_kernel(float* out, int iters, long long* cycles) {
    __shared__ __align__(16) float sa[64 * 32], sb[64 * 64];
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) sa[i] = 1.0f + 1e-6f * i;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sb[i] = 1.0f - 1e-6f * i;
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    const int lm = (threadIdx.x >> 3) & 3, ln = threadIdx.x & 7;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = sa[i + lm * 8]; b[i] = sb[i + ln * 8]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int k = 0; k < 64; k++) {
            if (MODE == 1) {
                const float4 a0 = *reinterpret_cast<const float4*>(sa + k * 32 + 4 * lm);
                const float4 a1 = *reinterpret_cast<const float4*>(sa + k * 32 + 16 + 4 * lm);
                const float4 b0 = *reinterpret_cast<const float4*>(sb + k * 64 + 4 * ln);
                const float4 b1 = *reinterpret_cast<const float4*>(sb + k * 64 + 32 + 4 * ln);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Packed fp32 (Blackwell fma.rn.f32x2 / FFMA2): acc[i][jp] (float2) += {a_i, a_i} * {b_2jp, b_2jp+1}
template <int MODE>  // 0: registers, 1: operands re-read from shared memory every k
__global__ void __launch_bounds__(512, 1) fma2_kernel(float* out, int iters, long long* cycles) {
    __shared__ __align__(16) float sa[64 * 32], sb[64 * 64];
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) sa[i] = 1.0f + 1e-6f * i;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sb[i] = 1.0f - 1e-6f * i;
    __syncthreads();
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);
    const int lm = (threadIdx.x >> 3) & 3, ln = threadIdx.x & 7;
    float a[8];
    float2 b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = sa[i + lm * 8];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = make_float2(sb[2 * j + ln * 8], sb[2 * j + 1 + ln * 8]);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int k = 0; k < 64; k++) {
            if (MODE == 1) {
                const float4 a0 = *reinterpret_cast<const float4*>(sa + k * 32 + 4 * lm);
                const float4 a1 = *reinterpret_cast<const float4*>(sa + k * 32 + 16 + 4 * lm);
                const float4 b0 = *reinterpret_cast<const float4*>(sb + k * 64 + 4 * ln);
                const float4 b1 = *reinterpret_cast<const float4*>(sb + k * 64 + 32 + 4 * ln);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                b[0] = make_float2(b0.x, b0.y); b[1] = make_float2(b0.z, b0.w); b[2] = make_float2(b1.x, b1.y); b[3] = make_float2(b1.z, b1.w);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(aa, b[j], acc[i][j]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int threads, int sms) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 2000;
    if (MODE < 2) fma_kernel<MODE & 1><<<sms, threads>>>(out, 10, cyc); else fma2_kernel<MODE & 1><<<sms, threads>>>(out, 10, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (MODE < 2) fma_kernel<MODE & 1><<<sms, threads>>>(out, iters, cyc); else fma2_kernel<MODE & 1><<<sms, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double fma = (double)iters * 64 * 64 * threads;            // per SM
    double per_clk = fma / (double)h[0];
    double tflops = 2.0 * fma * sms / (ms * 1e-3) / 1e12;
    printf("mode=%d threads=%4d : %.1f FMA/clk/SM (clock64), %.2f TFLOP/s chip (events, %.3f ms), implied clock %.0f MHz\n", MODE, threads,
           per_clk, tflops, ms, (double)h[0] / (ms * 1e-3) / 1e6);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int t : {128, 256, 512}) run<0>(t, p.multiProcessorCount);
    for (int t : {128, 256, 512}) run<1>(t, p.multiProcessorCount);
    printf("packed fp32 (fma.rn.f32x2):\n");
    for (int t : {128, 256, 512}) run<2>(t, p.multiProcessorCount);
    for (int t : {128, 256, 512}) run<3>(t, p.multiProcessorCount);
    return 0;
}
