// umma_f16_unit.cu -- standalone check of the tcgen05 building blocks of the fp16 split-precision kernel (svad_h16.cuh):
//   A (weights)      [M x 64] fp16, K-major, SWIZZLE_128B (rows of 128 B, 16-byte chunk ^= row & 7), M = 128 or 64
//   B (activations)  [K rows][32 n] fp16, MN-major, SWIZZLE_64B (rows of 64 B, 16-byte chunk ^= (row >> 1) & 3), atoms of 8 rows;
//                    several N atoms at the descriptor's LBO stride (hi | lo rows, or overlapping STFT frames 128 rows apart)
//   D in TMEM, fp32, kind::f16, K = 16 per instruction
// Reports max |D - ref| for (1) N = 32, (2) N = 64 via LBO, (3) N = 128 with overlapping row windows (the STFT operand),
// (4) M = 64 and its TMEM lane map, (5) the 3-product split against fp64, (6) cycles per instruction for the shapes used.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/umma_f16_unit tools/umma_f16_unit.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}
// D fp32 (1 << 4), A / B fp16 (format 0), A K-major, B MN-major (1 << 16)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Job {
    int M, N, nk;            // nk = number of K=16 steps
    int a_off[3];            // byte offsets of up to 3 A tiles in smem (product p uses a_off[p])
    int b_off[3];            // byte offsets of the B row blocks
    int lbo;                 // B: N-atom stride in bytes
    int nprod;               // products accumulated per k-step (1 or 3)
    int a_kstep, b_kstep;    // descriptor-low-word increments per k-step (A: 2 = 32 B; B: 64 = 16 rows x 64 B)
    int a_tile_bytes;        // A advances by this many bytes every 4 k-steps (next [M x 64] tile)
    int reps;                // timing: repeat the whole job
};

__global__ void __launch_bounds__(128, 1) umma_f16(const unsigned char* img, int img_bytes, Job j, float* d_out, long long* clk) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < img_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(img)[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
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
    const uint32_t idesc = idesc_f16(j.M, j.N);
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        if (j.reps > 1) {   // timing: 4 instructions per asm block (descriptors advance in registers), like the kernel issues them
            const uint64_t ad = make_desc(su32(smem + j.a_off[0]), 16, 1024, 2);
            const uint64_t bd = make_desc(su32(smem + j.b_off[0]), (uint32_t)j.lbo, 512, 4);
            for (int r = 0; r < j.reps; r++)
                asm volatile("{\n.reg .pred t;\n.reg .b64 a, b;\nsetp.eq.u32 t, %3, %3;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, t;\n"
                             "add.s64 a, %1, 2; add.s64 b, %2, 64;\ntcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n"
                             "add.s64 a, %1, 4; add.s64 b, %2, 128;\ntcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n"
                             "add.s64 a, %1, 6; add.s64 b, %2, 192;\ntcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n}\n"
                             ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        } else {
        for (int r = 0; r < j.reps; r++)
            for (int ks = 0; ks < j.nk; ks++)
                for (int p = 0; p < j.nprod; p++) {
                    const uint64_t ad = make_desc(su32(smem + j.a_off[p] + (ks >> 2) * j.a_tile_bytes), 16, 1024, 2) + (uint64_t)((ks & 3) * j.a_kstep);
                    const uint64_t bd = make_desc(su32(smem + j.b_off[p]), (uint32_t)j.lbo, 512, 4) + (uint64_t)(ks * j.b_kstep);
                    const uint32_t acc = (r | ks | p) ? 1u : 0u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(ad),
                                 "l"(bd), "r"(idesc), "r"(acc)
                                 : "memory");
                }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < j.N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) d_out[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// ---- host layouts
// A[r][k] (r < M, k < 64 per tile; ntile tiles along K) -> K-major SWIZZLE_128B tiles of M x 64 fp16
static void pack_a(const float* A, int M, int K, __half* dst) {
    for (int t = 0; t < K / 64; t++)
        for (int r = 0; r < M; r++)
            for (int k = 0; k < 64; k++) {
                const int chunk = (k / 8) ^ (r % 8);
                dst[(size_t)t * M * 64 + (r / 8) * 512 + (r % 8) * 64 + chunk * 8 + (k % 8)] = __float2half(A[(size_t)r * K + t * 64 + k]);
            }
}
// X[row][n] (n < 32) -> MN-major SWIZZLE_64B rows of 64 B: element (row, n) at row*32 + (((n/8) ^ ((row>>1)&3)) * 8) + n%8
static void pack_b(const float* X, int rows, __half* dst) {
    for (int r = 0; r < rows; r++)
        for (int n = 0; n < 32; n++) dst[(size_t)r * 32 + ((((n / 8) ^ ((r >> 1) & 3))) * 8) + (n % 8)] = __float2half(X[(size_t)r * 32 + n]);
}
static float h2f(float v) { return __half2float(__float2half(v)); }

static std::vector<float> run(const std::vector<unsigned char>& img, const Job& j, long long* cycles = nullptr) {
    unsigned char* dimg; float* dd; long long* dclk;
    cudaMalloc(&dimg, img.size()); cudaMalloc(&dd, 128 * 256 * 4); cudaMalloc(&dclk, 8);
    cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * 256 * 4);
    cudaFuncSetAttribute(umma_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)img.size());
    umma_f16<<<1, 128, img.size()>>>(dimg, (int)img.size(), j, dd, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<float> d(128 * 256);
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    long long c = 0; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
    if (cycles) *cycles = c;
    cudaFree(dimg); cudaFree(dd); cudaFree(dclk);
    return d;
}

int main() {
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    int fails = 0;
    // ---------------- (1) M=128, N=32, K=64: one A tile, one B block of 64 rows
    {
        std::vector<float> A(128 * 64), X(64 * 32);
        for (auto& v : A) v = h2f(nd(rng));
        for (auto& v : X) v = h2f(nd(rng));
        std::vector<unsigned char> img(16384 + 4096);
        pack_a(A.data(), 128, 64, (__half*)img.data());
        pack_b(X.data(), 64, (__half*)(img.data() + 16384));
        Job j{128, 32, 4, {0, 0, 0}, {16384, 0, 0}, 4096, 1, 2, 64, 16384, 1};
        auto d = run(img, j);
        double e = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 32; n++) {
                double ref = 0;
                for (int k = 0; k < 64; k++) ref += (double)A[m * 64 + k] * X[k * 32 + n];
                e = fmax(e, fabs(d[m * 256 + n] - ref));
            }
        printf("(1) M=128 N=32 K=64 (A K-major SW128, B MN-major SW64): max|D-ref| = %.3e %s\n", e, e < 1e-3 ? "OK" : "FAIL");
        fails += !(e < 1e-3);
    }
    // ---------------- (2) N=64 through LBO: second atom column block = another 64-row block 8 KB above
    {
        std::vector<float> A(128 * 64), X0(64 * 32), X1(64 * 32);
        for (auto& v : A) v = h2f(nd(rng));
        for (auto& v : X0) v = h2f(nd(rng));
        for (auto& v : X1) v = h2f(nd(rng));
        std::vector<unsigned char> img(16384 + 8192 + 4096);
        pack_a(A.data(), 128, 64, (__half*)img.data());
        pack_b(X0.data(), 64, (__half*)(img.data() + 16384));
        pack_b(X1.data(), 64, (__half*)(img.data() + 16384 + 8192));
        Job j{128, 64, 4, {0, 0, 0}, {16384, 0, 0}, 8192, 1, 2, 64, 16384, 1};
        auto d = run(img, j);
        double e = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 64; n++) {
                double ref = 0;
                for (int k = 0; k < 64; k++) ref += (double)A[m * 64 + k] * (n < 32 ? X0[k * 32 + n] : X1[k * 32 + n - 32]);
                e = fmax(e, fabs(d[m * 256 + n] - ref));
            }
        printf("(2) N=64 as two N atoms at LBO = 8192: max|D-ref| = %.3e %s\n", e, e < 1e-3 ? "OK" : "FAIL");
        fails += !(e < 1e-3);
    }
    // ---------------- (3) STFT operand: xp[640 rows][32], four overlapping 256-row windows 128 rows apart, N=128, K=256 (4 A tiles)
    {
        std::vector<float> A(128 * 256), X(640 * 32);
        for (auto& v : A) v = h2f(nd(rng) * 0.1f);
        for (auto& v : X) v = h2f(nd(rng));
        std::vector<unsigned char> img(4 * 16384 + 640 * 64);
        pack_a(A.data(), 128, 256, (__half*)img.data());
        pack_b(X.data(), 640, (__half*)(img.data() + 65536));
        Job j{128, 128, 16, {0, 0, 0}, {65536, 0, 0}, 128 * 64, 1, 2, 64, 16384, 1};
        auto d = run(img, j);
        double e = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 128; n++) {
                const int f = n / 32, s = n % 32;
                double ref = 0;
                for (int k = 0; k < 256; k++) ref += (double)A[m * 256 + k] * X[(128 * f + k) * 32 + s];
                e = fmax(e, fabs(d[m * 256 + n] - ref));
            }
        printf("(3) STFT operand: N=128 = 4 frame windows at LBO = 128 rows, K=256: max|D-ref| = %.3e %s\n", e, e < 2e-3 ? "OK" : "FAIL");
        fails += !(e < 2e-3);
    }
    // ---------------- (4) M=64: rows of the 64-row A tile; accumulator lane map
    {
        std::vector<float> A(64 * 64), X(64 * 32);
        for (auto& v : A) v = h2f(nd(rng));
        for (auto& v : X) v = h2f(nd(rng));
        std::vector<unsigned char> img(8192 + 4096);
        pack_a(A.data(), 64, 64, (__half*)img.data());
        pack_b(X.data(), 64, (__half*)(img.data() + 8192));
        Job j{64, 32, 4, {0, 0, 0}, {8192, 0, 0}, 4096, 1, 2, 64, 8192, 1};
        auto d = run(img, j);
        double e = 0;
        for (int m = 0; m < 64; m++)
            for (int n = 0; n < 32; n++) {
                double ref = 0;
                for (int k = 0; k < 64; k++) ref += (double)A[m * 64 + k] * X[k * 32 + n];
                const int lane = 32 * (m / 16) + (m % 16);
                e = fmax(e, fabs(d[lane * 256 + n] - ref));
            }
        printf("(4) M=64 (row r in TMEM lane 32*(r/16) + r%%16): max|D-ref| = %.3e %s\n", e, e < 1e-3 ? "OK" : "FAIL");
        fails += !(e < 1e-3);
    }
    // ---------------- (5) split precision: w.x ~= wh.xh + wh.xl + wl.xh, scaled operands, one accumulator; K = 256
    {
        const int K = 256;
        std::vector<float> W(128 * K), X(K * 32);
        for (auto& v : W) v = nd(rng) * 0.2f;
        for (auto& v : X) v = fabsf(nd(rng)) * 3.0f;
        const float sw = 4096.f, sx = 64.f;
        std::vector<float> Wh(W.size()), Wl(W.size()), Xh(X.size()), Xl(X.size());
        for (size_t i = 0; i < W.size(); i++) { Wh[i] = h2f(W[i] * sw); Wl[i] = h2f(W[i] * sw - Wh[i]); }
        for (size_t i = 0; i < X.size(); i++) { Xh[i] = h2f(X[i] * sx); Xl[i] = h2f(X[i] * sx - Xh[i]); }
        const int a_bytes = 4 * 16384, b_bytes = K * 64;
        std::vector<unsigned char> img(2 * a_bytes + 2 * b_bytes);
        pack_a(Wh.data(), 128, K, (__half*)img.data());
        pack_a(Wl.data(), 128, K, (__half*)(img.data() + a_bytes));
        pack_b(Xh.data(), K, (__half*)(img.data() + 2 * a_bytes));
        pack_b(Xl.data(), K, (__half*)(img.data() + 2 * a_bytes + b_bytes));
        Job j{128, 32, K / 16, {0, 0, a_bytes}, {2 * a_bytes, 2 * a_bytes + b_bytes, 2 * a_bytes}, 4096, 3, 2, 64, 16384, 1};
        auto d = run(img, j);
        double e3 = 0, e32 = 0, mx = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 32; n++) {
                double ex = 0; float f32 = 0.f;
                for (int k = 0; k < K; k++) { ex += (double)W[m * K + k] * X[k * 32 + n]; f32 = fmaf(W[m * K + k], X[k * 32 + n], f32); }
                e3 = fmax(e3, fabs((double)d[m * 256 + n] / (sw * sx) - ex)); e32 = fmax(e32, fabs(f32 - ex)); mx = fmax(mx, fabs(ex));
            }
        printf("(5) fp16 split, 3 products, K=256: max|D-exact| = %.3e (plain fp32 fmaf chain %.3e, max|D| %.1f) %s\n", e3, e32, mx, e3 < 2e-4 ? "OK" : "FAIL");
        fails += !(e3 < 2e-4);
    }
    // ---------------- (6) cycles per instruction (operands zero), one issuing thread, 256 instructions
    {
        struct { int M, N; } shapes[] = {{128, 32}, {128, 64}, {128, 96}, {128, 128}, {128, 256}, {64, 32}, {64, 64}, {64, 128}};
        for (auto s : shapes) {
            std::vector<unsigned char> img(16384 + 65536, 0);
            Job j{s.M, s.N, 4, {0, 0, 0}, {16384, 0, 0}, 4096, 1, 2, 64, 0, 256};
            long long c = 0;
            run(img, j, &c);
            printf("(6) kind::f16 M=%3d N=%3d K=16: %6.1f cycles per instruction (%lld cycles / 1024 instructions, 4 per asm block)\n", s.M, s.N, c / 1024.0, c);
        }
    }
    printf(fails ? "FAILED (%d)\n" : "ALL OK\n", fails);
    return fails != 0;
}
