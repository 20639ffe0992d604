// umma_unit.cu -- standalone check of the tcgen05 building blocks the tensor-core path uses:
//   A (weights)      [128 x 32] tf32, K-major, SWIZZLE_128B canonical layout (16 KB tile)
//   B (activations)  [32 k x 32 n] tf32, MN-major, SWIZZLE_128B (rows of 32 floats, chunk ^= k & 7), 4 KB tile
//   D in TMEM [128 lanes x 32 columns] fp32, kind::tf32, M=128 N=32 K=8 per instruction
// Reports (a) max |D - ref| for refs built from TRUNCATED and from RNA-rounded inputs (how does the tensor core
// read fp32 containers?), (b) the 3xTF32 (hi/lo split) error against fp64.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

// smem: A tiles [ntile][16 KB] | B tiles [ntile][4 KB]; each "tile" is one K=32 block; D += sum over tiles
__global__ void __launch_bounds__(128, 1) umma_test(const float* a_tiles, const float* b_tiles, int ntile, float* d_out, uint32_t idesc, int b_kmajor, int mode, uint32_t* dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = reinterpret_cast<float*>(smem + (size_t)ntile * 16384);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < ntile * 4096; i += blockDim.x) sa[i] = a_tiles[i];
    for (int i = threadIdx.x; i < ntile * 1024; i += blockDim.x) sb[i] = b_tiles[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tmem_base)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) dbg[0] = tm;
    if (mode == 1) {   // TMEM st/ld round trip only: lane*1000 + column
        const int w = threadIdx.x >> 5;
        for (int j = 0; j < 32; j++) {
            const uint32_t v = __float_as_uint((float)(threadIdx.x * 1000 + j));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((uint32_t)(w * 32) << 16) + j), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0 && mode == 0) {
        for (int t = 0; t < ntile; t++) {
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t ad = make_desc(su32(sa + t * 4096) + ks * 32, 16, 1024);
                const uint64_t bd = b_kmajor == 1 ? make_desc(su32(sb + t * 1024) + ks * 32, 16, 1024)
                                  : b_kmajor == 2 ? make_desc(su32(sb + t * 1024) + ks * 1024, 4096, 512, 1)   // MN-major, 128B swizzle with 32B base
                                                  : make_desc(su32(sb + t * 1024) + ks * 1024, 4096, 1024);
                const uint32_t acc = (t | ks) ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm),
                    "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    }
    if (threadIdx.x == 0 && mode == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&bar)) : "memory");
    // everyone waits for the MMAs
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; j++) d_out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

// ---- host layouts
static void pack_a(const float* A /*[128][32]*/, float* tile /*4096 floats*/) {   // K-major SW128: row r at (r/8)*1024 + (r%8)*128 B, 16B chunk c ^= r%8
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < 32; k++) {
            const int chunk = (k / 4) ^ (r % 8);
            tile[(r / 8) * 256 + (r % 8) * 32 + chunk * 4 + (k % 4)] = A[r * 32 + k];
        }
}
static void pack_b(const float* B /*[32 k][32 n]*/, float* tile /*1024 floats*/) {  // MN-major SW128: k row at (k/8)*1024 + (k%8)*128 B, chunk ^= k%8
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) {
            const int chunk = (n / 4) ^ (k % 8);
            tile[(k / 8) * 256 + (k % 8) * 32 + chunk * 4 + (n % 4)] = B[k * 32 + n];
        }
}
static void pack_b_kmajor(const float* B /*[32 k][32 n]*/, float* tile) {   // B^T rows n: (n/8)*1024 + (n%8)*128, chunk (k/4) ^ n%8
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) tile[(n / 8) * 256 + (n % 8) * 32 + (((k / 4) ^ (n % 8)) * 4) + (k % 4)] = B[k * 32 + n];
}
static void pack_b_mn32(const float* B /*[32 k][32 n]*/, float* tile) {   // MN-major SW128_BASE32B: row k at k*128 B, 32B chunk (n/8) ^ (k%4)
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) tile[k * 32 + (((n / 8) ^ (k % 4)) * 8) + (n % 8)] = B[k * 32 + n];
}
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static int g_bk = 0, g_mode = 0;
static std::vector<float> run(const std::vector<std::vector<float>>& As, const std::vector<std::vector<float>>& Bs) {
    const int nt = (int)As.size();
    std::vector<float> at(nt * 4096), bt(nt * 1024), d(128 * 32);
    for (int t = 0; t < nt; t++) { pack_a(As[t].data(), at.data() + t * 4096); if (g_bk == 1) pack_b_kmajor(Bs[t].data(), bt.data() + t * 1024); else if (g_bk == 2) pack_b_mn32(Bs[t].data(), bt.data() + t * 1024); else pack_b(Bs[t].data(), bt.data() + t * 1024); }
    float *da, *db, *dd;
    cudaMalloc(&da, at.size() * 4); cudaMalloc(&db, bt.size() * 4); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, at.data(), at.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)nt * (16384 + 4096);
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint32_t* dbg; cudaMalloc(&dbg, 16); cudaMemset(dbg, 0xff, 16);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((g_bk == 1 ? 0u : 1u) << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    umma_test<<<1, 128, smem>>>(da, db, nt, dd, idesc, g_bk, g_mode, dbg);
    uint32_t hd[4]; cudaMemcpy(hd, dbg, 16, cudaMemcpyDeviceToHost); printf("  [tmem_base=0x%08x idesc=0x%08x b_kmajor=%d mode=%d]\n", hd[0], idesc, g_bk, g_mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return d;
}

int main() {
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> A(128 * 32), B(32 * 32);
    for (auto& v : A) v = nd(rng);
    for (auto& v : B) v = nd(rng);
    g_mode = 1; { auto r = run({A}, {B}); printf("st/ld roundtrip: D[0][0]=%g D[1][2]=%g D[37][5]=%g D[127][31]=%g (want 0, 1002, 37005, 127031)\n", r[0], r[34], r[37 * 32 + 5], r[127 * 32 + 31]); }
    g_mode = 0;
  for (g_bk = 2; g_bk >= 1; g_bk--) {
    // (a) raw fp32 containers in, single tile
    auto d = run({A}, {B});
    double e_tr = 0, e_rn = 0, e_ex = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 32; n++) {
            double tr = 0, rn = 0, ex = 0;
            for (int k = 0; k < 32; k++) {
                tr += (double)trunc_tf32(A[m * 32 + k]) * trunc_tf32(B[k * 32 + n]);
                rn += (double)rna_tf32(A[m * 32 + k]) * rna_tf32(B[k * 32 + n]);
                ex += (double)A[m * 32 + k] * B[k * 32 + n];
            }
            e_tr = fmax(e_tr, fabs(d[m * 32 + n] - tr)); e_rn = fmax(e_rn, fabs(d[m * 32 + n] - rn)); e_ex = fmax(e_ex, fabs(d[m * 32 + n] - ex));
        }
    printf("raw fp32 inputs: max|D-ref| truncated-inputs ref %.3e, rna-inputs ref %.3e, exact ref %.3e  (D[0][0]=%f D[5][7]=%f)\n", e_tr, e_rn, e_ex, d[0], d[5 * 32 + 7]);
    // (b) 3xTF32: tiles (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) accumulated in one launch
    std::vector<float> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    for (size_t i = 0; i < A.size(); i++) { Ah[i] = rna_tf32(A[i]); Al[i] = rna_tf32(A[i] - Ah[i]); }
    for (size_t i = 0; i < B.size(); i++) { Bh[i] = rna_tf32(B[i]); Bl[i] = rna_tf32(B[i] - Bh[i]); }
    auto d3 = run({Ah, Al, Ah}, {Bh, Bh, Bl});
    double e3 = 0, e32 = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 32; n++) {
            double ex = 0; float f32 = 0.f;
            for (int k = 0; k < 32; k++) { ex += (double)A[m * 32 + k] * B[k * 32 + n]; f32 = fmaf(A[m * 32 + k], B[k * 32 + n], f32); }
            e3 = fmax(e3, fabs(d3[m * 32 + n] - ex)); e32 = fmax(e32, fabs(f32 - ex));
        }
    printf("3xTF32 (hi*hi + lo*hi + hi*lo): max|D-exact| %.3e   (plain fp32 fmaf chain: %.3e)\n", e3, e32);
  }
    return 0;
}
