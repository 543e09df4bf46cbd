// Standalone check of the tcgen05 path this repo plans to use for 1x1 convolutions:
//   D[128 px][N ch] = A[128 px][K] * B[N][K]^T   with  A = activations, MN-major (pixels contiguous), 128B swizzle
//                                                      B = weights, K-major, no swizzle (packed on the host)
// kind::tf32, accumulators in TMEM, one MMA per 8 input channels, optional 3xTF32 split (hi*hi + hi*lo + lo*hi).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- layouts -------------------------------------------------------------------------------------------
// A (MN-major, SWIZZLE_128B): atoms of [8 k][32 m] fp32 = 1 KB; byte offset of element (m, k):
__host__ __device__ inline uint32_t a_off(int m, int k, uint32_t lbo, uint32_t sbo) {
    return (k >> 3) * sbo + (m >> 5) * lbo + (k & 7) * 128 + ((((m & 31) >> 2) ^ (k & 7)) << 4) + (m & 3) * 4;
}
// B (K-major, no swizzle): [n/8][k/4][n%8][k%4]
__host__ __device__ inline uint32_t b_off(int n, int k, int K) {
    return (n >> 3) * (K / 4) * 128 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version 1 (Blackwell)
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}

template <int N, int K, bool SPLIT, int AMODE, bool DO_MMA>
__global__ void __launch_bounds__(128) umma_test(const float* __restrict__ A /*[128][K]*/, const float* __restrict__ B /*[N][K]*/,
                                                 float* __restrict__ D /*[128][N]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int ABYTES = (K / 8) * 4096;                 // 4 atoms along M per 8 k (same size in both layouts: 128*K*4)
    constexpr int BBYTES = N * K * 4;
    unsigned char* sAhi = smem;
    unsigned char* sAlo = sAhi + ABYTES;
    unsigned char* sBhi = sAlo + ABYTES;
    unsigned char* sBlo = sBhi + BBYTES;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K;
        const float v = A[i];
        float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);   // tf32 = top 19 bits
        float lo = v - hi;
        if (!SPLIT) { hi = v; lo = 0.f; }
        const uint32_t off = (AMODE == 0 || AMODE == 3) ? a_off(m, k, 1024, 4096)
                           : AMODE == 1 ? b_off(m, k, K)
                           : AMODE == 4 ? (uint32_t)((k >> 3) * 4096 + (m >> 5) * 1024 + ((k >> 2) & 1) * 512 + (k & 3) * 128 + (((((m & 31) >> 3) ^ (k & 3)) << 5)) + (m & 7) * 4)
                           : (uint32_t)((k >> 3) * 4096 + (m >> 2) * 128 + (k & 7) * 16 + (m & 3) * 4);
        *reinterpret_cast<float*>(sAhi + off) = hi;
        *reinterpret_cast<float*>(sAlo + off) = lo;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float v = B[i];
        float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        float lo = v - hi;
        if (!SPLIT) { hi = v; lo = 0.f; }
        *reinterpret_cast<float*>(sBhi + b_off(n, k, K)) = hi;
        *reinterpret_cast<float*>(sBlo + b_off(n, k, K)) = lo;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) printf("tmem base 0x%x  smem A 0x%x B 0x%x\n", tmem, smem_u32(sAhi), smem_u32(sBhi));
    {   // TMEM store/load sanity: column c of lane l := 1000*l + c, overwritten by the MMA afterwards (accumulate = 0)
        uint32_t w[16];
        for (int c0 = 0; c0 < N; c0 += 16) {
            for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(1000.f * (warp * 32 + lane) + c0 + j);
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
                           "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (!DO_MMA) goto epilogue;

    if (tid == 0) {
        // instruction descriptor: D=f32, A=B=tf32, A MN-major, B K-major, N, M=128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMODE != 1 ? 1u : 0u) << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t accum = 0;
        const int npass = SPLIT ? 3 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const unsigned char* sa = (pass == 2) ? sAlo : sAhi;      // hi*hi, hi*lo, lo*hi
            const unsigned char* sb = (pass == 1) ? sBlo : sBhi;
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t adesc = AMODE == 0 ? make_desc(smem_u32(sa) + kb * 4096, 1024, 4096, 2 /*SWIZZLE_128B*/)
                                     : AMODE == 3 ? make_desc(smem_u32(sa) + kb * 4096, 4096, 1024, 2)
                                     : AMODE == 4 ? make_desc(smem_u32(sa) + kb * 4096, 1024 /*MN atoms*/, 512 /*K atoms of 4 rows*/, 1 /*SWIZZLE_128B_BASE32B*/)
                                     : AMODE == 2 ? make_desc(smem_u32(sa) + kb * 4096, 4096 /*K blocks*/, 128 /*MN blocks of 4*/, 0)
                                                  : make_desc(smem_u32(sa) + kb * 256, 128, (K / 4) * 128, 0);
                const uint64_t bdesc = make_desc(smem_u32(sb) + kb * 256, 128, (K / 4) * 128, 0 /*no swizzle*/);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
                accum = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // everyone waits for the MMAs
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
epilogue:
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (= pixels), 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

template <int N, int K, bool SPLIT, int AMODE = 0, bool DO_MMA = true>
int run(const char* tag) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = 2 * (K / 8) * 4096 + 2 * N * K * 4 + 1024;
    cudaFuncSetAttribute(umma_test<N, K, SPLIT, AMODE, DO_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_test<N, K, SPLIT, AMODE, DO_MMA><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    printf("%s N=%d K=%d split=%d: max|err| %.3e  max|ref| %.3f  D[0][0]=%f D[1][1]=%f D[127][N-1]=%f\n", tag, N, K, (int)SPLIT, maxerr, maxref, D[0], D[N + 1], D[127 * N + N - 1]);
    return 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    switch (which) {
        case 0: return run<32, 16, false>("tf32");
        case 1: return run<32, 16, true>("3xtf32");
        case 2: return run<96, 16, true>("3xtf32");
        case 3: return run<16, 96, true>("3xtf32");
        case 4: return run<48, 224, true>("3xtf32");
        case 5: return run<32, 16, false, 0, false>("tmem st/ld only");
        case 6: return run<32, 16, false, 1, true>("tf32 A K-major noswizzle");
        case 7: return run<32, 16, false, 2, true>("tf32 A MN-major noswizzle");
        case 8: return run<32, 16, false, 3, true>("tf32 A MN-major SW128 lbo/sbo swapped");
        case 11: return run<32, 16, false, 4, true>("tf32 A MN-major SW128_BASE32B");
        case 12: return run<96, 16, true, 4, true>("3xtf32 A MN-major SW128_BASE32B");
        case 13: return run<16, 96, true, 4, true>("3xtf32 A MN-major SW128_BASE32B");
        case 9: return run<32, 8, false, 0, true>("tf32 A MN-major SW128 K=8");
        case 10: return run<32, 8, false, 2, true>("tf32 A MN-major noswizzle K=8");
    }
    return 0;
}
