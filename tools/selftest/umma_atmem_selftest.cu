// Standalone check of tcgen05.mma with the A operand in TENSOR MEMORY (the dense 3x3 group of yf_tcdense2.cuh):
//   D[128 px][N] = A[128 px][K] . B[N][K]^T,  kind::tf32, A written by the threads with tcgen05.st (lane = row m, one 32-bit column per k),
//   B = weights in shared memory (K-major core matrices, no swizzle), D in TMEM, K = 8 per MMA, several MMAs accumulated.
// Prints the maximum deviation from a float64 host evaluation (inputs are pre-rounded to tf32, so the products are exact).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_atmem_selftest umma_atmem_selftest.cu && ./umma_atmem_selftest
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint32_t b_off(int n, int k, int K) {          // [n/8][k/4][n%8][k%4] floats -> byte offset
    return (n >> 3) * (K / 4) * 128 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// D = f32, A = B = tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <int N, int K>
__global__ void __launch_bounds__(128) atmem_test(const float* __restrict__ A /*[128][K]*/, const float* __restrict__ B /*[N][K]*/, float* __restrict__ D /*[128][N]*/) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < N * K; i += 128) *reinterpret_cast<float*>(smem + b_off(i / K, i % K, K)) = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // A: thread = row m = tid (lane quarter = warp), columns ACOL + k
    constexpr int ACOL = 256;
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + ACOL;
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(A[tid * K + k0 + i]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(ta + k0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t db = make_desc(smem_u32(smem), 128, (K / 4) * 128);
        for (int kb = 0; kb < K / 8; ++kb) {
            const uint32_t acc = kb ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                         ::"r"(tmem), "r"(tmem + ACOL + kb * 8), "l"(db + (uint64_t)(kb * 16)), "r"(idesc_tf32(N)), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tmem + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(td + n0) : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + n0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static float tf32r(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }

template <int N, int K>
int run() {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(7);
    for (auto& v : A) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : B) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    atmem_test<N, K><<<1, 128, N * K * 4 + 1024>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * (double)B[n * K + k];
            worst = fmax(worst, fabs(r - (double)D[m * N + n]));
        }
    printf("A in TMEM: M=128 N=%d K=%d  max|D - ref| = %.3e  %s\n", N, K, worst, worst < 1e-5 ? "OK" : "MISMATCH");
    return worst < 1e-5 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<48, 8>();
    bad += run<32, 8>();
    bad += run<48, 24>();
    bad += run<64, 72>();
    return bad;
}
