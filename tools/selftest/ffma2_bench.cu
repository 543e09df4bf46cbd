// ffma2_bench.cu — microbenchmark behind the FFMA2 decision (DESIGN.md "Kernels"): issue cost of the register-tiled
// 1x1 inner loop on sm_100a with scalar FFMA (8 ch x 4 px per thread, 3 LDS.128 per k) against packed
// fma.rn.f32x2 (FFMA2; weights stored duplicated so a 64-bit register pair holds (w, w)), and the raw FMA rate of both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// raw: 32 independent accumulators, no memory
__global__ void __launch_bounds__(256) raw_ffma(float* out, int iters, float a, float b) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
    float w[8] = {a, a + 1, a + 2, a + 3, a + 4, a + 5, a + 6, a + 7};
    float x[4] = {b, b + 1, b + 2, b + 3};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[n * 4 + i] = fmaf(w[n], x[i], acc[n * 4 + i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) raw_ffma2(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 w[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) w[n] = make_float2(a + n, a + n);
    float2 x[2] = {make_float2(b, b + 1), make_float2(b + 2, b + 3)};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int i = 0; i < 2; ++i) acc[n * 2 + i] = __ffma2_rn(w[n], x[i], acc[n * 2 + i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// tile: the pw_accum inner loop — X [K][256 px] and W [K][8] (or duplicated [K][16]) in shared memory
constexpr int K = 96, PIX = 1024;
__global__ void __launch_bounds__(256) tile_ffma(const float* __restrict__ gx, const float* __restrict__ gw, float* out, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm;                // [K][PIX]
    float* W = sm + K * PIX / 4;  // dummy offset, replaced below
    X = sm; W = sm + K * 256;
    for (int i = threadIdx.x; i < K * 256; i += 256) X[i] = gx[i];
    for (int i = threadIdx.x; i < K * 8; i += 256) W[i] = gw[i];
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
    const float* xp = X + (threadIdx.x & 63) * 4;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            const float4 xv = ld4(xp + k * 256);
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int n4 = 0; n4 < 2; ++n4) {
                const float4 w = ld4(W + k * 8 + n4 * 4);
                const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[n4 * 4 + q][i] = fmaf(w4[q], x4[i], acc[n4 * 4 + q][i]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += acc[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) tile_ffma2(const float* __restrict__ gx, const float* __restrict__ gw, float* out, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm;
    float* W = sm + K * 256;      // [K][8][2] duplicated
    for (int i = threadIdx.x; i < K * 256; i += 256) X[i] = gx[i];
    for (int i = threadIdx.x; i < K * 16; i += 256) W[i] = gw[i >> 1];
    __syncthreads();
    float2 acc[8][2];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[n][i] = make_float2(0.f, 0.f);
    const float* xp = X + (threadIdx.x & 63) * 4;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            const float4 xv = ld4(xp + k * 256);
            const float2 x2[2] = {make_float2(xv.x, xv.y), make_float2(xv.z, xv.w)};
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                const float4 w = ld4(W + k * 16 + n2 * 4);      // (w0, w0, w1, w1)
                const float2 wa = make_float2(w.x, w.y), wb = make_float2(w.z, w.w);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    acc[n2 * 2][i] = __ffma2_rn(wa, x2[i], acc[n2 * 2][i]);
                    acc[n2 * 2 + 1][i] = __ffma2_rn(wb, x2[i], acc[n2 * 2 + 1][i]);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 2; ++i) s += acc[n][i].x + acc[n][i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// pairs along channels: (w[n], w[n+1]) natural, x duplicated with register moves
__global__ void __launch_bounds__(256) tile_ffma2_chpair(const float* __restrict__ gx, const float* __restrict__ gw, float* out, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm;
    float* W = sm + K * 256;
    for (int i = threadIdx.x; i < K * 256; i += 256) X[i] = gx[i];
    for (int i = threadIdx.x; i < K * 8; i += 256) W[i] = gw[i];
    __syncthreads();
    float2 acc[4][4];       // [channel pair][pixel]
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[n][i] = make_float2(0.f, 0.f);
    const float* xp = X + (threadIdx.x & 63) * 4;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            const float4 xv = ld4(xp + k * 256);
            const float2 xd[4] = {make_float2(xv.x, xv.x), make_float2(xv.y, xv.y), make_float2(xv.z, xv.z), make_float2(xv.w, xv.w)};
#pragma unroll
            for (int n4 = 0; n4 < 2; ++n4) {
                const float4 w = ld4(W + k * 8 + n4 * 4);
                const float2 wa = make_float2(w.x, w.y), wb = make_float2(w.z, w.w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[n4 * 2][i] = __ffma2_rn(wa, xd[i], acc[n4 * 2][i]);
                    acc[n4 * 2 + 1][i] = __ffma2_rn(wb, xd[i], acc[n4 * 2 + 1][i]);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += acc[n][i].x + acc[n][i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s  SMs %d  clock %d MHz\n", p.name, sms, clk_khz / 1000);
    float *out, *gx, *gw;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaMalloc(&gx, sizeof(float) * K * 256);
    cudaMalloc(&gw, sizeof(float) * K * 16);
    cudaMemset(gx, 0, sizeof(float) * K * 256);
    cudaMemset(gw, 0, sizeof(float) * K * 16);
    const int smem = (K * 256 + K * 16) * 4;
    cudaFuncSetAttribute(tile_ffma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(tile_ffma2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(tile_ffma2_chpair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const double hz = clk_khz * 1e3;
    for (int bps : {1, 2, 4}) {
        const int grid = sms * bps, iters = 20000;
        const double fma_raw = (double)grid * 256 * 32 * iters;
        float t1 = time_ms([&] { raw_ffma<<<grid, 256>>>(out, iters, 1.f, 2.f); });
        float t2 = time_ms([&] { raw_ffma2<<<grid, 256>>>(out, iters, 1.f, 2.f); });
        printf("raw   blocks/SM %d: FFMA %.1f FMA/clk/SM (%.1f TFLOP/s)   FFMA2 %.1f FMA/clk/SM (%.1f TFLOP/s)\n", bps,
               fma_raw / (t1 * 1e-3 * hz * sms), 2 * fma_raw / (t1 * 1e-3) * 1e-12, fma_raw / (t2 * 1e-3 * hz * sms), 2 * fma_raw / (t2 * 1e-3) * 1e-12);
    }
    for (int bps : {1, 2}) {
        const int grid = sms * bps, reps = 400;
        const double fma = (double)grid * 256 * 32 * K * reps;
        float t1 = time_ms([&] { tile_ffma<<<grid, 256, smem>>>(gx, gw, out, reps); });
        float t2 = time_ms([&] { tile_ffma2<<<grid, 256, smem>>>(gx, gw, out, reps); });
        float t3 = time_ms([&] { tile_ffma2_chpair<<<grid, 256, smem>>>(gx, gw, out, reps); });
        printf("tile  blocks/SM %d: FFMA %.1f   FFMA2(px pairs, dup w) %.1f   FFMA2(ch pairs, dup x) %.1f   FMA/clk/SM\n", bps,
               fma / (t1 * 1e-3 * hz * sms), fma / (t2 * 1e-3 * hz * sms), fma / (t3 * 1e-3 * hz * sms));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
