// Minimal check of the TMA / mbarrier / bulk-copy wrappers of yf_kernels.cuh on a real GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../yolo_fastest_b200/csrc/yf_kernels.cuh"
using namespace yf;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_bulk(const float* src, float* out, int n) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, n * 4); bulk_load(sm, src, n * 4, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}

__global__ void k_tma(const __grid_constant__ CUtensorMap tm, float* out, int n, int x, int y, int c, int b, int variant) {
    extern __shared__ __align__(128) float sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    float* sm = sm_raw;
    if (variant & 1) sm = (float*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) printf("smem dst offset 0x%x bar 0x%x\n", smem_u32(sm), smem_u32(&bar));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, n * 4);
        if (variant & 2) {
            asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(sm)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(c), "r"(b) : "memory");
        } else {
            tma_load_4d(sm, &tm, &bar, x, y, c, b);
        }
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    printf("variant %d\n", variant);
    const int B = 2, C = 3, H = 8, W = 16;
    std::vector<float> h(B * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMalloc(&o, 4096 * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    // bulk
    k_bulk<<<1, 64, 1024 * 4>>>(d, o, 256);
    CK(cudaDeviceSynchronize());
    std::vector<float> r(4096);
    CK(cudaMemcpy(r.data(), o, 256 * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 256; ++i) bad += r[i] != h[i];
    printf("bulk copy: %s\n", bad ? "MISMATCH" : "ok");
    // tensor
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, C, B};
    cuuint64_t strides[3] = {W * 4, W * H * 4, W * H * C * 4};
    const int bw = 8, bh = 4, bc = 3;
    cuuint32_t box[4] = {bw, bh, bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, (variant & 4) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)cr);
    const int n = bw * bh * bc;
    const int x0 = (variant & 8) ? 0 : -1, y0 = (variant & 8) ? 0 : -1, c0 = 0, b0 = 1;
    k_tma<<<1, 64, 2048 * 4>>>(tm, o, n, x0, y0, c0, b0, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("tma kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    CK(cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost));
    bad = 0;
    for (int c = 0; c < bc; ++c)
        for (int y = 0; y < bh; ++y)
            for (int x = 0; x < bw; ++x) {
                const int gx = x0 + x, gy = y0 + y;
                const float want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0.f : h[((b0 * C + c0 + c) * H + gy) * W + gx];
                bad += r[(c * bh + y) * bw + x] != want;
            }
    printf("tma tile: %s\n", bad ? "MISMATCH" : "ok");
    return bad != 0;
}
