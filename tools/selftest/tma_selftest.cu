// tma_selftest.cu — GPU check of the tensor-tile TMA wrappers of yolo_fastest_b200/csrc/yf_tma.cuh (cp.async.bulk.tensor.4d, UTMALDG):
//   1. halo boxes of a planar fp32 [B][C][H][W] tensor at NEGATIVE start coordinates and hanging over the right / bottom edge: elements
//      outside the tensor must arrive as zeros, everything else bit-exact.  The innermost start coordinate is kept 16-byte aligned:
//      `./tma_selftest box <x> <y> <bw> <bh>` runs ONE box in its own process (a fault poisons the context) and shows that x = 1, -1, 2, 3
//      raise "illegal instruction" on sm_100a while x = 0, 4, -4 and any y work (profiles/r02_tma_selftest.txt);
//   2. the same for a uint8 tensor (the raw network input);
//   3. the per-warp protocol of the warp-streaming kernels (yf_wirb.cuh): every warp owns two buffers and two mbarriers, one elected
//      lane arms the barrier and issues the copy, all lanes wait on the parity, __syncwarp() orders the reads before the refill;
//   4. streaming bandwidth of that protocol with the box shapes of the res1 / res2 kernels over a 335 MB tensor.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_selftest tma_selftest.cu     Run: ./tma_selftest
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../yolo_fastest_b200/csrc/yf_kernels.cuh"
#include "../../yolo_fastest_b200/csrc/yf_tma.cuh"
using namespace yf;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e__), __LINE__); return 1; } } while (0)

// ---- 1/2: one box per block --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_box(const __grid_constant__ CUtensorMap tm, T* out, int n, const int4* coords) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int4 c = coords[blockIdx.x];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, n * (int)sizeof(T));
        tma_load4(sm, &tm, &bar, c.x, c.y, c.z, c.w);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[(size_t)blockIdx.x * n + i] = sm[i];
}

template <typename T>
int check_boxes(const char* what, int B, int C, int H, int W, int bw, int bh, int bc, const std::vector<int4>& coords) {
    std::vector<T> h((size_t)B * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (T)(sizeof(T) == 1 ? (i * 7 + 3) % 251 + 1 : i + 1);      // never zero inside the tensor
    T *d, *o;
    int4* dc;
    const int n = bw * bh * bc;
    CK(cudaMalloc(&d, h.size() * sizeof(T)));
    CK(cudaMalloc(&o, coords.size() * n * sizeof(T)));
    CK(cudaMalloc(&dc, coords.size() * sizeof(int4)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dc, coords.data(), coords.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CUtensorMap tm;
    const int rc = tma_make_map4(&tm, d, (int)sizeof(T), B, C, H, W, bw, bh, bc);
    if (rc) { printf("%s: cuTensorMapEncodeTiled failed (%d)\n", what, rc); return 1; }
    k_box<T><<<(int)coords.size(), 128, n * sizeof(T) + 128>>>(tm, o, n, dc);
    CK(cudaDeviceSynchronize());
    std::vector<T> r(coords.size() * n);
    CK(cudaMemcpy(r.data(), o, r.size() * sizeof(T), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (size_t k = 0; k < coords.size(); ++k) {
        const int4 c = coords[k];
        int badk = 0, zeros = 0;
        for (int cc = 0; cc < bc; ++cc)
            for (int y = 0; y < bh; ++y)
                for (int x = 0; x < bw; ++x) {
                    const int gx = c.x + x, gy = c.y + y, gc = c.z + cc;
                    T want = 0;
                    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gc >= 0 && gc < C) want = h[(((size_t)c.w * C + gc) * H + gy) * W + gx];
                    else ++zeros;
                    if (r[k * n + (cc * bh + y) * bw + x] != want) ++badk;
                }
        printf("  %s box {%d,%d,%d} at (x=%d, y=%d, c=%d, b=%d): %d elements outside the tensor, %d mismatches\n", what, bw, bh, bc, c.x, c.y, c.z, c.w, zeros, badk);
        bad += badk;
    }
    cudaFree(d); cudaFree(o); cudaFree(dc);
    return bad;
}

// ---- 3/4: per-warp double-buffered streaming ------------------------------------------------------------------------------------
// Every warp walks units (image, row band, column strip) of a [B][C][H][W] fp32 tensor; a unit is NCH boxes of RC rows starting at
// (x0 - 4, y0 - 1 + c * RC) — the aligned column left of the 1-pixel halo.  The warp sums everything it received (zeros outside) into one double per warp.
template <int BW, int RC, int C>
__global__ void __launch_bounds__(512, 1)
k_stream(const __grid_constant__ CUtensorMap tm, double* sums, int H, int W, int OW, int R, int nstrips, int nbands, int total_units) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ __align__(8) uint64_t bars[16][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr int BOXF = BW * RC * C;
    float* buf = reinterpret_cast<float*>(sm_raw) + (size_t)warp * 2 * BOXF;
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    mbar_fence_init();
    __syncthreads();
    const int NCH = (R + 2 + RC - 1) / RC;
    const int gw = blockIdx.x * nw + warp, tw = gridDim.x * nw;
    const int my_units = gw < total_units ? (total_units - gw + tw - 1) / tw : 0;
    const int total_chunks = my_units * NCH;
    auto issue = [&](int gi) {
        const int u = gw + (gi / NCH) * tw, c = gi % NCH;
        const int strip = u % nstrips, band = (u / nstrips) % nbands, b = u / (nstrips * nbands);
        if (lane == 0) {
            mbar_expect_tx(&bars[warp][gi & 1], BOXF * 4);
            tma_load4(buf + (gi & 1) * BOXF, &tm, &bars[warp][gi & 1], strip * OW - 4, band * R - 1 + c * RC, 0, b);
        }
    };
    int gi = 0;
    for (; gi < 2 && gi < total_chunks; ++gi) issue(gi);
    double s = 0.0;
    for (int g = 0; g < total_chunks; ++g) {
        mbar_wait(&bars[warp][g & 1], (g >> 1) & 1);
        const float* p = buf + (g & 1) * BOXF;
        float a = 0.f;
        for (int i = lane * 4; i < BOXF; i += 128) { const float4 v = ld4(p + i); a += (v.x + v.y) + (v.z + v.w); }
        s += (double)a;
        __syncwarp();
        if (gi < total_chunks) { issue(gi); ++gi; }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sums[gw] = s;
}

template <int BW, int RC, int C>
int stream_case(const char* what, int B, int H, int W, int OW, int R, int nwarps, bool verify) {
    const size_t n = (size_t)B * C * H * W;
    float* d;
    CK(cudaMalloc(&d, n * 4));
    std::vector<float> h;
    if (verify) {
        h.resize(n);
        for (size_t i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) % 17);     // small integers: the sums are exact in float
        CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    } else {
        CK(cudaMemset(d, 0, n * 4));
    }
    CUtensorMap tm;
    if (tma_make_map4(&tm, d, 4, B, C, H, W, BW, RC, C)) { printf("%s: encode failed\n", what); return 1; }
    const int nstrips = (W + OW - 1) / OW, nbands = (H + R - 1) / R, total = B * nstrips * nbands;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm, tw = grid * nwarps;
    const size_t smem = (size_t)nwarps * 2 * BW * RC * C * 4 + 128;
    CK(cudaFuncSetAttribute(k_stream<BW, RC, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double* ds;
    CK(cudaMalloc(&ds, tw * sizeof(double)));
    CK(cudaMemset(ds, 0, tw * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        if (rep == 2) cudaEventRecord(e0);
        k_stream<BW, RC, C><<<grid, nwarps * 32, smem>>>(tm, ds, H, W, OW, R, nstrips, nbands, total);
    }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    int bad = 0;
    if (verify) {
        std::vector<double> got(tw), want(tw, 0.0);
        CK(cudaMemcpy(got.data(), ds, tw * sizeof(double), cudaMemcpyDeviceToHost));
        const int NCH = (R + 2 + RC - 1) / RC;
        for (int u = 0; u < total; ++u) {
            const int strip = u % nstrips, band = (u / nstrips) % nbands, b = u / (nstrips * nbands);
            double s = 0.0;
            for (int c = 0; c < C; ++c)
                for (int y = band * R - 1; y < band * R - 1 + NCH * RC; ++y)
                    for (int x = strip * OW - 4; x < strip * OW - 4 + BW; ++x)
                        if (x >= 0 && x < W && y >= 0 && y < H) s += h[(((size_t)b * C + c) * H + y) * W + x];
            want[u % tw] += s;
        }
        for (int i = 0; i < tw; ++i) if (got[i] != want[i]) ++bad;
        printf("  %s: %d units over %d warps, %d warp sums differ\n", what, total, tw, bad);
    } else {
        printf("  %s: [%d][%d][%d][%d] fp32 (%.0f MB), box {%d,%d,%d}, band %d rows, %d warps/SM: %.3f ms = %.0f GB/s of tensor bytes\n", what, B, C, H, W,
               n * 4e-6, BW, RC, C, R, nwarps, ms, n * 4.0 / (ms * 1e-3) / 1e9);
    }
    cudaFree(d); cudaFree(ds);
    return bad;
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc >= 6 && !strcmp(argv[1], "box")) {        // one box per process: ./tma_selftest box <x> <y> <bw> <bh>  (a fault poisons the context)
        const int x = atoi(argv[2]), y = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]);
        bad = check_boxes<float>("fp32", 2, 4, 40, 72, bw, bh, 4, {{x, y, 0, 0}});
        printf(bad ? "FAIL\n" : "PASS\n");
        return bad ? 1 : 0;
    }
    printf("1. fp32 halo boxes (planar [2][4][40][72])\n");
    bad += check_boxes<float>("fp32", 2, 4, 40, 72, 40, 6, 4,
                              {{-4, -1, 0, 0}, {28, -1, 0, 1}, {60, 35, 0, 1}, {4, 5, 0, 0}, {-8, -3, 0, 1}, {36, 17, 0, 0}, {68, 39, 0, 0}, {4, 2, 2, 1}});
    printf("2. uint8 halo boxes (planar [2][1][64][96])\n");
    bad += check_boxes<unsigned char>("uint8", 2, 1, 64, 96, 96, 13, 1, {{-16, -3, 0, 0}, {32, 50, 0, 1}, {64, -1, 0, 1}, {16, 60, 0, 0}});
    printf("3. per-warp double-buffered streaming, verified sums\n");
    bad += stream_case<40, 6, 4>("res1 shape", 3, 64, 96, 32, 10, 16, true);
    bad += stream_case<24, 6, 8>("res2 shape", 3, 40, 48, 16, 10, 12, true);
    printf("4. streaming bandwidth\n");
    stream_case<40, 6, 4>("res1 shape", 256, 256, 320, 32, 28, 16, false);
    stream_case<24, 6, 8>("res2 shape", 256, 128, 160, 16, 28, 12, false);
    stream_case<40, 6, 4>("res1 shape, 8 warps", 256, 256, 320, 32, 28, 8, false);
    printf(bad ? "FAIL: %d mismatches\n" : "PASS: tensor-tile TMA loads halo boxes at negative (16-byte aligned) coordinates and over the edges with zero fill (%d mismatches)\n", bad);
    return bad ? 1 : 0;
}
