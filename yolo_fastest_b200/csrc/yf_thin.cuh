// yf_thin.cuh — register-resident inverted-residual kernel for the THIN blocks at the top of the network
// (res1_1: 4 -> 8 -> 4 channels at half resolution, yolo_fastest.py:52-66,83,155), where the generic engine of
// yf_kernels.cuh spends its time on barriers and shared-memory round trips rather than on arithmetic.
//
// A thread owns a strip of 4 output columns x RH output rows and ALL output channels of it (accumulators in registers).
// It walks the mid channels in groups of MPAR; for each group it slides down the RH + 2 halo rows of its strip:
//   e   = relu(W1 . x + b1) for the 6 halo columns of the row (zero outside the image: the depthwise conv pads ITS input)
//   d   = relu(dw3x3(e rows r-2..r) + bd)            sliding window of three e rows in registers
//   acc += W2[m] (x) d                                1x1 projection, outer product into the register tile
// The expanded activation never leaves the register file; only the block INPUT tile lives in shared memory (cp.async, double
// buffered across the tiles of a persistent CTA) and is re-read once per mid-channel group. The price is recomputing the
// expand on the strip's own halo (6/4 columns x (RH+2)/RH rows), cheap for CIN = 4; there is no block-level barrier inside
// a tile. Output = acc + b2 (+ x, the residual, yolo_fastest.py:65), 128-bit stores.
//
// Packed weights (floats): CMID x { [W1: CIN][b1][Wd: 9][bd][W2: COUT] padded to WM }, then [b2: COUT].
#pragma once
#include "yf_kernels.cuh"

namespace yf {

template <int CIN_, int CMID_, int COUT_, int TH_, int TW_, int RH_, int MPAR_, int MINB_, bool RES_>
struct ThinCfg {
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, TH = TH_, TW = TW_, RH = RH_, MPAR = MPAR_, MINB = MINB_;
    static constexpr bool RES = RES_;
    static constexpr int NSTRIP = TW / 4, NSEG = TH / RH, NT = NSTRIP * NSEG;
    static constexpr int XROWS = TH + 2, XW = TW + 8;                 // tile column c sits at index c + 4: halo columns at 3 and TW + 4
    static constexpr int XS1 = rup(CIN * XROWS * XW, 32);
    static constexpr int WM = rup(CIN + COUT + 11, 4);                // floats per mid channel
    static constexpr int OFF_W1 = 0, OFF_B1 = CIN, OFF_WD = CIN + 1, OFF_BD = CIN + 10, OFF_W2 = CIN + 11;
    static constexpr int OFF_B2 = CMID * WM;
    static constexpr int WFLOATS = rup(OFF_B2 + COUT, 4);
    static constexpr int SMEM_FLOATS = 2 * XS1 + rup(WFLOATS, 32);
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static_assert(TW % 4 == 0 && TH % RH == 0 && NT % 32 == 0 && CMID % MPAR == 0, "bad thin tiling");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert(SMEM_BYTES <= 227 * 1024, "thin tile too large");
};

template <class C, bool BORDER>
__device__ __forceinline__ void thin_tile(const float* __restrict__ Xs, const float* __restrict__ Ws, float (&acc)[C::RH][C::COUT][4],
                                          int g, int seg, uint32_t rowmask, uint32_t colmask) {
    constexpr int RH = C::RH, CIN = C::CIN, COUT = C::COUT, MPAR = C::MPAR;
    const float* xt = Xs + (seg * RH) * C::XW + 4 * g + 3;           // halo row 0, halo column 0 of this strip (channel 0)
#pragma unroll 1
    for (int m0 = 0; m0 < C::CMID; m0 += MPAR) {
        float wv[MPAR][C::WM];
#pragma unroll
        for (int p = 0; p < MPAR; ++p)
#pragma unroll
            for (int q = 0; q < C::WM / 4; ++q) {
                const float4 t = ld4(Ws + (m0 + p) * C::WM + 4 * q);
                wv[p][4 * q] = t.x; wv[p][4 * q + 1] = t.y; wv[p][4 * q + 2] = t.z; wv[p][4 * q + 3] = t.w;
            }
        float win[MPAR][3][6];
#pragma unroll
        for (int p = 0; p < MPAR; ++p)
#pragma unroll
            for (int j = 0; j < 6; ++j) win[p][1][j] = win[p][2][j] = 0.f;
#pragma unroll
        for (int r = 0; r < RH + 2; ++r) {
            float e[MPAR][6];
#pragma unroll
            for (int p = 0; p < MPAR; ++p)
#pragma unroll
                for (int j = 0; j < 6; ++j) e[p][j] = wv[p][C::OFF_B1];
#pragma unroll
            for (int k = 0; k < CIN; ++k) {
                const float* xrow = xt + (k * C::XROWS + r) * C::XW;
                const float xl = xrow[0];
                const float4 xm = ld4(xrow + 1);
                const float xr = xrow[5];
                const float xv[6] = {xl, xm.x, xm.y, xm.z, xm.w, xr};
#pragma unroll
                for (int p = 0; p < MPAR; ++p)
#pragma unroll
                    for (int j = 0; j < 6; ++j) e[p][j] = fmaf(wv[p][C::OFF_W1 + k], xv[j], e[p][j]);
            }
#pragma unroll
            for (int p = 0; p < MPAR; ++p)
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    float v = fmaxf(e[p][j], 0.f);
                    if (BORDER) v = (((rowmask >> r) & 1u) && ((colmask >> j) & 1u)) ? v : 0.f;
                    win[p][0][j] = win[p][1][j];
                    win[p][1][j] = win[p][2][j];
                    win[p][2][j] = v;
                }
            if (r >= 2) {
#pragma unroll
                for (int p = 0; p < MPAR; ++p) {
                    float d[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = wv[p][C::OFF_BD];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int i = 0; i < 4; ++i) d[i] = fmaf(wv[p][C::OFF_WD + dy * 3 + dx], win[p][dy][i + dx], d[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
#pragma unroll
                    for (int n = 0; n < COUT; ++n)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[r - 2][n][i] = fmaf(wv[p][C::OFF_W2 + n], d[i], acc[r - 2][n][i]);
                }
            }
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
thin_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
            int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NT = C::NT, RH = C::RH;
    extern __shared__ __align__(128) float smem[];
    float* Xs0 = smem;
    float* Ws = smem + 2 * C::XS1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid % C::NSTRIP, seg = tid / C::NSTRIP;

    auto origin = [&](int tile, int& b, int& oy0, int& ox0) {
        const int tx = tile % tiles_x;
        const int r = tile / tiles_x;
        oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW; b = r / tiles_y;
    };
    // halo tile [CIN][TH + 2][TW + 2] of image b -> dst (column j of the halo at index j + 3), zero outside the image
    auto stage_tile = [&](int tile, float* dst) {
        int b, oy0, ox0;
        origin(tile, b, oy0, ox0);
        const float* src = x + (size_t)b * C::CIN * H * W;
        const bool vec = (W & 3) == 0;             // 16-byte copies for the tile's own columns (ox0 and the smem index are 16B aligned)
        for (int row = warp; row < C::CIN * C::XROWS; row += NT / 32) {
            const int k = row / C::XROWS, r = row - k * C::XROWS;
            const int gy = oy0 - 1 + r;
            const bool rowok = (unsigned)gy < (unsigned)H;
            const float* srow = src + ((size_t)k * H + (rowok ? gy : 0)) * W;
            float* drow = dst + row * C::XW + 3;      // halo column j at drow[j]; tile column c at drow[c + 1]
            if (vec) {
#pragma unroll
                for (int q0 = 0; q0 < C::TW / 4; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < C::TW / 4) {
                        const int gx = ox0 + 4 * q;
                        const bool ok = rowok && gx < W;
                        cp_async16(drow + 1 + 4 * q, ok ? srow + gx : src, ok ? 16 : 0);
                    }
                }
                if (lane >= 30) {                      // the two halo columns
                    const int j = lane == 30 ? 0 : C::TW + 1;
                    const int gx = ox0 - 1 + j;
                    const bool ok = rowok && (unsigned)gx < (unsigned)W;
                    cp_async4(drow + j, ok ? srow + gx : src, ok ? 4 : 0);
                }
            } else {
#pragma unroll
                for (int j0 = 0; j0 < C::TW + 2; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < C::TW + 2) {
                        const int gx = ox0 - 1 + j;
                        const bool ok = rowok && (unsigned)gx < (unsigned)W;
                        cp_async4(drow + j, ok ? srow + gx : src, ok ? 4 : 0);
                    }
                }
            }
        }
        cp_async_commit();
    };

    int tile = blockIdx.x;
    if (tile < total_tiles) stage_tile(tile, Xs0);
    copy_f4<NT>(Ws, wts, C::WFLOATS);
    for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
        float* Xs = Xs0 + (it & 1) * C::XS1;
        int tb, oy0, ox0;
        origin(tile, tb, oy0, ox0);
        cp_async_wait_all();
        __syncthreads();                   // this tile's input has landed; the other buffer (previous tile) is free
        if (tile + (int)gridDim.x < total_tiles) stage_tile(tile + gridDim.x, Xs0 + ((it + 1) & 1) * C::XS1);

        // which of this strip's halo rows / columns lie inside the image
        uint32_t rowmask = 0, colmask = 0;
#pragma unroll
        for (int r = 0; r < RH + 2; ++r) rowmask |= ((unsigned)(oy0 + seg * RH + r - 1) < (unsigned)H ? 1u : 0u) << r;
#pragma unroll
        for (int j = 0; j < 6; ++j) colmask |= ((unsigned)(ox0 + 4 * g + j - 1) < (unsigned)W ? 1u : 0u) << j;
        const bool border = rowmask != (1u << (RH + 2)) - 1u || colmask != 63u;

        float acc[RH][C::COUT][4];
#pragma unroll
        for (int r = 0; r < RH; ++r)
#pragma unroll
            for (int n = 0; n < C::COUT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[r][n][i] = 0.f;
        if (__any_sync(0xffffffffu, border)) thin_tile<C, true>(Xs, Ws, acc, g, seg, rowmask, colmask);    // warp-uniform choice
        else thin_tile<C, false>(Xs, Ws, acc, g, seg, rowmask, colmask);

        const int gx0 = ox0 + 4 * g;
#pragma unroll
        for (int r = 0; r < RH; ++r) {
            const int gy = oy0 + seg * RH + r;
            if (gy < H && gx0 < W) {
#pragma unroll
                for (int n = 0; n < C::COUT; ++n) {
                    const float b = Ws[C::OFF_B2 + n];
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = acc[r][n][i] + b;
                    if (C::RES) {
                        const float4 xc = ld4(Xs + (n * C::XROWS + seg * RH + r + 1) * C::XW + 4 * g + 4);
                        v[0] += xc.x; v[1] += xc.y; v[2] += xc.z; v[3] += xc.w;
                    }
                    store_px4(y + (((size_t)tb * C::COUT + n) * H + gy) * W, gx0, W, v);
                }
            }
        }
    }
}

}  // namespace yf
