// yf_tcdense.cuh — the dense group conv1_8 (1x1 4 -> 24, ReLU) -> conv1_9 (dense 3x3 stride 2, 24 -> 24, ReLU) -> conv2_1 (1x1 24 -> 8,
// linear), yolo_fastest.py:86-89,158-160, with the 3x3 as an implicit GEMM on the tensor cores (22% of the network's MACs):
//
//   O9[out px][n] = sum over taps t = (ky, kx) and channels c of  E[c][2 oy + ky - 1][2 ox + kx - 1] . W9[n][c][ky][kx]       K = 9 x 24 = 216
//
// A tile is 8 x 16 output pixels = exactly one 128-row MMA tile. The A operand is written tap by tap by the worker threads (im2col in
// shared memory): thread = (output pixel, kx); for each ky it reads the 4 input channels of its input pixel from the staged x tile,
// evaluates conv1_8 for the 8 channels of the current channel block (4 FMAs each — cheaper than staging E and copying it: an
// MN-major swizzled operand cannot be a shifted window of another array), splits the result into hi/lo (3xTF32) and stores it at
// its pixel's row of the operand block of tap t. One step = one channel block = 9 taps x 8 channels = 9 K-blocks (72 KB with hi and
// lo; double buffered); per step the tensor-core thread issues 9 x { D_hi . [W9hi | W9lo] (N = 64), D_lo . W9hi (N = 32) }; all
// 27 weight blocks (55 KB) are resident. The accumulators are double buffered in TMEM, so the epilogue of tile t (thread = pixel:
// sum of the two column groups + b9, ReLU, then conv2_1 as 192 FMAs in registers, 8 channel stores) overlaps the MMAs of tile t + 1.
// Packed weights (floats): [27 x (64 x 8 K-major): rows 0..23 W9hi, 32..55 W9lo][w8: 24 x 4][b8: 24][b9: 24][w21: 8 x 24][b21: 8].
#pragma once
#include "yf_tcpw.cuh"

#ifndef YF_DSPLIT
#define YF_DSPLIT 1      // 1: one accumulator per channel block (hi.hi chains of 9 MMAs), 0: one per tile (chains of 27)
#endif

namespace yf {

template <int NWW_>
struct DenseTcCfg {
    static constexpr int NWW = NWW_, NTW = NWW * 32, NT = NTW + 32;
    static constexpr int TH = 8, TW = 16, OPIX = TH * TW;
    static constexpr int RH = 2 * TH + 1, RWP = TW + 1, XW = 36;                     // staged x tile [4][RH][XW]: RWP column PAIRS from column 2 ox0 - 2
    static constexpr int XS1 = rup(4 * RH * XW, 32);
    static constexpr int KBLK = 4 * 256;                                            // floats of one K-block (8 channels x 128 pixels)
    static constexpr int DA1 = 9 * KBLK;                                            // one operand chunk (hi or lo): 9 taps
    static constexpr int WRES = 27 * 64 * 8;                                        // resident B operands
    static constexpr int OFF_W8 = WRES, OFF_B8 = OFF_W8 + 96, OFF_B9 = OFF_B8 + 24, OFF_W21 = OFF_B9 + 24, OFF_B21 = OFF_W21 + 192;   // all multiples of 4
    static constexpr int WFLOATS = rup(OFF_B21 + 8, 4);
    static constexpr int TCOLS = 512;                                               // 2 tile buffers x 3 channel-block accumulators x 64 columns
    static constexpr int WPAD = rup(WFLOATS, 32);
    static constexpr int SMEM_FLOATS = 4 * DA1 + WPAD + 2 * XS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static_assert(NTW == 3 * OPIX, "one worker thread per (output pixel, kx)");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
dense_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts,
                int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NTW = C::NTW, NWW = C::NWW;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Dbuf = base;                          // [2 buffers][hi | lo][9 taps][8 ch x 128 px]
    float* Wr = Dbuf + 4 * C::DA1;               // resident weights
    float* Xs0 = Wr + C::WPAD;                    // [2 buffers][4][RH][XW]
    __shared__ __align__(8) uint64_t wres, dfull[2], dfree[2], ofull[2], ofree[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&wres, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], NWW); mbar_init(&dfree[i], 1); mbar_init(&ofull[i], 1); mbar_init(&ofree[i], NWW); }
        mbar_fence_init();
    }
    if (warp == NWW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ntile = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == NWW) {
        // ================= tensor-core warp =================
        if (lane == 0 && ntile > 0) {
            constexpr uint32_t IDESC_A = umma_idesc_tf32(64), IDESC_B = umma_idesc_tf32(32);
            mbar_expect_tx(&wres, C::WFLOATS * 4);
            bulk_load(Wr, wts, C::WFLOATS * 4, &wres);
            const uint64_t dd0 = umma_desc(smem_u32(Dbuf), 1024, 512, 1);
            const uint64_t dw0 = umma_desc(smem_u32(Wr), 128, 256, 0);
            mbar_wait(&wres, 0);
            int d = 0;
            for (int ti = 0; ti < ntile; ++ti) {
                const int ob = ti & 1;
                if (ti >= 2) mbar_wait(&ofree[ob], ((ti >> 1) - 1) & 1);       // the tile that used this accumulator buffer has been read out
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb, ++d) {
                    const int b = d & 1;
                    mbar_wait(&dfull[b], (d >> 1) & 1);
                    tc_fence_after();
                    const uint64_t db = dd0 + (uint64_t)((uint32_t)(b * 2 * C::DA1 * 4) >> 4);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint64_t wb = dw0 + (uint64_t)(((cb * 9 + t) * 512 * 4) >> 4);
                        // one accumulator per channel block: the tensor core accumulates with truncation, so the rounding bias grows
                        // with the length of an accumulation chain — 18 MMAs here instead of 54; the epilogue adds the three in fp32
                        umma_tf32(tmem + ob * 192 + YF_DSPLIT * cb * 64, db + (uint64_t)(t * C::KBLK * 4 / 16), wb, IDESC_A, (t | ((1 - YF_DSPLIT) * cb)) ? 1u : 0u);
                        // the lo . hi correction joins the hi . lo correction (columns 32..63): the main term hi . hi keeps the shortest chain
                        umma_tf32(tmem + ob * 192 + YF_DSPLIT * cb * 64 + 32, db + (uint64_t)((C::DA1 + t * C::KBLK) * 4 / 16), wb, IDESC_B, 1u);
                    }
                    umma_commit(&dfree[b]);
                }
                umma_commit(&ofull[ob]);
            }
        }
    } else {
        // ================= worker warps: thread = (output pixel m, kx) =================
        const int m = tid % C::OPIX, kx = tid / C::OPIX;
        const int oyl = m / C::TW, oxl = m - oyl * C::TW;
        const int tpi = tiles_x * tiles_y;
        const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;
        auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            b = tile / tpi;
            const int t = tile - b * tpi;
            const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
            oy0 = ty * C::TH; ox0 = (t - ty * tiles_x) * C::TW;
        };
        const size_t plane = (size_t)Hin * Win;
        // x tile [4][RH][2 RWP] of image b at input rows 2 oy0 - 1 .., columns 2 ox0 - 2 .. -> dst, zero outside the image (8-byte cp.async)
        auto stage_x = [&](int b, int oy0, int ox0, float* dst) {
            const float* src = x + (size_t)b * 4 * plane;
            const bool even = (Win & 1) == 0;                        // column pairs never straddle the image edge
            for (int idx = tid; idx < 4 * C::RH * C::RWP; idx += NTW) {
                const int k = idx / (C::RH * C::RWP), rem = idx - k * (C::RH * C::RWP);
                const int r = rem / C::RWP, jp = rem - r * C::RWP;
                const int gy = 2 * oy0 - 1 + r, gx = 2 * ox0 - 2 + 2 * jp;
                const bool rowok = (unsigned)gy < (unsigned)Hin;
                float* d = dst + (k * C::RH + r) * C::XW + 2 * jp;
                const float* sp = src + (size_t)k * plane + (size_t)(rowok ? gy : 0) * Win;
                if (even) {
                    const bool ok = rowok && (unsigned)gx < (unsigned)Win;
                    cp_async8(d, ok ? sp + gx : src, ok ? 8 : 0);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool ok = rowok && (unsigned)(gx + e) < (unsigned)Win;
                        cp_async4(d + e, ok ? sp + gx + e : src, ok ? 4 : 0);
                    }
                }
            }
            cp_async_commit();
        };
        int tb = 0, oy0 = 0, ox0 = 0, nb = 0, noy0 = 0, nox0 = 0;
        if (ntile > 0) {
            origin(0, nb, noy0, nox0);
            stage_x(nb, noy0, nox0, Xs0);
        }
        mbar_wait(&wres, 0);                                         // conv1_8 / bias / conv2_1 weights live behind the B operands
        // this thread's operand address pieces: pixel row m inside a K-block
        const int obase = (m >> 5) * 256 + (m & 7), mcx = (m & 31) >> 3;
        // epilogue of tile te (thread = pixel; the three warps that share a TMEM lane quarter split the 8 output channels):
        // O9 -> + b9, ReLU -> conv2_1 (24 -> 8) in registers -> HBM. Runs one step late (after the first step of the next tile), so the
        // tensor core never waits for it and the workers never wait for the last MMAs of a tile.
        auto epilogue = [&](int te, int eb, int ey0, int ex0) {
            const int ob = te & 1;
            mbar_wait(&ofull[ob], (te >> 1) & 1);
            tc_fence_after();
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ob * 192;
            float v[24];
#pragma unroll
            for (int n = 0; n < 24; ++n) v[n] = Wr[C::OFF_B9 + n];
#pragma unroll
            for (int cb = 0; cb < (YF_DSPLIT ? 3 : 1); ++cb) {
                float a0[16], a1[8], l0[16], l1[8];
                tmem_ld16(ta + cb * 64, a0); tmem_ld8(ta + cb * 64 + 16, a1); tmem_ld16(ta + cb * 64 + 32, l0); tmem_ld8(ta + cb * 64 + 48, l1);
#pragma unroll
                for (int n = 0; n < 16; ++n) v[n] += a0[n] + l0[n];
#pragma unroll
                for (int n = 0; n < 8; ++n) v[16 + n] += a1[n] + l1[n];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ofree[ob]);
#pragma unroll
            for (int n = 0; n < 24; ++n) v[n] = fmaxf(v[n], 0.f);
            const int gy = ey0 + oyl, gx = ex0 + oxl;                // this thread's pixel m (every kx group holds all 128 pixels)
            if (gy < Hout && gx < Wout) {
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int j = kx * 3 + jj;                       // channels {0,1,2}, {3,4,5}, {6,7}
                    if (j < 8) {
                        float r = Wr[C::OFF_B21 + j];
#pragma unroll
                        for (int n4 = 0; n4 < 6; ++n4) {
                            const float4 w = ld4(Wr + C::OFF_W21 + j * 24 + 4 * n4);
                            r = fmaf(w.x, v[4 * n4], r); r = fmaf(w.y, v[4 * n4 + 1], r);
                            r = fmaf(w.z, v[4 * n4 + 2], r); r = fmaf(w.w, v[4 * n4 + 3], r);
                        }
                        y[(((size_t)eb * 8 + j) * Hout + gy) * Wout + gx] = r;
                    }
                }
            }
        };
        int d = 0;
        for (int ti = 0; ti < ntile; ++ti) {
            const int pb = tb, poy0 = oy0, pox0 = ox0;               // previous tile (its epilogue is still pending)
            tb = nb; oy0 = noy0; ox0 = nox0;
            const bool have_next = ti + 1 < ntile;
            if (have_next) origin(ti + 1, nb, noy0, nox0);
            const float* Xs = Xs0 + (ti & 1) * C::XS1;
            cp_async_wait_all();
            named_bar_sync<1, NTW>();                                // this tile's x has landed for everyone; the other buffer is free
            if (have_next) stage_x(nb, noy0, nox0, Xs0 + ((ti + 1) & 1) * C::XS1);
            // the three input pixels (ky = 0..2) of this thread: tile coordinates (2 oyl + ky, 2 oxl + kx)
            float xv[3][4];
            bool in[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int r = 2 * oyl + ky, j = 2 * oxl + kx;
                in[ky] = (unsigned)(2 * oy0 - 1 + r) < (unsigned)Hin && (unsigned)(2 * ox0 - 1 + j) < (unsigned)Win;
#pragma unroll
                for (int k = 0; k < 4; ++k) xv[ky][k] = Xs[(k * C::RH + r) * C::XW + j + 1];
            }
#pragma unroll 1
            for (int cb = 0; cb < 3; ++cb, ++d) {
                const int b = d & 1;
                float4 w8r[8];
                float b8r[8];
#pragma unroll
                for (int cl = 0; cl < 8; ++cl) { w8r[cl] = ld4(Wr + C::OFF_W8 + (cb * 8 + cl) * 4); b8r[cl] = Wr[C::OFF_B8 + cb * 8 + cl]; }
                if (d >= 2) mbar_wait(&dfree[b], ((d >> 1) - 1) & 1);
                float* Dh = Dbuf + b * 2 * C::DA1;
#pragma unroll
                for (int cl = 0; cl < 8; ++cl) {
                    const float4 w = w8r[cl];
                    const float bb = b8r[cl];
                    const int oc = obase + ((cl >> 2) & 1) * 128 + (cl & 3) * 32 + ((mcx ^ (cl & 3)) << 3);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        float e = fmaf(w.x, xv[ky][0], bb);
                        e = fmaf(w.y, xv[ky][1], e);
                        e = fmaf(w.z, xv[ky][2], e);
                        e = fmaf(w.w, xv[ky][3], e);
                        e = in[ky] ? fmaxf(e, 0.f) : 0.f;            // conv1_9 zero-pads ITS input, the activation
                        const float hi = tf32_hi(e);
                        const int o = (ky * 3 + kx) * C::KBLK + oc;
                        Dh[o] = hi;
                        Dh[C::DA1 + o] = e - hi;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dfull[b]);
                if (cb == 0 && ti > 0) epilogue(ti - 1, pb, poy0, pox0);
            }
        }
        if (ntile > 0) epilogue(ntile - 1, tb, oy0, ox0);
    }
    __syncthreads();
    if (warp == NWW) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
