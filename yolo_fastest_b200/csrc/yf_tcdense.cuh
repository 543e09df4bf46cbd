// yf_tcdense.cuh — the dense group conv1_8 (1x1 4 -> 24, ReLU) -> conv1_9 (dense 3x3 stride 2, 24 -> 24, ReLU) -> conv2_1 (1x1 24 -> 8,
// linear), yolo_fastest.py:86-89,158-160, with the 3x3 as an implicit GEMM on the tensor cores (22% of the network's MACs):
//
//   O9[out px][n] = sum over taps t = (ky, kx) and channels c of  E[c][2 oy + ky - 1][2 ox + kx - 1] . W9[n][c][ky][kx]       K = 9 x 24 = 216
//
// No im2col: the A operand of every tap is a SHIFTED WINDOW of one array. E = relu(conv1_8(x)) is computed once per input pixel
// and stored (hi and lo parts, 3xTF32) as [4-channel group][column parity][input row][column / 2][4 channels]. That is the K-major,
// un-swizzled operand layout: a core matrix = 8 consecutive pixels x 16 bytes (4 channels), the K-adjacent core matrix one
// channel-group plane pair further (LBO), the next 8-pixel row group TWO input rows further (SBO) — so with a tile of 16 x 8 output
// pixels (row group = one output row) the operand of tap (ky, kx) is the same array read from start address
// base + parity(kx) plane + ky rows + (kx >> 1) pixels: the descriptor's 16-byte start granularity is exactly one pixel.
// (An MN-major swizzled operand, as used by the other kernels, cannot start at an odd pixel; the first version of this kernel
// therefore copied every tap into its own operand block — 2.25 stores per E value, 21 k instructions per tile — and was worker-bound.)
//
// One step = one block of 8 channels (2 channel groups) of E for the whole input tile, 3 steps per tile, 3-deep ring; per step the
// tensor-core thread issues 9 taps x { E_hi . [W9hi | W9lo] (N = 48), E_lo . W9hi (N = 32, into the correction columns) } into
// the step's own accumulator (chains of 9 MMAs: the tensor core accumulates with truncation); all 27 weight blocks (55 KB) are
// resident. The accumulators of a tile are double buffered in TMEM, so the epilogue of tile t (thread = pixel: sum of the three
// accumulators + b9, ReLU, then conv2_1 as FMAs in registers) runs one step late and overlaps the MMAs of tile t + 1.
// tcgen05 dispatch floor (128 N / 256 cycles per K = 8 MMA): 27 x (24 + 16) = 1080 cycles per tile, a fifth of the ~5.1 k cycles a tile
// takes — the kernel is bound by its producer warps, not by the pipe.
// Packed weights (floats): [27 x (64 x 8 K-major): rows 0..23 W9hi, 24..47 W9lo][w8: 24 x 4][b8: 24][b9: 24][w21 transposed: 24 x 8][b21: 8].
#pragma once
#include "yf_tcpw.cuh"

namespace yf {

#if defined(YF_TC_TRACE) && defined(YF_DENSE_TRACE)
#define DTRACE(tile, ev) do { if (blockIdx.x == 0 && (tile) < 64) g_tc_trace[(tile) * 16 + (ev)] = clock64(); } while (0)
#else
#define DTRACE(tile, ev) do { } while (0)
#endif

// 32 consecutive columns of this thread's lane (load + wait in ONE statement: the registers are only valid after the wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
template <int NWW_>
struct DenseTcCfg {
    static constexpr int NWW = NWW_, NTW = NWW * 32;                                // producer warps / threads
    static constexpr int NEW = 4;                                                   // epilogue warps (one per TMEM lane quarter)
    static constexpr int NT = NTW + NEW * 32 + 32;                                  // + the tensor-core warp
    static constexpr int TH = 16, TW = 8, OPIX = TH * TW;
    static constexpr int RH = 2 * TH + 1, RW = 2 * TW + 1;                          // input rows / columns feeding a tile
    static constexpr int XWP = TW + 1, XW = 2 * XWP + 2;                            // staged x tile [4][RH][XW]: XWP column pairs from column 2 ox0 - 2
    static constexpr int XS1 = rup(4 * RH * XW, 32);
    static constexpr int XH = TW + 1;                                               // E entries per row and parity (even columns: TW + 1, odd: TW)
    static constexpr int ROW = XH * 4;                                              // floats per E row
    static constexpr int PLANE = rup(RH * ROW, 32) + 16;                            // floats per (channel group, parity) plane; 64 B mod 128 B, so the
                                                                                    // two parities of a warp's 128-bit stores fall on different banks
    static constexpr int HL = 4 * PLANE;                                            // hi (or lo) part of a slot: 2 channel groups x 2 parities
    static constexpr int SLOT = 2 * HL, NS = 3;                                     // ring of 3 slots
    static constexpr int WRES = 27 * 64 * 8;                                        // resident B operands
    static constexpr int OFF_W8 = WRES, OFF_B8 = OFF_W8 + 96, OFF_B9 = OFF_B8 + 24, OFF_W21 = OFF_B9 + 24, OFF_B21 = OFF_W21 + 192;   // all multiples of 4
    static constexpr int WFLOATS = rup(OFF_B21 + 8, 4);
    static constexpr int WPAD = rup(WFLOATS, 32);
    static constexpr int NMAIN = 48, OBUF = 3 * NMAIN + 32;                         // columns of a step accumulator (hi.hi | hi.lo) / of a tile buffer
    static constexpr int TCOLS = 512;                                               // 2 tile buffers x (3 step accumulators x 48 + 32 correction columns)
    static constexpr int NPIX_IN = RH * RW;
    static constexpr int NTH = NTW / 2, IPT = cdiv(NPIX_IN, NTH);                   // half of the workers per channel group; input pixels per thread and step
    static constexpr int NSTG = 4 * RH * XWP, SPT = cdiv(NSTG, NTW);                // x-tile staging copies (8 bytes each) per tile / per producer thread
    static constexpr int SMEM_FLOATS = NS * SLOT + WPAD + 2 * XS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static_assert(NEW * 32 == OPIX && (PLANE * 4) % 128 == 64, "epilogue mapping / bank layout");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
dense_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts,
                int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NTW = C::NTW, NWW = C::NWW;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Ebuf = base;                          // [NS slots][hi | lo][2 channel groups][2 parities][RH][XH][4]
    float* Wr = Ebuf + C::NS * C::SLOT;          // resident weights
    float* Xs0 = Wr + C::WPAD;                   // [2 buffers][4][RH][XW]
    __shared__ __align__(8) uint64_t wres, dfull[C::NS], dfree[C::NS], ofull[2], ofree[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&wres, 1);
        for (int i = 0; i < C::NS; ++i) { mbar_init(&dfull[i], NWW); mbar_init(&dfree[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&ofull[i], 1); mbar_init(&ofree[i], C::NEW); }
        mbar_fence_init();
        mbar_expect_tx(&wres, C::WFLOATS * 4);                   // before the dependency wait: weights are not activations
        bulk_load(Wr, wts, C::WFLOATS * 4, &wres);
    }
    if (warp == NWW + C::NEW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
    const uint32_t tmem = tmem_slot;
    const int ntile = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == NWW + C::NEW) {
        // ================= tensor-core warp =================
        if (ntile > 0 && elect_one()) {
            // instruction descriptors: A is K-major here (bit 15 clear), unlike the MN-major operands of the other kernels
            constexpr uint32_t IDESC_A = umma_idesc_tf32(C::NMAIN) & ~(1u << 15), IDESC_B = umma_idesc_tf32(32) & ~(1u << 15);
            // A: core matrices of 8 pixels x 16 B; K-adjacent one = next channel group (2 planes further, LBO), next row group = two input rows (SBO)
            const uint64_t de0 = umma_desc(smem_u32(Ebuf), 2 * C::PLANE * 4, 2 * C::ROW * 4, 0);
            const uint64_t dw0 = umma_desc(smem_u32(Wr), 128, 256, 0);
            mbar_wait(&wres, 0);
            int d = 0;
            for (int ti = 0; ti < ntile; ++ti) {
                const int ob = ti & 1;
                if (ti >= 2) mbar_wait(&ofree[ob], ((ti >> 1) - 1) & 1);       // the tile that used these accumulators has been read out
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb, ++d) {
                    const int sl = d % C::NS;
                    mbar_wait(&dfull[sl], (d / C::NS) & 1);
                    tc_fence_after();
                    DTRACE(ti, 8 + cb);
                    const uint64_t db = de0 + (uint64_t)(((uint32_t)sl * C::SLOT * 4) >> 4);
                    // two independent accumulation chains, interleaved so consecutive MMAs never wait for each other's accumulator:
                    //   main  (one per step, 9 MMAs): E_hi . [W9hi | W9lo]  -> columns cb * 48 + [0, 48)
                    //   corr  (one per tile, 27 MMAs): E_lo . W9hi          -> columns 144 + [0, 32)
                    // (the tensor core accumulates with truncation: the main term gets the short chain, the small correction the long one)
                    const uint32_t acc = tmem + ob * C::OBUF + cb * C::NMAIN, corr = tmem + ob * C::OBUF + 3 * C::NMAIN;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int ky = t / 3, kx = t % 3;
                        const uint64_t wb = dw0 + (uint64_t)(((cb * 9 + t) * 512 * 4) >> 4);
                        const uint64_t ao = (uint64_t)((((kx & 1) * C::PLANE + ky * C::ROW + (kx >> 1) * 4) * 4) >> 4);
                        umma_tf32(acc, db + ao, wb, IDESC_A, t ? 1u : 0u);
                        umma_tf32(corr, db + ao + (uint64_t)((C::HL * 4) >> 4), wb, IDESC_B, (cb | t) ? 1u : 0u);
                    }
                    umma_commit(&dfree[sl]);
                    DTRACE(ti, 11 + cb);
                }
                umma_commit(&ofull[ob]);
            }
        }
    } else {
        const int tpi = tiles_x * tiles_y;
        const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;
        auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            b = tile / tpi;
            const int t = tile - b * tpi;
            const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
            oy0 = ty * C::TH; ox0 = (t - ty * tiles_x) * C::TW;
        };
        if (warp >= NWW) {
            // ================= epilogue warps (thread = pixel): the three step accumulators (main + hi.lo columns) + the tile's lo.hi
            // correction + b9, ReLU, conv2_1 (24 -> 8) as FMAs in registers -> HBM. Runs off the producers' path: one tile behind. ====
            mbar_wait(&wres, 0);
            const int q = warp - NWW, em = q * 32 + lane, eoy = em >> 3, eox = em & 7;
            for (int ti = 0; ti < ntile; ++ti) {
                int eb, ey0, ex0;
                origin(ti, eb, ey0, ex0);
                const int ob = ti & 1;
                mbar_wait(&ofull[ob], (ti >> 1) & 1);
                tc_fence_after();
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + ob * C::OBUF;
                if (em == 0) DTRACE(ti, 5);
                float v[24];
                {
                    uint32_t c[32];
                    tmem_ld32(ta + 3 * C::NMAIN, c);                                          // the tile's lo . hi correction accumulator
#pragma unroll
                    for (int n = 0; n < 24; ++n) v[n] = __uint_as_float(c[n]) + Wr[C::OFF_B9 + n];
                }
#pragma unroll
                for (int cb = 0; cb < 3; ++cb) {
                    uint32_t a[32];
                    float a2[16];                                          // columns [0, 24) E_hi . W9hi, [24, 48) E_hi . W9lo
                    tmem_ld32(ta + cb * C::NMAIN, a);
                    tmem_ld16(ta + cb * C::NMAIN + 32, a2);
#pragma unroll
                    for (int n = 0; n < 24; ++n) v[n] += __uint_as_float(a[n]);
#pragma unroll
                    for (int n = 0; n < 8; ++n) v[n] += __uint_as_float(a[24 + n]);
#pragma unroll
                    for (int n = 0; n < 16; ++n) v[8 + n] += a2[n];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ofree[ob]);
                if (em == 0) DTRACE(ti, 7);
                float ps[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) ps[j] = Wr[C::OFF_B21 + j];
#pragma unroll
                for (int n = 0; n < 24; ++n) {
                    const float vn = fmaxf(v[n], 0.f);
                    const float4 wa = ld4(Wr + C::OFF_W21 + n * 8), wb = ld4(Wr + C::OFF_W21 + n * 8 + 4);      // [n][j]
                    ps[0] = fmaf(wa.x, vn, ps[0]); ps[1] = fmaf(wa.y, vn, ps[1]); ps[2] = fmaf(wa.z, vn, ps[2]); ps[3] = fmaf(wa.w, vn, ps[3]);
                    ps[4] = fmaf(wb.x, vn, ps[4]); ps[5] = fmaf(wb.y, vn, ps[5]); ps[6] = fmaf(wb.z, vn, ps[6]); ps[7] = fmaf(wb.w, vn, ps[7]);
                }
                const int gy = ey0 + eoy, gx = ex0 + eox;
                if (gy < Hout && gx < Wout) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[(((size_t)eb * 8 + j) * Hout + gy) * Wout + gx] = ps[j];
                }
                if (em == 0) DTRACE(ti, 14);
            }
        } else {
            // ================= producer warps =================
            const size_t plane = (size_t)Hin * Win;
            // x tile [4][RH][2 XWP] of image b at input rows 2 oy0 - 1 .., columns 2 ox0 - 2 .. -> dst, zero outside the image (8-byte
            // cp.async); the copies of a thread are the same every tile: (source offset, destination, row, column pair) are precomputed
            int sg_src[C::SPT], sg_dst[C::SPT], sg_rj[C::SPT];
#pragma unroll
            for (int i = 0; i < C::SPT; ++i) {
                const int idx = min(tid + i * NTW, C::NSTG - 1);
                const int k = idx / (C::RH * C::XWP), rem = idx - k * (C::RH * C::XWP);
                const int r = rem / C::XWP, jp = rem - r * C::XWP;
                sg_src[i] = (int)(k * plane) + r * Win + 2 * jp;
                sg_dst[i] = (k * C::RH + r) * C::XW + 2 * jp;
                sg_rj[i] = r | (jp << 8);
            }
            const bool even = (Win & 1) == 0;                        // column pairs never straddle the image edge
            auto stage_x = [&](int b, int oy0, int ox0, float* dst) {
                const float* src = x + (size_t)b * 4 * plane;
                const int o0 = (2 * oy0 - 1) * Win + 2 * ox0 - 2;
#pragma unroll
                for (int i = 0; i < C::SPT; ++i) {
                    if (tid + i * NTW < C::NSTG) {
                        const int gy = 2 * oy0 - 1 + (sg_rj[i] & 255), gx = 2 * ox0 - 2 + 2 * (sg_rj[i] >> 8);
                        const bool rowok = (unsigned)gy < (unsigned)Hin;
                        if (even) {
                            const bool ok = rowok && (unsigned)gx < (unsigned)Win;
                            cp_async8(dst + sg_dst[i], ok ? src + o0 + sg_src[i] : src, ok ? 8 : 0);
                        } else {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const bool ok = rowok && (unsigned)(gx + e) < (unsigned)Win;
                                cp_async4(dst + sg_dst[i] + e, ok ? src + o0 + sg_src[i] + e : src, ok ? 4 : 0);
                            }
                        }
                    }
                }
                cp_async_commit();
            };
            int oy0 = 0, ox0 = 0, nb = 0, noy0 = 0, nox0 = 0;
            if (ntile > 0) {
                origin(0, nb, noy0, nox0);
                stage_x(nb, noy0, nox0, Xs0);
            }
            mbar_wait(&wres, 0);                                     // the conv1_8 weights live behind the B operands
            // this thread's producer items (fixed for the whole kernel): channel group cg = tid / NTH, input pixels pi = (tid % NTH) + i * NTH
            // -> E entry (cg, parity j & 1, r, j >> 1)
            const int cg = tid / C::NTH, tl = tid - cg * C::NTH;
            int it_x[C::IPT], it_e[C::IPT], it_rj[C::IPT];
#pragma unroll
            for (int i = 0; i < C::IPT; ++i) {
                const int pi = min(tl + i * C::NTH, C::NPIX_IN - 1);
                const int r = pi / C::RW, j = pi - r * C::RW;
                it_rj[i] = r | (j << 8);
                it_x[i] = r * C::XW + j + 1;                         // staged column index: the tile's column j sits at j + 1
                it_e[i] = (cg * 2 + (j & 1)) * C::PLANE + r * C::ROW + (j >> 1) * 4;
            }
            int d = 0;
            for (int ti = 0; ti < ntile; ++ti) {
                oy0 = noy0; ox0 = nox0;
                const bool have_next = ti + 1 < ntile;
                if (have_next) origin(ti + 1, nb, noy0, nox0);
                const float* Xs = Xs0 + (ti & 1) * C::XS1;
                if (tid == 0) DTRACE(ti, 0);
                cp_async_wait_all();
                named_bar_sync<1, NTW>();                            // this tile's x has landed for everyone; the other buffer is free
                if (tid == 0) DTRACE(ti, 1);
                if (have_next) stage_x(nb, noy0, nox0, Xs0 + ((ti + 1) & 1) * C::XS1);
                // this thread's input pixels: the 4 input channels and whether the pixel lies inside the image
                float xv[C::IPT][4];
                bool in[C::IPT];
#pragma unroll
                for (int i = 0; i < C::IPT; ++i) {
                    in[i] = (unsigned)(2 * oy0 - 1 + (it_rj[i] & 255)) < (unsigned)Hin && (unsigned)(2 * ox0 - 1 + (it_rj[i] >> 8)) < (unsigned)Win;
#pragma unroll
                    for (int k = 0; k < 4; ++k) xv[i][k] = Xs[k * C::RH * C::XW + it_x[i]];
                }
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb, ++d) {
                    const int sl = d % C::NS;
                    float4 w8k[4];                                   // conv1_8 weights of this thread's 4 channels in this step: w8k[k] = the 4 channels' k-th weight
#pragma unroll
                    for (int k = 0; k < 4; ++k) w8k[k] = ld4(Wr + C::OFF_W8 + (cb * 2 + cg) * 16 + k * 4);
                    const float4 b8v = ld4(Wr + C::OFF_B8 + cb * 8 + cg * 4);
                    if (d >= C::NS) mbar_wait(&dfree[sl], ((d / C::NS) - 1) & 1);
                    if (tid == 0 && cb == 0) DTRACE(ti, 6);
                    float* Eh = Ebuf + sl * C::SLOT;
#pragma unroll
                    for (int i = 0; i < C::IPT; ++i) {
                        if (tl + i * C::NTH < C::NPIX_IN) {
                            float hi[4], lo[4];
                            float ev[4] = {b8v.x, b8v.y, b8v.z, b8v.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {             // packed FMAs over channel pairs (fmaf rounding per lane), k ascending like the scalar loop
                                ffma2(w8k[k].x, w8k[k].y, xv[i][k], ev[0], ev[1]);
                                ffma2(w8k[k].z, w8k[k].w, xv[i][k], ev[2], ev[3]);
                            }
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const float e = in[i] ? fmaxf(ev[cc], 0.f) : 0.f;      // conv1_9 zero-pads ITS input, the activation
                                hi[cc] = tf32_hi(e);
                                lo[cc] = e - hi[cc];
                            }
                            st4(Eh + it_e[i], make_float4(hi[0], hi[1], hi[2], hi[3]));
                            st4(Eh + C::HL + it_e[i], make_float4(lo[0], lo[1], lo[2], lo[3]));
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dfull[sl]);
                    if (tid == 0) DTRACE(ti, 2 + cb);
                }
            }
        }
    }
    __syncthreads();
    if (warp == NWW + C::NEW) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
