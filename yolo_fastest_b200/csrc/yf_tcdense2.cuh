// yf_tcdense2.cuh — the dense group conv1_8 -> conv1_9 (dense 3x3 stride 2) -> conv2_1 with the A operand of the implicit GEMM in
// TENSOR MEMORY (yolo_fastest.py:86-89,158-160; same arithmetic and the same packed weights as yf_tcdense.cuh).
//
// yf_tcdense.cuh keeps E = relu(conv1_8(x)) in shared memory and lets every tap read it through a shifted descriptor. Measured
// (profiles/r02_dense_phase_trace.txt, ncu): that kernel is bound by shared-memory bandwidth — 9 taps x 3 channel blocks x (E_hi + E_lo)
// re-read 216 KB of operand per 128-pixel tile, next to 108 KB of producer stores. Here shared memory carries only the weights:
//
//   tcgen05.mma  D[out px][n] += A[out px][k] . B[n][k]     A in TMEM: lane = output pixel, one 32-bit column per k (validated by
//                                                           tools/selftest/umma_atmem_selftest.cu), B = a resident weight block
//
// A producer thread owns ONE output pixel (its TMEM lane) and one kernel row ky: per step (8 channels) it evaluates conv1_8 + ReLU at
// the three input pixels (2 oy + ky - 1, 2 ox + kx - 1) from registers (the 4 input channels of its 3 pixels are loaded once per
// tile), splits the 24 values into hi | lo and writes them with tcgen05.st into the step's A buffer: columns (ky * 3 + kx) * 8 + c.
// E is recomputed 2.25x (every input pixel feeds 2.25 taps on average) — 4 FMAs per value, cheaper than any exchange. The small
// weights (conv1_8, biases, conv2_1) are kernel parameters: constant-bank operands, no loads.
//
// TMEM columns: 2 A buffers x (72 hi + 72 lo) | 3 step accumulators x 48 (E_hi . [W9hi | W9lo]) + 32 (E_lo . W9hi): 464 of 512.
// Warps: 12 producers = 4 lane quarters x 3 kernel rows, 4 epilogue warps (thread = pixel, as in yf_tcdense.cuh), the tensor-core warp.
#pragma once
#include <type_traits>
#include "yf_tcdense.cuh"
#include "yf_tma.cuh"

namespace yf {

struct DenseSmall {                        // conv1_8 transposed [k][c] (channel pairs adjacent: packed FMAs), its bias, conv1_9's bias,
    float w8[4][24], b8[24], b9[24], w21[24][8], b21[8];      // conv2_1 transposed [n][j], its bias
};

__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// (r0, r1) = (a0, a1) + (b0, b1) in one issue slot (add.rn.f32x2), bit-identical to two scalar adds
__device__ __forceinline__ void add2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
    const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
    r0 = r.x; r1 = r.y;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}
// 24 consecutive columns of this thread's lane under one wait
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, float (&v)[24]) {
    uint32_t r[24];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%25];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(taddr), "r"(taddr + 16) : "memory");
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] = __uint_as_float(r[i]);
}

// NSPLIT = 1: 12 producer warps (8 channels per thread, tap and step); 2: 24 warps, each half of the step's channels (measured
// slower, 0.784 against 0.668 ms: the kernel is bound by instruction issue, not by latency — kept as the A/B switch YF_DENSE_TA_SPLIT)
template <int NSPLIT_ = 1>
struct DenseTaCfgT {
    static constexpr int NSPLIT = NSPLIT_, CPT = 8 / NSPLIT;
    static constexpr int NWW = 12 * NSPLIT, NTW = NWW * 32, NEW = 4, NT = NTW + NEW * 32 + 32;
    static constexpr int TH = 16, TW = 8, OPIX = TH * TW;
    static constexpr int RH = 2 * TH + 1, RW = 2 * TW + 1;
    static constexpr int XW = 24, XOFF = 3;                                         // x box [4][RH][XW] from the 16-byte aligned column 2 ox0 - 4 (yf_tma.cuh): tile column j at XOFF + j
    static constexpr int XS1 = 4 * RH * XW, NXB = 3;                                // three boxes in flight
    static constexpr int WRES = 27 * 64 * 8;                                        // the B blocks of pack_dense_tc
    static constexpr int NMAIN = 48, ACOLS = 72, ABUF = 2 * ACOLS, TM_ACC = 2 * ABUF, OBUF = 3 * NMAIN + 32, TCOLS = 512;
    static constexpr int SMEM_FLOATS = WRES + NXB * XS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static_assert(TM_ACC + OBUF <= TCOLS && NEW * 32 == OPIX, "TMEM columns / epilogue mapping");
    static_assert((XS1 * 4) % 128 == 0 && (WRES * 4) % 128 == 0, "TMA destination alignment");
    static_assert(NSPLIT == 1 || NSPLIT == 2, "channel split");
};
#ifndef YF_DENSE_TA_SPLIT
#define YF_DENSE_TA_SPLIT 1
#endif
using DenseTaCfg = DenseTaCfgT<YF_DENSE_TA_SPLIT>;

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
dense_ta_kernel(const __grid_constant__ CUtensorMap xmap, float* __restrict__ y, const float* __restrict__ wts, const __grid_constant__ DenseSmall sw,
                int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NTW = C::NTW, NWW = C::NWW;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Wr = base;                            // 27 resident B operands
    float* Xs0 = Wr + C::WRES;                   // [NXB boxes][4][RH][XW]
    __shared__ __align__(8) uint64_t wres, afull[2], afree[2], ofull, ofree, xfull[C::NXB], xfree[C::NXB];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&wres, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], NWW); mbar_init(&afree[i], 1); }
        mbar_init(&ofull, 1); mbar_init(&ofree, C::NEW);
        for (int i = 0; i < C::NXB; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], NWW); }
        mbar_fence_init();
        mbar_expect_tx(&wres, C::WRES * 4);                      // before the dependency wait: weights are not activations
        bulk_load(Wr, wts, C::WRES * 4, &wres);
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

    const int tpi = tiles_x * tiles_y;
    const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;
    const uint32_t inv_tpi = (uint32_t)((0x100000000ull + (uint64_t)tpi - 1ull) / (uint64_t)tpi);      // tile / tpi as one multiply + fix-up
    auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        b = (int)__umulhi((uint32_t)tile, inv_tpi);
        if (b * tpi > tile) --b;                                 // the rounded-up reciprocal can overshoot by one once tile * tpi >= 2^32
        const int t = tile - b * tpi;
        const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
        oy0 = ty * C::TH; ox0 = (t - ty * tiles_x) * C::TW;
    };
    if (warp == NWW + C::NEW) {
        // ================= tensor-core warp (also keeps the x boxes coming: tensor-tile TMA, three tiles in flight) =================
        if (ntile > 0 && elect_one()) {
            constexpr uint32_t IDESC_A = umma_idesc_tf32(C::NMAIN) & ~(1u << 15), IDESC_B = umma_idesc_tf32(32) & ~(1u << 15);
            const uint64_t dw0 = umma_desc(smem_u32(Wr), 128, 256, 0);
            // the x box of tile ti: [4][RH][XW] from the aligned column 2 ox0 - 4, zero outside the image
            auto issue_x = [&](int ti) {
                int b, oy0, ox0;
                origin(ti, b, oy0, ox0);
                mbar_expect_tx(&xfull[ti % C::NXB], C::XS1 * 4);
                tma_load4(Xs0 + (ti % C::NXB) * C::XS1, &xmap, &xfull[ti % C::NXB], 2 * ox0 - 4, 2 * oy0 - 1, 0, b);
            };
            tma_prefetch_desc(&xmap);
            for (int i = 0; i < C::NXB && i < ntile; ++i) issue_x(i);
            mbar_wait(&wres, 0);
            int d = 0;
            for (int ti = 0; ti < ntile; ++ti) {
                if (ti + C::NXB < ntile) {                                      // every producer warp has taken tile ti out of its box: refill it
                    mbar_wait(&xfree[ti % C::NXB], (ti / C::NXB) & 1);
                    issue_x(ti + C::NXB);
                }
                if (ti >= 1) mbar_wait(&ofree, (ti - 1) & 1);                   // the accumulators of tile ti - 1 have been read out
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb, ++d) {
                    const int buf = d & 1;
                    mbar_wait(&afull[buf], (d >> 1) & 1);
                    tc_fence_after();
                    DTRACE(ti, 8 + cb);
                    // main (9 MMAs per step): E_hi . [W9hi | W9lo] -> this step's 48 columns; corr (27 per tile): E_lo . W9hi -> 32 columns
                    const uint32_t acc = tmem + C::TM_ACC + cb * C::NMAIN, corr = tmem + C::TM_ACC + 3 * C::NMAIN;
                    const uint32_t ahi = tmem + buf * C::ABUF, alo = ahi + C::ACOLS;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint64_t wb = dw0 + (uint64_t)(((cb * 9 + t) * 512 * 4) >> 4);
                        umma_tf32_ta(acc, ahi + t * 8, wb, IDESC_A, t ? 1u : 0u);
                        umma_tf32_ta(corr, alo + t * 8, wb, IDESC_B, (cb | t) ? 1u : 0u);
                    }
                    umma_commit(&afree[buf]);
                    DTRACE(ti, 11 + cb);
                }
                umma_commit(&ofull);
            }
        }
    } else {
        if (warp >= NWW) {
            // ================= epilogue warps (thread = pixel): step accumulators + correction + b9, ReLU, conv2_1 in registers -> HBM ====
            const int q = warp - NWW, em = q * 32 + lane, eoy = em >> 3, eox = em & 7;
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + C::TM_ACC;
            for (int ti = 0; ti < ntile; ++ti) {
                int eb, ey0, ex0;
                origin(ti, eb, ey0, ex0);
                mbar_wait(&ofull, ti & 1);
                tc_fence_after();
                if (em == 0) DTRACE(ti, 5);
                float v[24], t24[24];
                tmem_ld24(ta + 3 * C::NMAIN, t24);                                   // the tile's lo . hi correction accumulator
#pragma unroll
                for (int n = 0; n < 24; n += 2) add2(t24[n], t24[n + 1], sw.b9[n], sw.b9[n + 1], v[n], v[n + 1]);
#pragma unroll
                for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                                    // columns [0, 24) E_hi . W9hi, [24, 48) E_hi . W9lo
                        tmem_ld24(ta + cb * C::NMAIN + h * 24, t24);
#pragma unroll
                        for (int n = 0; n < 24; n += 2) add2(v[n], v[n + 1], t24[n], t24[n + 1], v[n], v[n + 1]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ofree);
                if (em == 0) DTRACE(ti, 7);
                float ps[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) ps[j] = sw.b21[j];
#pragma unroll
                for (int n = 0; n < 24; ++n) {
                    const float vn = fmaxf(v[n], 0.f);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) ffma2(sw.w21[n][j], sw.w21[n][j + 1], vn, ps[j], ps[j + 1]);
                }
                const int gy = ey0 + eoy, gx = ex0 + eox;
                if (gy < Hout && gx < Wout) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[(((size_t)eb * 8 + j) * Hout + gy) * Wout + gx] = ps[j];
                }
                if (em == 0) DTRACE(ti, 14);
            }
        } else {
            // ================= producer warps: lane quarter q (output pixels 32 q .. 32 q + 31), kernel row ky =================
            const int q = warp & 3, ky = (warp >> 2) % 3, half = warp / 12;     // lane quarter, kernel row, which CPT channels of a step
            const int m = q * 32 + lane, oyl = m >> 3, oxl = m & 7;
            const int r = 2 * oyl + ky;                                          // input row / columns 2 oxl + kx of the tile
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + ky * 24;
            // this thread's three input pixels (4 channels each) of a tile and whether they lie inside the image
            auto fetch = [&](int ti, float (&xv)[3][4], float (&inm)[3], int& oy0, int& ox0) {
                int b;
                origin(ti, b, oy0, ox0);
                const float* Xs = Xs0 + (ti % C::NXB) * C::XS1;
                mbar_wait(&xfull[ti % C::NXB], (ti / C::NXB) & 1);
                const bool rowin = (unsigned)(2 * oy0 - 1 + r) < (unsigned)Hin;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int j = 2 * oxl + kx;
                    inm[kx] = (rowin && (unsigned)(2 * ox0 - 1 + j) < (unsigned)Win) ? 1.f : 0.f;   // conv1_9 zero-pads ITS input, the activation
#pragma unroll
                    for (int k = 0; k < 4; ++k) xv[kx][k] = Xs[k * C::RH * C::XW + r * C::XW + j + C::XOFF];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xfree[ti % C::NXB]);
            };
            int d = 0;
            float xv[3][4], inm[3], xn[3][4], inmn[3];
            int oy0 = 0, ox0 = 0, noy0 = 0, nox0 = 0;
            if (ntile > 0) fetch(0, xn, inmn, noy0, nox0);
            for (int ti = 0; ti < ntile; ++ti) {
                if (tid == 0) DTRACE(ti, 0);
                oy0 = noy0; ox0 = nox0;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    inm[kx] = inmn[kx];
#pragma unroll
                    for (int k = 0; k < 4; ++k) xv[kx][k] = xn[kx][k];
                }
                // the next tile's pixels are fetched now (their box landed long ago) and consumed after this tile's steps
                if (ti + 1 < ntile) fetch(ti + 1, xn, inmn, noy0, nox0);
                if (tid == 0) DTRACE(ti, 1);
                // interior tiles (no input pixel of the tile lies outside the image: all but the top / left tiles) skip the bias mask
                const bool interior = oy0 > 0 && ox0 > 0 && 2 * oy0 + C::RH - 2 < Hin && 2 * ox0 + C::RW - 2 < Win;
                auto steps = [&](auto tag) {
                    constexpr bool INTERIOR = decltype(tag)::value;
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb, ++d) {
                        const int buf = d & 1;
                        if (d >= 2) mbar_wait(&afree[buf], ((d >> 1) - 1) & 1);  // the MMAs that read this A buffer have completed
                        tc_fence_after();
                        if (tid == 0 && cb == 0) DTRACE(ti, 6);
                        const uint32_t tb = ta + buf * C::ABUF;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            float hi[C::CPT], lo[C::CPT];
                            auto pairs = [&](auto htag) {
                                constexpr int H = decltype(htag)::value;
#pragma unroll
                                for (int i = 0; i < C::CPT; i += 2) {
                                    const int c = cb * 8 + H * C::CPT + i;
                                    // outside the image x is 0 (zero-filled) and the bias is masked: e = relu(0) = 0
                                    float e0 = INTERIOR ? sw.b8[c] : sw.b8[c] * inm[kx], e1 = INTERIOR ? sw.b8[c + 1] : sw.b8[c + 1] * inm[kx];
#pragma unroll
                                    for (int k = 0; k < 4; ++k) ffma2(sw.w8[k][c], sw.w8[k][c + 1], xv[kx][k], e0, e1);   // k ascending, fmaf rounding per lane
                                    e0 = fmaxf(e0, 0.f); e1 = fmaxf(e1, 0.f);
                                    hi[i] = tf32_hi(e0); hi[i + 1] = tf32_hi(e1);
                                    lo[i] = e0; lo[i + 1] = e1;
                                    ffma2(hi[i], hi[i + 1], -1.f, lo[i], lo[i + 1]);     // lo = e - hi, exact either way
                                }
                            };
                            if (C::NSPLIT == 1 || half == 0) pairs(std::integral_constant<int, 0>{});
                            else pairs(std::integral_constant<int, C::NSPLIT - 1>{});
                            if constexpr (C::NSPLIT == 1) {
                                tmem_st8(tb + kx * 8, hi);
                                tmem_st8(tb + C::ACOLS + kx * 8, lo);
                            } else {
                                tmem_st4(tb + kx * 8 + half * 4, hi);
                                tmem_st4(tb + C::ACOLS + kx * 8 + half * 4, lo);
                            }
                        }
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&afull[buf]);
                        if (tid == 0) DTRACE(ti, 2 + cb);
                    }
                };
                if (interior) steps(std::true_type{});
                else steps(std::false_type{});
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWW + C::NEW) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
