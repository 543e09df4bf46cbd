// yf_kernels.cuh — fused forward kernels of the YOLO-Fastest hot path for sm_100a.
//
// Data layout: activations are planar fp32 [B][C][H][W] in HBM — the reference's own NCHW
// (yolo_fastest.py:150-218).  A CTA owns one spatial tile of one image; every stage keeps its
// operands in shared memory as [channel][tile pixel], so the 32 lanes of a warp walk consecutive
// pixels (conflict-free 128-bit LDS, coalesced 128-bit STG) while weights are warp-broadcast.
// Each fused group keeps its intermediate activations on chip:
//
//   pw_halo   1x1 conv + bias + ReLU over the halo tile         (conv_norm_relu k=1, yolo_fastest.py:16-26)
//   dw_stage  depthwise KxK stride S + bias + ReLU, sliding window in registers   (groups=C, :57,81,95,...)
//   pw_accum  1x1 conv accumulated over mid-channel chunks into per-thread register tiles (:58,82,...)
//
// The mid channels of an inverted-residual block are processed in chunks of MC so shared memory
// holds only [MC][halo] + [MC][tile] floats whatever the expansion width (8..224).
// BatchNorm is folded into the weights on the host (eval-mode BN is affine).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifndef YF_FFMA2
#define YF_FFMA2 1
#endif

namespace yf {

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int rup(int a, int b) { return cdiv(a, b) * b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// Tile geometry of a KSxKS stride-S stage producing a TH x TW output tile.
template <int KS_, int S_, int TH_, int TW_>
struct Geo {
    static constexpr int KS = KS_, S = S_, TH = TH_, TW = TW_;
    static constexpr int P = (KS - 1) / 2;            // zero padding (k-1)//2, yolo_fastest.py:17-19
    static constexpr int IH = (TH - 1) * S + KS;      // halo tile rows
    static constexpr int IW = (TW - 1) * S + KS;      // halo tile columns actually needed
    static constexpr int IWS = rup(IW, 4);            // row stride in smem (float4 aligned)
    static constexpr int IPIX = IH * IWS;
    static constexpr int OPIX = TH * TW;
    // Stride-2 stages keep their INPUT rows (E) de-interleaved by column parity: [even cols (EW) | odd cols (TW)].
    // Four adjacent outputs then read three conflict-free 128-bit words per row instead of a stride-2 gather
    // (which is a 2-way/8-way bank conflict with the lanes of a warp on consecutive output groups).
    static constexpr bool SPLIT = (S == 2);
    static constexpr int EW = SPLIT ? rup(TW + 1, 4) : 0;
    static_assert(TW % 4 == 0, "tile width must be a multiple of 4");
    static_assert(!SPLIT || (KS == 3 && EW + TW == IWS), "parity split is laid out for 3x3 stride 2");
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// Two FMAs in one issue slot (fma.rn.f32x2 -> FFMA2, sm_100+): (a0, a1) += (w0, w1) * x. Per-lane rounding is the scalar fmaf's, so
// results are bit-identical; the register-tiled 1x1 loops below are issue-bound, and pairing the accumulators over adjacent output
// channels (weights arrive as natural pairs from 128-bit loads) raises their FMA rate by ~15% (tools/selftest/ffma2_bench.cu).
__device__ __forceinline__ void ffma2(float w0, float w1, float x, float& a0, float& a1) {
#if YF_FFMA2
    const float2 r = __ffma2_rn(make_float2(w0, w1), make_float2(x, x), make_float2(a0, a1));
    a0 = r.x; a1 = r.y;
#else
    a0 = fmaf(w0, x, a0); a1 = fmaf(w1, x, a1);
#endif
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Cooperative copy of n floats (n % 4 == 0, both 16B aligned) global -> shared.
template <int NT>
__device__ __forceinline__ void copy_f4(float* __restrict__ dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += NT * 4) st4(dst + i, __ldg(reinterpret_cast<const float4*>(src + i)));
}

// ---- async-proxy primitives (sm_90+/sm_100a): mbarrier, TMA tensor tiles, bulk copies ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Programmatic dependent launch (yf_api.cu: launch_k): a kernel launched with the programmatic-serialization attribute starts as soon as
// every CTA of its predecessor has executed pdl_trigger() (or exited) — its set-up (barriers, TMEM allocation, weights into shared
// memory / registers) overlaps the predecessor's tail — and must execute pdl_wait() before it touches any activation: the wait returns
// when the predecessor grid has completed and its writes are visible. Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// 4-byte cp.async with zero fill (src_bytes = 0 writes zeros and does not touch src): the block-cooperative engine of round 1 stages
// its halo tiles with LDGSTS (they start 1-2 floats left of a 16-byte boundary); the round-2 engines use tensor-tile TMA boxes that
// start at the aligned column instead (yf_tma.cuh).
__device__ __forceinline__ void cp_async4(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// 16-byte form (both addresses 16B aligned); src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 1-D bulk copy global -> shared (bytes % 16 == 0, both 16B aligned), completion on an mbarrier.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Load an [nch][RH][RW] rectangle (zero outside the image, zero in the pad columns >= RW) of a
// planar image into smem [nch][RH][RWS]. src points at channel 0 of the image.
template <int RH, int RW, int RWS, int NT>
__device__ __forceinline__ void load_rect(float* __restrict__ dst, const float* __restrict__ src, int nch, int nch_valid,
                                          int Hin, int Win, int iy0, int ix0) {
    constexpr int RPIX = RH * RWS;
    for (int idx = threadIdx.x; idx < nch * RPIX; idx += NT) {
        const int c = idx / RPIX;
        const int rem = idx - c * RPIX;
        const int r = rem / RWS;
        const int j = rem - r * RWS;
        const int gy = iy0 + r, gx = ix0 + j;
        float v = 0.f;
        if (c < nch_valid && j < RW && (unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win)
            v = __ldg(src + ((size_t)c * Hin + gy) * Win + gx);
        dst[idx] = v;
    }
}

// Asynchronous form of load_rect: warps walk (channel, row) pairs, lanes walk columns, every element is one
// zero-filling 4-byte cp.async. The caller commits the group and waits (cp_async_wait_all + __syncthreads) before use.
template <int RH, int RW, int RWS, int NT>
__device__ __forceinline__ void load_rect_async(float* __restrict__ dst, const float* __restrict__ src, int nch, int nch_valid,
                                                int Hin, int Win, int iy0, int ix0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    for (int row = warp; row < nch * RH; row += NW) {
        const int c = row / RH;
        const int r = row - c * RH;
        const int gy = iy0 + r;
        const bool rowok = c < nch_valid && (unsigned)gy < (unsigned)Hin;
        const float* srow = src + ((size_t)c * Hin + (rowok ? gy : 0)) * Win;
        float* drow = dst + row * RWS;
#pragma unroll
        for (int j0 = 0; j0 < RWS; j0 += 32) {
            const int j = j0 + lane;
            if (j < RWS) {
                const int gx = ix0 + j;
                const bool ok = rowok && j < RW && (unsigned)gx < (unsigned)Win;
                cp_async4(drow + j, ok ? srow + gx : src, ok ? 4 : 0);
            }
        }
    }
}

// Stage 1: E[m][p] = mask(p) * relu(sum_k W[k][m] X[k][p] + b[m]) over the whole halo tile.
// mask(p) = pixel p lies inside the image: the depthwise conv that follows zero-pads ITS input,
// i.e. the activation, not the bias (yolo_fastest.py:17-22).
// Thread item = 4 consecutive halo pixels x PN mid channels.  W is [K][MC] in smem.
// DUAL additionally writes the tile-owned pixels to `skip` (global, [MC..][Hin][Win] of this image).
template <class G, int K, int MC, int PN, int NT, bool DUAL>
__device__ __forceinline__ void pw_halo(const float* __restrict__ Xs, const float* __restrict__ W,
                                        const float* __restrict__ bias, float* __restrict__ Es,
                                        int iy0, int ix0, int Hin, int Win,
                                        float* __restrict__ skip, int skip_valid) {
    static_assert(PN % 4 == 0 && MC % PN == 0, "bad PN");
    constexpr int NPG = G::IPIX / 4;
    constexpr int NCG = MC / PN;
    for (int item = threadIdx.x; item < NPG * NCG; item += NT) {
        const int cg = item / NPG;
        const int pg = item - cg * NPG;
        const int p0 = pg * 4;
        const int r = p0 / G::IWS;
        const int j0 = p0 - r * G::IWS;
        const int gy = iy0 + r, gx0 = ix0 + j0;
        float* e = Es + cg * PN * G::IPIX + p0;
        const bool rowok = (unsigned)gy < (unsigned)Hin;
        bool m[4];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m[i] = rowok && (j0 + i < G::IW) && ((unsigned)(gx0 + i) < (unsigned)Win);
            any |= m[i];
        }
        float acc[PN][4];
        if (any) {
#pragma unroll
            for (int n4 = 0; n4 < PN / 4; ++n4) {
                const float4 b = ld4(bias + cg * PN + n4 * 4);
                const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[n4 * 4 + q][i] = bb[q];
            }
            const float* xp = Xs + p0;
            const float* wp = W + cg * PN;
#pragma unroll 8
            for (int k = 0; k < K; ++k) {
                const float4 xv = ld4(xp + k * G::IPIX);
                const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int n4 = 0; n4 < PN / 4; ++n4) {
                    const float4 w = ld4(wp + k * MC + n4 * 4);
                    const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int q = 0; q < 4; q += 2)
#pragma unroll
                        for (int i = 0; i < 4; ++i) ffma2(w4[q], w4[q + 1], x4[i], acc[n4 * 4 + q][i], acc[n4 * 4 + q + 1][i]);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < PN; ++n) {
            float4 o;
            o.x = (any && m[0]) ? fmaxf(acc[n][0], 0.f) : 0.f;
            o.y = (any && m[1]) ? fmaxf(acc[n][1], 0.f) : 0.f;
            o.z = (any && m[2]) ? fmaxf(acc[n][2], 0.f) : 0.f;
            o.w = (any && m[3]) ? fmaxf(acc[n][3], 0.f) : 0.f;
            if (G::SPLIT) {
                float* er = Es + (cg * PN + n) * G::IPIX + r * G::IWS + (j0 >> 1);
                *reinterpret_cast<float2*>(er) = make_float2(o.x, o.z);                                  // columns j0, j0+2
                if ((j0 >> 1) < G::TW) *reinterpret_cast<float2*>(er + G::EW) = make_float2(o.y, o.w);   // columns j0+1, j0+3
            } else {
                st4(e + n * G::IPIX, o);
            }
            if (DUAL && any) {
                // pixels this tile owns (not halo): rows/cols [P, P + T*S)
                const bool rown = r >= G::P && r < G::P + G::TH * G::S;
                if (rown && (cg * PN + n) < skip_valid) {
                    float* sp = skip + ((size_t)(cg * PN + n) * Hin + gy) * Win + gx0;
                    const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j0 + i;
                        if (m[i] && j >= G::P && j < G::P + G::TW * G::S) sp[i] = ov[i];
                    }
                }
            }
        }
    }
}

template <int NV>
__device__ __forceinline__ void load_row(float (&dst)[NV], const float* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
        const float4 v = ld4(p + 4 * i);
        dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
    if (NV % 4 >= 2) {
        const float2 v = *reinterpret_cast<const float2*>(p + (NV / 4) * 4);
        dst[(NV / 4) * 4 + 0] = v.x; dst[(NV / 4) * 4 + 1] = v.y;
    }
    if (NV % 2 == 1) dst[NV - 1] = p[NV - 1];
}

// The NV = 3*S + KS input columns feeding 4 adjacent outputs (output group g) of one input row.
// S == 1: contiguous. S == 2 (parity-split row, EW even columns then the odd ones): v[2i+dx] for i<4, dx<3.
template <class G>
__device__ __forceinline__ void load_window(float (&dst)[3 * G::S + G::KS], const float* __restrict__ row, int g) {
    if (G::SPLIT) {
        const float4 a = ld4(row + 4 * g);
        const float4 b = ld4(row + 4 * g + 4);
        const float4 c = ld4(row + G::EW + 4 * g);
        dst[0] = a.x; dst[2] = a.y; dst[4] = a.z; dst[6] = a.w; dst[8] = b.x;
        dst[1] = c.x; dst[3] = c.y; dst[5] = c.z; dst[7] = c.w;
    } else {
        load_row<3 * G::S + G::KS>(dst, row + 4 * g * G::S);
    }
}

// Stage 2: depthwise KSxKS stride S + bias + ReLU:  D[m][oy][ox] = relu(sum W[m][dy][dx] E[m][oy*S+dy][ox*S+dx] + b[m]).
// Thread item = one channel x a strip of 4 output columns x RH output rows; the KS input rows of
// the window live in registers and slide down (S new row loads per output row).
template <class G, int MC, int RH, int NT>
__device__ __forceinline__ void dw_stage(const float* __restrict__ Es, const float* __restrict__ Wd,
                                         const float* __restrict__ bd, float* __restrict__ Ds) {
    constexpr int KS = G::KS, S = G::S;
    static_assert(G::TH % RH == 0, "RH must divide TH");
    static_assert(KS > S, "window must overlap");
    constexpr int NSTRIP = G::TW / 4, NSEG = G::TH / RH;
    constexpr int NV = 3 * S + KS;   // input columns feeding 4 adjacent outputs
    constexpr int NITEM = MC * NSEG * NSTRIP;
    for (int item = threadIdx.x; item < NITEM; item += NT) {
        const int m = item / (NSEG * NSTRIP);
        const int rem = item - m * (NSEG * NSTRIP);
        const int seg = rem / NSTRIP;
        const int g = rem - seg * NSTRIP;
        float w[KS * KS];
#pragma unroll
        for (int t = 0; t < KS * KS; ++t) w[t] = Wd[m * KS * KS + t];
        const float b = bd[m];
        const float* e = Es + m * G::IPIX + (seg * RH * S) * G::IWS;
        float* d = Ds + m * G::OPIX + (seg * RH) * G::TW + 4 * g;
        float win[KS][NV];
#pragma unroll
        for (int dd = 0; dd < KS - S; ++dd) load_window<G>(win[dd + S], e + dd * G::IWS, g);
#pragma unroll
        for (int oy = 0; oy < RH; ++oy) {
#pragma unroll
            for (int dd = 0; dd < KS - S; ++dd)
#pragma unroll
                for (int v = 0; v < NV; ++v) win[dd][v] = win[dd + S][v];
#pragma unroll
            for (int dd = KS - S; dd < KS; ++dd) load_window<G>(win[dd], e + (oy * S + dd) * G::IWS, g);
            float a[4] = {b, b, b, b};
#pragma unroll
            for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = fmaf(w[dy * KS + dx], win[dy][i * S + dx], a[i]);
            st4(d + oy * G::TW, make_float4(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f), fmaxf(a[2], 0.f), fmaxf(a[3], 0.f)));
        }
    }
}

// Stage 3: acc[n][p] += sum_{m < KC} W[m][n] D[m][p]; thread item = 4 consecutive pixels x PN out
// channels, IPT items per thread, accumulators persist in registers across chunks.
template <int OPIX, int KC, int N, int PN, int IPT, int NT>
__device__ __forceinline__ void pw_accum(const float* __restrict__ Ds, const float* __restrict__ W,
                                         float (&acc)[IPT][PN][4]) {
    static_assert(PN % 4 == 0 && N % PN == 0 && OPIX % 4 == 0, "bad pw_accum tiling");
    constexpr int NPG = OPIX / 4;
    constexpr int NCG = N / PN;
#pragma unroll
    for (int it = 0; it < IPT; ++it) {
        const int item = threadIdx.x + it * NT;
        if (item < NPG * NCG) {
            const int cg = item / NPG;
            const int pg = item - cg * NPG;
            const float* dp = Ds + pg * 4;
            const float* wp = W + cg * PN;
#pragma unroll 8
            for (int m = 0; m < KC; ++m) {
                const float4 dv = ld4(dp + m * OPIX);
                const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int n4 = 0; n4 < PN / 4; ++n4) {
                    const float4 w = ld4(wp + m * N + n4 * 4);
                    const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int q = 0; q < 4; q += 2)
#pragma unroll
                        for (int i = 0; i < 4; ++i) ffma2(w4[q], w4[q + 1], d4[i], acc[it][n4 * 4 + q][i], acc[it][n4 * 4 + q + 1][i]);
                }
            }
        }
    }
}

// Store 4 consecutive pixels of one output channel row (guarded; 128-bit when aligned).
__device__ __forceinline__ void store_px4(float* __restrict__ rowp, int gx0, int Wout, const float (&v)[4]) {
    if (((Wout & 3) == 0) && gx0 + 3 < Wout) {
        st4(rowp + gx0, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (gx0 + i < Wout) rowp[gx0 + i] = v[i];
    }
}

struct TileId { int b, ty, tx; };
__device__ __forceinline__ TileId tile_id(int tiles_x, int tiles_y) {
    int t = blockIdx.x;
    TileId id;
    id.tx = t % tiles_x; t /= tiles_x;
    id.ty = t % tiles_y;
    id.b = t / tiles_y;
    return id;
}

// ---------------------------------------------------------------------------------------------
// IRB engine: [1x1 expand + ReLU] -> depthwise KxK (stride S) + ReLU -> 1x1 project (+bias)
//             (+ residual) (+ ReLU) [-> 1x1 head with bias]
// Covers BasicResBlock (yolo_fastest.py:52-66), the strided transition triples (:94-97,111-114),
// conv4_2/conv4_3/conv5_1 with the conv4_2 skip output (:121-124,190), and with EXPAND=false the
// neck pairs dw5x5 -> 1x1 (:133-136,143-146) including the biased head convs (:138,148).
// Packed weights (floats): NCHUNK x { [W1: CIN*MC][b1: MC] (EXPAND only) [Wd: MC*KS*KS][bd: MC][W2: MC*COUT] },
// then [b2: COUT]. Head groups (HEADN > 0) carry no W2/b2: after the chunk blocks come the composed head weights
// [Wh': CMIDP*headp][bh': headp], headp = rup(headn, 4), headn = num_anchors * (5 + num_cls) given at run time.
// ---------------------------------------------------------------------------------------------
template <int CIN_, int CMID_, int COUT_, int KS_, int S_, int TH_, int TW_, int MC_, int PN1_, int PN3_, int RH_,
          int NT_, int MINB_, bool EXPAND_, bool RES_, bool RELU_OUT_, bool DUAL_, int HEADN_ = 0, int XBUF_ = 1>
struct IrbCfg {
    static constexpr int XBUF = XBUF_;   // input-tile buffers: 2 = next tile prefetched during the whole tile, 1 = during its last chunk
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, MC = MC_, PN1 = PN1_, PN3 = PN3_, RH = RH_;
    static constexpr int NT = NT_, MINB = MINB_, HEADN = HEADN_;
    static constexpr bool EXPAND = EXPAND_, RES = RES_, RELU_OUT = RELU_OUT_, DUAL = DUAL_;
    using G = Geo<KS_, S_, TH_, TW_>;
    static constexpr int KK = KS_ * KS_;
    static constexpr int CMIDP = rup(CMID, MC);
    static constexpr int NCHUNK = CMIDP / MC;
    // HEADN > 0: the group ends in a biased 1x1 head conv (yolo_fastest.py:138,148). The 1x1 before it (conv5_6 / conv4_1_5) has no
    // activation, so the two linear maps are composed on the host into ONE [CMID][headp] matrix (yf_api.cu: pack_irb) and the
    // engine runs depthwise -> composed head: the depthwise output of ALL mid channels stays in smem (Os) and is contracted once.
    static constexpr bool HEADC = HEADN > 0;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_B1 = OFF_W1 + (EXPAND ? CIN * MC : 0);
    static constexpr int OFF_WD = OFF_B1 + (EXPAND ? MC : 0);
    static constexpr int OFF_BD = OFF_WD + MC * KK;
    static constexpr int OFF_W2 = OFF_BD + MC;
    static constexpr int CB = OFF_W2 + (HEADC ? 0 : MC * COUT);              // floats per chunk block
    static constexpr int OFF_B2 = NCHUNK * CB;
    static constexpr int OFF_WH = HEADC ? OFF_B2 : OFF_B2 + COUT;             // composed head weights [CMIDP][headp], then bias [headp] (runtime headn)
    // shared memory (floats); every region starts 128B aligned
    static constexpr int XS1 = EXPAND ? rup(CIN * G::IPIX, 32) : 0;   // one input halo tile [CIN][IH][IWS]
    static constexpr int XS = XBUF * XS1;
    static constexpr int ES = rup(MC * G::IPIX, 32);
    static constexpr int EBUF = HEADC ? 2 : 1;                         // staged-chunk buffers (HEADC has no 1x1 phase to hide the next chunk's copy behind)
    static constexpr int DS = HEADC ? rup(CMIDP * G::OPIX, 32) : rup(MC * G::OPIX, 32);
    static constexpr int ACT = XS + EBUF * ES + DS;
    static constexpr int WS1 = rup(CB, 32);
    static constexpr int SMEM_FLOATS = ACT + 2 * WS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static constexpr int XBOX_C = EXPAND ? CIN : MC;            // channels staged per tile load
    static constexpr int NPG3 = G::OPIX / 4, NCG3 = COUT / PN3;
    static constexpr int IPT = cdiv(NPG3 * NCG3, NT);
    static_assert(MC % 4 == 0 && COUT % 4 == 0 && CB % 4 == 0, "alignment");
    static_assert(!RES || (CIN == COUT && S_ == 1 && EXPAND), "residual needs same shape");
    static_assert(EXPAND || (CMID == CIN && CIN % MC == 0), "dw-first groups have CMID == CIN, a multiple of MC");
    static_assert(!HEADC || !EXPAND, "head groups are depthwise-first");
    static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
    static_assert(XBUF == 1 || XBUF == 2, "one or two input-tile buffers");
};

// Persistent CTAs (grid <= #SM x MINB) loop over tiles. While a tile is computed, the input halo tile of the NEXT
// tile (cp.async into the other X buffer) and the weight block of the NEXT mid-channel chunk (cp.async.bulk ->
// mbarrier, double buffered) are in flight, so no warp waits on a global load in steady state.
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
irb_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ skip, const float* __restrict__ wts,
           int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles, int headn) {
    using G = typename C::G;
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];     // weight buffers
    float* Xs0 = smem;
    float* Es0 = smem + C::XS;
    float* Ds = Es0 + C::EBUF * C::ES;             // HEADC: depthwise output of all mid channels [CMIDP][OPIX]
    float* Ws = smem + C::ACT;
    const int tid = threadIdx.x;

    auto tile_origin = [&](int tile, int& b, int& oy0, int& ox0) {
        const int tx = tile % tiles_x;
        const int r = tile / tiles_x;
        oy0 = (r % tiles_y) * G::TH;
        ox0 = tx * G::TW;
        b = r / tiles_y;
    };
    // all threads: stage channels [c0, c0 + XBOX_C) of the tile's halo rectangle into dst (asynchronously)
    auto stage_tile = [&](int tile, int c0, float* dst) {
        int b, oy0, ox0;
        tile_origin(tile, b, oy0, ox0);
        load_rect_async<G::IH, G::IW, G::IWS, NT>(dst, x + ((size_t)b * C::CIN + c0) * Hin * Win, C::XBOX_C, C::CIN - c0, Hin, Win,
                                                  oy0 * G::S - G::P, ox0 * G::S - G::P);
        cp_async_commit();
    };
    auto issue_w = [&](int chunk, int buf) {      // one thread
        mbar_expect_tx(&bars[buf], C::CB * 4);
        bulk_load(Ws + buf * C::WS1, wts + (size_t)chunk * C::CB, C::CB * 4, &bars[buf]);
    };

    int tile = blockIdx.x;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tile < total_tiles) {
        if (tid == 0) issue_w(0, 0);
        stage_tile(tile, 0, C::EXPAND ? Xs0 : Es0);
    }

    uint32_t q = 0;      // running chunk sequence number of this CTA (weight buffer q & 1, parity (q >> 1) & 1)
    uint32_t eq = 0;     // running staged-chunk number (HEADC: staging buffer eq & 1)
    for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
        int tb, oy0, ox0;
        tile_origin(tile, tb, oy0, ox0);
        const int iy0 = oy0 * G::S - G::P, ix0 = ox0 * G::S - G::P;
        float* Xs = Xs0 + (C::XBUF == 2 ? (it & 1) * C::XS1 : 0);
        const bool have_next = tile + (int)gridDim.x < total_tiles;

        float acc[C::IPT][C::PN3][4];
#pragma unroll
        for (int i = 0; i < C::IPT; ++i)
#pragma unroll
            for (int n = 0; n < C::PN3; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][n][j] = 0.f;

        for (int c = 0; c < C::NCHUNK; ++c, ++q) {
            const int wbuf = (C::NCHUNK == 1) ? 0 : (int)(q & 1);
            const float* Wc = Ws + wbuf * C::WS1;
            if (!C::EXPAND || c == 0) cp_async_wait_all();   // this thread's share of the staged tile / E chunk has landed
            __syncthreads();   // (A) staged data visible to all; everything of the previous chunk / tile has been consumed
            if (tid == 0 && C::NCHUNK > 1 && (c + 1 < C::NCHUNK || have_next)) issue_w(c + 1 < C::NCHUNK ? c + 1 : 0, (int)((q + 1) & 1));
            if (C::EXPAND && C::XBUF == 2 && c == 0 && have_next) stage_tile(tile + gridDim.x, 0, Xs0 + ((it + 1) & 1) * C::XS1);
            if (C::RES && c == 0) {
                // out = project(...) + x (yolo_fastest.py:65): start the accumulators from the residual so the input
                // tile is dead after the last expand stage and its buffer can take the next tile
#pragma unroll
                for (int i = 0; i < C::IPT; ++i) {
                    const int item = tid + i * NT;
                    if (item < C::NPG3 * C::NCG3) {
                        const int cg = item / C::NPG3;
                        const int p0 = (item - cg * C::NPG3) * 4;
                        const int oy = p0 / G::TW, ox = p0 - oy * G::TW;
#pragma unroll
                        for (int n = 0; n < C::PN3; ++n) {
                            const float* xr = Xs + (cg * C::PN3 + n) * G::IPIX + (oy + G::P) * G::IWS + ox + G::P;
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][n][j] = xr[j];
                        }
                    }
                }
            }
            if (C::NCHUNK > 1) mbar_wait(&bars[wbuf], (q >> 1) & 1);
            else if (it == 0) mbar_wait(&bars[0], 0);
            float* Es = Es0 + (C::HEADC ? (eq & 1) * C::ES : 0);
            if (C::HEADC) {
                // the next channel chunk (or chunk 0 of the next tile) flies into the other staging buffer during this chunk's depthwise
                float* En = Es0 + ((eq + 1) & 1) * C::ES;
                if (c + 1 < C::NCHUNK) stage_tile(tile, (c + 1) * C::MC, En);
                else if (have_next) stage_tile(tile + gridDim.x, 0, En);
                ++eq;
                dw_stage<G, C::MC, C::RH, NT>(Es, Wc + C::OFF_WD, Wc + C::OFF_BD, Ds + c * C::MC * G::OPIX);
                continue;
            }
            if (C::EXPAND) {
                float* sk = C::DUAL ? skip + ((size_t)tb * C::CMID + c * C::MC) * Hin * Win : nullptr;
                pw_halo<G, C::CIN, C::MC, C::PN1, NT, C::DUAL>(Xs, Wc + C::OFF_W1, Wc + C::OFF_B1, Es, iy0, ix0, Hin, Win,
                                                               sk, C::CMID - c * C::MC);
                __syncthreads();   // (B) E complete; after the last chunk the input tile is dead
                if (C::XBUF == 1 && c == C::NCHUNK - 1 && have_next) stage_tile(tile + gridDim.x, 0, Xs0);
            }
            dw_stage<G, C::MC, C::RH, NT>(Es, Wc + C::OFF_WD, Wc + C::OFF_BD, Ds);
            __syncthreads();   // (C) D complete, E free
            if (!C::EXPAND) {
                // the next E chunk (next channel chunk of this tile, or chunk 0 of the next tile) flies during the 1x1 below
                if (c + 1 < C::NCHUNK) stage_tile(tile, (c + 1) * C::MC, Es);
                else if (have_next) stage_tile(tile + gridDim.x, 0, Es);
            }
            pw_accum<G::OPIX, C::MC, C::COUT, C::PN3, C::IPT, NT>(Ds, Wc + C::OFF_W2, acc);
        }

        if (C::HEADC) {
            // composed head: out[n] = sum_m Wh'[m][n] * D[m] + bh'[n] over ALL mid channels (the 1x1 before the head and the head
            // itself, yolo_fastest.py:136-138 / 146-148, folded into one matrix on the host); weights are read through L1.
            __syncthreads();               // D of every chunk is complete
            const int headp = (headn + 3) & ~3;
            const float* Wh = wts + C::OFF_WH;
            const float* bh = Wh + C::CMIDP * headp;
            constexpr int NPG = G::OPIX / 4;
            const int NCGH = headp / 4;
            for (int item = tid; item < NPG * NCGH; item += NT) {
                const int cg = item / NPG;
                const int pg = item - cg * NPG;
                const int p0 = pg * 4;
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bh + cg * 4));
                float a[4][4] = {{bb.x, bb.x, bb.x, bb.x}, {bb.y, bb.y, bb.y, bb.y}, {bb.z, bb.z, bb.z, bb.z}, {bb.w, bb.w, bb.w, bb.w}};
#pragma unroll 8
                for (int k = 0; k < C::CMIDP; ++k) {
                    const float4 ov = ld4(Ds + k * G::OPIX + p0);
                    const float4 w = __ldg(reinterpret_cast<const float4*>(Wh + k * headp + cg * 4));
                    const float o4[4] = {ov.x, ov.y, ov.z, ov.w};
                    const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int qq = 0; qq < 4; qq += 2)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ffma2(w4[qq], w4[qq + 1], o4[j], a[qq][j], a[qq + 1][j]);
                }
                const int oy = p0 / G::TW, ox = p0 - oy * G::TW;
                const int gy = oy0 + oy, gx0 = ox0 + ox;
                if (gy < Hout) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int ch = cg * 4 + qq;
                        if (ch < headn) store_px4(y + (((size_t)tb * headn + ch) * Hout + gy) * Wout, gx0, Wout, a[qq]);
                    }
                }
            }
            continue;                      // the next tile's first sync (A) orders these reads before its depthwise writes D again
        }
        const float* b2 = wts + C::OFF_B2;
#pragma unroll
        for (int i = 0; i < C::IPT; ++i) {
            const int item = tid + i * NT;
            if (item < C::NPG3 * C::NCG3) {
                const int cg = item / C::NPG3;
                const int pg = item - cg * C::NPG3;
                const int p0 = pg * 4;
                const int oy = p0 / G::TW, ox = p0 - oy * G::TW;
                const int gy = oy0 + oy, gx0 = ox0 + ox;
#pragma unroll
                for (int n = 0; n < C::PN3; ++n) {
                    const int ch = cg * C::PN3 + n;
                    const float b = __ldg(b2 + ch);
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = acc[i][n][j] + b;
                    if (C::RELU_OUT) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (gy < Hout) store_px4(y + (((size_t)tb * C::COUT + ch) * Hout + gy) * Wout, gx0, Wout, v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stem group: conv0 (dense 3x3 s2, in_ch=1 -> 8, ReLU) -> conv1_2 (1x1 8->8 ReLU) -> conv1_3 (dw3x3 ReLU)
// -> conv1_4 (1x1 8->4 linear)   (yolo_fastest.py:78-82,151-154).  Input [B,1,H,W], output [B,4,H/2,W/2].
// Packed weights: [W0: CIN*9*8 ([ci][tap][c])][b0: 8][W1: 8*8 ([k][m])][b1: 8][Wd: 8*9][bd: 8][W2: 8*4 ([m][n])][b2: 4]
// With U8IN the input is uint8 and (x-128)/255 is applied on load (detect.py:123-124).
// ---------------------------------------------------------------------------------------------
template <int TH_, int TW_, int NT_, int MINB_, int CIN_ = 1>
struct StemCfg {
    static constexpr int NT = NT_, MINB = MINB_, CIN = CIN_;          // CIN: image channels (1 = the shipped models, 3 = colour input)
    using G = Geo<3, 1, TH_, TW_>;
    static constexpr int RH = 2 * G::IH + 1;          // raw input rows feeding the conv0 halo tile
    static constexpr int RW = 2 * G::IWS + 1;         // raw input columns
    static constexpr int REW = G::IWS + 4;            // raw rows are stored parity-split: [even cols (REW) | odd cols (IWS)]
    static constexpr int RWS = REW + G::IWS;
    static constexpr int OFF_W0 = 0, OFF_B0 = 72 * CIN, OFF_W1 = OFF_B0 + 8, OFF_B1 = OFF_W1 + 64, OFF_WD = OFF_B1 + 8, OFF_BD = OFF_WD + 72,
                         OFF_W2 = OFF_BD + 8, OFF_B2 = OFF_W2 + 32;
    static constexpr int WFLOATS = OFF_B2 + 4;
    static constexpr int RS1 = RH * RWS, RS = CIN * RS1;              // raw rectangle of one channel / of all channels
    static constexpr int NPRE = cdiv(RS, NT_);
    static constexpr int XS = 8 * G::IPIX, ES = 8 * G::IPIX, DS = 8 * G::OPIX;
    static constexpr int WPAD = rup(WFLOATS, 4);
    static constexpr int SMEM_FLOATS = RS + XS + ES + DS + WPAD + 256;   // + the uint8 -> (x - 128) / 255 table
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static constexpr int IPT = cdiv(G::OPIX / 4, NT);
    static_assert(SMEM_BYTES <= 227 * 1024, "stem tile too large");
};

template <class C, bool U8IN>
__global__ void __launch_bounds__(C::NT, C::MINB)
stem_kernel(const void* __restrict__ xin, float* __restrict__ y, const float* __restrict__ wts,
            int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles) {
    using G = typename C::G;
    constexpr int NT = C::NT;
    constexpr int NPRE = C::NPRE;               // raw-tile elements per thread
    extern __shared__ __align__(128) float smem[];
    float* Rs = smem;
    float* Xs = Rs + C::RS;
    float* Es = Xs + C::XS;
    float* Ds = Es + C::ES;
    float* Ws = Ds + C::DS;
    float* Lut = Ws + C::WPAD;       // U8IN: (u - 128) / 255 for u = 0..255, the reference's fp32 division done once (detect.py:124)
    if (U8IN) {
        for (int i = threadIdx.x; i < 256; i += NT) Lut[i] = ((float)i - 128.0f) / 255.0f;
        __syncthreads();
    }
    // fetch this thread's share of a tile's raw rectangle into registers (normalising uint8 through the table)
    auto fetch_raw = [&](int tile, float (&pre)[NPRE]) {
        const int tx = tile % tiles_x;
        const int r0 = tile / tiles_x;
        const int ry0 = 2 * ((r0 % tiles_y) * G::TH - 1) - 1, rx0 = 2 * (tx * G::TW - 1) - 1, b = r0 / tiles_y;
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
            const int idx = threadIdx.x + k * NT;
            const int ci = C::CIN == 1 ? 0 : idx / C::RS1, i1 = idx - ci * C::RS1;
            const int r = i1 / C::RWS, jj = i1 - r * C::RWS;
            const int j = jj < C::REW ? 2 * jj : 2 * (jj - C::REW) + 1;     // raw column held at split position jj
            const int gy = ry0 + r, gx = rx0 + j;
            float v = 0.f;
            if (idx < C::RS && j < C::RW && (unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win) {
                const size_t off = (((size_t)b * C::CIN + ci) * Hin + gy) * Win + gx;
                if (U8IN) v = Lut[__ldg(reinterpret_cast<const unsigned char*>(xin) + off)];
                else v = __ldg(reinterpret_cast<const float*>(xin) + off);
            }
            pre[k] = v;
        }
    };
    for (int i = threadIdx.x; i < C::WFLOATS; i += NT) Ws[i] = __ldg(wts + i);
    float pre[NPRE];
    int tile = blockIdx.x;
    if (tile < total_tiles) fetch_raw(tile, pre);
    for (; tile < total_tiles; tile += gridDim.x) {
    const int tx_ = tile % tiles_x;
    const int r0_ = tile / tiles_x;
    const int oy0 = (r0_ % tiles_y) * G::TH, ox0 = tx_ * G::TW, tb = r0_ / tiles_y;
    const int iy0 = oy0 - 1, ix0 = ox0 - 1;            // halo origin in the H/2 map
    __syncthreads();                                   // previous tile done with Rs/Xs/Es/Ds
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
        const int idx = threadIdx.x + k * NT;
        if (idx < C::RS) Rs[idx] = pre[k];
    }
    __syncthreads();
    if (tile + (int)gridDim.x < total_tiles) fetch_raw(tile + gridDim.x, pre);   // lands while this tile is computed
    // conv0 over the halo tile: item = 4 consecutive halo pixels x 8 channels
    {
        constexpr int NPG = G::IPIX / 4;
        const float* W0 = Ws + C::OFF_W0;
        for (int pg = threadIdx.x; pg < NPG; pg += NT) {
            const int p0 = pg * 4;
            const int r = p0 / G::IWS, j0 = p0 - r * G::IWS;
            float a[8][4];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) a[c][i] = Ws[C::OFF_B0 + c];
#pragma unroll
            for (int ci = 0; ci < C::CIN; ++ci)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                float v[9];   // raw columns 2*j0 .. 2*j0+8 of this row: v[2i+dx] feeds output i, tap dx
                const float* rp = Rs + ci * C::RS1 + (2 * r + dy) * C::RWS + j0;
                {
                    const float4 a = ld4(rp), b = ld4(rp + 4), c = ld4(rp + C::REW);
                    v[0] = a.x; v[2] = a.y; v[4] = a.z; v[6] = a.w; v[8] = b.x;
                    v[1] = c.x; v[3] = c.y; v[5] = c.z; v[7] = c.w;
                }
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 wa = ld4(W0 + (ci * 9 + dy * 3 + dx) * 8);
                    const float4 wb = ld4(W0 + (ci * 9 + dy * 3 + dx) * 8 + 4);
                    const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int c = 0; c < 8; c += 2)
#pragma unroll
                        for (int i = 0; i < 4; ++i) ffma2(w8[c], w8[c + 1], v[2 * i + dx], a[c][i], a[c + 1][i]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st4(Xs + c * G::IPIX + p0, make_float4(fmaxf(a[c][0], 0.f), fmaxf(a[c][1], 0.f), fmaxf(a[c][2], 0.f), fmaxf(a[c][3], 0.f)));
        }
    }
    __syncthreads();
    pw_halo<G, 8, 8, 8, NT, false>(Xs, Ws + C::OFF_W1, Ws + C::OFF_B1, Es, iy0, ix0, Hout, Wout, nullptr, 0);
    __syncthreads();
    dw_stage<G, 8, G::TH, NT>(Es, Ws + C::OFF_WD, Ws + C::OFF_BD, Ds);
    __syncthreads();
    float acc[C::IPT][4][4];
#pragma unroll
    for (int it = 0; it < C::IPT; ++it)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[it][n][i] = 0.f;
    pw_accum<G::OPIX, 8, 4, 4, C::IPT, NT>(Ds, Ws + C::OFF_W2, acc);
#pragma unroll
    for (int it = 0; it < C::IPT; ++it) {
        const int pg = threadIdx.x + it * NT;
        if (pg < G::OPIX / 4) {
            const int p0 = pg * 4;
            const int oy = p0 / G::TW, ox = p0 - oy * G::TW;
            const int gy = oy0 + oy, gx0 = ox0 + ox;
            if (gy < Hout) {
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float b = Ws[C::OFF_B2 + n];
                    const float v[4] = {acc[it][n][0] + b, acc[it][n][1] + b, acc[it][n][2] + b, acc[it][n][3] + b};
                    store_px4(y + (((size_t)tb * 4 + n) * Hout + gy) * Wout, gx0, Wout, v);
                }
            }
        }
    }
    }   // tile loop
}

// ---------------------------------------------------------------------------------------------
// Dense group: conv1_8 (1x1 4->24 ReLU) -> conv1_9 (dense 3x3 s2 24->24 ReLU) -> conv2_1 (1x1 24->8 linear)
// (yolo_fastest.py:86-89,158-160).  22% of the network's MACs sit in conv1_9.
// The 24 mid channels are processed in chunks of MC; conv1_9 accumulates in registers
// (item = 4 output pixels x 8 output channels).
// Packed weights: NCHUNK x { [W1: 4*MC][b1: MC][Wc: MC*3*3(dy)*(3 cg)*(3 dx * 8 n)] } then
//                 [b9: 24][W3: 24*8 ([m][n])][b3: 8]
// ---------------------------------------------------------------------------------------------
template <int TH_, int TW_, int MC_, int NT_, int MINB_, int PXI_ = 1>
struct DenseCfg {
    static constexpr int NT = NT_, MINB = MINB_, MC = MC_, PXI = PXI_;   // PXI pixel groups per thread share one weight fetch
    using G = Geo<3, 2, TH_, TW_>;
    static constexpr int CM = 24;
    static constexpr int NCHUNK = CM / MC;
    static constexpr int OFF_W1 = 0, OFF_B1 = 4 * MC, OFF_WC = OFF_B1 + MC;
    static constexpr int CB = OFF_WC + MC * 3 * 3 * 24;          // [c][dy][cg][dx][8]
    static constexpr int OFF_B9 = NCHUNK * CB, OFF_W3 = OFF_B9 + 24, OFF_B3 = OFF_W3 + 24 * 8;
    static constexpr int WFLOATS = OFF_B3 + 8;
    static constexpr int XS = rup(4 * G::IPIX, 32);
    static constexpr int ES = rup(cmax(MC * G::IPIX, 24 * G::OPIX), 32);   // E chunk; reused for the conv1_9 output D [24][OPIX]
    static constexpr int SMEM_FLOATS = XS + ES + rup(WFLOATS, 32);
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static constexpr int NPG = G::OPIX / 4;
    static constexpr int PGT = NPG / PXI;              // conv1_9: thread = (channel group of 8) x (slot of PXI pixel groups)
    static constexpr int IPT9 = PXI;
    static constexpr int IPT3 = cdiv(NPG * 2, NT);     // conv2_1 items: px-groups x 2 channel groups of 4
    static_assert(CM % MC == 0 && MC % 4 == 0 && WFLOATS % 4 == 0, "bad MC");
    static_assert(NPG % PXI == 0 && 3 * PGT <= NT, "conv1_9 thread mapping does not fit the CTA");
    static_assert(SMEM_BYTES <= 227 * 1024, "dense tile too large");
};

// Persistent: the 22 KB of weights are fetched once per CTA (one bulk copy); the 4-channel input halo tile of the next
// tile is staged with cp.async as soon as the last expand stage of the current tile has consumed the buffer.
// conv1_9 inner loop: for every (input channel, kernel row) the 3x8 weights are read once and applied to all of the
// thread's pixel groups (IPT9 x 4 pixels x 8 output channels of accumulators).
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
dense_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts,
             int Hin, int Win, int Hout, int Wout, int tiles_x, int tiles_y, int total_tiles) {
    using G = typename C::G;
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t wbar;
    float* Xs = smem;
    float* Es = Xs + C::XS;
    float* Ds = Es;                    // alias: D is written after the last chunk's conv has consumed E
    float* Ws = Es + C::ES;
    const int tid = threadIdx.x;
    auto stage_tile = [&](int tile) {
        const int tx = tile % tiles_x;
        const int r = tile / tiles_x;
        const int oy0 = (r % tiles_y) * G::TH, ox0 = tx * G::TW, b = r / tiles_y;
        load_rect_async<G::IH, G::IW, G::IWS, NT>(Xs, x + (size_t)b * 4 * Hin * Win, 4, 4, Hin, Win, oy0 * 2 - 1, ox0 * 2 - 1);
        cp_async_commit();
    };
    int tile = blockIdx.x;
    if (tid == 0) { mbar_init(&wbar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tile < total_tiles) {
        if (tid == 0) { mbar_expect_tx(&wbar, C::WFLOATS * 4); bulk_load(Ws, wts, C::WFLOATS * 4, &wbar); }
        stage_tile(tile);
    }
    // conv1_9 thread geometry (fixed across tiles): channel group my_cg, pixel groups slot + i * PGT
    const bool conv_on = tid < 3 * C::PGT;
    const int my_cg = conv_on ? tid / C::PGT : 0;
    const int my_slot = conv_on ? tid - my_cg * C::PGT : 0;
    int it_row[C::IPT9], it_g[C::IPT9];
#pragma unroll
    for (int i = 0; i < C::IPT9; ++i) {
        const int p0 = (my_slot + i * C::PGT) * 4;
        it_row[i] = p0 / G::TW;
        it_g[i] = (p0 - it_row[i] * G::TW) >> 2;
    }
    for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
        const int tx = tile % tiles_x;
        const int rr = tile / tiles_x;
        const int oy0 = (rr % tiles_y) * G::TH, ox0 = tx * G::TW, tb = rr / tiles_y;
        const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
        const bool have_next = tile + (int)gridDim.x < total_tiles;

        float acc[C::IPT9][8][4];
#pragma unroll
        for (int i = 0; i < C::IPT9; ++i)
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][n][j] = 0.f;

        for (int c = 0; c < C::NCHUNK; ++c) {
            const float* Wc = Ws + c * C::CB;
            if (c == 0) cp_async_wait_all();
            __syncthreads();   // staged tile visible; E / D of the previous chunk / tile consumed
            if (c == 0 && it == 0) mbar_wait(&wbar, 0);
            pw_halo<G, 4, C::MC, (C::MC % 8 == 0 ? 8 : 4), NT, false>(Xs, Wc + C::OFF_W1, Wc + C::OFF_B1, Es, iy0, ix0, Hin, Win, nullptr, 0);
            __syncthreads();
            if (c == C::NCHUNK - 1 && have_next) stage_tile(tile + gridDim.x);     // X is dead: next tile flies during the conv
            if (conv_on) {
#pragma unroll 1
                for (int m = 0; m < C::MC; ++m) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const float* w = Wc + C::OFF_WC + my_cg * 24 + (m * 3 + dy) * 72;
                        float w8[3][8];
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float4 wa = ld4(w + dx * 8);
                            const float4 wb = ld4(w + dx * 8 + 4);
                            w8[dx][0] = wa.x; w8[dx][1] = wa.y; w8[dx][2] = wa.z; w8[dx][3] = wa.w;
                            w8[dx][4] = wb.x; w8[dx][5] = wb.y; w8[dx][6] = wb.z; w8[dx][7] = wb.w;
                        }
#pragma unroll
                        for (int i = 0; i < C::IPT9; ++i) {
                            float v[9];
                            load_window<G>(v, Es + m * G::IPIX + (it_row[i] * 2 + dy) * G::IWS, it_g[i]);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int n = 0; n < 8; n += 2)
#pragma unroll
                                    for (int j = 0; j < 4; ++j) ffma2(w8[dx][n], w8[dx][n + 1], v[2 * j + dx], acc[i][n][j], acc[i][n + 1][j]);
                        }
                    }
                }
            }
        }
        __syncthreads();       // every thread is done reading E before D overwrites it
        // conv1_9 bias + ReLU -> Ds[24][OPIX]
        if (conv_on) {
#pragma unroll
            for (int i = 0; i < C::IPT9; ++i) {
                const int pg = my_slot + i * C::PGT;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float b = Ws[C::OFF_B9 + my_cg * 8 + n];
                    st4(Ds + (my_cg * 8 + n) * G::OPIX + pg * 4,
                        make_float4(fmaxf(acc[i][n][0] + b, 0.f), fmaxf(acc[i][n][1] + b, 0.f),
                                    fmaxf(acc[i][n][2] + b, 0.f), fmaxf(acc[i][n][3] + b, 0.f)));
                }
            }
        }
        __syncthreads();
        float a3[C::IPT3][4][4];
#pragma unroll
        for (int i = 0; i < C::IPT3; ++i)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) a3[i][n][j] = 0.f;
        pw_accum<G::OPIX, 24, 8, 4, C::IPT3, NT>(Ds, Ws + C::OFF_W3, a3);
#pragma unroll
        for (int i = 0; i < C::IPT3; ++i) {
            const int item = tid + i * NT;
            if (item < C::NPG * 2) {
                const int cg = item / C::NPG;
                const int pg = item - cg * C::NPG;
                const int p0 = pg * 4;
                const int oy = p0 / G::TW, ox = p0 - oy * G::TW;
                const int gy = oy0 + oy, gx0 = ox0 + ox;
                if (gy < Hout) {
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const int ch = cg * 4 + n;
                        const float b = Ws[C::OFF_B3 + ch];
                        const float v[4] = {a3[i][n][0] + b, a3[i][n][1] + b, a3[i][n][2] + b, a3[i][n][3] + b};
                        store_px4(y + (((size_t)tb * 8 + ch) * Hout + gy) * Wout, gx0, Wout, v);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Plain 1x1 conv + bias (+ReLU) over a flat pixel tile (conv5_2, yolo_fastest.py:132,200).
// Packed weights: [W: K*N ([k][n])][b: N]
// ---------------------------------------------------------------------------------------------
template <int K_, int N_, int PIXT_, int PN_, int NT_, bool RELU_>
struct PwCfg {
    static constexpr int K = K_, N = N_, PIXT = PIXT_, PN = PN_, NT = NT_;
    static constexpr bool RELU = RELU_;
    static constexpr int WFLOATS = K * N + N;
    static constexpr int SMEM_FLOATS = K * PIXT + WFLOATS;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static constexpr int IPT = cdiv((PIXT / 4) * (N / PN), NT);
};

template <class C>
__global__ void __launch_bounds__(C::NT)
pw_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int HW, int tiles) {
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) float smem[];
    float* Xs = smem;
    float* Ws = Xs + C::K * C::PIXT;
    const int b = blockIdx.x / tiles;
    const int p0t = (blockIdx.x - b * tiles) * C::PIXT;
    const float* xb = x + (size_t)b * C::K * HW;
    copy_f4<NT>(Ws, wts, C::WFLOATS);
    pdl_trigger();
    pdl_wait();
    for (int idx = threadIdx.x; idx < C::K * C::PIXT; idx += NT) {
        const int k = idx / C::PIXT, p = idx - k * C::PIXT;
        Xs[idx] = (p0t + p < HW) ? __ldg(xb + (size_t)k * HW + p0t + p) : 0.f;
    }
    __syncthreads();
    float acc[C::IPT][C::PN][4];
#pragma unroll
    for (int it = 0; it < C::IPT; ++it)
#pragma unroll
        for (int n = 0; n < C::PN; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[it][n][i] = 0.f;
    pw_accum<C::PIXT, C::K, C::N, C::PN, C::IPT, NT>(Xs, Ws, acc);
    constexpr int NPG = C::PIXT / 4;
#pragma unroll
    for (int it = 0; it < C::IPT; ++it) {
        const int item = threadIdx.x + it * NT;
        if (item < NPG * (C::N / C::PN)) {
            const int cg = item / NPG;
            const int pg = item - cg * NPG;
#pragma unroll
            for (int n = 0; n < C::PN; ++n) {
                const int ch = cg * C::PN + n;
                const float bsv = Ws[C::K * C::N + ch];
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[i] = acc[it][n][i] + bsv;
                    if (C::RELU) v[i] = fmaxf(v[i], 0.f);
                }
                float* yp = y + ((size_t)b * C::N + ch) * HW;
                const int p = p0t + pg * 4;
                if (((HW & 3) == 0) && p + 3 < HW) st4(yp + p, make_float4(v[0], v[1], v[2], v[3]));
                else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (p + i < HW) yp[p + i] = v[i];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused upsample + concat + 1x1: deconv5_1 (ConvTranspose2d k2 s2 96->96 + BN + ReLU, yolo_fastest.py:42-48,140,208)
// -> cat((conv4_2[136], deconv5_1[96]), 1) (:209) -> conv4_1_1 (1x1 232->96 + ReLU, :142).
// The upsampled tensor and the concatenation never exist in HBM:
//   phase A  U[m][2y+py][2x+px] = relu(sum_c Wt[c][py][px][m] * P[c][y][x] + bt[m]) for the tile's parent
//            pixels — a 1x1 GEMM over parent pixels with one weight set per output parity — kept in smem;
//   phase B  out = relu(Wa . skip + Wb . U + b), K streamed in chunks: skip channels staged from HBM,
//            up channels read straight from the smem U tile.
// Packed weights: [Wab: (144 + 96) * 96 ([k][n]; rows 136..143 zero)][b: 96][Wt: 96(c) * 4(py*2+px) * 96(m)][bt: 96]
// Persistent; all weight chunks arrive by bulk copy (double buffered, mbarrier), the parent tile, the skip-channel
// chunks and the next tile's parent tile by cp.async, so both phases compute while their next operands fly.
// ---------------------------------------------------------------------------------------------
template <int TH_, int TW_, int KC_, int KCA_, int NT_, int MINB_>
struct UpCatCfg {
    static constexpr int TH = TH_, TW = TW_, KC = KC_, KCA = KCA_, NT = NT_, MINB = MINB_;
    static constexpr int CS = 136, CU = 96, N = 96;
    static constexpr int OPIX = TH * TW, PH = TH / 2, PW = TW / 2, PPIX = PH * PW;
    static constexpr int CSP = rup(CS, KC);
    static constexpr int OFF_WAB = 0, OFF_B = (CSP + CU) * N, OFF_WT = OFF_B + N, OFF_BT = OFF_WT + CU * 4 * CU;
    static constexpr int WFLOATS = OFF_BT + CU;
    static constexpr int PN = 8;
    static constexpr int IPT = cdiv((OPIX / 4) * (N / PN), NT);               // phase B items per thread
    static constexpr int IPTA = cdiv((PPIX / 4) * 4 * (CU / PN), NT);         // phase A items per thread
    static constexpr int NCA = CU / KCA;                                      // phase A chunks
    static constexpr int NCS = CSP / KC, NCB = NCS + CU / KC;                 // phase B chunks (skip part, total)
    static constexpr int PS = rup(CU * PPIX, 32);   // parent tile [96][PPIX]
    static constexpr int US = rup(CU * OPIX, 32);   // upsampled tile [96][OPIX]
    static constexpr int WB1 = rup(cmax(KCA * 4 * CU, KC * N), 32);    // one weight buffer
    static constexpr int BS1 = rup(KC * OPIX, 32);                      // one skip-chunk buffer
    static constexpr int SMEM_FLOATS = PS + US + 2 * WB1 + 2 * BS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
    static_assert(TH % 2 == 0 && TW % 4 == 0 && CU % KC == 0 && CU % KCA == 0 && KC % 4 == 0 && PPIX % 4 == 0, "bad upcat tiling");
    static_assert(SMEM_BYTES <= 227 * 1024, "upcat tile too large");
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB)
upcat_kernel(const float* __restrict__ skip /*[B,136,H,W]*/, const float* __restrict__ low /*[B,96,H/2,W/2]*/,
             float* __restrict__ y /*[B,96,H,W]*/, const float* __restrict__ wts, int H, int W, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    float* Ps = smem;
    float* Us = Ps + C::PS;
    float* Wb = Us + C::US;            // 2 weight buffers
    float* Bs = Wb + 2 * C::WB1;       // 2 skip-chunk buffers
    const int tid = threadIdx.x;
    const int Hl = H / 2, Wl = W / 2;
    auto origin = [&](int tile, int& b, int& oy0, int& ox0) {
        const int tx = tile % tiles_x;
        const int r = tile / tiles_x;
        oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW; b = r / tiles_y;
    };
    auto stage_parent = [&](int tile) {
        int b, oy0, ox0;
        origin(tile, b, oy0, ox0);
        load_rect_async<C::PH, C::PW, C::PW, NT>(Ps, low + (size_t)b * C::CU * Hl * Wl, C::CU, C::CU, Hl, Wl, oy0 / 2, ox0 / 2);
        cp_async_commit();
    };
    auto stage_skip = [&](int tile, int chunk, float* dst) {
        int b, oy0, ox0;
        origin(tile, b, oy0, ox0);
        load_rect_async<C::TH, C::TW, C::TW, NT>(dst, skip + ((size_t)b * C::CS + chunk * C::KC) * H * W, C::KC, C::CS - chunk * C::KC,
                                                 H, W, oy0, ox0);
        cp_async_commit();
    };
    // weight chunk sequence of one tile: NCA phase-A blocks of Wt, then NCB phase-B row blocks of Wab
    auto issue_w = [&](int s, int buf) {     // one thread
        const float* src = s < C::NCA ? wts + C::OFF_WT + (size_t)s * C::KCA * 4 * C::CU : wts + C::OFF_WAB + (size_t)(s - C::NCA) * C::KC * C::N;
        const uint32_t bytes = (s < C::NCA ? C::KCA * 4 * C::CU : C::KC * C::N) * 4;
        mbar_expect_tx(&bars[buf], bytes);
        bulk_load(Wb + buf * C::WB1, src, bytes, &bars[buf]);
    };
    int tile = blockIdx.x;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (tile < total_tiles) {
        if (tid == 0) issue_w(0, 0);
        stage_parent(tile);
    }
    uint32_t q = 0;
    for (; tile < total_tiles; tile += gridDim.x) {
        int tb, oy0, ox0;
        origin(tile, tb, oy0, ox0);
        const bool have_next = tile + (int)gridDim.x < total_tiles;
        // ---- phase A: item = 4 consecutive parent pixels x one output parity x 8 up channels ----------------
        {
            constexpr int NPG = C::PPIX / 4, NCG = C::CU / C::PN;
            float u[C::IPTA][C::PN][4];
#pragma unroll
            for (int it = 0; it < C::IPTA; ++it)
#pragma unroll
                for (int n = 0; n < C::PN; ++n)
#pragma unroll
                    for (int i = 0; i < 4; ++i) u[it][n][i] = 0.f;
            for (int ca = 0; ca < C::NCA; ++ca, ++q) {
                if (ca == 0) cp_async_wait_all();                        // parent tile landed
                __syncthreads();                                         // previous chunk / tile consumed
                if (ca == 0) stage_skip(tile, 0, Bs);                    // first skip chunk flies during phase A
                if (tid == 0) issue_w(ca + 1, (int)((q + 1) & 1));       // ca + 1 == NCA is the first phase-B block
                mbar_wait(&bars[q & 1], (q >> 1) & 1);
                const float* Wt = Wb + (q & 1) * C::WB1;
#pragma unroll
                for (int it = 0; it < C::IPTA; ++it) {
                    const int item = tid + it * NT;
                    if (item < NPG * 4 * NCG) {
                        const int pg = item % NPG;
                        const int par = (item / NPG) & 3;
                        const int cg = item / (NPG * 4);
                        const float* pp = Ps + (size_t)ca * C::KCA * C::PPIX + pg * 4;
                        const float* wp = Wt + par * C::CU + cg * C::PN;
#pragma unroll
                        for (int k = 0; k < C::KCA; ++k) {
                            const float4 pv = ld4(pp + k * C::PPIX);
                            const float p4[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                            for (int n4 = 0; n4 < C::PN / 4; ++n4) {
                                const float4 w = ld4(wp + k * 4 * C::CU + n4 * 4);
                                const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                                    for (int i = 0; i < 4; ++i) u[it][n4 * 4 + qq][i] = fmaf(w4[qq], p4[i], u[it][n4 * 4 + qq][i]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < C::IPTA; ++it) {
                const int item = tid + it * NT;
                if (item < NPG * 4 * NCG) {
                    const int pg = item % NPG;
                    const int par = (item / NPG) & 3;
                    const int cg = item / (NPG * 4);
#pragma unroll
                    for (int n = 0; n < C::PN; ++n) {
                        const int mch = cg * C::PN + n;
                        const float b = __ldg(wts + C::OFF_BT + mch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = pg * 4 + i;
                            const int py = p / C::PW, px = p - py * C::PW;
                            Us[mch * C::OPIX + (2 * py + (par >> 1)) * C::TW + 2 * px + (par & 1)] = fmaxf(u[it][n][i] + b, 0.f);
                        }
                    }
                }
            }
        }
        // ---- phase B: 1x1 over the concatenation, K in blocks of KC ------------------------------------------
        float acc[C::IPT][C::PN][4];
#pragma unroll
        for (int it = 0; it < C::IPT; ++it)
#pragma unroll
            for (int n = 0; n < C::PN; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[it][n][i] = 0.f;
        for (int cb = 0; cb < C::NCB; ++cb, ++q) {
            if (cb <= C::NCS) cp_async_wait_all();     // skip chunk cb (and, at cb == 0, nothing else pending) has landed
            __syncthreads();                           // previous block consumed; at cb == 0: U complete, parent tile dead
            if (cb == 0 && have_next) stage_parent(tile + gridDim.x);
            if (cb + 1 < C::NCS) stage_skip(tile, cb + 1, Bs + ((cb + 1) & 1) * C::BS1);
            if (tid == 0 && (cb + 1 < C::NCB || have_next)) issue_w(cb + 1 < C::NCB ? C::NCA + cb + 1 : 0, (int)((q + 1) & 1));
            mbar_wait(&bars[q & 1], (q >> 1) & 1);
            const float* src = cb < C::NCS ? Bs + (cb & 1) * C::BS1 : Us + (size_t)(cb - C::NCS) * C::KC * C::OPIX;
            pw_accum<C::OPIX, C::KC, C::N, C::PN, C::IPT, NT>(src, Wb + (q & 1) * C::WB1, acc);
        }
        constexpr int NPG = C::OPIX / 4;
#pragma unroll
        for (int it = 0; it < C::IPT; ++it) {
            const int item = tid + it * NT;
            if (item < NPG * (C::N / C::PN)) {
                const int cg = item / NPG;
                const int pg = item - cg * NPG;
                const int p0 = pg * 4;
                const int oy = p0 / C::TW, ox = p0 - oy * C::TW;
                const int gy = oy0 + oy, gx0 = ox0 + ox;
                if (gy < H) {
#pragma unroll
                    for (int n = 0; n < C::PN; ++n) {
                        const int ch = cg * C::PN + n;
                        const float b = __ldg(wts + C::OFF_B + ch);
                        const float v[4] = {fmaxf(acc[it][n][0] + b, 0.f), fmaxf(acc[it][n][1] + b, 0.f),
                                            fmaxf(acc[it][n][2] + b, 0.f), fmaxf(acc[it][n][3] + b, 0.f)};
                        store_px4(y + (((size_t)tb * C::N + ch) * H + gy) * W, gx0, W, v);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Two chained 1x1 convs: relu(W1 x + b1) -> W2 . + b2 (linear). YoloFastest_lite's forward goes conv3_2 -> conv3_4 directly (it skips
// the depthwise conv3_3, yolo_fastest.py:335-337), so that group has no spatial stage at all: one thread = 4 consecutive pixels, the
// K inputs and N accumulators live in registers, the mid channels are produced and consumed one at a time; weights in smem.
// Packed weights: [W1: K x M (k-major)][b1: M][W2: M x N (m-major)][b2: N]
// ---------------------------------------------------------------------------------------------
template <int K_, int M_, int N_, int NT_>
struct PwPwCfg {
    static constexpr int K = K_, M = M_, N = N_, NT = NT_;
    static constexpr int OFF_W1 = 0, OFF_B1 = K * M, OFF_W2 = OFF_B1 + M, OFF_B2 = OFF_W2 + M * N;
    static constexpr int WFLOATS = rup(OFF_B2 + N, 4);
    static_assert(N % 4 == 0 && M % 2 == 0, "tiling");
};

template <class C>
__global__ void __launch_bounds__(C::NT)
pwpw_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int HW, long long total_groups) {
    __shared__ __align__(16) float Ws[C::WFLOATS];
    for (int i = threadIdx.x; i < C::WFLOATS; i += C::NT) Ws[i] = __ldg(wts + i);
    __syncthreads();
    const int gpi = HW / 4;                                    // pixel groups per image (HW % 4 == 0)
    for (long long gidx = (long long)blockIdx.x * C::NT + threadIdx.x; gidx < total_groups; gidx += (long long)gridDim.x * C::NT) {
        const int b = (int)(gidx / gpi);
        const int p0 = (int)(gidx - (long long)b * gpi) * 4;
        const float* xb = x + (size_t)b * C::K * HW + p0;
        float xv[C::K][4];
#pragma unroll
        for (int k = 0; k < C::K; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)k * HW));
            xv[k][0] = v.x; xv[k][1] = v.y; xv[k][2] = v.z; xv[k][3] = v.w;
        }
        float acc[C::N][4];
#pragma unroll
        for (int n = 0; n < C::N; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[n][i] = Ws[C::OFF_B2 + n];
#pragma unroll 2
        for (int m = 0; m < C::M; ++m) {
            float e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = Ws[C::OFF_B1 + m];
#pragma unroll
            for (int k = 0; k < C::K; ++k) {
                const float w = Ws[C::OFF_W1 + k * C::M + m];
#pragma unroll
                for (int i = 0; i < 4; ++i) e[i] = fmaf(w, xv[k][i], e[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = fmaxf(e[i], 0.f);
#pragma unroll
            for (int n4 = 0; n4 < C::N / 4; ++n4) {
                const float4 w = ld4(Ws + C::OFF_W2 + m * C::N + 4 * n4);
                const float w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 4; q += 2)
#pragma unroll
                    for (int i = 0; i < 4; ++i) ffma2(w4[q], w4[q + 1], e[i], acc[4 * n4 + q][i], acc[4 * n4 + q + 1][i]);
            }
        }
        float* yb = y + (size_t)b * C::N * HW + p0;
#pragma unroll
        for (int n = 0; n < C::N; ++n) st4(yb + (size_t)n * HW, make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]));
    }
}

// u8 -> (x - 128) / 255 fp32 (detect.py:123-124) for callers that want the normalised tensor itself.
__global__ void u8_normalize_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = ((float)src[i] - 128.0f) / 255.0f;
}

}  // namespace yf
