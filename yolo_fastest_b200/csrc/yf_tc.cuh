// yf_tc.cuh — tensor-core (tcgen05) variant of the inverted-residual engine for the wide blocks, where the two
// 1x1 convolutions are real dense contractions (K = 16..48 in, N = 96..224 mid, yolo_fastest.py:52-66).
//
//   S1  E[halo px][MC]  = X[halo px][CIN] . W1[CIN][MC]      tcgen05.mma kind::tf32, M = 128 pixels per MMA tile
//   dw  D = relu(dw3x3(relu(E + b1)) + bd)                     CUDA cores, sliding window (as in yf_kernels.cuh)
//   S3  O[out px][COUT] += D[out px][MC] . W2[MC][COUT]        tcgen05.mma, accumulators stay in TMEM across the chunks
//
// fp32 parity is kept with the 3xTF32 split: every operand v is stored as hi = tf32(v) and lo = v - hi and each
// product is issued as hi*hi + hi*lo + lo*hi (fp32 accumulation in TMEM); measured error 8e-7 relative
// (tools/selftest/umma_selftest.cu), i.e. fp32 grade, where a single TF32 pass would give 1e-3.
//
// Operand layouts (validated by the selftest):
//   A = activations, MN-major (pixels contiguous) — for 32-bit MN-major operands the only legal shared-memory layout
//       is SWIZZLE_128B_BASE32B: atoms of [4 channels][32 pixels] fp32 = 512 B, the 32-byte chunk index XOR-ed with
//       (channel & 3); two atoms per MMA (K = 8), LBO = stride between 32-pixel atoms, SBO = stride between 4-channel atoms.
//   B = weights, K-major, no swizzle, packed on the host as [n/8][k/4][n%8][k%4] (core matrices of 8 rows x 16 bytes).
//   D = TMEM, lane = pixel, column = output channel; the epilogues read it with tcgen05.ld.32x32b (thread = pixel).
#pragma once
#include "yf_kernels.cuh"

namespace yf {

__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
// float index of activation element (pixel m, channel k) in an A operand region; kblk = floats per 8-channel block
__device__ __forceinline__ int a_idx(int m, int k, int kblk) {
    return (k >> 3) * kblk + (m >> 5) * 256 + ((k >> 2) & 1) * 128 + (k & 3) * 32 + ((((m & 31) >> 3) ^ (k & 3)) << 3) + (m & 7);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}
// instruction descriptor: D = f32, A = B = tf32, A MN-major, B K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int pow2_ge(int v) { int p = 32; while (p < v) p <<= 1; return p; }

template <int CIN_, int CMID_, int COUT_, int TH_, int TW_, int MC_, int RH_, int NT_, bool RES_>
struct IrbTcCfg {
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, MC = MC_, RH = RH_, NT = NT_;
    static constexpr bool RES = RES_;
    using G = Geo<3, 1, TH_, TW_>;
    static constexpr int CMIDP = rup(CMID, MC), NCHUNK = CMIDP / MC;
    static constexpr int MT1 = cdiv(G::IPIX, 128), MT3 = cdiv(G::OPIX, 128);      // 128-pixel MMA tiles of the halo / output tile
    static constexpr int NP3 = rup(COUT, 16);                                     // N of the project MMA (M = 128 needs N % 16 == 0)
    static constexpr int KB1 = MT1 * 1024, KB3 = MT3 * 1024;                      // floats per 8-channel block of an A region
    static constexpr int XA1 = (CIN / 8) * KB1, DA1 = (MC / 8) * KB3;             // floats of one A region (hi or lo)
    static constexpr int TM_E = 0, TM_O = MT1 * MC;                               // TMEM columns: expand accumulators, project accumulators
    static constexpr int TCOLS = pow2_ge(TM_O + MT3 * NP3);
    // weight block of one chunk (floats)
    static constexpr int OFF_W1H = 0, OFF_W1L = MC * CIN, OFF_B1 = 2 * MC * CIN, OFF_WD = OFF_B1 + MC, OFF_BD = OFF_WD + MC * 9;
    static constexpr int OFF_W2H = OFF_BD + MC, OFF_W2L = OFF_W2H + NP3 * MC, CB = OFF_W2L + NP3 * MC;
    static constexpr int OFF_B2 = NCHUNK * CB;
    static constexpr int WFLOATS = OFF_B2 + COUT;
    static constexpr int XRAW = rup(CIN * G::IPIX, 32), ES = rup(MC * G::IPIX, 32), WS1 = rup(CB, 32);
    static constexpr int SMEM_FLOATS = 2 * XA1 + 2 * DA1 + XRAW + ES + 2 * WS1;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;     // + slack to align the dynamic window to 1 KB
    static_assert(CIN % 8 == 0 && MC % 16 == 0 && NT % 128 == 0 && CB % 4 == 0, "tcgen05 tiling constraints");
    static_assert(TCOLS <= 512, "TMEM columns");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
};

// depthwise 3x3 s1 + bias + ReLU from E [MC][halo] into the A-operand layout of the project MMA (hi and lo parts)
template <class G, int MC, int RH, int NT, int KB3>
__device__ __forceinline__ void dw_stage_split(const float* __restrict__ Es, const float* __restrict__ Wd, const float* __restrict__ bd,
                                               float* __restrict__ DAhi, float* __restrict__ DAlo) {
    static_assert(G::TH % RH == 0 && G::S == 1 && G::KS == 3, "3x3 stride 1");
    constexpr int NSTRIP = G::TW / 4, NSEG = G::TH / RH;
    constexpr int NITEM = MC * NSEG * NSTRIP;
    for (int item = threadIdx.x; item < NITEM; item += NT) {
        const int m = item / (NSEG * NSTRIP);
        const int rem = item - m * (NSEG * NSTRIP);
        const int seg = rem / NSTRIP;
        const int g = rem - seg * NSTRIP;
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = Wd[m * 9 + t];
        const float b = bd[m];
        const float* e = Es + m * G::IPIX + (seg * RH) * G::IWS;
        float win[3][6];
#pragma unroll
        for (int dd = 0; dd < 2; ++dd) load_window<G>(win[dd + 1], e + dd * G::IWS, g);
#pragma unroll
        for (int oy = 0; oy < RH; ++oy) {
#pragma unroll
            for (int dd = 0; dd < 2; ++dd)
#pragma unroll
                for (int v = 0; v < 6; ++v) win[dd][v] = win[dd + 1][v];
            load_window<G>(win[2], e + (oy + 2) * G::IWS, g);
            float a[4] = {b, b, b, b};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = fmaf(w[dy * 3 + dx], win[dy][i + dx], a[i]);
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = fmaxf(a[i], 0.f);
                hi[i] = tf32_hi(v);
                lo[i] = v - hi[i];
            }
            const int p0 = (seg * RH + oy) * G::TW + 4 * g;          // 4 consecutive output pixels stay inside one 32-byte chunk
            const int o = a_idx(p0, m, KB3);
            st4(DAhi + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
            st4(DAlo + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
irbtc_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
             int tiles_x, int tiles_y, int total_tiles) {
    using G = typename C::G;
    constexpr int NT = C::NT, NW = NT / 32;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* XAhi = base;
    float* XAlo = XAhi + C::XA1;
    float* DAhi = XAlo + C::XA1;
    float* DAlo = DAhi + C::DA1;
    float* Xraw = DAlo + C::DA1;
    float* Es = Xraw + C::XRAW;
    float* Ws = Es + C::ES;
    __shared__ __align__(8) uint64_t wbar[2], mbar1, mbar3;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, wgrp = warp >> 2;     // TMEM lane quarter this warp may access; warps sharing a quarter split the MMA tiles
    constexpr int NGRP = NW / 4;

    auto origin = [&](int tile, int& b, int& oy0, int& ox0) {
        const int tx = tile % tiles_x;
        const int r = tile / tiles_x;
        oy0 = (r % tiles_y) * G::TH; ox0 = tx * G::TW; b = r / tiles_y;
    };
    auto stage_tile = [&](int tile) {
        int b, oy0, ox0;
        origin(tile, b, oy0, ox0);
        load_rect_async<G::IH, G::IW, G::IWS, NT>(Xraw, x + (size_t)b * C::CIN * H * W, C::CIN, C::CIN, H, W, oy0 - 1, ox0 - 1);
        cp_async_commit();
    };
    auto issue_w = [&](int chunk, int buf) {
        mbar_expect_tx(&wbar[buf], C::CB * 4);
        bulk_load(Ws + buf * C::WS1, wts + (size_t)chunk * C::CB, C::CB * 4, &wbar[buf]);
    };

    if (tid == 0) {
        mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); mbar_init(&mbar1, 1); mbar_init(&mbar3, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the pad pixels (m >= IPIX / OPIX) of the operand regions are never written again: zero everything once
    for (int i = tid * 4; i < 2 * C::XA1 + 2 * C::DA1; i += NT * 4) st4(XAhi + i, make_float4(0.f, 0.f, 0.f, 0.f));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;

    int tile = blockIdx.x;
    if (tile < total_tiles) {
        if (tid == 0) issue_w(0, 0);
        stage_tile(tile);
    }
    constexpr uint32_t IDESC1 = umma_idesc_tf32(C::MC), IDESC3 = umma_idesc_tf32(C::NP3);
    uint32_t q = 0, ph1 = 0, ph3 = 0;
    for (; tile < total_tiles; tile += gridDim.x) {
        int tb, oy0, ox0;
        origin(tile, tb, oy0, ox0);
        const int iy0 = oy0 - 1, ix0 = ox0 - 1;
        const bool have_next = tile + (int)gridDim.x < total_tiles;
        cp_async_wait_all();
        __syncthreads();                       // raw tile landed; the previous tile is completely done
        for (int idx = tid; idx < C::CIN * G::IPIX; idx += NT) {
            const int k = idx / G::IPIX, m = idx - k * G::IPIX;
            const float v = Xraw[idx];
            const float hi = tf32_hi(v);
            const int o = a_idx(m, k, C::KB1);
            XAhi[o] = hi;
            XAlo[o] = v - hi;
        }
        fence_proxy_async();
        __syncthreads();                       // operand A of the expand MMA is visible to the tensor core; Xraw is free
        if (have_next) stage_tile(tile + gridDim.x);

        for (int c = 0; c < C::NCHUNK; ++c, ++q) {
            const int wbuf = (int)(q & 1);
            const float* Wc = Ws + wbuf * C::WS1;
            mbar_wait(&wbar[wbuf], (q >> 1) & 1);
            // ---- S1: expand on the tensor core ---------------------------------------------------------------
            if (tid == 0) {
                tc_fence_after();
                const uint32_t w1h = smem_u32(Wc + C::OFF_W1H), w1l = smem_u32(Wc + C::OFF_W1L);
                const uint32_t xh = smem_u32(XAhi), xl = smem_u32(XAlo);
#pragma unroll 1
                for (int mt = 0; mt < C::MT1; ++mt) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a0 = (pass == 2 ? xl : xh) + mt * 4096;
                        const uint32_t b0 = (pass == 1 ? w1l : w1h);
#pragma unroll
                        for (int kb = 0; kb < C::CIN / 8; ++kb)
                            umma_tf32(tmem + C::TM_E + mt * C::MC, umma_desc(a0 + kb * C::KB1 * 4, 1024, 512, 1),
                                      umma_desc(b0 + kb * 256, 128, (C::CIN / 4) * 128, 0), IDESC1, (pass | kb) ? 1u : 0u);
                    }
                }
                umma_commit(&mbar1);
            }
            mbar_wait(&mbar1, ph1); ph1 ^= 1;
            tc_fence_after();
            // ---- S1 epilogue: TMEM -> bias + ReLU + zero outside the image -> E[ch][halo pixel] ------------------
            for (int mt = wgrp; mt < C::MT1; mt += NGRP) {
                const int pix = mt * 128 + quarter * 32 + lane;
                const int r = pix / G::IWS, j = pix - r * G::IWS;
                const bool ok = pix < G::IPIX && j < G::IW && (unsigned)(iy0 + r) < (unsigned)H && (unsigned)(ix0 + j) < (unsigned)W;
#pragma unroll
                for (int c0 = 0; c0 < C::MC; c0 += 16) {
                    float v[16];
                    tmem_ld16(tmem + lane_base + C::TM_E + mt * C::MC + c0, v);
                    if (pix < G::IPIX) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) Es[(c0 + i) * G::IPIX + pix] = ok ? fmaxf(v[i] + Wc[C::OFF_B1 + c0 + i], 0.f) : 0.f;
                    }
                }
            }
            tc_fence_before();
            __syncthreads();                   // E complete; the expand accumulators may be overwritten by the next chunk
            if (c > 0) { mbar_wait(&mbar3, ph3); ph3 ^= 1; }      // project MMA of the previous chunk has consumed D and its weights
            if (tid == 0 && (c + 1 < C::NCHUNK || have_next)) issue_w(c + 1 < C::NCHUNK ? c + 1 : 0, (int)((q + 1) & 1));
            // ---- depthwise on CUDA cores, output split into the project MMA's operand ------------------------------
            dw_stage_split<G, C::MC, C::RH, NT, C::KB3>(Es, Wc + C::OFF_WD, Wc + C::OFF_BD, DAhi, DAlo);
            fence_proxy_async();
            __syncthreads();                   // D complete and visible to the tensor core
            // ---- S3: project, accumulating over the chunks in TMEM -------------------------------------------------
            if (tid == 0) {
                tc_fence_after();
                const uint32_t w2h = smem_u32(Wc + C::OFF_W2H), w2l = smem_u32(Wc + C::OFF_W2L);
                const uint32_t dh = smem_u32(DAhi), dl = smem_u32(DAlo);
#pragma unroll 1
                for (int mt = 0; mt < C::MT3; ++mt) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a0 = (pass == 2 ? dl : dh) + mt * 4096;
                        const uint32_t b0 = (pass == 1 ? w2l : w2h);
#pragma unroll
                        for (int kb = 0; kb < C::MC / 8; ++kb)
                            umma_tf32(tmem + C::TM_O + mt * C::NP3, umma_desc(a0 + kb * C::KB3 * 4, 1024, 512, 1),
                                      umma_desc(b0 + kb * 256, 128, (C::MC / 4) * 128, 0), IDESC3, (c | pass | kb) ? 1u : 0u);
                    }
                }
                umma_commit(&mbar3);
            }
        }
        mbar_wait(&mbar3, ph3); ph3 ^= 1;
        tc_fence_after();
        // ---- output epilogue: TMEM -> + bias (+ residual, yolo_fastest.py:65) -> HBM -------------------------------------
        for (int mt = wgrp; mt < C::MT3; mt += NGRP) {
            const int pix = mt * 128 + quarter * 32 + lane;
            const int oy = pix / G::TW, ox = pix - oy * G::TW;
            const int gy = oy0 + oy, gx = ox0 + ox;
            const bool ok = pix < G::OPIX && gy < H && gx < W;
#pragma unroll
            for (int c0 = 0; c0 < C::COUT; c0 += 16) {
                float v[16];
                tmem_ld16(tmem + lane_base + C::TM_O + mt * C::NP3 + c0, v);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (c0 + i < C::COUT) {
                            const size_t o = (((size_t)tb * C::COUT + c0 + i) * H + gy) * W + gx;
                            float r = v[i] + __ldg(wts + C::OFF_B2 + c0 + i);
                            if (C::RES) r += __ldg(x + o);
                            y[o] = r;
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
}

}  // namespace yf
