// yf_tcirb2.cuh — tensor-core inverted-residual engine for the WIDEST blocks (res5_*: 48 -> 224 -> 48 at 1/32 resolution,
// yolo_fastest.py:52-66,125-130), where yf_tc.cuh does not fit: the split input operand alone is 96 KB. Same step machine as
// yf_tcup.cuh — everything is chunked by 16 channels and one weight-ring slot feeds one step:
//
//   A  E[halo px][mid half] = X[halo px][CIN] . W1      K = CIN in chunks of 16, N = CMID / NH per half (TMEM holds one half of E)
//   per 16 mid channels:  TMEM E -> + b1, ReLU, zero outside the image -> smem (double buffered) -> depthwise 3x3 -> split operand
//   B  O[out px][COUT] += D[out px][16] . W2             accumulated over all CMID / 16 chunks in TMEM
//   out = O + b2 + x (residual, yolo_fastest.py:65)
//
// Step order of a tile: A(half 0) x CIN/16 | CMID/NH/16 mid chunks | A(half 1) x CIN/16 | mid chunks | ...; the X operand (hi and lo
// of all input channels, resident) of the next tile is written when the last expand GEMM of this tile is complete.
// Packed weights (floats): STEPS slots of SLOT floats in step order { A: [W1hi: NA x 16 K-major][W1lo] | mid: [W2hi: COUTP x 16][W2lo]
// [b1: 16][Wd: 16*9][bd: 16] }, then [b2: COUT].
#pragma once
#include "yf_tc.cuh"

namespace yf {

template <int CIN_, int CMID_, int COUT_, int TH_, int TW_, int NH_, int RH_, int NWW_, bool RES_>
struct IrbTc2Cfg {
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, NH = NH_, RH = RH_, NWW = NWW_;
    static constexpr bool RES = RES_;
    static constexpr int NTW = NWW * 32, NT = NTW + 32, NWB = 3, MC = 16;
    using G = Geo<3, 1, TH_, TW_>;
    static constexpr int CMIDP = rup(CMID, MC * NH);
    static constexpr int NA = CMIDP / NH;                                   // N of the expand GEMM (one half)
    static constexpr int NMC = NA / MC;                                     // mid chunks per half
    static constexpr int NKA = CIN / MC;                                    // K chunks of the expand GEMM
    static constexpr int COUTP = rup(COUT, 16);
    static constexpr int MT1 = cdiv(G::IPIX, 128), MT3 = cdiv(G::OPIX, 128);
    static constexpr int NG1 = cdiv(G::IPIX, 32), NG3 = cdiv(G::OPIX, 32);
    static constexpr int KB1 = NG1 * 256, KB3 = NG3 * 256;
    static constexpr int XA1 = (CIN / 8) * KB1;                             // resident X operand (hi or lo), floats
    static constexpr int DA1 = (MC / 8) * KB3;                              // project operand chunk (hi or lo)
    static constexpr int EPS = e_stride(G::IPIX, G::TW / 4), ES1 = rup(MC * EPS, 32);
    static constexpr int STEPS = NH * (NKA + NMC), DSTEPS = NH * NMC;
    static constexpr int OFF_W2 = 0, OFF_B1 = 2 * COUTP * MC, OFF_WD = OFF_B1 + MC, OFF_BD = OFF_WD + MC * 9;
    static constexpr int SLOT = rup(cmax(2 * NA * MC, OFF_BD + MC), 32);
    static constexpr int OFF_B2 = STEPS * SLOT;
    static constexpr int WFLOATS = OFF_B2 + COUT;
    // project accumulators per 128-pixel tile: [hi.hi of even chunks | hi.hi of odd chunks | both correction terms of all chunks].
    // The tensor core accumulates with truncation, so the error of an accumulator grows with its MMA chain: the main term gets two
    // short chains (CMID / 32 steps each) and never shares an accumulator with the small terms; the three are summed in the epilogue
    static constexpr int NACC = 3;
    static constexpr int TM_O = 0, TM_E = MT3 * NACC * COUTP, TCOLS = pow2_ge(TM_E + MT1 * NA);
    static constexpr int SMEM_FLOATS = 4 * DA1 + 2 * XA1 + 2 * ES1 + NWB * SLOT;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static constexpr int NITEM_X = (CIN / 8) * G::IPIX;
    static_assert(CIN % MC == 0 && NA % 16 == 0 && NA <= 256 && CMIDP % (MC * NH) == 0, "tiling");
    static_assert(TCOLS <= 512, "TMEM columns");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
irbtc2_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
              int tiles_x, int tiles_y, int total_tiles) {
    using G = typename C::G;
    constexpr int NTW = C::NTW, NWW = C::NWW;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Dbuf = base;                          // [2 buffers][hi | lo][DA1]
    float* Xbuf = Dbuf + 4 * C::DA1;             // [hi | lo][XA1]
    float* Es0 = Xbuf + 2 * C::XA1;              // [2 buffers][MC][EPS]
    float* Ws = Es0 + 2 * C::ES1;                // weight ring
    __shared__ __align__(8) uint64_t wbar[C::NWB], sdone[C::NWB], dfull[2], dfree[2], xfull, efull, efree, ofull, ofree;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < C::NWB; ++i) { mbar_init(&wbar[i], 1); mbar_init(&sdone[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], NWW); mbar_init(&dfree[i], 1); }
        mbar_init(&xfull, NWW); mbar_init(&efull, 1); mbar_init(&efree, NWW); mbar_init(&ofull, 1); mbar_init(&ofree, NWW);
        mbar_fence_init();
    }
    if (warp == NWW) {
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

    if (warp == NWW) {
        // ================= tensor-core warp =================
        if (ntile > 0 && elect_one()) {
            constexpr uint32_t IDESC_A = umma_idesc_tf32(C::NA), IDESC_B = umma_idesc_tf32(C::COUTP);
            const int S = ntile * C::STEPS;
            const uint64_t dd0 = umma_desc(smem_u32(Dbuf), 1024, 512, 1);
            const uint64_t dx0 = umma_desc(smem_u32(Xbuf), 1024, 512, 1);
            const uint64_t dwa = umma_desc(smem_u32(Ws), 128, (C::MC / 4) * 128, 0);
            int q = 0, d = 0;                    // weight-ring step, operand-chunk step (both global)
            auto issue_w = [&](int s) {
                const int slot = s % C::NWB;
                mbar_expect_tx(&wbar[slot], C::SLOT * 4);
                bulk_load(Ws + slot * C::SLOT, wts + (size_t)(s % C::STEPS) * C::SLOT, C::SLOT * 4, &wbar[slot]);
            };
            auto step_begin = [&]() -> uint64_t {
                mbar_wait(&wbar[q % C::NWB], (q / C::NWB) & 1);
                return dwa + (uint64_t)(((uint32_t)(q % C::NWB) * C::SLOT * 4) >> 4);
            };
            auto step_end = [&]() {
                umma_commit(&sdone[q % C::NWB]);
                if (q + 2 < S) {
                    if (q >= 1) mbar_wait(&sdone[(q - 1) % C::NWB], ((q - 1) / C::NWB) & 1);
                    issue_w(q + 2);
                }
                ++q;
            };
            auto gemm_a = [&]() {                         // E (one half of the mid channels) = X . W1 over CIN / 16 K chunks
#pragma unroll 1
                for (int kc = 0; kc < C::NKA; ++kc) {
                    const uint64_t wb = step_begin();
                    tc_fence_after();
#pragma unroll 1
                    for (int mt = 0; mt < C::MT1; ++mt) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                            for (int kb = 0; kb < C::MC / 8; ++kb)
                                umma_tf32(tmem + C::TM_E + mt * C::NA,
                                          dx0 + (uint64_t)(mt * 256) + (uint64_t)(((pass == 2 ? C::XA1 : 0) + (kc * (C::MC / 8) + kb) * C::KB1) * 4 / 16),
                                          wb + (uint64_t)(((pass == 1 ? C::NA * C::MC : 0) * 4 + kb * 256) / 16), IDESC_A, (kc | pass | kb) ? 1u : 0u);
                    }
                    step_end();
                }
                umma_commit(&efull);
            };
            auto gemm_b_chunk = [&](int k) {              // O += D chunk . W2 chunk (k = chunk index inside the tile)
                const int b = d & 1;
                mbar_wait(&dfull[b], (d >> 1) & 1);
                const uint64_t wb = step_begin();
                tc_fence_after();
                const uint64_t db = dd0 + (uint64_t)((uint32_t)(b * 2 * C::DA1 * 4) >> 4);
#pragma unroll 1
                for (int mt = 0; mt < C::MT3; ++mt) {
                    const uint32_t o0 = tmem + C::TM_O + mt * C::NACC * C::COUTP;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                        for (int kb = 0; kb < C::MC / 8; ++kb)
                            umma_tf32(o0 + (pass == 0 ? (k & 1) : 2) * C::COUTP,
                                      db + (uint64_t)(mt * 256) + (uint64_t)(((pass == 2 ? C::DA1 : 0) + kb * C::KB3) * 4 / 16),
                                      wb + (uint64_t)(((pass == 1 ? C::COUTP * C::MC : 0) * 4 + kb * 256) / 16), IDESC_B,
                                      (kb == 0 && ((pass == 0 && k < 2) || (pass == 1 && k == 0))) ? 0u : 1u);
                }
                umma_commit(&dfree[b]);
                step_end();
                ++d;
            };
            issue_w(0);
            issue_w(1);
            int eh = 0;                          // expand-half counter (global): efull / efree phase
            for (int ti = 0; ti < ntile; ++ti) {
                mbar_wait(&xfull, ti & 1);
                for (int h = 0; h < C::NH; ++h, ++eh) {
                    if (eh > 0) mbar_wait(&efree, (eh - 1) & 1);               // the previous half of E has been read out of TMEM
                    gemm_a();
                    if (h == 0 && ti > 0) mbar_wait(&ofree, (ti - 1) & 1);     // the previous tile's output has been read out
                    for (int c = 0; c < C::NMC; ++c) gemm_b_chunk(h * C::NMC + c);
                }
                umma_commit(&ofull);
            }
        }
    } else {
        // ================= worker warps =================
        const int quarter = warp & 3, wq = warp >> 2;
        const int nwq = (NWW - quarter + 3) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int tpi = tiles_x * tiles_y;
        const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;
        auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            b = tile / tpi;
            const int t = tile - b * tpi;
            const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
            oy0 = ty * G::TH; ox0 = (t - ty * tiles_x) * G::TW;
        };
        const size_t plane = (size_t)H * W;
        // input halo tile of all CIN channels -> resident split operand (item = 8 channels of one halo pixel; zero outside the image)
        auto put_x = [&](int b, int oy0, int ox0) {
            const float* xb = x + (size_t)b * C::CIN * plane;
            for (int item = tid; item < C::NITEM_X; item += NTW) {
                const int kh = item / G::IPIX, m = item - kh * G::IPIX;
                const int r = m / G::IWS, j = m - r * G::IWS;
                const int gy = oy0 - 1 + r, gx = ox0 - 1 + j;
                const bool ok = j < G::IW && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                const float* px = xb + (size_t)(kh * 8) * plane + (ok ? (size_t)gy * W + gx : 0);
                float v[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) v[kk] = __ldg(px + kk * plane);
                const int ob = kh * C::KB1 + (m >> 5) * 256 + (m & 7), mc = (m & 31) >> 3;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float vv = ok ? v[kk] : 0.f;
                    const float hi = tf32_hi(vv);
                    const int o = ob + ((kk >> 2) & 1) * 128 + (kk & 3) * 32 + ((mc ^ (kk & 3)) << 3);
                    Xbuf[o] = hi;
                    Xbuf[C::XA1 + o] = vv - hi;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&xfull);
        };

        int tb = 0, oy0 = 0, ox0 = 0, nb = 0, noy0 = 0, nox0 = 0;
        if (ntile > 0) {
            origin(0, nb, noy0, nox0);
            put_x(nb, noy0, nox0);
        }
        int q = 0, d = 0, eh = 0;                // weight-ring step, operand-chunk step, expand-half counter (all global)
        for (int ti = 0; ti < ntile; ++ti) {
            tb = nb; oy0 = noy0; ox0 = nox0;
            const int iy0 = oy0 - 1, ix0 = ox0 - 1;
            const bool have_next = ti + 1 < ntile;
            if (have_next) origin(ti + 1, nb, noy0, nox0);
#pragma unroll 1
            for (int h = 0; h < C::NH; ++h, ++eh) {
                q += C::NKA;                                         // the expand steps of this half consume ring slots too
                mbar_wait(&efull, eh & 1);
                tc_fence_after();
                if (h == C::NH - 1 && have_next) put_x(nb, noy0, nox0);   // every expand GEMM of this tile is complete: X is free
#pragma unroll 1
                for (int j = 0; j < C::NMC; ++j, ++q, ++d) {
                    const float* Wc = Ws + (q % C::NWB) * C::SLOT;
                    float* Es = Es0 + (d & 1) * C::ES1;
                    mbar_wait(&wbar[q % C::NWB], (q / C::NWB) & 1);
                    // ---- TMEM E (16 columns of this chunk) -> + b1, ReLU, zero outside the image -> Es[ch][halo pixel] ----
                    for (int mt = wq; mt < C::MT1; mt += nwq) {
                        if (mt * 128 + quarter * 32 >= G::IPIX) break;
                        const int pix = mt * 128 + quarter * 32 + lane;
                        const int r = pix / G::IWS, jj = pix - r * G::IWS;
                        const bool ok = jj < G::IW && (unsigned)(iy0 + r) < (unsigned)H && (unsigned)(ix0 + jj) < (unsigned)W;
                        float bb[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 b4 = ld4(Wc + C::OFF_B1 + 4 * i);
                            bb[4 * i] = b4.x; bb[4 * i + 1] = b4.y; bb[4 * i + 2] = b4.z; bb[4 * i + 3] = b4.w;
                        }
                        float v[16];
                        tmem_ld16(tmem + lane_base + C::TM_E + mt * C::NA + j * C::MC, v);
                        const float keep = ok ? 0.f : -INFINITY;
                        if (pix < G::IPIX) {
                            float* ep = Es + pix;
#pragma unroll
                            for (int i = 0; i < 16; ++i) ep[i * C::EPS] = fmaxf(v[i] + (bb[i] + keep), 0.f);
                        }
                    }
                    if (j == C::NMC - 1) {                           // last chunk of this half: TMEM E may be overwritten
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&efree);
                    }
                    named_bar_sync<1, NTW>();                        // E chunk complete (and everyone is done with the chunk two back)
                    const int b = d & 1;
                    if (d >= 2) mbar_wait(&dfree[b], ((d >> 1) - 1) & 1);
                    float* Dh = Dbuf + b * 2 * C::DA1;
                    dw_stage_split<G, C::MC, C::RH, NTW, C::KB3, C::EPS>(tid, Es, Wc + C::OFF_WD, Wc + C::OFF_BD, Dh, Dh + C::DA1);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dfull[b]);
                }
            }
            // ---- epilogue: TMEM O -> + b2 (+ residual) -> HBM (thread = pixel) ----
            mbar_wait(&ofull, ti & 1);
            tc_fence_after();
            for (int u = wq; u < C::MT3 * (C::COUTP / 16); u += nwq) {
                const int mt = u / (C::COUTP / 16), c0 = (u - mt * (C::COUTP / 16)) * 16;
                if (mt * 128 + quarter * 32 >= G::OPIX) break;
                const int pix = mt * 128 + quarter * 32 + lane;
                const int oy = pix / G::TW, ox = pix - oy * G::TW;
                const int gy = oy0 + oy, gx = ox0 + ox;
                float v[16], v1[16], v2[16];
                tmem_ld16(tmem + lane_base + C::TM_O + mt * C::NACC * C::COUTP + c0, v);
                tmem_ld16(tmem + lane_base + C::TM_O + mt * C::NACC * C::COUTP + C::COUTP + c0, v1);
                tmem_ld16(tmem + lane_base + C::TM_O + mt * C::NACC * C::COUTP + 2 * C::COUTP + c0, v2);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = (v[i] + v1[i]) + v2[i];
                if (pix < G::OPIX && gy < H && gx < W) {
                    const size_t o = ((size_t)tb * C::COUT + c0) * plane + (size_t)gy * W + gx;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (c0 + i < C::COUT) {
                            float r = v[i] + __ldg(wts + C::OFF_B2 + c0 + i);
                            if (C::RES) r += __ldg(x + o + i * plane);
                            y[o + i * plane] = r;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ofree);
        }
    }
    __syncthreads();
    if (warp == NWW) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
