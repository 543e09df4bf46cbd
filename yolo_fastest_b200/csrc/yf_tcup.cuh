// yf_tcup.cuh — fused upsample + concat + 1x1 on the tensor cores: deconv5_1 (ConvTranspose2d k2 s2 96 -> 96 + BN + ReLU,
// yolo_fastest.py:42-48,140,208) -> cat((conv4_2[136], deconv5_1[96]), 1) (:209) -> conv4_1_1 (1x1 232 -> 96 + BN + ReLU, :142).
// These are the widest contractions of the network (K = 96 -> N = 4 x 96, and K = 232 -> N = 96).
//
// An output tile is 8 x 20 pixels (its parent tile 4 x 10). Two GEMMs, both 3xTF32 (hi*hi + hi*lo + lo*hi), fp32 accumulators in TMEM:
//   A  U[parent px][(parity, m)] = P[parent px][c] . Wt[c][(parity, m)]      M = 128 (40 parent pixels used), K = 96, N = 192 per
//      channel half (TMEM holds O, 192 columns, and ONE half of U, 192 columns); every parity of a parent pixel is one column group
//   B  O[out px][n] = [skip | U][out px][k] . Wab[k][n]                       M = 2 x 128 (160 used), K = 136 + 96 in chunks of 16, N = 96
// The upsampled tensor and the concatenation never exist in HBM — U goes TMEM -> registers (bias, ReLU, split) -> the B operand
// chunk of GEMM B, written at the four output pixels of each parent pixel.
//
// Warp-specialised like yf_tc.cuh. Step order of a tile (one weight-ring slot each; the tensor pipe executes them in this order):
//   A(half 0) x 6 K-chunks | 9 skip chunks | 3 U chunks (up channels 0..47) | A(half 1) x 6 | 3 U chunks (48..95)
// Workers: skip chunks (global -> split -> operand, double buffered) run ahead while GEMM A executes; U chunks wait for `ufull`;
// GEMM A of the second half waits for `ufree` (the first half has been read out of TMEM); the parent tile of the NEXT tile is
// split into its operand as soon as GEMM A of the second half is complete; the output epilogue waits for `ofull`.
// Packed weights (floats): 27 slots of SLOT floats in step order { A: [hi: 192 x 16 K-major][lo]  |  B: [hi: 96 x 16][lo] }, then
// [bt: 96] (deconv bias) and [b: 96].
#pragma once
#include "yf_tc.cuh"

namespace yf {

template <int NWW_>
struct UpCatTcCfg {
    static constexpr int NWW = NWW_, NTW = NWW * 32, NT = NTW + 32, NWB = 3;
    static constexpr int CS = 136, CU = 96, N = 96, TH = 8, TW = 20, OPIX = TH * TW, PH = TH / 2, PW = TW / 2, PPIX = PH * PW;
    static constexpr int MC = 16;                                           // K chunk
    static constexpr int NG3 = cdiv(OPIX, 32), KB3 = NG3 * 256, DA1 = (MC / 8) * KB3;      // GEMM B operand chunk (hi or lo), floats
    static constexpr int MTO = cdiv(OPIX, 128);
    static constexpr int NGP = cdiv(PPIX, 32), KBP = NGP * 256, PA1 = (CU / 8) * KBP;      // parent operand (hi or lo), all 96 channels
    static constexpr int NHALF = 48 * 4;                                    // N of GEMM A per channel half
    static constexpr int NSKIP = cdiv(CS, MC), NUPH = 48 / MC;              // 9 skip chunks, 3 up chunks per half
    static constexpr int NKA = CU / MC;                                     // 6 K chunks of GEMM A
    static constexpr int STEPS = 2 * NKA + NSKIP + 2 * NUPH;                // 27 weight slots per tile
    static constexpr int DSTEPS = NSKIP + 2 * NUPH;                         // 15 operand chunks per tile
    static constexpr int SLOT = 2 * NHALF * MC;                             // floats per ring slot (the B-step blocks use half of it)
    static constexpr int OFF_BT = STEPS * SLOT, OFF_B = OFF_BT + CU;
    static constexpr int WFLOATS = OFF_B + N;
    static constexpr int TM_O = 0, TM_U = MTO * N, TCOLS = pow2_ge(TM_U + NHALF);
    static constexpr int SK_ITEMS = MC * TH * (TW / 4), SK_PER = cdiv(SK_ITEMS, NTW);   // 128-bit skip loads per chunk / per worker thread
    static constexpr int SMEM_FLOATS = 4 * DA1 + 2 * PA1 + NWB * SLOT;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static_assert(TCOLS <= 512, "TMEM columns");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
upcat_tc_kernel(const float* __restrict__ skip /*[B,136,H,W]*/, const float* __restrict__ low /*[B,96,H/2,W/2]*/,
                float* __restrict__ y /*[B,96,H,W]*/, const float* __restrict__ wts, int H, int W, int tiles_x, int tiles_y, int total_tiles) {
    constexpr int NTW = C::NTW, NWW = C::NWW;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Dbuf = base;                          // [2 buffers][hi | lo][DA1]
    float* Pbuf = Dbuf + 4 * C::DA1;             // [hi | lo][PA1]
    float* Ws = Pbuf + 2 * C::PA1;               // weight ring
    __shared__ __align__(8) uint64_t wbar[C::NWB], sdone[C::NWB], dfull[2], dfree[2], pfull, ufull, ufree, ofull, ofree;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < C::NWB; ++i) { mbar_init(&wbar[i], 1); mbar_init(&sdone[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], NWW); mbar_init(&dfree[i], 1); }
        mbar_init(&pfull, NWW); mbar_init(&ufull, 1); mbar_init(&ufree, NWW); mbar_init(&ofull, 1); mbar_init(&ofree, NWW);
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
    const int Hp = H / 2, Wp = W / 2;

    if (warp == NWW) {
        // ================= tensor-core warp =================
        if (ntile > 0 && elect_one()) {
            constexpr uint32_t IDESC_A = umma_idesc_tf32(C::NHALF), IDESC_B = umma_idesc_tf32(C::N);
            const int S = ntile * C::STEPS;
            const uint64_t dd0 = umma_desc(smem_u32(Dbuf), 1024, 512, 1);
            const uint64_t dp0 = umma_desc(smem_u32(Pbuf), 1024, 512, 1);
            const uint64_t dwa = umma_desc(smem_u32(Ws), 128, (C::MC / 4) * 128, 0);
            int q = 0;                           // weight-ring step (global), d: operand-chunk step (global)
            int d = 0;
            auto issue_w = [&](int s) {
                const int slot = s % C::NWB;
                mbar_expect_tx(&wbar[slot], C::SLOT * 4);
                bulk_load(Ws + slot * C::SLOT, wts + (size_t)(s % C::STEPS) * C::SLOT, C::SLOT * 4, &wbar[slot]);
            };
            auto step_begin = [&]() -> uint64_t {         // wait for this step's weights; returns their descriptor base
                mbar_wait(&wbar[q % C::NWB], (q / C::NWB) & 1);
                return dwa + (uint64_t)(((uint32_t)(q % C::NWB) * C::SLOT * 4) >> 4);
            };
            auto step_end = [&]() {                       // commit, refill the ring two steps ahead
                umma_commit(&sdone[q % C::NWB]);
                if (q + 2 < S) {
                    if (q >= 1) mbar_wait(&sdone[(q - 1) % C::NWB], ((q - 1) / C::NWB) & 1);
                    issue_w(q + 2);
                }
                ++q;
            };
            auto gemm_a = [&]() {                         // U (one channel half) = P . Wt over 6 K chunks
#pragma unroll 1
                for (int kc = 0; kc < C::NKA; ++kc) {
                    const uint64_t wb = step_begin();
                    tc_fence_after();
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                        for (int kb = 0; kb < C::MC / 8; ++kb)
                            umma_tf32(tmem + C::TM_U, dp0 + (uint64_t)(((pass == 2 ? C::PA1 : 0) + (kc * (C::MC / 8) + kb) * C::KBP) * 4 / 16),
                                      wb + (uint64_t)(((pass == 1 ? C::NHALF * C::MC : 0) * 4 + kb * 256) / 16), IDESC_A, (kc | pass | kb) ? 1u : 0u);
                    step_end();
                }
                umma_commit(&ufull);
            };
            auto gemm_b_chunk = [&](bool first) {         // O += operand chunk . Wab chunk
                const int b = d & 1;
                mbar_wait(&dfull[b], (d >> 1) & 1);
                const uint64_t wb = step_begin();
                tc_fence_after();
                const uint64_t db = dd0 + (uint64_t)((uint32_t)(b * 2 * C::DA1 * 4) >> 4);
#pragma unroll 1
                for (int mt = 0; mt < C::MTO; ++mt) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                        for (int kb = 0; kb < C::MC / 8; ++kb)
                            umma_tf32(tmem + C::TM_O + mt * C::N, db + (uint64_t)(mt * 256) + (uint64_t)(((pass == 2 ? C::DA1 : 0) + kb * C::KB3) * 4 / 16),
                                      wb + (uint64_t)(((pass == 1 ? C::N * C::MC : 0) * 4 + kb * 256) / 16), IDESC_B, (first && pass == 0 && kb == 0) ? 0u : 1u);
                }
                umma_commit(&dfree[b]);
                step_end();
                ++d;
            };
            issue_w(0);
            issue_w(1);
            for (int ti = 0; ti < ntile; ++ti) {
                mbar_wait(&pfull, ti & 1);                                     // parent operand of this tile is in smem
                if (ti > 0) mbar_wait(&ufree, (2 * ti - 1) & 1);               // previous tile's second half has been read out of TMEM
                gemm_a();
                if (ti > 0) mbar_wait(&ofree, (ti - 1) & 1);                   // previous tile's output has been read out of TMEM
                for (int c = 0; c < C::NSKIP; ++c) gemm_b_chunk(c == 0);
                for (int c = 0; c < C::NUPH; ++c) gemm_b_chunk(false);
                mbar_wait(&ufree, (2 * ti) & 1);                               // first half read out
                gemm_a();
                for (int c = 0; c < C::NUPH; ++c) gemm_b_chunk(false);
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
            oy0 = ty * C::TH; ox0 = (t - ty * tiles_x) * C::TW;
        };
        const size_t plane = (size_t)H * W, pplane = (size_t)Hp * Wp;
        const float* bt = wts + C::OFF_BT;
        const float* bo = wts + C::OFF_B;

        // parent tile [96][4 x 10] of image b -> split operand of GEMM A (item = 8 channels of one parent pixel); the values are
        // fetched into registers early (unconditional clamped loads) and written once the operand buffer is free
        constexpr int NITEM_P = (C::CU / 8) * C::PPIX, NIT_P = (NITEM_P + NTW - 1) / NTW;
        float pr[NIT_P * 8];
        uint32_t pok = 0;
        auto fetch_parent = [&](int b, int oy0, int ox0) {
            pok = 0;
#pragma unroll
            for (int i = 0; i < NIT_P; ++i) {
                const int item = min(tid + i * NTW, NITEM_P - 1);
                const int kh = item / C::PPIX, pp = item - kh * C::PPIX;
                const int yy = pp / C::PW, xx = pp - yy * C::PW;
                const int gy = oy0 / 2 + yy, gx = ox0 / 2 + xx;
                const bool ok = gy < Hp && gx < Wp;
                pok |= (ok ? 1u : 0u) << i;
                const float* src = low + ((size_t)b * C::CU + kh * 8) * pplane + (size_t)(ok ? gy : 0) * Wp + (ok ? gx : 0);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) pr[i * 8 + kk] = __ldg(src + kk * pplane);
            }
        };
        auto put_parent = [&]() {
#pragma unroll
            for (int i = 0; i < NIT_P; ++i) {
                const int item = tid + i * NTW;
                if (item < NITEM_P) {
                    const int kh = item / C::PPIX, pp = item - kh * C::PPIX;
                    const int ob = kh * C::KBP + (pp >> 5) * 256 + (pp & 7), mc = (pp & 31) >> 3;
                    const bool ok = (pok >> i) & 1u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const float v = ok ? pr[i * 8 + kk] : 0.f;
                        const float hi = tf32_hi(v);
                        const int o = ob + ((kk >> 2) & 1) * 128 + (kk & 3) * 32 + ((mc ^ (kk & 3)) << 3);
                        Pbuf[o] = hi;
                        Pbuf[C::PA1 + o] = v - hi;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull);
        };
        // skip chunk: item = (channel, row, 4-pixel group) -> one 128-bit load; two items per thread, fetched TWO chunks ahead into
        // alternating register buffers (one chunk step is shorter than the HBM/L2 latency)
        constexpr int SK_ITEMS = C::SK_ITEMS, SK_PER = C::SK_PER;
        float4 skA[SK_PER], skB[SK_PER];
        auto fetch_skip = [&](int b, int oy0, int ox0, int c, float4 (&sk)[SK_PER]) {
#pragma unroll
            for (int k = 0; k < SK_PER; ++k) {
                const int item = min(tid + k * NTW, SK_ITEMS - 1);
                const int ch = item / (C::TH * (C::TW / 4)), rem = item - ch * (C::TH * (C::TW / 4));
                const int row = rem / (C::TW / 4), g = rem - row * (C::TW / 4);
                const int cg = min(c * C::MC + ch, C::CS - 1), gy = min(oy0 + row, H - 1), gx = min(ox0 + 4 * g, W - 4);
                sk[k] = __ldg(reinterpret_cast<const float4*>(skip + ((size_t)b * C::CS + cg) * plane + (size_t)gy * W + gx));
            }
        };
        auto put_skip = [&](int c, float* Dh, const float4 (&sk)[SK_PER]) {
#pragma unroll
            for (int k = 0; k < SK_PER; ++k) {
                const int item = tid + k * NTW;
                if (item < SK_ITEMS) {
                    const int ch = item / (C::TH * (C::TW / 4)), rem = item - ch * (C::TH * (C::TW / 4));
                    const int row = rem / (C::TW / 4), g = rem - row * (C::TW / 4);
                    const bool ok = c * C::MC + ch < C::CS;           // channels 136..143 of the last chunk are zero (their weights too)
                    const float v[4] = {ok ? sk[k].x : 0.f, ok ? sk[k].y : 0.f, ok ? sk[k].z : 0.f, ok ? sk[k].w : 0.f};
                    float hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { hi[i] = tf32_hi(v[i]); lo[i] = v[i] - hi[i]; }
                    const int o = a_idx(row * C::TW + 4 * g, ch, C::KB3);
                    st4(Dh + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
                    st4(Dh + C::DA1 + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
                }
            }
        };
        // up chunk j of half h: TMEM U columns (parity, 16 channels) -> + bt, ReLU -> split -> the 4 output pixels of each parent pixel
        auto put_up = [&](int h, int j, float* Dh) {
            const int pp = quarter * 32 + lane;
            if (quarter * 32 < C::PPIX) {                            // warp-uniform: this quarter holds parent pixels
                const int yy = pp / C::PW, xx = pp - yy * C::PW;
                for (int u = wq; u < 8; u += nwq) {                  // unit = (parity, 8-channel half of the chunk)
                    const int par = u >> 1, hh = u & 1;
                    float v[8];
                    tmem_ld8(tmem + lane_base + C::TM_U + par * 48 + j * C::MC + hh * 8, v);
                    if (pp < C::PPIX) {
                        const int p = (2 * yy + (par >> 1)) * C::TW + 2 * xx + (par & 1);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float r = fmaxf(v[e] + __ldg(bt + h * 48 + j * C::MC + hh * 8 + e), 0.f);
                            const float hi = tf32_hi(r);
                            const int o = a_idx(p, hh * 8 + e, C::KB3);
                            Dh[o] = hi;
                            Dh[C::DA1 + o] = r - hi;
                        }
                    }
                }
            }
        };

        int tb = 0, oy0 = 0, ox0 = 0, nb = 0, noy0 = 0, nox0 = 0;
        if (ntile > 0) {
            origin(0, nb, noy0, nox0);
            fetch_parent(nb, noy0, nox0);
            fetch_skip(nb, noy0, nox0, 0, skA);
            fetch_skip(nb, noy0, nox0, 1, skB);
            put_parent();
        }
        int d = 0;                                                   // operand-chunk step (global)
        auto chunk_begin = [&]() -> float* {
            const int b = d & 1;
            if (d >= 2) mbar_wait(&dfree[b], ((d >> 1) - 1) & 1);
            return Dbuf + b * 2 * C::DA1;
        };
        auto chunk_end = [&]() {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dfull[d & 1]);
            ++d;
        };
        for (int ti = 0; ti < ntile; ++ti) {
            tb = nb; oy0 = noy0; ox0 = nox0;
            const bool have_next = ti + 1 < ntile;
            if (have_next) origin(ti + 1, nb, noy0, nox0);
#pragma unroll 1
            for (int c = 0; c < C::NSKIP; c += 2) {                  // chunks c (buffer A) and c + 1 (buffer B)
                float* Dh = chunk_begin();
                put_skip(c, Dh, skA);
                chunk_end();
                if (c + 2 < C::NSKIP) fetch_skip(tb, oy0, ox0, c + 2, skA);
                if (c + 1 < C::NSKIP) {
                    Dh = chunk_begin();
                    put_skip(c + 1, Dh, skB);
                    chunk_end();
                    if (c + 3 < C::NSKIP) fetch_skip(tb, oy0, ox0, c + 3, skB);
                }
            }
            if (have_next) {                                         // the next tile's first loads travel during the up chunks and the epilogue
                fetch_skip(nb, noy0, nox0, 0, skA);
                fetch_skip(nb, noy0, nox0, 1, skB);
                fetch_parent(nb, noy0, nox0);
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                mbar_wait(&ufull, (2 * ti + h) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < C::NUPH; ++j) {
                    float* Dh = chunk_begin();
                    put_up(h, j, Dh);
                    chunk_end();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ufree);
            }
            if (have_next) put_parent();                             // GEMM A of the second half is complete: the parent operand is free
            // ---- epilogue: TMEM O -> + b, ReLU -> HBM (thread = pixel) ----
            mbar_wait(&ofull, ti & 1);
            tc_fence_after();
            for (int u = wq; u < C::MTO * (C::N / 16); u += nwq) {
                const int mt = u / (C::N / 16), c0 = (u - mt * (C::N / 16)) * 16;
                if (mt * 128 + quarter * 32 >= C::OPIX) break;
                const int pix = mt * 128 + quarter * 32 + lane;
                const int oy = pix / C::TW, ox = pix - oy * C::TW;
                const int gy = oy0 + oy, gx = ox0 + ox;
                float v[16];
                tmem_ld16(tmem + lane_base + C::TM_O + mt * C::N + c0, v);
                if (pix < C::OPIX && gy < H && gx < W) {
                    float* yp = y + ((size_t)tb * C::N + c0) * plane + (size_t)gy * W + gx;
#pragma unroll
                    for (int i = 0; i < 16; ++i) yp[i * plane] = fmaxf(v[i] + __ldg(bo + c0 + i), 0.f);
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
