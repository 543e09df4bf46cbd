// yf_tcpw.cuh — depthwise KxK (+ReLU) followed by a WIDE 1x1 on the tensor cores: the neck pairs conv5_3 -> conv5_4
// (96 -> 128) and conv4_1_2 -> conv4_1_3 (96 -> 96), yolo_fastest.py:133-134,143-144,201-202,212-213. Here the 1x1 is a real
// dense contraction (K = 96, N = 96..128), so every tcgen05.mma carries N = 96..128 columns of work per fetched operand row.
//
//   dw  D[m] = relu(dwKxK(X[m]) + bd[m])                 CUDA cores, sliding window in registers, one chunk of MC channels
//   pw  O[out px][N] += D[out px][MC] . W[MC][N]          tcgen05.mma kind::tf32 (3xTF32: hi*hi + hi*lo + lo*hi), accumulators in TMEM
//   out = O + b (no activation: conv_norm, yolo_fastest.py:29-38)
//
// Same warp-specialised skeleton as yf_tc.cuh: NWW worker warps stage the input chunk (cp.async, double buffered), run the
// depthwise and write its output as the split MMA operand (double buffered); one extra warp issues the MMAs and the weight-block
// bulk copies (3-deep ring). The MMA of chunk q runs while the workers compute chunk q + 1.
// The head groups (conv5_5 -> conv5_6 -> head_5, conv4_1_4 -> conv4_1_5 -> head_4) use the same kernel with the two 1x1s composed
// into one [C][headn] matrix on the host (see yf_kernels.cuh, HEADC) when headn <= 32: N = 32 columns, `nout` = headn stored.
// Packed weights (floats): NCHUNK x { [Whi: NP x MC K-major][Wlo][Wd: MC*KK][bd: MC] }, then [b: N].
#pragma once
#include "yf_tc.cuh"

namespace yf {

__device__ __forceinline__ void cp_async8(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}

template <int C_, int N_, int KS_, int TH_, int TW_, int MC_, int RH_, int NWW_, bool RELU_, bool NRT_ = false>
struct DwPwTcCfg {
    static constexpr int C = C_, N = N_, MC = MC_, RH = RH_, NWW = NWW_;
    static constexpr bool RELU = RELU_;
    static constexpr bool NRT = NRT_;             // the number of stored output channels is a run-time argument (head groups)
    static constexpr int NTW = NWW * 32, NT = NTW + 32, NWB = 3;
    using G = Geo<KS_, 1, TH_, TW_>;
    static constexpr int KK = KS_ * KS_;
    static constexpr int NCHUNK = C / MC;
    static constexpr int MT3 = cdiv(G::OPIX, 128), NG3 = cdiv(G::OPIX, 32);
    static constexpr int KB3 = NG3 * 256, DA1 = (MC / 8) * KB3;              // floats of one operand region (hi or lo)
    static constexpr int NP = rup(N, 16);
    static constexpr int EPS = e_stride(G::IPIX, G::TW / 4);
    static constexpr int XS1 = rup(MC * EPS, 32);
    static constexpr int OFF_WH = 0, OFF_WL = NP * MC, OFF_WD = 2 * NP * MC, OFF_BD = OFF_WD + MC * KK, CB = rup(OFF_BD + MC, 32);
    static constexpr int OFF_B = NCHUNK * CB;
    static constexpr int WFLOATS = OFF_B + N;
    static constexpr int TCOLS = pow2_ge(MT3 * NP);
    static constexpr int SMEM_FLOATS = 4 * DA1 + 2 * XS1 + NWB * CB;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    static constexpr int HPAIR = G::IWS / 2;                                  // 8-byte pieces per staged row
    static_assert(C % MC == 0 && MC % 8 == 0 && NWW >= 4 && G::P % 2 == 0 && TW_ % 4 == 0, "tiling constraints");
    static_assert(TCOLS <= 512 && NP <= 256, "TMEM columns / MMA N");
    static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
    static_assert(XS1 * 4 >= 4096, "over-read guard behind the operand regions");
};

// depthwise KSxKS s1 + bias + ReLU from a staged chunk [MC][halo] into the A-operand layout of the 1x1 MMA (hi and lo parts)
template <class G, int MC, int RH, int NT, int KB3, int EPS>
__device__ __forceinline__ void dwk_stage_split(int tid, const float* __restrict__ Xs, const float* __restrict__ Wd, const float* __restrict__ bd,
                                                float* __restrict__ DAhi, float* __restrict__ DAlo) {
    constexpr int KS = G::KS, NV = 3 + KS;
    static_assert(G::TH % RH == 0 && G::S == 1, "stride 1");
    constexpr int NSTRIP = G::TW / 4, NSEG = G::TH / RH;
    constexpr int NITEM = MC * NSEG * NSTRIP;
    for (int item = tid; item < NITEM; item += NT) {
        const int m = item / (NSEG * NSTRIP);
        const int rem = item - m * (NSEG * NSTRIP);
        const int seg = rem / NSTRIP;
        const int g = rem - seg * NSTRIP;
        float w[KS * KS];
#pragma unroll
        for (int t = 0; t < KS * KS; ++t) w[t] = Wd[m * KS * KS + t];
        const float b = bd[m];
        const float* e = Xs + m * EPS + (seg * RH) * G::IWS;
        float win[KS][NV];
#pragma unroll
        for (int dd = 0; dd < KS - 1; ++dd) load_window<G>(win[dd + 1], e + dd * G::IWS, g);
#pragma unroll
        for (int oy = 0; oy < RH; ++oy) {
#pragma unroll
            for (int dd = 0; dd < KS - 1; ++dd)
#pragma unroll
                for (int v = 0; v < NV; ++v) win[dd][v] = win[dd + 1][v];
            load_window<G>(win[KS - 1], e + (oy + KS - 1) * G::IWS, g);
            float a[4] = {b, b, b, b};
#pragma unroll
            for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = fmaf(w[dy * KS + dx], win[dy][i + dx], a[i]);
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = fmaxf(a[i], 0.f);
                hi[i] = tf32_hi(v);
                lo[i] = v - hi[i];
            }
            const int p0 = (seg * RH + oy) * G::TW + 4 * g;
            const int o = a_idx(p0, m, KB3);
            st4(DAhi + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
            st4(DAlo + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
dwpw_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
               int tiles_x, int tiles_y, int total_tiles, int nout /* output channels actually stored, <= N (head groups: run time) */) {
    using G = typename C::G;
    constexpr int NTW = C::NTW, NWW = C::NWW, NCH = C::NCHUNK;
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Dbuf = base;                          // [2 buffers][hi | lo][DA1]
    float* Xs0 = Dbuf + 4 * C::DA1;              // [2 buffers][MC][EPS]
    float* Ws = Xs0 + 2 * C::XS1;                // ring of chunk weight blocks
    __shared__ __align__(8) uint64_t wbar[C::NWB], dfull[2], dfree[2], ofull, ofree;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < C::NWB; ++i) mbar_init(&wbar[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], NWW); mbar_init(&dfree[i], 1); }
        mbar_init(&ofull, 1); mbar_init(&ofree, NWW);
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
    const int S = ntile * NCH;

    if (warp == NWW) {
        // ================= tensor-core warp =================
        if (S > 0 && elect_one()) {
            constexpr uint32_t IDESC = umma_idesc_tf32(C::NP);
            const uint32_t ws = smem_u32(Ws);
            auto issue_w = [&](int q) {
                const int slot = q % C::NWB;
                mbar_expect_tx(&wbar[slot], C::CB * 4);
                bulk_load(Ws + slot * C::CB, wts + (size_t)(q % NCH) * C::CB, C::CB * 4, &wbar[slot]);
            };
            const uint64_t dd0 = umma_desc(smem_u32(Dbuf), 1024, 512, 1);
            const uint64_t dw0 = umma_desc(ws, 128, (C::MC / 4) * 128, 0);
            issue_w(0);
            if (S > 1) issue_w(1);
            for (int q = 0; q < S; ++q) {
                const int b = q & 1, c = q % NCH;
                if (c == 0 && q > 0) mbar_wait(&ofree, ((q / NCH) - 1) & 1);     // the previous tile's accumulators have been read out
                mbar_wait(&dfull[b], (q >> 1) & 1);
                mbar_wait(&wbar[q % C::NWB], (q / C::NWB) & 1);
                tc_fence_after();
                const uint64_t wb = dw0 + (uint64_t)(((uint32_t)(q % C::NWB) * C::CB * 4) >> 4);
                const uint64_t db = dd0 + (uint64_t)((uint32_t)(b * 2 * C::DA1 * 4) >> 4);
#pragma unroll 1
                for (int mt = 0; mt < C::MT3; ++mt) {
                    const uint64_t mo = (uint64_t)(mt * 256);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                        for (int kb = 0; kb < C::MC / 8; ++kb)
                            umma_tf32(tmem + mt * C::NP, db + mo + (uint64_t)(((pass == 2 ? C::DA1 : 0) + kb * C::KB3) * 4 / 16),
                                      wb + (uint64_t)(((pass == 1 ? C::OFF_WL : C::OFF_WH) * 4 + kb * 256) / 16), IDESC,
                                      (c | pass | kb) ? 1u : 0u);
                    }
                }
                umma_commit(&dfree[b]);
                if (c == NCH - 1) umma_commit(&ofull);
                if (q + 2 < S) {
                    if (q >= 1) mbar_wait(&dfree[(q - 1) & 1], ((q - 1) >> 1) & 1);   // ring slot (q+2)%3 == (q-1)%3 is free
                    issue_w(q + 2);
                }
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
        const bool even = (W & 1) == 0;
        // channels [c0, c0 + MC) of the halo rectangle of tile (b, oy0, ox0) -> dst [MC][EPS], zero outside the image (8-byte cp.async:
        // the halo starts P = 2 columns left of a tile column that is a multiple of 4, and the image width is even)
        auto stage_chunk = [&](int b, int oy0, int ox0, int c0, float* dst) {
            const float* src = x + ((size_t)b * C::C + c0) * plane;
            constexpr int PER_CH = G::IH * C::HPAIR;
            for (int idx = tid; idx < C::MC * PER_CH; idx += NTW) {
                const int ch = idx / PER_CH, rem = idx - ch * PER_CH;
                const int r = rem / C::HPAIR, jp = rem - r * C::HPAIR;
                const int gy = oy0 - G::P + r, gx = ox0 - G::P + 2 * jp;
                const bool rowok = (unsigned)gy < (unsigned)H;
                float* d = dst + ch * C::EPS + r * G::IWS + 2 * jp;
                const float* s = src + (size_t)ch * plane + (size_t)(rowok ? gy : 0) * W;
                if (even) {
                    const bool ok = rowok && 2 * jp < G::IW && (unsigned)gx < (unsigned)W;
                    cp_async8(d, ok ? s + gx : src, ok ? 8 : 0);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool ok = rowok && 2 * jp + e < G::IW && (unsigned)(gx + e) < (unsigned)W;
                        cp_async4(d + e, ok ? s + gx + e : src, ok ? 4 : 0);
                    }
                }
            }
            cp_async_commit();
        };

        int tb = 0, oy0 = 0, ox0 = 0, nb = 0, noy0 = 0, nox0 = 0;
        if (ntile > 0) {
            origin(0, nb, noy0, nox0);
            stage_chunk(nb, noy0, nox0, 0, Xs0);
        }
        int q = 0;
        for (int ti = 0; ti < ntile; ++ti) {
            tb = nb; oy0 = noy0; ox0 = nox0;
            const bool have_next = ti + 1 < ntile;
            if (have_next) origin(ti + 1, nb, noy0, nox0);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++q) {
                const int b = q & 1;
                const float* Wc = Ws + (q % C::NWB) * C::CB;
                cp_async_wait_all();
                named_bar_sync<1, NTW>();              // chunk q has landed for everyone; everyone is done with chunk q - 1
                if (c + 1 < NCH) stage_chunk(tb, oy0, ox0, (c + 1) * C::MC, Xs0 + ((q + 1) & 1) * C::XS1);
                else if (have_next) stage_chunk(nb, noy0, nox0, 0, Xs0 + ((q + 1) & 1) * C::XS1);
                mbar_wait(&wbar[q % C::NWB], (q / C::NWB) & 1);
                if (q >= 2) mbar_wait(&dfree[b], ((q >> 1) - 1) & 1);            // the MMA that read this operand buffer is complete
                float* Dh = Dbuf + b * 2 * C::DA1;
                dwk_stage_split<G, C::MC, C::RH, NTW, C::KB3, C::EPS>(tid, Xs0 + (q & 1) * C::XS1, Wc + C::OFF_WD, Wc + C::OFF_BD, Dh, Dh + C::DA1);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dfull[b]);
            }
            // ---- epilogue: TMEM -> + bias -> HBM (thread = pixel, 16 output channels per unit) --------------------------------
            mbar_wait(&ofull, ti & 1);
            tc_fence_after();
            for (int u = wq; u < C::MT3 * (C::NP / 16); u += nwq) {
                const int mt = u / (C::NP / 16), c0 = (u - mt * (C::NP / 16)) * 16;
                if (mt * 128 + quarter * 32 >= G::OPIX) break;
                const int pix = mt * 128 + quarter * 32 + lane;
                const int oy = pix / G::TW, ox = pix - oy * G::TW;
                const int gy = oy0 + oy, gx = ox0 + ox;
                float v[16];
                tmem_ld16(tmem + lane_base + mt * C::NP + c0, v);
                if (pix < G::OPIX && gy < H && gx < W) {
                    const int nch = C::NRT ? nout : C::N;
                    float* yp = y + ((size_t)tb * nch + c0) * plane + (size_t)gy * W + gx;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (c0 + i < nch) {
                            float r = v[i] + __ldg(wts + C::OFF_B + c0 + i);
                            if (C::RELU) r = fmaxf(r, 0.f);
                            yp[i * plane] = r;
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
