// yf_tct.cuh — "channel-lane" tensor-core inverted-residual block (res3_3..6: 16 -> 96 -> 16, yolo_fastest.py:52-66).
//
// yf_tc.cuh computes E[pixel][mid] (TMEM lane = pixel) and has to push every E value through shared memory before the depthwise can
// see a pixel's neighbours; its workers spend their time on that round trip. Here the expand GEMM is issued TRANSPOSED:
//
//   S1  E^T[mid][halo px] = W1^T[mid][CIN + 1] . X^T[CIN + 1][halo px]   tcgen05.mma kind::tf32, M = 128 (mid channels), N = halo pixels
//
// so a TMEM lane is a mid CHANNEL and its columns are the tile's halo pixels, row after row. A worker thread owns one channel: it
// reads three halo rows of ITS channel straight from TMEM into registers (tcgen05.ld, 18 columns per row), keeps its nine depthwise
// weights in registers for the whole kernel and runs the 3x3 entirely in registers — E never exists in shared memory, there is no
// per-chunk weight traffic and no block-wide barrier. The expand bias rides in the GEMM as an extra "ones" input channel that is 1
// inside the image and 0 outside, which also makes E exactly 0 on the depthwise's zero padding: the read-out is one FMAX per value.
//
//   dw  D = relu(dw3x3(relu(E^T)) + bd)                                 registers; written as the project operand (hi | lo)
//   S3  O[out px][COUT] = D[out px][mid] . W2[mid][COUT]                as in yf_tc.cuh: A = D MN-major swizzled, B = [W2hi | W2lo]
//
// Tile = 8 x 16 output pixels (128 = one MMA tile of the project GEMM), halo 10 x 18 = 180 -> N = 192. TMEM: two E^T buffers
// (2 x 192 columns) + two O buffers (2 x 32): the expand GEMM of tile t + 1 runs while the workers are on tile t, the output
// epilogue of tile t - 1 runs behind the depthwise of tile t. 19 warps, warp % 4 = TMEM lane quarter: of the first 16, quarters 0..2
// (12 warps) are workers — channel = 32 * quarter + lane, warp / 4 = which pair of output rows — and quarter 3 holds two
// input-staging warps, one output-epilogue warp and the tensor-core warp; warps 16..18 are the output-epilogue warps of quarters 0..2. fp32 parity through 3xTF32 exactly as in yf_tc.cuh.
//
// Operand layouts: W1^T (A) and X^T (B) are both K-major, un-swizzled core matrices [row / 8][k / 4][row % 8][k % 4] — the layout the
// weight operands of every other kernel use (validated there); the input staging writes one 16-byte core-matrix row per
// (pixel, 4 channels). D and W2 as in yf_tc.cuh.
// Packed weights (floats): [W1hi: 128 x 24][W1lo: 128 x 24][W2: (hi 16 | lo 16) x 96][wd: 96 x 9][bd: 96][b2: 16].
#pragma once
#include "yf_tc.cuh"
#include "yf_tma.cuh"

namespace yf {

// one halo row of this thread's channel: 18 consecutive TMEM columns (load + wait in ONE statement: the registers are only valid after the wait)
__device__ __forceinline__ void tmem_ld_row18(uint32_t taddr, float (&v)[18]) {
    uint32_t r[18];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%18];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%16,%17}, [%19];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17])
        : "r"(taddr), "r"(taddr + 16) : "memory");
#pragma unroll
    for (int i = 0; i < 18; ++i) v[i] = fmaxf(__uint_as_float(r[i]), 0.f);       // ReLU; the bias came through the GEMM
}

// -DYF_TC_TRACE -DYF_TCT_TRACE: CTA 0 records clock64() per tile: slots 0..5 worker warp 0, 6..9 tensor-core thread, 10..13 staging warp 0
#if defined(YF_TC_TRACE) && defined(YF_TCT_TRACE)
#define TT_TRACE(tile, ev) do { if (blockIdx.x == 0 && (tile) < 64) g_tc_trace[(tile) * 16 + (ev)] = clock64(); } while (0)
#else
#define TT_TRACE(tile, ev) do { } while (0)
#endif

template <int CIN_, int CMID_, int COUT_, bool RES_>
struct IrbTtCfg {
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_;
    static constexpr bool RES = RES_;
    static constexpr int TH = 8, TW = 16, OPIX = TH * TW;                 // output tile: one 128-row MMA tile of the project GEMM
    static constexpr int HR = TH + 2, HW = TW + 2, HPIX = HR * HW;        // halo tile, pixel n = r * HW + j
    static constexpr int NPX = rup(HPIX, 16);                             // N of the expand MMA
    static constexpr int KX = CIN + 8;                                    // + the ones channel (and 7 zero channels: K advances by 8)
    static constexpr int NWARP = 19, NT = NWARP * 32;                     // 12 workers, 2 staging, tensor-core warp, 4 epilogue warps (one per TMEM lane quarter)
    static constexpr int NWORK = 12, NAUX = 2, NTA = NAUX * 32;         // worker warps; input-staging warps (a third quarter-3 warp runs that quarter's output epilogue)
    static constexpr int COUTP = rup(COUT, 16);
    static constexpr int KB3 = (OPIX / 32) * 256;                         // floats per 8-channel block of the operand D
    // packed weights
    static constexpr int OFF_W1H = 0, OFF_W1L = 128 * KX, OFF_W2 = 2 * 128 * KX, WRES = OFF_W2 + 2 * COUTP * CMID;
    static constexpr int OFF_WD = WRES, OFF_BD = OFF_WD + CMID * 9, OFF_B2 = OFF_BD + CMID, WFLOATS = rup(OFF_B2 + COUT, 32);
    // shared memory (floats)
    static constexpr int XH = NPX * KX, XL = NPX * CIN, DA = (CMID / 8) * KB3;
    static constexpr int RW = 24, RAW = CIN * HR * RW;                    // raw input box [CIN][HR][RW] from column ox0 - 4 (TMA boxes start 16-byte aligned)
    static constexpr int SMEM_FLOATS = WRES + XH + XL + 2 * DA + 2 * RAW;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    // TMEM columns
    static constexpr int TM_E = 0, TM_O = 2 * NPX, TCOLS = 512;
    static constexpr int NITEM = HPIX * (CIN / 4), NIT = cdiv(NITEM, NTA);       // staging items (halo pixel, 4 input channels) per tile / per staging thread
    static_assert(CMID % 32 == 0 && CMID <= 96, "worker quarters 0..2 hold the mid channels");
    static_assert(CIN % 8 == 0 && NPX <= 256 && TM_O + 2 * 2 * COUTP <= TCOLS, "MMA shape / TMEM columns");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert(WRES % 32 == 0 && (WRES * 4) % 1024 == 0, "operand alignment");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
    static_assert((RAW * 4) % 128 == 0 && RAW < 4096 && XH < 65536, "raw box alignment / packed staging offsets");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
irbt_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
            int tiles_x, int tiles_y, int total_tiles) {
    extern __shared__ unsigned char smem_raw[];
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Wr = base;                            // resident operands: W1hi | W1lo | W2
    float* DAhi = Wr + C::WRES;
    float* DAlo = DAhi + C::DA;
    float* Xhi = DAlo + C::DA;                   // [NPX / 8][KX / 4][8][4]
    float* Xlo = Xhi + C::XH;                    // [NPX / 8][CIN / 4][8][4]
    float* Raw = Xlo + C::XL;                    // [2][CIN][HR][RW]: the TMA boxes of the next two tiles
    __shared__ __align__(8) uint64_t wres, xfull, xfree, efull[2], efree[2], dfull, dfree, ofull[2], ofree[2], rawfull[2], rawfree[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, g = warp >> 2;       // TMEM lane quarter; rank inside the quarter

    if (tid == 0) {
        mbar_init(&wres, 1);
        mbar_init(&xfull, C::NAUX); mbar_init(&xfree, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&efull[i], 1); mbar_init(&efree[i], C::NWORK); mbar_init(&ofull[i], 1); mbar_init(&ofree[i], 4); mbar_init(&rawfull[i], 1); mbar_init(&rawfree[i], C::NAUX); }
        mbar_init(&dfull, C::NWORK); mbar_init(&dfree, 1);
        mbar_fence_init();
    }
    if (warp == 15) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the staged input: padding pixels [HPIX, NPX) and the 7 zero channels behind the ones channel are never written again
    for (int i = tid; i < (C::XH + C::XL) / 4; i += C::NT) st4(Xhi + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ntile = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tpi = tiles_x * tiles_y;
    const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;
    auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        b = tile / tpi;
        const int t = tile - b * tpi;
        const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
        oy0 = ty * C::TH; ox0 = (t - ty * tiles_x) * C::TW;
    };
    const size_t plane = (size_t)H * W;

    // Output epilogue of tile ti for the 32 pixels of TMEM lane quarter q (thread = pixel), in two halves so that the HBM latency of the
    // residual is not on the caller's path: epi_fetch (address + residual loads) is issued a phase early, epi_store reads O (hi +
    // correction columns), adds b2 and the residual and writes the tile.
    float er[C::RES ? C::COUT : 1];
    size_t eoff = 0;
    bool eok = false;
    auto epi_fetch = [&](int ti) {
        int b, oy0, ox0;
        origin(ti, b, oy0, ox0);
        const int p = q * 32 + lane, gy = oy0 + (p >> 4), gx = ox0 + (p & 15);
        eok = gy < H && gx < W;
        eoff = ((size_t)b * C::COUT * H + min(gy, H - 1)) * W + min(gx, W - 1);
        if (C::RES) {
#pragma unroll
            for (int i = 0; i < C::COUT; ++i) er[i] = __ldg(x + eoff + i * plane);
        }
    };
    auto epi_store = [&](int ti) {
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + C::TM_O + (ti & 1) * 2 * C::COUTP;
        float r[C::COUT];
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < C::COUT; c0 += 8) {
            float vh[8], vl[8];
            tmem_ld8(ta + c0, vh);
            tmem_ld8(ta + C::COUTP + c0, vl);
#pragma unroll
            for (int i = 0; i < 8; ++i) r[c0 + i] = (vh[i] + vl[i]) + __ldg(wts + C::OFF_B2 + c0 + i) + (C::RES ? er[c0 + i] : 0.f);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ofree[ti & 1]);
        if (eok) {
#pragma unroll
            for (int i = 0; i < C::COUT; ++i) y[eoff + i * plane] = r[i];
        }
    };

    if (warp == 15) {
        // ================= tensor-core warp =================
        if (ntile > 0 && elect_one()) {
            constexpr uint32_t IDESC_E = umma_idesc_tf32(C::NPX) & ~(1u << 15);          // A and B K-major
            constexpr uint32_t IDESC3A = umma_idesc_tf32(2 * C::COUTP), IDESC3B = umma_idesc_tf32(C::COUTP);
            mbar_expect_tx(&wres, C::WRES * 4);
            bulk_load(Wr, wts, C::WRES * 4, &wres);
            const uint64_t dw1h = umma_desc(smem_u32(Wr + C::OFF_W1H), 128, (C::KX / 4) * 128, 0);
            const uint64_t dw1l = umma_desc(smem_u32(Wr + C::OFF_W1L), 128, (C::KX / 4) * 128, 0);
            const uint64_t dxh = umma_desc(smem_u32(Xhi), 128, (C::KX / 4) * 128, 0);
            const uint64_t dxl = umma_desc(smem_u32(Xlo), 128, (C::CIN / 4) * 128, 0);
            const uint64_t ddh = umma_desc(smem_u32(DAhi), 1024, 512, 1), ddl = umma_desc(smem_u32(DAlo), 1024, 512, 1);
            const uint64_t dw2 = umma_desc(smem_u32(Wr + C::OFF_W2), 128, (C::CMID / 4) * 128, 0);
            auto expand = [&](int t) {
                const uint32_t acc = tmem + C::TM_E + (t & 1) * C::NPX;
#pragma unroll
                for (int kb = 0; kb < C::KX / 8; ++kb) umma_tf32(acc, dw1h + (uint64_t)(kb * 16), dxh + (uint64_t)(kb * 16), IDESC_E, kb ? 1u : 0u);
#pragma unroll
                for (int kb = 0; kb < C::KX / 8; ++kb) umma_tf32(acc, dw1l + (uint64_t)(kb * 16), dxh + (uint64_t)(kb * 16), IDESC_E, 1u);
#pragma unroll
                for (int kb = 0; kb < C::CIN / 8; ++kb) umma_tf32(acc, dw1h + (uint64_t)(kb * 16), dxl + (uint64_t)(kb * 16), IDESC_E, 1u);
                umma_commit(&xfree);
                umma_commit(&efull[t & 1]);
            };
            auto project = [&](int t) {
                const uint32_t acc = tmem + C::TM_O + (t & 1) * 2 * C::COUTP;
#pragma unroll
                for (int kb = 0; kb < C::CMID / 8; ++kb)
                    umma_tf32(acc, ddh + (uint64_t)(kb * C::KB3 * 4 / 16), dw2 + (uint64_t)(kb * 16), IDESC3A, kb ? 1u : 0u);
#pragma unroll
                for (int kb = 0; kb < C::CMID / 8; ++kb)
                    umma_tf32(acc + C::COUTP, ddl + (uint64_t)(kb * C::KB3 * 4 / 16), dw2 + (uint64_t)(kb * 16), IDESC3B, 1u);
                umma_commit(&dfree);
                umma_commit(&ofull[t & 1]);
            };
            mbar_wait(&wres, 0);
            mbar_wait(&xfull, 0);
            tc_fence_after();
            expand(0);
            for (int t = 0; t < ntile; ++t) {
                if (t + 1 < ntile) {
                    mbar_wait(&xfull, (t + 1) & 1);
                    if (t + 1 >= 2) mbar_wait(&efree[(t + 1) & 1], (((t + 1) >> 1) - 1) & 1);
                    tc_fence_after();
                    TT_TRACE(t, 6);
                    expand(t + 1);
                    TT_TRACE(t, 7);
                }
                mbar_wait(&dfull, t & 1);
                if (t >= 2) mbar_wait(&ofree[t & 1], ((t >> 1) - 1) & 1);
                tc_fence_after();
                TT_TRACE(t, 8);
                project(t);
                TT_TRACE(t, 9);
            }
        }
    } else if (warp >= 16 || (q == 3 && g == C::NAUX)) {
        // ================= output epilogue warps, one per TMEM lane quarter (warps 16, 17, 18 and 11): each waits for EVERY phase of
        // ofull in order (a parity wait that skipped a phase could pass on the phase before), and the next completion of the same
        // barrier needs the warp's own ofree arrival =================
        for (int t = 0; t < ntile; ++t) {
            epi_fetch(t);
            mbar_wait(&ofull[t & 1], (t >> 1) & 1);            // project MMA of tile t complete
            epi_store(t);
        }
    } else if (q == 3) {
        // ================= input staging warps =================
        // Staging: the raw halo box of tile t (16 channels x 10 rows x 24 columns, zero outside the image) arrives by TMA two tiles ahead;
        // these warps split it into the expand operand X^T (hi | lo, one 16-byte core-matrix row per pixel and 4 channels) and set the
        // ones channel. A thread's items (halo pixel n, channel group kg) are the same every tile: offsets precomputed.
        const int ta = g * 32 + lane;                          // staging thread 0..63
        uint32_t pk[C::NIT], pl[C::NIT];                       // raw offset | Xhi offset << 16;  Xlo offset | r << 16 | j << 24
#pragma unroll
        for (int i = 0; i < C::NIT; ++i) {
            const int item = min(ta + i * C::NTA, C::NITEM - 1);
            const int kg = item / C::HPIX, n = item - kg * C::HPIX;
            const int r = n / C::HW, j = n - r * C::HW;
            pk[i] = (uint32_t)(kg * 4 * C::HR * C::RW + r * C::RW + j + 3) | ((uint32_t)(((n >> 3) * (C::KX / 4) + kg) * 32 + (n & 7) * 4) << 16);
            pl[i] = (uint32_t)(((n >> 3) * (C::CIN / 4) + kg) * 32 + (n & 7) * 4) | ((uint32_t)r << 16) | ((uint32_t)j << 24);
        }
        auto issue = [&](int ti) {                             // one thread: the TMA box of tile ti
            int b, oy0, ox0;
            origin(ti, b, oy0, ox0);
            mbar_expect_tx(&rawfull[ti & 1], C::RAW * 4);
            tma_load4(Raw + (ti & 1) * C::RAW, &xmap, &rawfull[ti & 1], ox0 - 4, oy0 - 1, 0, b);
        };
        auto put = [&](int ti) {
            int b, oy0, ox0;
            origin(ti, b, oy0, ox0);
            const float* raw = Raw + (ti & 1) * C::RAW;
#pragma unroll
            for (int i = 0; i < C::NIT; ++i) {
                if (ta + i * C::NTA < C::NITEM) {
                    const float* rp = raw + (pk[i] & 0xFFFFu);
                    float hi[4], lo[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float v = rp[kk * C::HR * C::RW];
                        hi[kk] = tf32_hi(v);
                        lo[kk] = v - hi[kk];
                    }
                    float* xh = Xhi + (pk[i] >> 16);
                    st4(xh, make_float4(hi[0], hi[1], hi[2], hi[3]));
                    st4(Xlo + (pl[i] & 0xFFFFu), make_float4(lo[0], lo[1], lo[2], lo[3]));
                    if (ta + i * C::NTA < C::HPIX) {           // channel group 0 also sets the ones channel: 1 inside the image (carries the expand bias)
                        const int gy = oy0 - 1 + (int)((pl[i] >> 16) & 255u), gx = ox0 - 1 + (int)(pl[i] >> 24);
                        const bool ok = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                        st4(xh + (C::CIN / 4) * 32, make_float4(ok ? 1.f : 0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xfull); mbar_arrive(&rawfree[ti & 1]); }
        };
        {
            if (ta == 0) {
                tma_prefetch_desc(&xmap);
                if (ntile > 0) issue(0);
                if (ntile > 1) issue(1);
            }
            for (int t = 0; t < ntile; ++t) {
                if (ta == 0) TT_TRACE(t, 10);
                mbar_wait(&rawfull[t & 1], (t >> 1) & 1);      // the box of tile t has landed
                if (ta == 0) TT_TRACE(t, 11);
                if (t >= 1) mbar_wait(&xfree, (t - 1) & 1);    // the expand MMA of tile t - 1 has read X
                if (ta == 0) TT_TRACE(t, 12);
                put(t);
                if (ta == 0) TT_TRACE(t, 13);
                if (ta == 0 && t + 2 < ntile) {
                    mbar_wait(&rawfree[t & 1], (t >> 1) & 1);  // both staging warps are done with this buffer
                    issue(t + 2);
                }
            }
        }
    } else {
        // ================= worker warps: thread = mid channel, warp = (channel quarter, pair of output rows) =================
        const int c = q * 32 + lane;
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = __ldg(wts + C::OFF_WD + c * 9 + t);
        const float bd = __ldg(wts + C::OFF_BD + c);
        // operand D: a_idx(p, c) with p = 32 g + 16 rr + ox
        float* dh = DAhi + (c >> 3) * C::KB3 + g * 256 + ((c >> 2) & 1) * 128 + (c & 3) * 32;
        float* dl = dh + C::DA;
        const bool swp = (c >> 2) & 1;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int t = 0; t < ntile; ++t) {
            const uint32_t te = tmem + lane_base + C::TM_E + (t & 1) * C::NPX + (2 * g) * C::HW;
            if (tid == 0) TT_TRACE(t, 0);
            mbar_wait(&efull[t & 1], (t >> 1) & 1);
            tc_fence_after();
            if (tid == 0) TT_TRACE(t, 1);
            float e[3][18], a[2][16];
            tmem_ld_row18(te, e[0]);
            tmem_ld_row18(te + C::HW, e[1]);
            tmem_ld_row18(te + 2 * C::HW, e[2]);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr == 1) {
                    tmem_ld_row18(te + 3 * C::HW, e[0]);       // fourth halo row replaces the first
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&efree[t & 1]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) a[rr][i] = bd;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float (&er)[18] = e[(rr + dy) % 3];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[rr][i] = fmaf(w[dy * 3 + dx], er[i + dx], a[rr][i]);
                }
            }
            // everything above overlapped the project MMA of tile t - 1; only the operand stores have to wait for it
            if (tid == 0) TT_TRACE(t, 2);
            if (t >= 1) mbar_wait(&dfree, (t - 1) & 1);
            if (tid == 0) TT_TRACE(t, 3);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
                for (int o8 = 0; o8 < 2; ++o8) {
                    // Lanes c and c + 4 of a 128-bit store phase own the same banks (the operand's 32-byte chunks are swizzled by c & 3
                    // only): the odd 4-channel group writes the two pixel quads of an 8-pixel chunk in the opposite order.
                    float v[2][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p0 = fmaxf(a[rr][o8 * 8 + i], 0.f), p1 = fmaxf(a[rr][o8 * 8 + 4 + i], 0.f);
                        v[0][i] = swp ? p1 : p0;
                        v[1][i] = swp ? p0 : p1;
                    }
                    const int ch = ((2 * rr + o8) ^ (c & 3)) << 3;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { hi[i] = tf32_hi(v[h][i]); lo[i] = v[h][i] - hi[i]; }
                        const int o = ch + ((h ^ (int)swp) << 2);
                        st4(dh + o, make_float4(hi[0], hi[1], hi[2], hi[3]));
                        st4(dl + o, make_float4(lo[0], lo[1], lo[2], lo[3]));
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dfull);
            if (tid == 0) TT_TRACE(t, 4);
            if (tid == 0) TT_TRACE(t, 5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 15) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
