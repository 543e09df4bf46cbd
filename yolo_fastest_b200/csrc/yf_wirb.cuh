// yf_wirb.cuh — warp-streaming engine for the THIN inverted-residual groups (mid width <= 32): res1_1, res2_1, res2_2 and the
// stride-2 transition conv2_2 -> conv2_3 -> conv3_1 (yolo_fastest.py:52-66,94-97).  These groups move the most HBM bytes per FLOP
// of the network (4-8 channels at 1/2 and 1/4 resolution); the block-cooperative engine of yf_kernels.cuh spent its time on
// per-element cp.async staging, three block barriers per chunk and the shared-memory round trips of E and D.  Here:
//
//   * one WARP owns a unit = (image, band of R output rows, strip of OW output columns) and streams down its rows; warps never
//     synchronise with each other (no __syncthreads after set-up), a persistent CTA is just NW independent warps;
//   * the input halo rows arrive by tensor-tile TMA (cp.async.bulk.tensor.4d -> UTMALDG, yf_tma.cuh): boxes of RC rows x (OW + halo)
//     columns x all CIN channels at start coordinates (S*x0 - 4, S*y0 - 1 + c*RC) — negative at the image border; the column start is the
//     16-byte aligned one left of the 1-pixel halo because sm_100a faults on an unaligned innermost coordinate (yf_tma.cuh) —
//     zero-filled outside the image, double buffered per warp, completion on the warp's own mbarriers, issued by one elected lane;
//   * lane = (mid-channel PAIR pl, column strip sl): the 1x1 expand, ReLU and the depthwise 3x3 of its two channels over its SPX output
//     columns stay in REGISTERS.  Every FMA is a packed FFMA2 over the channel pair (fma.rn.f32x2).  The depthwise conv is evaluated in
//     scatter form: a new expanded row E[iy] is added into the accumulators of the output rows it touches (iy-1, iy, iy+1 for stride 1),
//     so no window of E rows is kept and nothing is recomputed vertically; horizontally a lane recomputes its 2 halo columns;
//   * only the depthwise output row D (CMID x OW floats) passes through warp-private shared memory, so that the 1x1 projection can
//     re-partition the lanes as (output channel n, pixel group q); one __syncwarp per output row (D is double buffered);
//   * zero padding of the depthwise INPUT (= the activation E, not the bias, yolo_fastest.py:17-22): rows outside the image are skipped
//     get a zero bias, so with the TMA-zero-filled x they expand to relu(0) = 0; of the columns only the two window-edge columns can lie
//     outside the image next to a stored output, and they are multiplied by a 0/1 mask.  No per-element predicate in the row loop.
//
// Packed weights (floats): [W1: CIN x CMID (k-major)][b1: CMID][Wd: 9 x CMID (tap-major)][bd: CMID][W2: COUT x CMID (n-major)][b2: COUT]
#pragma once
#include "yf_kernels.cuh"
#include "yf_tma.cuh"

namespace yf {

// (b - 128) / 255 of byte K of `word`, bit for bit the reference's fp32 division (detect.py:124), without a table: the byte goes into the
// mantissa of 2^23 (one PRMT), n = b - 128 exactly, q = n * (1 / 255) is within one ulp, and one residual step q + (n - 255 q) * (1 / 255)
// rounds correctly — checked exhaustively for the 256 byte values in exact rational arithmetic (tests/test_preprocess_cpu.py).
template <int K>
__device__ __forceinline__ float norm_u8(uint32_t word) {
    const float n = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | K)) - 8388736.0f;
    const float r = 1.0f / 255.0f;
    const float q = __fmul_rn(n, r);
    return __fmaf_rn(__fmaf_rn(-255.0f, q, n), r, q);
}

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fma2s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 mul2s(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }

// NV window values starting ONE float left of a 16-byte boundary: scalar, aligned 128-bit words, scalar tail
template <int NV>
__device__ __forceinline__ void load_win(float (&dst)[NV], const float* __restrict__ p) {
    dst[0] = p[0];
#pragma unroll
    for (int i = 0; i < (NV - 1) / 4; ++i) {
        const float4 v = ld4(p + 1 + 4 * i);
        dst[1 + 4 * i] = v.x; dst[2 + 4 * i] = v.y; dst[3 + 4 * i] = v.z; dst[4 + 4 * i] = v.w;
    }
#pragma unroll
    for (int i = 1 + ((NV - 1) / 4) * 4; i < NV; ++i) dst[i] = p[i];
}

template <int CIN_, int CMID_, int COUT_, int S_, int SPX_, int NSL_, int RC_, int NW_, bool RES_, bool W2REG_>
struct WirbCfg {
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, S = S_, SPX = SPX_, NSL = NSL_, RC = RC_, NW = NW_;
    static constexpr bool RES = RES_, W2REG = W2REG_;
    static constexpr int NPL = 32 / NSL;                       // lanes along the mid-channel pairs
    static constexpr int OW = NSL * SPX;                       // output columns of a warp row
    static constexpr int NC = S == 1 ? SPX + 2 : 2 * SPX + 1;  // expanded columns a lane needs for its SPX outputs
    static constexpr int XOFF = 4;                             // the box starts XOFF columns left of S*x0 (16-byte aligned start, yf_tma.cuh)
    static constexpr int XW = S == 1 ? OW + 8 : 2 * OW + 4;    // box width (floats): columns S*x0 - 4 .. S*(x0 + OW - 1) + 1, rounded to 16 bytes
    static constexpr int LSTEP = S * SPX;                      // window of strip sl starts at box column XOFF - 1 + LSTEP * sl
    static constexpr int XBOX = CIN * RC * XW;                 // floats of one box
    static constexpr int DWS = OW + 4;                         // D row stride
    static constexpr int DROW = CMID * DWS;
    static constexpr int PG = COUT * OW / 32;                  // projection item = 1 output channel x PG pixels
    static constexpr int NQ = OW / PG;
    static constexpr int WARP_FLOATS = 2 * XBOX + 2 * DROW;
    static constexpr int W2S = CMID + 4;                       // padded row of the shared projection matrix (conflict-free broadcast reads)
    static constexpr int OFF_W1 = 0, OFF_B1 = CIN * CMID, OFF_WD = OFF_B1 + CMID, OFF_BD = OFF_WD + 9 * CMID, OFF_W2 = OFF_BD + CMID,
                         OFF_B2 = OFF_W2 + COUT * CMID;
    static constexpr int WFLOATS = rup(OFF_B2 + COUT, 4);
    static constexpr int SMEM_BYTES = (NW * WARP_FLOATS + (W2REG ? 0 : COUT * W2S)) * 4 + 128;
    static constexpr int ROT = S == 1 ? 3 : 2;                 // accumulator rotation period (rows of E per period: 3 / 4)
    static_assert(NSL * NPL == 32 && CMID == 2 * NPL, "one mid-channel pair per lane");
    static_assert(SPX % 4 == 0 && (PG == 2 || PG == 4) && COUT * NQ == 32, "bad lane partition");
    static_assert(S == 1 ? RC % 6 == 0 : RC % 4 == 0, "chunk rows must be a multiple of the rotation periods");
    static_assert(!RES || (S == 1 && CIN == COUT && PG == 4), "residual needs same shape");
    static_assert((XBOX * 4) % 128 == 0 && (DROW * 4) % 16 == 0, "boxes are 128-byte aligned");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

// input rows a band of R output rows needs, and the boxes that hold them
template <class C> __host__ __device__ constexpr int wirb_nin(int R) { return C::S * R + 3 - C::S; }
template <class C> __host__ __device__ constexpr int wirb_nch(int R) { return (wirb_nin<C>(R) + C::RC - 1) / C::RC; }

template <class C> __host__ __device__ constexpr bool closes_every_row() { return C::S == 1; }

template <class C>
__global__ void __launch_bounds__(C::NW * 32, 1)
wirb_kernel(const __grid_constant__ CUtensorMap xmap, float* __restrict__ y, const float* __restrict__ wts,
            int Hin, int Win, int Hout, int Wout, int R, int nstrips, int nbands, int total_units) {
    constexpr int S = C::S, SPX = C::SPX, NC = C::NC, RC = C::RC, PG = C::PG, CMID = C::CMID, CIN = C::CIN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[C::NW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* smem = reinterpret_cast<float*>(smem_raw + ((128 - (tma_smem_u32(smem_raw) & 127)) & 127));
    float* Xb = smem + (size_t)warp * C::WARP_FLOATS;           // two input boxes [CIN][RC][XW]
    float* Db = Xb + 2 * C::XBOX;                               // two D rows [CMID][DWS]
    float* W2sh = smem + (size_t)C::NW * C::WARP_FLOATS;        // shared projection matrix [COUT][W2S] (unless it lives in registers)
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    if (threadIdx.x == 0) tma_prefetch_desc(&xmap);
    if (!C::W2REG)
        for (int i = threadIdx.x; i < C::COUT * CMID; i += C::NW * 32) W2sh[(i / CMID) * C::W2S + (i % CMID)] = __ldg(wts + C::OFF_W2 + i);
    mbar_fence_init();
    __syncthreads();                                            // the only block-wide barrier

    // ---- lane roles ------------------------------------------------------------------------------------------------------
    const int pl = lane / C::NSL, sl = lane % C::NSL;           // expand / depthwise: channel pair, column strip
    const int pn = lane / C::NQ, pq = lane % C::NQ;             // projection: output channel, pixel group
    float2 w1[CIN], wd[9];
#pragma unroll
    for (int k = 0; k < CIN; ++k) w1[k] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_W1 + k * CMID) + pl);
#pragma unroll
    for (int t = 0; t < 9; ++t) wd[t] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_WD + t * CMID) + pl);
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_B1) + pl);
    const float2 bd = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_BD) + pl);
    const float b2 = __ldg(wts + C::OFF_B2 + pn);
    float w2r[C::W2REG ? CMID : 1];
    if (C::W2REG) {
#pragma unroll
        for (int m = 0; m < CMID; ++m) w2r[m] = __ldg(wts + C::OFF_W2 + pn * CMID + m);
    }
    const float* w2s = W2sh + pn * C::W2S;
    pdl_trigger();
    pdl_wait();                                                 // weights are in registers / shared memory: now the predecessor's output may be read

    // ---- this warp's units and the flat sequence of their boxes ---------------------------------------------------------------
    const int NCH = wirb_nch<C>(R);
    const int gw = blockIdx.x * C::NW + warp, tw = gridDim.x * C::NW;
    const int my_units = gw < total_units ? (total_units - gw + tw - 1) / tw : 0;
    const int total_chunks = my_units * NCH;
    auto unit_origin = [&](int u, int& b, int& y0, int& x0) {
        const int strip = u % nstrips;
        const int t = u / nstrips;
        y0 = (t % nbands) * R;
        x0 = strip * C::OW;
        b = t / nbands;
    };
    auto issue = [&](int gi) {                                  // all lanes call; one issues
        int b, y0, x0;
        unit_origin(gw + (gi / NCH) * tw, b, y0, x0);
        const int c = gi % NCH;
        if (lane == 0) {
            mbar_expect_tx(&bars[warp][gi & 1], C::XBOX * 4);
            tma_load4(Xb + (gi & 1) * C::XBOX, &xmap, &bars[warp][gi & 1], S * x0 - C::XOFF, S * y0 - 1 + c * RC, 0, b);
        }
    };
    int gi = 0;
    for (; gi < 2 && gi < total_chunks; ++gi) issue(gi);

    float2 acc[C::ROT][SPX];
#pragma unroll
    for (int a = 0; a < C::ROT; ++a)
#pragma unroll
        for (int i = 0; i < SPX; ++i) acc[a][i] = make_float2(0.f, 0.f);
    float resv[3][PG];                                          // residual rows in flight: loaded with E[iy], consumed two rows later
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < PG; ++i) resv[a][i] = 0.f;
    float* pend = nullptr;                                      // output address of the row whose D is waiting in shared memory (null: none)

    // 1x1 projection of the D row written one step earlier (+ bias + residual) -> HBM. Deferred by one row so that its shared-memory
    // loads and FMA chains interleave with the expand of the next row instead of serialising behind the __syncwarp.
    auto project = [&](const float* D, const float (&res)[PG]) {
        float2 o[2][PG / 2];                                    // two accumulation chains per output pair (even / odd mid channels)
#pragma unroll
        for (int i = 0; i < PG / 2; ++i) {
            o[0][i] = make_float2(b2, b2);
            o[1][i] = C::RES ? make_float2(res[2 * i], res[2 * i + 1]) : make_float2(0.f, 0.f);
        }
        const float* dp = D + PG * pq;
#pragma unroll
        for (int m4 = 0; m4 < CMID / 4; ++m4) {
            float wv[4];
            if (C::W2REG) {
#pragma unroll
                for (int t = 0; t < 4; ++t) wv[t] = w2r[C::W2REG ? 4 * m4 + t : 0];
            } else {
                const float4 w = ld4(w2s + 4 * m4);
                wv[0] = w.x; wv[1] = w.y; wv[2] = w.z; wv[3] = w.w;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float* dm = dp + (4 * m4 + t) * C::DWS;
                if (PG == 4) {
                    const float4 dv = ld4(dm);
                    o[t & 1][0] = fma2s(make_float2(dv.x, dv.y), wv[t], o[t & 1][0]);
                    o[t & 1][PG / 2 - 1] = fma2s(make_float2(dv.z, dv.w), wv[t], o[t & 1][PG / 2 - 1]);
                } else {
                    const float2 dv = *reinterpret_cast<const float2*>(dm);
                    o[t & 1][0] = fma2s(dv, wv[t], o[t & 1][0]);
                }
            }
        }
        if (pend) {
            if (PG == 4) st4(pend, make_float4(o[0][0].x + o[1][0].x, o[0][0].y + o[1][0].y, o[0][PG / 2 - 1].x + o[1][PG / 2 - 1].x, o[0][PG / 2 - 1].y + o[1][PG / 2 - 1].y));
            else *reinterpret_cast<float2*>(pend) = make_float2(o[0][0].x + o[1][0].x, o[0][0].y + o[1][0].y);
        }
    };

    int g = 0;
    for (int k = 0; k < my_units; ++k) {
        int ub, y0, x0;
        unit_origin(gw + k * tw, ub, y0, x0);
        // Zero padding in x: of the columns whose E feeds a STORED output only the image column -1 (window column 0 of the first strip)
        // and, for stride 1, the image column Win (window column NC-1 of the lane that owns the last outputs) lie outside the image
        // (Wout is a multiple of SPX, checked on the host), so two multiplicative masks per row replace a mask per column.
        const float m_left = (S * x0 + C::LSTEP * sl) > 0 ? 1.f : 0.f;
        const float m_right = (S * x0 - 1 + C::LSTEP * sl + NC - 1) < Win ? 1.f : 0.f;
        const int ox = x0 + PG * pq;                            // first output column of this lane's projection item
        const bool col_ok = ox < Wout;                          // Wout % 4 == 0 and PG | 4: an item is entirely inside or outside
        float* ycur = y + (((size_t)ub * C::COUT + pn) * Hout + y0) * Wout + ox;       // output row the next close writes
        const int oy_end = min(y0 + R, Hout);

        for (int c = 0; c < NCH; ++c, ++g) {
            mbar_wait(&bars[warp][g & 1], (g >> 1) & 1);
            const float* Xs = Xb + (g & 1) * C::XBOX;
            // The host picks R so that the unit's input rows fill its boxes (S == 1: exactly; S == 2: one spare row that only feeds an
            // accumulator nobody closes), so every row of every box runs the same branch-free body.
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r = c * RC + rr;                      // input row of the unit; image row iy
                const int iy = S * y0 - 1 + r;
                // ---- projection of the row closed one step ago (independent of everything below) ------------------------------
                const bool closes = S == 1 ? true : (rr & 1) == 0;
                const int par = S == 1 ? (rr & 1) : ((rr >> 1) & 1);
                if (closes) project(Db + (par ^ 1) * C::DROW, resv[(rr + 1) % 3]);
                // ---- expand: E[iy] over this lane's window. A row outside the image is all zeros: kill the bias (x is TMA zero fill)
                const bool row_in = (unsigned)iy < (unsigned)Hin;
                const float2 b1r = row_in ? b1 : make_float2(0.f, 0.f);
                float2 e[NC];
#pragma unroll
                for (int kk = 0; kk < CIN; ++kk) {
                    float xv[NC];
                    load_win<NC>(xv, Xs + (kk * RC + rr) * C::XW + C::XOFF - 1 + C::LSTEP * sl);
#pragma unroll
                    for (int j = 0; j < NC; ++j) e[j] = fma2s(w1[kk], xv[j], kk == 0 ? b1r : e[j]);
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) e[j] = relu2(e[j]);
                e[0] = mul2s(e[0], m_left);
                if (S == 1) e[NC - 1] = mul2s(e[NC - 1], m_right);
                // ---- depthwise 3x3 in scatter form (accumulator roles are compile-time: RC is a multiple of the rotation period)
                //   S == 1: E[iy] closes output row iy-1 (dy = 2), feeds iy (dy = 1), opens iy+1 (dy = 0)
                //   S == 2: odd iy = 2*oy - 1 opens oy (dy = 0) and closes oy-1 (dy = 2); even iy = 2*oy feeds oy (dy = 1)
                const float2* fin;
                if (S == 1) {
                    float2* fn = acc[(rr + 1) % 3];
                    float2* mid = acc[(rr + 2) % 3];
                    float2* nw = acc[rr % 3];
#pragma unroll
                    for (int i = 0; i < SPX; ++i) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            fn[i] = fma2(wd[6 + dx], e[i + dx], fn[i]);
                            mid[i] = fma2(wd[3 + dx], e[i + dx], mid[i]);
                            nw[i] = fma2(wd[dx], e[i + dx], dx == 0 ? bd : nw[i]);
                        }
                    }
                    fin = fn;
                    if (C::RES) {
                        // out = project(...) + x (yolo_fastest.py:65): the residual of output row iy is this box row
                        const float4 rv = ld4(Xs + (pn * RC + rr) * C::XW + C::XOFF + PG * pq);
                        resv[rr % 3][0] = rv.x; resv[rr % 3][1] = rv.y; resv[rr % 3][PG - 2] = rv.z; resv[rr % 3][PG - 1] = rv.w;
                    }
                } else if ((rr & 1) == 0) {
                    float2* fn = acc[((rr >> 1) + 1) & 1];
                    float2* nw = acc[(rr >> 1) & 1];
#pragma unroll
                    for (int i = 0; i < SPX; ++i) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            fn[i] = fma2(wd[6 + dx], e[2 * i + dx], fn[i]);
                            nw[i] = fma2(wd[dx], e[2 * i + dx], dx == 0 ? bd : nw[i]);
                        }
                    }
                    fin = fn;
                } else {
                    float2* mid = acc[(rr >> 1) & 1];
#pragma unroll
                    for (int i = 0; i < SPX; ++i)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) mid[i] = fma2(wd[3 + dx], e[2 * i + dx], mid[i]);
                    fin = mid;
                }
                // ---- close an output row: depthwise ReLU -> D (warp-private smem); its projection runs during the next step ----
                if (closes) {
                    float* d0 = Db + par * C::DROW + (2 * pl) * C::DWS + SPX * sl;
#pragma unroll
                    for (int i4 = 0; i4 < SPX / 4; ++i4) {
                        const float2 a = relu2(fin[4 * i4]), b = relu2(fin[4 * i4 + 1]), cc = relu2(fin[4 * i4 + 2]), d = relu2(fin[4 * i4 + 3]);
                        st4(d0 + 4 * i4, make_float4(a.x, b.x, cc.x, d.x));
                        st4(d0 + C::DWS + 4 * i4, make_float4(a.y, b.y, cc.y, d.y));
                    }
                    const int oy = S == 1 ? iy - 1 : (iy - 1) >> 1;         // S == 2: iy = 2*oy' - 1 closes oy' - 1
                    const bool real = r >= 2 && oy < oy_end && col_ok;
                    pend = real ? ycur : nullptr;
                    if (r >= 2) ycur += Wout;
                    __syncwarp();
                }
            }
            if (!closes_every_row<C>()) __syncwarp();           // every lane is done with this box
            if (gi < total_chunks) { issue(gi); ++gi; }         // refill it with the box two steps ahead
        }
    }
    // the last closed row of this warp is still waiting in shared memory
    project(Db + ((C::S == 1 ? (RC - 1) & 1 : ((RC - 2) >> 1) & 1)) * C::DROW, resv[(RC - 2) % 3]);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Stem group on the same engine: conv0 (dense 3x3 s2, 1 -> 8, ReLU) -> conv1_2 (1x1 8 -> 8 ReLU) -> conv1_3 (dw3x3 ReLU) -> conv1_4
// (1x1 8 -> 4 linear)   (yolo_fastest.py:78-82,151-154).  Input [B,1,H,W] fp32 or uint8, output [B,4,H/2,W/2].
// A unit is (image, band of R output rows, strip of 32 output columns).  Per expanded row iy the warp first computes the conv0 row
// (8 channels x 34 columns: lane = (channel pair, 4 columns), the 2 extra halo columns by lanes 0..7) from three raw rows of the TMA
// box into a warp-private row buffer, then runs the expand / scatter-depthwise / deferred-projection row of wirb_kernel with that
// buffer as its 8-channel input.  Raw boxes: 13 rows (6 expanded rows need raw rows 2*iy0-1 .. 2*iy0+11; consecutive boxes share one
// row) x 72 fp32 columns from the aligned column 2*x0 - 4, or x 96 uint8 columns from 2*x0 - 16; uint8 boxes are normalised into an
// fp32 copy through a 256-entry table of the reference's own fp32 division (x - 128) / 255 (detect.py:124).
// Packed weights: [W0: 9 x 8 (tap-major)][b0: 8][W1: 8 x 8 (k-major)][b1: 8][Wd: 9 x 8][bd: 8][W2: 4 x 8 (n-major)][b2: 4]
// ---------------------------------------------------------------------------------------------------------------------------------
template <int NW_>
struct WstemCfg {
    static constexpr int NW = NW_, S = 1, RC = 6, SPX = 4, NSL = 8, OW = 32, NC = 6, CMID = 8, COUT = 4, PG = 4, NQ = 8;
    static constexpr int RAWH = 2 * RC + 1;                    // raw rows of one box
    static constexpr int RAWW = 72;                            // fp32 raw columns: image columns 2*x0 - 4 .. 2*x0 + 67
    static constexpr int RAWWU = 96;                           // uint8 box columns: image columns 2*x0 - 16 .. 2*x0 + 79
    static constexpr int RAWF = rup(RAWH * RAWW, 32);          // floats of one fp32 raw buffer (128-byte multiple)
    static constexpr int RAWU = rup(RAWH * RAWWU, 128);        // bytes of one uint8 box
    static constexpr int C0W = 40, C0F = 8 * C0W;              // conv0 row buffer [8][40] (two of them): column 0 = output column x0 - 1; 40:
                                                               // the stage-0 stores of a 128-bit phase (4 strips x 2 channel pairs) hit 8 different bank groups
    static constexpr int DWS = OW + 4, DROW = CMID * DWS;
    static constexpr int OFF_W0 = 0, OFF_B0 = 72, OFF_W1 = 80, OFF_B1 = 144, OFF_WD = 152, OFF_BD = 224, OFF_W2 = 232, OFF_B2 = 264;
    static constexpr int WFLOATS = 268;
    template <bool U8> __host__ __device__ static constexpr int warp_bytes() { return (U8 ? 2 * RAWU + RAWF * 4 : 2 * RAWF * 4) + 2 * C0F * 4 + 2 * DROW * 4; }
    template <bool U8> __host__ __device__ static constexpr int smem_bytes() { return NW * warp_bytes<U8>() + 128; }
    static_assert(smem_bytes<false>() <= 227 * 1024 && smem_bytes<true>() <= 227 * 1024, "does not fit shared memory");
};

template <class C, bool U8>
__global__ void __launch_bounds__(C::NW * 32, 1)
wstem_kernel(const __grid_constant__ CUtensorMap xmap, float* __restrict__ y, const float* __restrict__ wts,
             int Hin, int Win, int Hout, int Wout, int R, int nstrips, int nbands, int total_units) {
    constexpr int RC = C::RC, NC = C::NC, PG = C::PG, CMID = C::CMID, SPX = C::SPX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[C::NW][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* base = smem_raw + ((128 - (tma_smem_u32(smem_raw) & 127)) & 127);
    unsigned char* wb = base + (size_t)warp * C::template warp_bytes<U8>();
    unsigned char* Ub = wb;                                                     // U8: two uint8 boxes
    float* Rf = reinterpret_cast<float*>(wb + (U8 ? 2 * C::RAWU : 0));          // fp32 raw rows: U8 one normalised copy, else two boxes
    float* C0b = Rf + (U8 ? 1 : 2) * C::RAWF;                                   // two conv0 row buffers
    float* Db = C0b + 2 * C::C0F;
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    if (threadIdx.x == 0) tma_prefetch_desc(&xmap);
    mbar_fence_init();
    __syncthreads();

    const int pl = lane >> 3, sl = lane & 7;                    // expand / depthwise: channel pair, column strip
    const int pn = lane >> 3, pq = lane & 7;                    // projection: output channel, pixel group
    // stage 0 (conv0) uses its own partition: the 8 lanes of a 128-bit shared-memory phase are 4 strips x 2 channel pairs, so their
    // raw-row loads (32 bytes apart per strip) span 128 bytes — with 8 strips per phase they were 2-way bank conflicted
    const int s0 = ((lane >> 3) & 1) * 4 + (lane & 3), c0p = ((lane >> 4) << 1) | ((lane >> 2) & 1);
    float2 w0[9], w1[8], wd[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w0[t] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_W0 + t * 8) + c0p);
#pragma unroll
    for (int k = 0; k < 8; ++k) w1[k] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_W1 + k * 8) + pl);
#pragma unroll
    for (int t = 0; t < 9; ++t) wd[t] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_WD + t * 8) + pl);
    const float2 b0 = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_B0) + c0p);
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_B1) + pl);
    const float2 bd = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_BD) + pl);
    const float b2 = __ldg(wts + C::OFF_B2 + pn);
    float w2r[CMID];
#pragma unroll
    for (int m = 0; m < CMID; ++m) w2r[m] = __ldg(wts + C::OFF_W2 + pn * CMID + m);
    // the two extra conv0 columns (32, 33) of a row: lanes 0..7 = (channel pair, column)
    const int xp = (lane >> 1) & 3, xc = 32 + (lane & 1);
    float2 w0x[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w0x[t] = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_W0 + t * 8) + xp);
    const float2 b0x = __ldg(reinterpret_cast<const float2*>(wts + C::OFF_B0) + xp);
    pdl_trigger();
    pdl_wait();

    const int NCH = (R + 2) / RC;                               // the host picks R = NCH * RC - 2
    const int gw = blockIdx.x * C::NW + warp, tw = gridDim.x * C::NW;
    const int my_units = gw < total_units ? (total_units - gw + tw - 1) / tw : 0;
    const int total_chunks = my_units * NCH;
    auto unit_origin = [&](int u, int& b, int& y0, int& x0) {
        const int strip = u % nstrips;
        const int t = u / nstrips;
        y0 = (t % nbands) * R;
        x0 = strip * C::OW;
        b = t / nbands;
    };
    auto issue = [&](int gi) {
        int b, y0, x0;
        unit_origin(gw + (gi / NCH) * tw, b, y0, x0);
        const int c = gi % NCH;
        if (lane == 0) {
            const int ry = 2 * (y0 - 1 + c * RC) - 1;           // first raw row of the box
            if (U8) {
                mbar_expect_tx(&bars[warp][gi & 1], C::RAWH * C::RAWWU);
                tma_load4(Ub + (gi & 1) * C::RAWU, &xmap, &bars[warp][gi & 1], 2 * x0 - 16, ry, 0, b);
            } else {
                mbar_expect_tx(&bars[warp][gi & 1], C::RAWH * C::RAWW * 4);
                tma_load4(Rf + (gi & 1) * C::RAWF, &xmap, &bars[warp][gi & 1], 2 * x0 - 4, ry, 0, b);
            }
        }
    };
    int gi = 0;
    for (; gi < 2 && gi < total_chunks; ++gi) issue(gi);

    float2 acc[3][SPX];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < SPX; ++i) acc[a][i] = make_float2(0.f, 0.f);
    float* pend = nullptr;

    auto project = [&](const float* D) {
        float2 o[2][2];
        o[0][0] = o[0][1] = make_float2(b2, b2);
        o[1][0] = o[1][1] = make_float2(0.f, 0.f);
        const float* dp = D + PG * pq;
#pragma unroll
        for (int m = 0; m < CMID; ++m) {
            const float4 dv = ld4(dp + m * C::DWS);
            o[m & 1][0] = fma2s(make_float2(dv.x, dv.y), w2r[m], o[m & 1][0]);
            o[m & 1][1] = fma2s(make_float2(dv.z, dv.w), w2r[m], o[m & 1][1]);
        }
        if (pend) st4(pend, make_float4(o[0][0].x + o[1][0].x, o[0][0].y + o[1][0].y, o[0][1].x + o[1][1].x, o[0][1].y + o[1][1].y));
    };

    int g = 0;
    for (int k = 0; k < my_units; ++k) {
        int ub, y0, x0;
        unit_origin(gw + k * tw, ub, y0, x0);
        const float m_left = (x0 + SPX * sl) > 0 ? 1.f : 0.f;
        const float m_right = (x0 - 1 + SPX * sl + NC - 1) < Wout ? 1.f : 0.f;
        const int ox = x0 + PG * pq;
        const bool col_ok = ox < Wout;
        float* ycur = y + (((size_t)ub * C::COUT + pn) * Hout + y0) * Wout + ox;
        const int oy_end = min(y0 + R, Hout);

        for (int c = 0; c < NCH; ++c, ++g) {
            mbar_wait(&bars[warp][g & 1], (g >> 1) & 1);
            const float* Rs;
            if (U8) {
                // normalise the uint8 box into the fp32 copy: buffer column j = image column 2*x0 - 4 + j = box byte 12 + j
                // (the zero padding is zero AFTER normalisation, detect.py:124 + yolo_fastest.py:17-19: bytes the TMA zero-filled outside the
                // image must become 0.0, not (0 - 128) / 255; W is a multiple of 4, so a group of 4 columns is inside or outside as a whole)
                const unsigned char* ub8 = Ub + (g & 1) * C::RAWU;
                const int ry = 2 * (y0 - 1 + c * RC) - 1, rx = 2 * x0 - 4;
                for (int i = lane; i < C::RAWH * (C::RAWW / 4); i += 32) {
                    const int row = i / (C::RAWW / 4), q4 = i - row * (C::RAWW / 4);
                    const uint32_t v = *reinterpret_cast<const uint32_t*>(ub8 + row * C::RAWWU + 12 + 4 * q4);
                    const bool in = (unsigned)(ry + row) < (unsigned)Hin && (unsigned)(rx + 4 * q4) < (unsigned)Win;
                    st4(Rf + row * C::RAWW + 4 * q4, in ? make_float4(norm_u8<0>(v), norm_u8<1>(v), norm_u8<2>(v), norm_u8<3>(v))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f));
                }
                __syncwarp();
                Rs = Rf;
            } else {
                Rs = Rf + (g & 1) * C::RAWF;
            }
            // stage 0: conv0 row of box row `row` (columns x0 - 1 + [0, 34)) from raw rows 2*row .. 2*row + 2 -> C0 buffer `row & 1`.
            // It runs ONE ROW AHEAD of the pipeline below (row 0 of a box before the loop, row rr + 1 inside iteration rr), so its
            // loads and FMAs interleave with the expand / depthwise of the current row and one __syncwarp per row covers both D and C0.
            auto stage0 = [&](int row) {
                const int iy0 = y0 - 1 + c * RC + row;
                const float mrow = (unsigned)iy0 < (unsigned)Hout ? 1.f : 0.f;     // rows outside the map are zero (the depthwise pads E = f(conv0))
                float* C0 = C0b + (row & 1) * C::C0F;
                float2 a0[4];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    // conv0 column c reads buffer columns 1 + 2c .. 3 + 2c; c = 4*s0 + i
                    const float* rp = Rs + (2 * row + dy) * C::RAWW + 8 * s0;
                    const float4 va = ld4(rp), vb = ld4(rp + 4);
                    const float2 vc = *reinterpret_cast<const float2*>(rp + 8);
                    const float v[10] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) a0[i] = fma2s(w0[dy * 3 + dx], v[1 + 2 * i + dx], (dy == 0 && dx == 0) ? b0 : a0[i]);
                }
                float2 q[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = mul2s(relu2(a0[i]), mrow);
                st4(C0 + (2 * c0p) * C::C0W + 4 * s0, make_float4(q[0].x, q[1].x, q[2].x, q[3].x));
                st4(C0 + (2 * c0p + 1) * C::C0W + 4 * s0, make_float4(q[0].y, q[1].y, q[2].y, q[3].y));
                if (lane < 8) {
                    float2 ax = b0x;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const float* rp = Rs + (2 * row + dy) * C::RAWW + 1 + 2 * xc;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) ax = fma2s(w0x[dy * 3 + dx], rp[dx], ax);
                    }
                    ax = mul2s(relu2(ax), mrow);
                    C0[(2 * xp) * C::C0W + xc] = ax.x;
                    C0[(2 * xp + 1) * C::C0W + xc] = ax.y;
                }
            };
            stage0(0);
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r = c * RC + rr;
                const int iy = y0 - 1 + r;                      // expanded row (H/2 map)
                const int par = rr & 1;
                project(Db + (par ^ 1) * C::DROW);
                if (rr + 1 < RC) stage0(rr + 1);
                const bool row_in = (unsigned)iy < (unsigned)Hout;
                const float* C0 = C0b + (rr & 1) * C::C0F;
                // ---- expand (conv1_2) over this lane's window of the conv0 row ------------------------------------------------
                const float2 b1r = row_in ? b1 : make_float2(0.f, 0.f);
                float2 e[NC];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float* cp = C0 + kk * C::C0W + 4 * sl;
                    const float4 va = ld4(cp);
                    const float2 vb = *reinterpret_cast<const float2*>(cp + 4);
                    const float xv[NC] = {va.x, va.y, va.z, va.w, vb.x, vb.y};
#pragma unroll
                    for (int j = 0; j < NC; ++j) e[j] = fma2s(w1[kk], xv[j], kk == 0 ? b1r : e[j]);
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) e[j] = relu2(e[j]);
                e[0] = mul2s(e[0], m_left);
                e[NC - 1] = mul2s(e[NC - 1], m_right);
                // ---- depthwise 3x3 (conv1_3) in scatter form ----------------------------------------------------------------------
                float2* fn = acc[(rr + 1) % 3];
                float2* mid = acc[(rr + 2) % 3];
                float2* nw = acc[rr % 3];
#pragma unroll
                for (int i = 0; i < SPX; ++i) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        fn[i] = fma2(wd[6 + dx], e[i + dx], fn[i]);
                        mid[i] = fma2(wd[3 + dx], e[i + dx], mid[i]);
                        nw[i] = fma2(wd[dx], e[i + dx], dx == 0 ? bd : nw[i]);
                    }
                }
                // ---- close output row iy - 1 ----------------------------------------------------------------------------------------
                {
                    float* d0 = Db + par * C::DROW + (2 * pl) * C::DWS + SPX * sl;
                    const float2 a = relu2(fn[0]), b = relu2(fn[1]), cc = relu2(fn[2]), d = relu2(fn[3]);
                    st4(d0, make_float4(a.x, b.x, cc.x, d.x));
                    st4(d0 + C::DWS, make_float4(a.y, b.y, cc.y, d.y));
                    const int oy = iy - 1;
                    const bool real = r >= 2 && oy < oy_end && col_ok;
                    pend = real ? ycur : nullptr;
                    if (r >= 2) ycur += Wout;
                    __syncwarp();
                }
            }
            if (gi < total_chunks) { issue(gi); ++gi; }
        }
    }
    project(Db + ((RC - 1) & 1) * C::DROW);
}

}  // namespace yf
