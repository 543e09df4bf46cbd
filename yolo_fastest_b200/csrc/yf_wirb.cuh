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
//     (they contribute nothing in scatter form), columns outside are killed by starting the expand accumulator from b1 * mask with
//     the TMA-zero-filled x: relu(0) = 0.  No per-element predicate anywhere in the row loop.
//
// Packed weights (floats): [W1: CIN x CMID (k-major)][b1: CMID][Wd: 9 x CMID (tap-major)][bd: CMID][W2: COUT x CMID (n-major)][b2: COUT]
#pragma once
#include "yf_kernels.cuh"
#include "yf_tma.cuh"

namespace yf {

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
    static_assert(!RES || (S == 1 && CIN == COUT), "residual needs same shape");
    static_assert((XBOX * 4) % 128 == 0 && (DROW * 4) % 16 == 0, "boxes are 128-byte aligned");
    static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

// input rows a band of R output rows needs, and the boxes that hold them
template <class C> __host__ __device__ constexpr int wirb_nin(int R) { return C::S * R + 3 - C::S; }
template <class C> __host__ __device__ constexpr int wirb_nch(int R) { return (wirb_nin<C>(R) + C::RC - 1) / C::RC; }

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

    // ---- this warp's units and the flat sequence of their boxes ---------------------------------------------------------------
    const int NCH = wirb_nch<C>(R), NIN = wirb_nin<C>(R);
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
    float resv[2][PG];
#pragma unroll
    for (int i = 0; i < PG; ++i) resv[0][i] = resv[1][i] = 0.f;

    int g = 0;
    for (int k = 0; k < my_units; ++k) {
        int ub, y0, x0;
        unit_origin(gw + k * tw, ub, y0, x0);
        // column masks of this lane's window: 1 inside the image
        float mc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int gx = S * x0 - 1 + C::LSTEP * sl + j;
            mc[j] = ((unsigned)gx < (unsigned)Win) ? 1.f : 0.f;
        }
        const int ox = x0 + PG * pq;                            // first output column of this lane's projection item
        float* yrow = y + (((size_t)ub * C::COUT + pn) * Hout) * Wout + ox;
        const bool col_ok = ox < Wout;                          // Wout % 4 == 0 and PG | 4: an item is entirely inside or outside
        const int oy_end = min(y0 + R, Hout);

        for (int c = 0; c < NCH; ++c, ++g) {
            mbar_wait(&bars[warp][g & 1], (g >> 1) & 1);
            const float* Xs = Xb + (g & 1) * C::XBOX;
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                const int r = c * RC + rr;                      // input row of the unit; image row iy
                const int iy = S * y0 - 1 + r;
                if (r < NIN) {                                  // warp-uniform
                    // accumulator roles of this row (compile-time: RC is a multiple of the rotation period)
                    //   S == 1: E[iy] closes output row iy-1 (dy = 2), feeds iy (dy = 1), opens iy+1 (dy = 0)
                    //   S == 2: odd iy = 2*oy - 1 opens oy (dy = 0) and closes oy-1 (dy = 2); even iy = 2*oy feeds oy (dy = 1)
                    constexpr bool dummy = false; (void)dummy;
                    const bool row_in = (unsigned)iy < (unsigned)Hin;
                    float2 e[NC];
                    if (row_in) {
#pragma unroll
                        for (int j = 0; j < NC; ++j) e[j] = mul2s(b1, mc[j]);
#pragma unroll
                        for (int kk = 0; kk < CIN; ++kk) {
                            float xv[NC];
                            load_win<NC>(xv, Xs + (kk * RC + rr) * C::XW + C::XOFF - 1 + C::LSTEP * sl);
#pragma unroll
                            for (int j = 0; j < NC; ++j) e[j] = fma2s(w1[kk], xv[j], e[j]);
                        }
#pragma unroll
                        for (int j = 0; j < NC; ++j) e[j] = relu2(e[j]);
                    }
                    if (S == 1) {
                        float2* fin = acc[(rr + 1) % 3];
                        float2* mid = acc[(rr + 2) % 3];
                        float2* nw = acc[rr % 3];
                        if (row_in) {
#pragma unroll
                            for (int i = 0; i < SPX; ++i) {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) {
                                    fin[i] = fma2(wd[6 + dx], e[i + dx], fin[i]);
                                    mid[i] = fma2(wd[3 + dx], e[i + dx], mid[i]);
                                    nw[i] = fma2(wd[dx], e[i + dx], dx == 0 ? bd : nw[i]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < SPX; ++i) nw[i] = bd;
                        }
                        if (C::RES) {
                            // out = project(...) + x (yolo_fastest.py:65): the residual of output row iy is this box row; it is
                            // consumed one row later, when E[iy+1] has closed the row
#pragma unroll
                            for (int i = 0; i < PG; ++i) resv[rr & 1][i] = Xs[(pn * RC + rr) * C::XW + C::XOFF + PG * pq + i];
                        }
                    } else {
                        if ((rr & 1) == 0) {
                            float2* fin = acc[((rr >> 1) + 1) & 1];
                            float2* nw = acc[(rr >> 1) & 1];
                            if (row_in) {
#pragma unroll
                                for (int i = 0; i < SPX; ++i) {
#pragma unroll
                                    for (int dx = 0; dx < 3; ++dx) {
                                        fin[i] = fma2(wd[6 + dx], e[2 * i + dx], fin[i]);
                                        nw[i] = fma2(wd[dx], e[2 * i + dx], dx == 0 ? bd : nw[i]);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < SPX; ++i) nw[i] = bd;
                            }
                        } else if (row_in) {
                            float2* mid = acc[(rr >> 1) & 1];
#pragma unroll
                            for (int i = 0; i < SPX; ++i)
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) mid[i] = fma2(wd[3 + dx], e[2 * i + dx], mid[i]);
                        }
                    }
                    // ---- close an output row: depthwise ReLU -> D (warp-private smem) -> 1x1 projection -> HBM ---------------
                    const bool closes = S == 1 ? true : (rr & 1) == 0;
                    const int oy = S == 1 ? iy - 1 : (iy - 1) / 2;          // S == 2: iy = 2*oy' - 1 closes oy' - 1 = (iy - 1) / 2
                    if (closes && r >= 2 && oy < oy_end) {
                        const float2* fin = S == 1 ? acc[(rr + 1) % 3] : acc[((rr >> 1) + 1) & 1];
                        const int par = S == 1 ? (rr & 1) : ((rr >> 1) & 1);
                        float* D = Db + par * C::DROW;
                        {
                            float* d0 = D + (2 * pl) * C::DWS + SPX * sl;
#pragma unroll
                            for (int i4 = 0; i4 < SPX / 4; ++i4) {
                                const float2 a = relu2(fin[4 * i4]), b = relu2(fin[4 * i4 + 1]), cc = relu2(fin[4 * i4 + 2]), d = relu2(fin[4 * i4 + 3]);
                                st4(d0 + 4 * i4, make_float4(a.x, b.x, cc.x, d.x));
                                st4(d0 + C::DWS + 4 * i4, make_float4(a.y, b.y, cc.y, d.y));
                            }
                        }
                        __syncwarp();
                        float2 o[PG / 2];
#pragma unroll
                        for (int i = 0; i < PG / 2; ++i) {
                            o[i] = make_float2(b2, b2);
                            if (C::RES) { o[i].x += resv[(rr + 1) & 1][2 * i]; o[i].y += resv[(rr + 1) & 1][2 * i + 1]; }
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
                                    o[0] = fma2s(make_float2(dv.x, dv.y), wv[t], o[0]);
                                    o[PG / 2 - 1] = fma2s(make_float2(dv.z, dv.w), wv[t], o[PG / 2 - 1]);
                                } else {
                                    const float2 dv = *reinterpret_cast<const float2*>(dm);
                                    o[0] = fma2s(dv, wv[t], o[0]);
                                }
                            }
                        }
                        if (col_ok) {
                            float* yp = yrow + (size_t)oy * Wout;
                            if (PG == 4) st4(yp, make_float4(o[0].x, o[0].y, o[PG / 2 - 1].x, o[PG / 2 - 1].y));
                            else *reinterpret_cast<float2*>(yp) = o[0];
                        }
                    }
                }
            }
            __syncwarp();                                       // every lane is done with this box
            if (gi < total_chunks) { issue(gi); ++gi; }         // refill it with the box two steps ahead
        }
    }
}

}  // namespace yf
