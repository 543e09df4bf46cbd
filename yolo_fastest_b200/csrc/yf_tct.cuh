// yf_tct.cuh — "channel-lane" tensor-core inverted-residual block (yolo_fastest.py:52-66): res3_1..6, conv3_2..3_4, res4_1..4.
//
// yf_tc.cuh computes E[pixel][mid] (TMEM lane = pixel) and has to push every E value through shared memory before the depthwise can
// see a pixel's neighbours; its workers spend their time on that round trip. Here the expand GEMM is issued TRANSPOSED:
//
//   S1  E^T[mid][halo px] = W1^T[mid][CIN + 1] . X^T[CIN + 1][halo px]   tcgen05.mma kind::tf32, M = 128 mid channels per MMA tile, N = halo pixels
//
// so a TMEM lane is a mid CHANNEL and its columns are the tile's halo pixels, row after row. A worker thread owns one channel: it
// reads the halo rows of ITS channel straight from TMEM into registers (tcgen05.ld, 10 columns per row), keeps its nine depthwise
// weights in registers for the whole kernel and runs the 3x3 entirely in registers — E never exists in shared memory, there is no
// per-chunk weight traffic and no block-wide barrier. The expand bias rides in the GEMM as an extra "ones" input channel that is 1
// inside the image and 0 outside, which also makes E exactly 0 on the depthwise's zero padding: the read-out is one FMAX per value.
//
//   dw  D = relu(dw3x3(relu(E^T)) + bd)                                 registers; written as the project operand (hi | lo)
//   S3  O[out px][COUT] = D[out px][mid] . W2[mid][COUT]                as in yf_tc.cuh: A = D MN-major swizzled, B = [W2hi | W2lo]
//
// Tile = TH x 8 output pixels, TH = 16 (128 pixels = one full MMA tile of the project GEMM) or 8 (64 pixels, for the blocks whose
// operands would not fit otherwise); halo (TH + 2) x 10 pixels = N of the expand MMA (rounded up to 16). More than 128 mid channels
// take a second M tile (its own TMEM columns). TMEM holds two E^T buffers — the expand GEMM of tile t + 1 runs while the workers are
// on tile t — and one or two O buffers. The raw input box of a tile ([CIN][TH + 2][16] from column ox0 - 4, zero outside the image)
// arrives by TMA two tiles ahead; two staging warps split it into the expand operand (hi | lo).
//
// Warps: warp % 4 is the TMEM lane quarter a warp may touch, so roles are dealt per quarter (IrbTtCfg::role): a worker warp =
// (M tile, channel quarter, group of 4 output rows), channel = 128 mt + 32 quarter + lane, 32 output pixels per thread and tile; one
// output-epilogue warp per 32 pixels of the tile (thread = pixel); two staging warps and the tensor-core warp in the free slots.
// fp32 parity through 3xTF32 exactly as in yf_tc.cuh.
//
// Operand layouts: W1^T (A) and X^T (B) are both K-major, un-swizzled core matrices [row / 8][k / 4][row % 8][k % 4] — the layout the
// weight operands of every other kernel use; the staging writes one 16-byte core-matrix row per (pixel, 4 channels). The last M
// tile stores only its real rows (rounded up to 8): the MMA reads on into whatever follows, into TMEM lanes nobody looks at.
// D and W2 as in yf_tc.cuh (a 64-pixel tile stores two of the four 32-pixel atoms of a channel block; same over-read).
// Packed weights (floats): [W1hi: rows x KX][W1lo][W2: (hi COUTP | lo COUTP) x CMID][pad][wd: CMID x 9][bd: CMID][b2: COUT].
#pragma once
#include "yf_tc.cuh"
#include "yf_tma.cuh"

namespace yf {

// one halo row of this thread's channel: 10 consecutive TMEM columns (load + wait in ONE statement: the registers are only valid
// after the wait), ReLU applied — the bias came through the GEMM
__device__ __forceinline__ void tmem_ld_row10(uint32_t taddr, float (&v)[10]) {
    uint32_t r[10];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%10];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%8,%9}, [%11];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9])
        : "r"(taddr), "r"(taddr + 8) : "memory");
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = fmaxf(__uint_as_float(r[i]), 0.f);
}

__device__ __forceinline__ void st4g(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// the same for a stride-2 block: 17 columns per input row
__device__ __forceinline__ void tmem_ld_row17(uint32_t taddr, float (&v)[17]) {
    uint32_t r[17];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%16}, [%18];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16])
        : "r"(taddr), "r"(taddr + 16) : "memory");
#pragma unroll
    for (int i = 0; i < 17; ++i) v[i] = fmaxf(__uint_as_float(r[i]), 0.f);
}

// -DYF_TC_TRACE -DYF_TCT_TRACE: CTA 0 records clock64() per tile: slots 0..4 worker warp 0, 6..9 tensor-core thread, 10..13 staging thread 0
#if defined(YF_TC_TRACE) && defined(YF_TCT_TRACE)
#ifndef YF_TCT_TRACE_CMID
#define YF_TCT_TRACE_CMID 96
#endif
#ifndef YF_TCT_TRACE_S
#define YF_TCT_TRACE_S 1
#endif
#define TT_TRACE(tile, ev) do { if (C::CMID == YF_TCT_TRACE_CMID && C::S == YF_TCT_TRACE_S && blockIdx.x == 0 && (tile) < 64) g_tc_trace[(tile) * 16 + (ev)] = clock64(); } while (0)
#else
#define TT_TRACE(tile, ev) do { } while (0)
#endif

enum : int { TT_WORKER = 0, TT_EPI = 1, TT_STAGE = 2, TT_MMA = 3, TT_IDLE = 4 };

template <int CIN_, int CMID_, int COUT_, int TH_, bool RES_, int TQ_ = 0, int OCC_ = 1, int S_ = 1, bool SKIP_ = false, bool RELU_OUT_ = false>
struct IrbTtCfg {
    static constexpr bool RELU_OUT = RELU_OUT_;                           // ReLU after the projection (conv5_1 is a conv_norm_relu, yolo_fastest.py:124)
    static constexpr bool SKIP = SKIP_;                                   // also write relu(expand) to HBM (conv4_2, the neck's skip tensor; stride-2 blocks only)
    static constexpr int S = S_;                                          // stride of the depthwise (2: the downsampling blocks, yolo_fastest.py:98-100)
    static constexpr int OCC = OCC_;                                      // CTAs per SM (shared memory, TMEM columns and registers are sized for it)
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_;
    static constexpr bool RES = RES_;
    static constexpr int TH = TH_, TW = 8, OPIX = TH * TW;
    static constexpr int G = TH / 4;                                      // groups of 4 output rows = worker warps per (M tile, quarter)
    static constexpr int HR = S * (TH - 1) + 3, HW = S * (TW - 1) + 3, HPIX = HR * HW;      // input halo tile, pixel n = r * HW + j
    static constexpr int NSPL = cdiv(HPIX, 256);                          // expand MMAs per K block (N <= 256 each)
    static constexpr int NPX = rup(HPIX, 16 * NSPL), NCH = NPX / NSPL;    // halo pixels incl. padding; N of one expand MMA
    static constexpr int KX = CIN + 8;                                    // + the ones channel (and 7 zero channels: K advances by 8)
    static constexpr int NMT = cdiv(CMID, 128);                           // M tiles of the expand GEMM
    // The second M tile's channels 128.. sit in TMEM lane quarter TQ.. of that tile (its A operand simply starts 32 TQ rows early, on
    // rows of the first tile), so that its workers fall into a quarter with free warps and the block stays at 16 warps.
    static constexpr int TQ = TQ_;
    static constexpr int COUTP = rup(COUT, 16);
    static constexpr int NA = OPIX / 32, KB3 = NA * 256;                  // 32-pixel atoms / floats per 8-channel block of the operand D
    static constexpr int NAUX = S == 2 ? 3 : 2, NTA = NAUX * 32;          // staging warps
    // packed weights
    static constexpr int W1ROWS = (NMT - 1) * 128 + rup(CMID - (NMT - 1) * 128, 8);
    static constexpr int OFF_W1H = 0, OFF_W1L = W1ROWS * KX, OFF_W2 = 2 * W1ROWS * KX, WRES = rup(OFF_W2 + 2 * COUTP * CMID, 256);
    static constexpr int OFF_WD = WRES, OFF_BD = OFF_WD + CMID * 9, OFF_B2 = OFF_BD + CMID, WFLOATS = rup(OFF_B2 + COUT, 32);
    // shared memory (floats)
    static constexpr int XH = NPX * KX, XL = NPX * CIN, DA = (CMID / 8) * KB3;
    static constexpr int RW = rup(HW + 3, 4), RAW = CIN * HR * RW;        // raw input box [CIN][HR][RW] from column S ox0 - 4 (TMA boxes start 16-byte aligned)
    static constexpr int SMEM_FLOATS = WRES + 2 * DA + XH + XL + 2 * RAW;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;
    // TMEM columns: one or two E buffers, then one or two O buffers
    static constexpr int EB = NMT * NPX;
    static constexpr int NEB = (2 * EB + 2 * COUTP <= 512 / OCC) ? 2 : 1, TM_O = NEB * EB;
    static constexpr int TCOLS = 512 / OCC;
    static constexpr int NOB = (TM_O + 2 * 2 * COUTP <= TCOLS) ? 2 : 1;
    static constexpr int NITEM = HPIX * (CIN / 4), NIT = cdiv(NITEM, NTA);       // staging items (halo pixel, 4 input channels) per tile / per staging thread

    // ---- roles. Worker warps of quarter q: G per M tile that has channels in that quarter.
    __host__ __device__ static constexpr int wcnt(int q) {
        int n = 0;
        if (q * 32 < CMID && q * 32 < 128) n += G;
        if (NMT == 2 && q >= TQ && 128 + (q - TQ) * 32 < CMID) n += G;
        return n;
    }
    static constexpr int NWORK = wcnt(0) + wcnt(1) + wcnt(2) + wcnt(3);
    __host__ __device__ static constexpr int base_role(int w) {                               // worker / epilogue / free slot (TT_IDLE)
        const int q = w & 3, g = w >> 2;
        return g < wcnt(q) ? TT_WORKER : (g == wcnt(q) && q < NA) ? TT_EPI : TT_IDLE;
    }
    __host__ __device__ static constexpr int role(int w) {                                    // the free slots, in warp order: staging, staging, tensor-core warp
        if (base_role(w) != TT_IDLE) return base_role(w);
        int nfree = 0;
        for (int v = 0; v < w; ++v) nfree += base_role(v) == TT_IDLE ? 1 : 0;
        return nfree < NAUX ? TT_STAGE : nfree == NAUX ? TT_MMA : TT_IDLE;
    }
    __host__ __device__ static constexpr int stage_rank(int w) {
        int n = 0;
        for (int v = 0; v < w; ++v) n += role(v) == TT_STAGE ? 1 : 0;
        return n;
    }
    __host__ __device__ static constexpr int nwarp() {                                        // up to the last warp that has a job
        int last = 0;
        for (int w = 0; w < 32; ++w) if (role(w) != TT_IDLE) last = w;
        return last + 1;
    }
    static constexpr int NWARP = nwarp(), NT = NWARP * 32;

    static_assert(TH == 4 || TH == 8 || TH == 16, "tile = 32, 64 or 128 pixels");
    static_assert(!SKIP || S == 2, "the dual output is written by the stride-2 worker loop");
    static_assert(CMID % 8 == 0 && NMT <= 2 && (TQ == 0 || NMT == 2) && TQ * 32 + CMID - 128 <= 128, "mid channels: whole 8-channel operand blocks, at most two M tiles");
    static_assert(CIN % 8 == 0 && COUT % 8 == 0 && NCH % 16 == 0 && NCH <= 256 && TM_O + NOB * 2 * COUTP <= TCOLS, "MMA shape / TMEM columns");
    static_assert((S == 1 || S == 2) && (S == 1 || !RES), "stride");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert((RAW * 4) % 128 == 0 && RAW < 65536 && XH < 65536, "raw box alignment / packed staging offsets");
    static_assert(OCC * (SMEM_BYTES + 1024) <= 227 * 1024 + 1024, "does not fit shared memory");
    static_assert(NT <= 1024, "too many warps");
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::OCC)
irbt_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ x, float* __restrict__ y, float* __restrict__ skip, const float* __restrict__ wts,
            int Hin, int Win, int H, int W, int tiles_x, int tiles_y, int total_tiles) {
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
    const int role = C::role(warp);

    if (tid == 0) {
        mbar_init(&wres, 1);
        mbar_init(&xfull, C::NAUX); mbar_init(&xfree, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&efull[i], 1); mbar_init(&efree[i], C::NWORK); mbar_init(&ofull[i], 1); mbar_init(&ofree[i], C::NA);
            mbar_init(&rawfull[i], 1); mbar_init(&rawfree[i], C::NAUX);
        }
        mbar_init(&dfull, C::NWORK); mbar_init(&dfree, 1);
        mbar_fence_init();
        mbar_expect_tx(&wres, C::WRES * 4);                     // resident operands: issued before the dependency wait (weights are not activations)
        bulk_load(Wr, wts, C::WRES * 4, &wres);
    }
    if (role == TT_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the staged input: padding pixels [HPIX, NPX) and the 7 zero channels behind the ones channel are never written again
    for (int i = tid; i < (C::XH + C::XL) / 4; i += C::NT) st4(Xhi + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
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

    if (role == TT_MMA) {
        // ================= tensor-core warp =================
        if (ntile > 0 && elect_one()) {
            constexpr uint32_t IDESC_E = umma_idesc_tf32(C::NCH) & ~(1u << 15);          // A and B K-major
            constexpr uint32_t IDESC3A = umma_idesc_tf32(2 * C::COUTP), IDESC3B = umma_idesc_tf32(C::COUTP);
            const uint64_t dw1h = umma_desc(smem_u32(Wr + C::OFF_W1H), 128, (C::KX / 4) * 128, 0);
            const uint64_t dw1l = umma_desc(smem_u32(Wr + C::OFF_W1L), 128, (C::KX / 4) * 128, 0);
            const uint64_t dxh = umma_desc(smem_u32(Xhi), 128, (C::KX / 4) * 128, 0);
            const uint64_t dxl = umma_desc(smem_u32(Xlo), 128, (C::CIN / 4) * 128, 0);
            const uint64_t ddh = umma_desc(smem_u32(DAhi), 1024, 512, 1), ddl = umma_desc(smem_u32(DAlo), 1024, 512, 1);
            const uint64_t dw2 = umma_desc(smem_u32(Wr + C::OFF_W2), 128, (C::CMID / 4) * 128, 0);
            auto expand = [&](int t) {
#pragma unroll
                for (int mt = 0; mt < C::NMT; ++mt) {
                    const uint64_t mo = (uint64_t)((mt * (128 - 32 * C::TQ) * C::KX * 4) >> 4);   // this M tile's rows of W1
#pragma unroll
                    for (int ns = 0; ns < C::NSPL; ++ns) {
                        const uint32_t acc = tmem + (t % C::NEB) * C::EB + mt * C::NPX + ns * C::NCH;
                        const uint64_t xho = (uint64_t)(((ns * C::NCH / 8) * (C::KX / 4) * 128) >> 4);      // this chunk's pixels of X^T
                        const uint64_t xlo = (uint64_t)(((ns * C::NCH / 8) * (C::CIN / 4) * 128) >> 4);
                        // the two small correction terms first, the main term last: the tensor core accumulates with truncation, and every
                        // add onto an accumulator that already holds the main term costs up to one ulp of it
#pragma unroll
                        for (int kb = 0; kb < C::CIN / 8; ++kb) umma_tf32(acc, dw1h + mo + (uint64_t)(kb * 16), dxl + xlo + (uint64_t)(kb * 16), IDESC_E, kb ? 1u : 0u);
#pragma unroll
                        for (int kb = 0; kb < C::KX / 8; ++kb) umma_tf32(acc, dw1l + mo + (uint64_t)(kb * 16), dxh + xho + (uint64_t)(kb * 16), IDESC_E, 1u);
#pragma unroll
                        for (int kb = 0; kb < C::KX / 8; ++kb) umma_tf32(acc, dw1h + mo + (uint64_t)(kb * 16), dxh + xho + (uint64_t)(kb * 16), IDESC_E, 1u);
                    }
                }
                umma_commit(&xfree);
                umma_commit(&efull[t % C::NEB]);
            };
            auto project = [&](int t) {
                const uint32_t acc = tmem + C::TM_O + (t % C::NOB) * 2 * C::COUTP;
#pragma unroll
                for (int kb = 0; kb < C::CMID / 8; ++kb)
                    umma_tf32(acc, ddh + (uint64_t)(kb * C::KB3 * 4 / 16), dw2 + (uint64_t)(kb * 16), IDESC3A, kb ? 1u : 0u);
#pragma unroll
                for (int kb = 0; kb < C::CMID / 8; ++kb)
                    umma_tf32(acc + C::COUTP, ddl + (uint64_t)(kb * C::KB3 * 4 / 16), dw2 + (uint64_t)(kb * 16), IDESC3B, 1u);
                umma_commit(&dfree);
                umma_commit(&ofull[t % C::NOB]);
            };
            mbar_wait(&wres, 0);
            mbar_wait(&xfull, 0);
            tc_fence_after();
            expand(0);
            for (int t = 0; t < ntile; ++t) {
                if (t + 1 < ntile) {
                    mbar_wait(&xfull, (t + 1) & 1);
                    if (t + 1 >= C::NEB) mbar_wait(&efree[(t + 1) % C::NEB], (((t + 1) / C::NEB) - 1) & 1);
                    tc_fence_after();
                    TT_TRACE(t, 6);
                    expand(t + 1);
                    TT_TRACE(t, 7);
                }
                mbar_wait(&dfull, t & 1);
                if (t >= C::NOB) mbar_wait(&ofree[t % C::NOB], ((t / C::NOB) - 1) & 1);
                tc_fence_after();
                TT_TRACE(t, 8);
                project(t);
                TT_TRACE(t, 9);
            }
        }
    } else if (role == TT_EPI) {
        // ================= output epilogue warps, one per 32 pixels of the tile (thread = pixel): O hi + correction columns + b2
        // (+ residual, yolo_fastest.py:65) -> HBM. Each waits for EVERY phase of its ofull barrier in order (a parity wait that skipped a
        // phase could pass on the phase before); the next completion of the same barrier needs the warp's own ofree arrival. =========
        const int p = q * 32 + lane;
        for (int t = 0; t < ntile; ++t) {
            int b, oy0, ox0;
            origin(t, b, oy0, ox0);
            const int gy = oy0 + (p >> 3), gx = ox0 + (p & 7);
            const bool ok = gy < H && gx < W;
            const size_t off = ((size_t)b * C::COUT * H + min(gy, H - 1)) * W + min(gx, W - 1);
            float r[C::COUT];
#pragma unroll
            for (int i = 0; i < C::COUT; ++i) r[i] = (C::RES ? __ldg(x + off + i * plane) : 0.f) + __ldg(wts + C::OFF_B2 + i);
            const int ob = t % C::NOB;
            mbar_wait(&ofull[ob], (t / C::NOB) & 1);           // project MMA of tile t complete
            tc_fence_after();
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + C::TM_O + ob * 2 * C::COUTP;
#pragma unroll
            for (int c0 = 0; c0 < C::COUT; c0 += 8) {
                float vh[8], vl[8];
                tmem_ld8(ta + c0, vh);
                tmem_ld8(ta + C::COUTP + c0, vl);
#pragma unroll
                for (int i = 0; i < 8; ++i) r[c0 + i] += vh[i] + vl[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ofree[ob]);
            if (ok) {
#pragma unroll
                for (int i = 0; i < C::COUT; ++i) y[off + i * plane] = C::RELU_OUT ? fmaxf(r[i], 0.f) : r[i];
            }
        }
    } else if (role == TT_STAGE) {
        // ================= input staging warps: the raw halo box of tile t arrives by TMA two tiles ahead; these warps split it into
        // the expand operand X^T (hi | lo, one 16-byte core-matrix row per pixel and 4 channels) and set the ones channel. A thread's
        // items (halo pixel n, channel group kg) are the same every tile: offsets precomputed. =================
        const int ta = C::stage_rank(warp) * 32 + lane;
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
            tma_load4(Raw + (ti & 1) * C::RAW, &xmap, &rawfull[ti & 1], C::S * ox0 - 4, C::S * oy0 - 1, 0, b);
        };
        if (ta == 0) {
            tma_prefetch_desc(&xmap);
            if (ntile > 0) issue(0);
            if (ntile > 1) issue(1);
        }
        for (int t = 0; t < ntile; ++t) {
            if (ta == 0) TT_TRACE(t, 10);
            mbar_wait(&rawfull[t & 1], (t >> 1) & 1);          // the box of tile t has landed
            if (ta == 0) TT_TRACE(t, 11);
            if (t >= 1) mbar_wait(&xfree, (t - 1) & 1);        // the expand MMA of tile t - 1 has read X
            if (ta == 0) TT_TRACE(t, 12);
            int b, oy0, ox0;
            origin(t, b, oy0, ox0);
            const float* raw = Raw + (t & 1) * C::RAW;
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
                        const int gy = C::S * oy0 - 1 + (int)((pl[i] >> 16) & 255u), gx = C::S * ox0 - 1 + (int)(pl[i] >> 24);
                        const bool ok = (unsigned)gy < (unsigned)Hin && (unsigned)gx < (unsigned)Win;
                        st4(xh + (C::CIN / 4) * 32, make_float4(ok ? 1.f : 0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xfull); mbar_arrive(&rawfree[t & 1]); }
            if (ta == 0) TT_TRACE(t, 13);
            if (ta == 0 && t + 2 < ntile) {
                mbar_wait(&rawfree[t & 1], (t >> 1) & 1);      // both staging warps are done with this buffer
                issue(t + 2);
            }
        }
    } else if (role == TT_WORKER) {
        // ================= worker warps: thread = mid channel, warp = (M tile, channel quarter, group of 4 output rows) =================
        // the g-th worker warp of quarter q: M tiles that have channels in this quarter come first to last, G row groups each
        const int mt = g / C::G, rg = g - mt * C::G;
        const int c = mt ? 128 + (q - C::TQ) * 32 + lane : q * 32 + lane;      // (the first M tile is full when there is a second, so tile index = g / G)
        const bool cvalid = c < C::CMID;
        const int cl = min(c, C::CMID - 1);
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = __ldg(wts + C::OFF_WD + cl * 9 + t);
        const float bd = __ldg(wts + C::OFF_BD + cl);
        // operand D: a_idx(p, c) with p = 32 rg + 8 rr + ox: atom rg, 32-byte chunk rr ^ (c & 3)
        float* dh = DAhi + (cl >> 3) * C::KB3 + rg * 256 + ((cl >> 2) & 1) * 128 + (cl & 3) * 32;
        float* dl = dh + C::DA;
        const bool swp = (cl >> 2) & 1;
        const uint32_t te0 = tmem + ((uint32_t)(q * 32) << 16) + mt * C::NPX + (C::S * 4 * rg) * C::HW;
        for (int t = 0; t < ntile; ++t) {
            const uint32_t te = te0 + (t % C::NEB) * C::EB;
            if (tid == 0) TT_TRACE(t, 0);
            mbar_wait(&efull[t % C::NEB], (t / C::NEB) & 1);
            tc_fence_after();
            if (tid == 0) TT_TRACE(t, 1);
            float a[4][8];
            if (C::S == 1) {
                float e[6][10];
#pragma unroll
                for (int r = 0; r < 6; ++r) tmem_ld_row10(te + r * C::HW, e[r]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&efree[t % C::NEB]);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[rr][i] = bd;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int i = 0; i < 8; ++i) a[rr][i] = fmaf(w[dy * 3 + dx], e[rr + dy][i + dx], a[rr][i]);
                }
            } else {
                // stride 2: output row rr reads input rows 2 rr .. 2 rr + 2 (17 columns each); three rows live at a time, row r in slot r % 3
                float e[3][17];
                tmem_ld_row17(te, e[0]);
                int sb = 0, soy0 = 0, sox0 = 0;
                if (C::SKIP) origin(t, sb, soy0, sox0);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    tmem_ld_row17(te + (2 * rr + 1) * C::HW, e[(2 * rr + 1) % 3]);
                    tmem_ld_row17(te + (2 * rr + 2) * C::HW, e[(2 * rr + 2) % 3]);
                    if (C::SKIP) {
                        // input rows 1..8 and columns 1..16 of the halo tile are the pixels this tile OWNS (each input pixel is owned by
                        // exactly one tile): relu(expand) of this thread's channel goes out as 64-byte row segments
#pragma unroll
                        for (int k = 1; k <= 2; ++k) {
                            const int r = 2 * rr + k, gy = 2 * (soy0 + 4 * rg) - 1 + r;
                            if (cvalid && gy < Hin) {
                                float* sp = skip + (((size_t)sb * C::CMID + c) * Hin + gy) * Win + 2 * sox0;
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4)
                                    if (2 * sox0 + 4 * j4 < Win)
                                        st4g(sp + 4 * j4, make_float4(e[r % 3][1 + 4 * j4], e[r % 3][2 + 4 * j4], e[r % 3][3 + 4 * j4], e[r % 3][4 + 4 * j4]));
                            }
                        }
                    }
                    if (rr == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&efree[t % C::NEB]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[rr][i] = bd;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int i = 0; i < 8; ++i) a[rr][i] = fmaf(w[dy * 3 + dx], e[(2 * rr + dy) % 3][2 * i + dx], a[rr][i]);
                }
            }
            // everything above overlapped the project MMA of tile t - 1; only the operand stores have to wait for it
            if (tid == 0) TT_TRACE(t, 2);
            if (t >= 1) mbar_wait(&dfree, (t - 1) & 1);
            if (tid == 0) TT_TRACE(t, 3);
            if (cvalid) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    // Lanes c and c + 4 of a 128-bit store phase own the same banks (the operand's 32-byte chunks are swizzled by c & 3
                    // only): the odd 4-channel group writes the two pixel quads of a row in the opposite order.
                    float v[2][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p0 = fmaxf(a[rr][i], 0.f), p1 = fmaxf(a[rr][4 + i], 0.f);
                        v[0][i] = swp ? p1 : p0;
                        v[1][i] = swp ? p0 : p1;
                    }
                    const int ch = (rr ^ (cl & 3)) << 3;
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
            if (tid == 0) TT_TRACE(t, 5);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dfull);
            if (tid == 0) TT_TRACE(t, 4);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (role == TT_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
