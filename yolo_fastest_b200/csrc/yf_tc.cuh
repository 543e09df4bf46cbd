// yf_tc.cuh — tensor-core (tcgen05) inverted-residual engine for the wide blocks, where the two 1x1 convolutions
// are real dense contractions (K = 16..48 in, N = 96..224 mid, yolo_fastest.py:52-66).
//
//   S1  E[halo px][MC]  = X[halo px][CIN] . W1[CIN][MC]      tcgen05.mma kind::tf32, M = 128 pixels per MMA tile
//   dw  D = relu(dw3x3(relu(E + b1)) + bd)                     CUDA cores, sliding window (as in yf_kernels.cuh)
//   S3  O[out px][COUT] += D[out px][MC] . W2[MC][COUT]        tcgen05.mma, accumulators stay in TMEM across the chunks
//
// fp32 parity is kept with the 3xTF32 split: every operand v is stored as hi = tf32(v) and lo = v - hi and each
// product is issued as hi*hi + hi*lo + lo*hi (fp32 accumulation in TMEM); measured error 8e-7 relative
// (tools/selftest/umma_selftest.cu), i.e. fp32 grade, where a single TF32 pass would give 1e-3.
//
// Warp-specialised, one CTA per SM: NWW worker warps run the CUDA-core phases of a chunk of MC mid channels
// (TMEM -> bias/ReLU/mask -> E in smem; depthwise -> D as MMA operand), one extra warp owns the tensor core: a single
// thread issues every tcgen05.mma and every weight-block bulk copy and talks to the workers through mbarriers only.
// The stream of chunks s = 0, 1, ... (all tiles of this persistent CTA back to back) is software pipelined:
//   expand MMA of chunk s+1 runs during the depthwise phase of chunk s     (trigger: e1free, TMEM E read out)
//   project MMA of chunk s  runs during the TMEM read-out of chunk s+1     (trigger: dfull, operand D written)
//   weight block of chunk s+2 is in flight (3-deep ring), the next tile's input is fetched to registers one phase ahead.
//
// Operand layouts (validated by the selftest):
//   A = activations, MN-major (pixels contiguous) — for 32-bit MN-major operands the only legal shared-memory layout
//       is SWIZZLE_128B_BASE32B: atoms of [4 channels][32 pixels] fp32 = 512 B, the 32-byte chunk index XOR-ed with
//       (channel & 3); two atoms per MMA (K = 8), LBO = stride between 32-pixel atoms, SBO = stride between 4-channel atoms.
//   B = weights, K-major, no swizzle, packed on the host as [n/8][k/4][n%8][k%4] (core matrices of 8 rows x 16 bytes).
//       The project weights are packed as [W2hi ; W2lo] along N, so D_hi . [W2hi | W2lo] is ONE pass over the operand D
//       (N = 2*COUTP) and D_lo . W2hi the second; the output epilogue adds the two column groups.
//   D = TMEM, lane = pixel, column = output channel; the epilogues read it with tcgen05.ld.32x32b (thread = pixel).
#pragma once
#include "yf_kernels.cuh"

namespace yf {

// hi part of the 3xTF32 split: round to nearest (ties away) at 10 mantissa bits, as cvt.rna.tf32.f32 does for finite
// values; written as integer add + mask (2 instructions — the cvt expands to 4 with its inf/nan guard)
__device__ __forceinline__ float tf32_hi(float v) {
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
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
// true in exactly one lane of a converged warp; unlike (lane == 0) the compiler knows the branch holds a single thread, so it issues
// the uniform-datapath tcgen05 instructions directly instead of wrapping each in an ELECT / BRA.U.ANY loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int ID, int N>
__device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }

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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int pow2_ge(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// -DYF_TC_TRACE: CTA 0 records clock64() at its phase boundaries (worker warp 0 and the tensor-core thread), 16 slots per chunk
// step for the first 64 steps; read back through yf_debug_trace (tools/tc_trace.py). Compiled out otherwise.
#ifdef YF_TC_TRACE
__device__ long long g_tc_trace[16 * 64];
#ifndef YF_TC_TRACE_CMID
#define YF_TC_TRACE_CMID 96          // which instantiation records (mid channels): 96 = res3_3..6, 136 = res4_1..4
#endif
#define TC_TRACE(step, ev) do { if (C::CMID == YF_TC_TRACE_CMID && blockIdx.x == 0 && (step) < 64) g_tc_trace[(step) * 16 + (ev)] = clock64(); } while (0)
#else
#define TC_TRACE(step, ev) do { } while (0)
#endif

// row stride of one E channel (floats): >= halo pixels, and chosen so the lanes of a 128-bit shared-memory phase that spill
// into the next channel continue on the banks after this channel's row (stride == row floats mod 32; 16-pixel rows, where
// the phase is shared by channels m and m + 2: stride % 16 == 8)
constexpr int e_stride(int ipix, int nstrip) {
    int v = ipix;
    if (nstrip == 4) { while (v % 16 != 8) v += 4; }
    else { while (v % 32 != (4 * nstrip) % 32) v += 4; }
    return v;
}

template <int CIN_, int CMID_, int COUT_, int TH_, int TW_, int MC_, int RH_, int NWW_, bool RES_, bool E1ALL_ = true, int OCC_ = 1>
struct IrbTcCfg {
    static constexpr int OCC = OCC_;              // CTAs per SM the kernel is compiled for (register cap, TMEM columns)
    static constexpr int CIN = CIN_, CMID = CMID_, COUT = COUT_, MC = MC_, RH = RH_, NWW = NWW_;
    static constexpr int NTW = NWW * 32;          // worker threads
    static constexpr int NT = NTW + 32;           // + the tensor-core warp
    static constexpr int NWB = 3;                 // weight-block ring
    static constexpr bool RES = RES_;
    // E1ALL: ONE expand MMA per tile over all mid channels (N = CMIDP; its weights stay resident in smem), so the operand X is
    // fetched by the tensor core once per tile instead of once per chunk. Otherwise one expand MMA per chunk (N = MC).
    static constexpr bool E1ALL = E1ALL_;
    using G = Geo<3, 1, TH_, TW_>;
    static constexpr int CMIDP = rup(CMID, MC), NCHUNK = CMIDP / MC;
    static constexpr int NE = E1ALL ? NCHUNK : 1;                                 // chunks per expand MMA
    static constexpr int N1 = NE * MC;                                            // N of the expand MMA
    static constexpr int MT1 = cdiv(G::IPIX, 128), MT3 = cdiv(G::OPIX, 128);      // 128-pixel MMA tiles of the halo / output tile
    static constexpr int NG1 = cdiv(G::IPIX, 32), NG3 = cdiv(G::OPIX, 32);        // 32-pixel operand atoms actually stored
    static constexpr int COUTP = rup(COUT, 16);                                   // N of the lo pass (M = 128 needs N % 16 == 0)
    static constexpr int KB1 = NG1 * 256, KB3 = NG3 * 256;                        // floats per 8-channel block of an A region
    static constexpr int XA1 = (CIN / 8) * KB1, DA1 = (MC / 8) * KB3;             // floats of one A region (hi or lo)
    static constexpr int TM_E = 0, TM_O = MT1 * N1;                               // TMEM columns: expand accumulators, project accumulators
    static constexpr int TCOLS = pow2_ge(TM_O + MT3 * 2 * COUTP);
    static constexpr int EPS = e_stride(G::IPIX, G::TW / 4);                      // E channel stride
    // packed weights (floats): [W1hi | W1lo over all CMIDP channels] (E1ALL only), then NCHUNK chunk blocks, then b2
    static constexpr int W1RES = E1ALL ? 2 * CMIDP * CIN : 0;
    static constexpr int W1C = E1ALL ? 0 : 2 * MC * CIN;                          // expand weights inside a chunk block otherwise
    static constexpr int OFF_W1H = 0, OFF_W1L = E1ALL ? CMIDP * CIN : MC * CIN;
    static constexpr int OFF_W2 = W1C, OFF_B1 = OFF_W2 + 2 * COUTP * MC;
    static constexpr int OFF_WD = OFF_B1 + MC, OFF_BD = OFF_WD + MC * 9, CB = rup(OFF_BD + MC, 32);
    static constexpr int OFF_CH = W1RES;                                          // first chunk block
    static constexpr int OFF_B2 = OFF_CH + NCHUNK * CB;
    static constexpr int WFLOATS = OFF_B2 + COUT;
    static constexpr int ES = rup(MC * EPS, 32);
    static constexpr int SMEM_FLOATS = 2 * XA1 + 2 * DA1 + ES + NWB * CB + W1RES;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + 1024;     // + slack to align the dynamic window to 1 KB
    static constexpr int NITEM_X = (CIN / 8) * G::IPIX;           // input staging items (8 channels of one halo pixel) per tile
    static constexpr int NIT = cdiv(NITEM_X, NTW);                // ... per worker thread
    static_assert(CIN % 8 == 0 && MC % 16 == 0 && NWW >= 4, "tcgen05 tiling constraints");
    static_assert(N1 % 16 == 0 && N1 <= 256, "expand MMA N");
    static_assert(TCOLS * OCC <= 512, "TMEM columns");
    static_assert(!RES || CIN == COUT, "residual needs same shape");
    static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
    static_assert(W1RES % 32 == 0, "resident weights: 128-byte granularity");
    // the last MMA tile of an operand region reads up to 3 atoms past the stored ones (results land in unused TMEM lanes);
    // those reads must stay inside the dynamic window: something at least 3 KB long follows every region
    static_assert(ES * 4 >= 4096, "over-read guard");
};

// depthwise 3x3 s1 + bias + ReLU from E [MC][halo] into the A-operand layout of the project MMA (hi and lo parts)
template <class G, int MC, int RH, int NT, int KB3, int EPS>
__device__ __forceinline__ void dw_stage_split(int tid, const float* __restrict__ Es, const float* __restrict__ Wd, const float* __restrict__ bd,
                                               float* __restrict__ DAhi, float* __restrict__ DAlo) {
    static_assert(G::TH % RH == 0 && G::S == 1 && G::KS == 3, "3x3 stride 1");
    constexpr int NSTRIP = G::TW / 4, NSEG = G::TH / RH;
    constexpr int NITEM = MC * NSEG * NSTRIP;
    // Lanes of one 128-bit shared-memory phase (8 lanes) must hit 8 different 16-byte bank groups. With >= 8 strips per row the
    // strips of one channel do; with 4 strips (16-pixel rows) the second half of the phase takes channel m + 2, whose operand
    // rows are XOR-swizzled onto the other two 32-byte chunks (a_idx) and whose E rows start 16 banks further on.
    constexpr bool PAIR = (NSTRIP == 4) && (MC % 4 == 0) && ((2 * EPS) % 32 == 16);
    for (int item = tid; item < NITEM; item += NT) {
        int m, seg, g;
        if (PAIR) {
            g = item & 3;
            const int k2 = (item >> 2) & 1;
            const int rest = item >> 3;
            seg = rest % NSEG;
            const int ab = rest / NSEG;                 // (m >> 2) * 2 + (m & 1)
            m = (ab >> 1) * 4 + k2 * 2 + (ab & 1);
        } else {
            m = item / (NSEG * NSTRIP);
            const int rem = item - m * (NSEG * NSTRIP);
            seg = rem / NSTRIP;
            g = rem - seg * NSTRIP;
        }
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = Wd[m * 9 + t];
        const float b = bd[m];
        const float* e = Es + m * EPS + (seg * RH) * G::IWS;
        float win[3][6], nxt[6];
        load_window<G>(win[1], e, g);
        load_window<G>(win[2], e + G::IWS, g);
        load_window<G>(nxt, e + 2 * G::IWS, g);
#pragma unroll
        for (int oy = 0; oy < RH; ++oy) {
#pragma unroll
            for (int v = 0; v < 6; ++v) { win[0][v] = win[1][v]; win[1][v] = win[2][v]; win[2][v] = nxt[v]; }
            if (oy + 1 < RH) load_window<G>(nxt, e + (oy + 3) * G::IWS, g);      // next row's window is in flight during this row's FMAs
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
__global__ void __launch_bounds__(C::NT, C::OCC)
irbtc_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ wts, int H, int W,
             int tiles_x, int tiles_y, int total_tiles) {
    using G = typename C::G;
    constexpr int NTW = C::NTW, NWW = C::NWW, NE = C::NE;
    extern __shared__ unsigned char smem_raw[];
    // align the dynamic window to 1 KB by OFFSET, so the pointers keep their shared-memory provenance (LDS/STS, not generic LD/ST)
    float* base = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float* XAhi = base;
    float* XAlo = XAhi + C::XA1;
    float* DAhi = XAlo + C::XA1;
    float* DAlo = DAhi + C::DA1;
    float* Es = DAlo + C::DA1;
    float* Ws = Es + C::ES;                      // ring of chunk weight blocks
    float* W1s = Ws + C::NWB * C::CB;            // resident expand weights (E1ALL)
    __shared__ __align__(8) uint64_t wbar[C::NWB], w1bar, xfull, e1full, e1free, dfull, dfree;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < C::NWB; ++i) mbar_init(&wbar[i], 1);
        mbar_init(&w1bar, 1);
        mbar_init(&xfull, NWW); mbar_init(&e1full, 1); mbar_init(&e1free, NWW); mbar_init(&dfull, NWW); mbar_init(&dfree, 1);
        mbar_fence_init();
    }
    if (warp == NWW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int ntile = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int S = ntile * C::NCHUNK;                     // chunk steps of this CTA

    if (warp == NWW) {
        // ================= tensor-core warp: one thread issues every MMA and every weight copy =================
        if (S > 0 && elect_one()) {
            constexpr uint32_t IDESC1 = umma_idesc_tf32(C::N1), IDESC3A = umma_idesc_tf32(2 * C::COUTP), IDESC3B = umma_idesc_tf32(C::COUTP);
            const uint32_t xh = smem_u32(XAhi), xl = smem_u32(XAlo), dh = smem_u32(DAhi), dl = smem_u32(DAlo), ws = smem_u32(Ws);
            auto issue_w = [&](int s) {
                const int buf = s % C::NWB;
                mbar_expect_tx(&wbar[buf], C::CB * 4);
                bulk_load(Ws + buf * C::CB, wts + C::OFF_CH + (size_t)(s % C::NCHUNK) * C::CB, C::CB * 4, &wbar[buf]);
            };
            // descriptor bases; an MMA adds its byte offset >> 4 to the 14-bit start-address field (never carries: smem < 256 KB)
            const uint64_t dxh = umma_desc(xh, 1024, 512, 1), dxl = umma_desc(xl, 1024, 512, 1);
            const uint64_t ddh = umma_desc(dh, 1024, 512, 1), ddl = umma_desc(dl, 1024, 512, 1);
            const uint64_t dw1 = umma_desc(C::E1ALL ? smem_u32(W1s) : ws, 128, (C::CIN / 4) * 128, 0);
            const uint64_t dw2 = umma_desc(ws + C::OFF_W2 * 4, 128, (C::MC / 4) * 128, 0);
            auto mma1 = [&](int s) {                     // expand MMA of the group of NE chunks starting at step s -> TMEM E
                const uint64_t wb = dw1 + (C::E1ALL ? (uint64_t)0 : (uint64_t)(((uint32_t)(s % C::NWB) * C::CB * 4) >> 4));
#pragma unroll 1
                for (int mt = 0; mt < C::MT1; ++mt) {
                    const uint64_t mo = (uint64_t)(mt * 256);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                        for (int kb = 0; kb < C::CIN / 8; ++kb)
                            umma_tf32(tmem + C::TM_E + mt * C::N1, (pass == 2 ? dxl : dxh) + mo + (uint64_t)(kb * C::KB1 * 4 / 16),
                                      wb + (uint64_t)(((pass == 1 ? C::OFF_W1L : C::OFF_W1H) * 4 + kb * 256) / 16), IDESC1, (pass | kb) ? 1u : 0u);
                    }
                }
                umma_commit(&e1full);
            };
            auto mma2 = [&](int s) {                     // project MMA of step s, accumulating over the chunks of a tile
                const uint64_t wb = dw2 + (uint64_t)(((uint32_t)(s % C::NWB) * C::CB * 4) >> 4);
                const uint32_t first = (s % C::NCHUNK) == 0 ? 0u : 1u;
#pragma unroll 1
                for (int mt = 0; mt < C::MT3; ++mt) {
                    const uint64_t mo = (uint64_t)(mt * 256);
#pragma unroll
                    for (int kb = 0; kb < C::MC / 8; ++kb)
                        umma_tf32(tmem + C::TM_O + mt * 2 * C::COUTP, ddh + mo + (uint64_t)(kb * C::KB3 * 4 / 16), wb + (uint64_t)(kb * 16), IDESC3A, kb ? 1u : first);
#pragma unroll
                    for (int kb = 0; kb < C::MC / 8; ++kb)
                        // D_lo . W2hi joins the other correction term D_hi . W2lo in the second column group: the tensor core accumulates
                        // with truncation, so the main term D_hi . W2hi gets the shortest possible accumulation chain
                        umma_tf32(tmem + C::TM_O + mt * 2 * C::COUTP + C::COUTP, ddl + mo + (uint64_t)(kb * C::KB3 * 4 / 16), wb + (uint64_t)(kb * 16), IDESC3B, 1u);
                }
                umma_commit(&dfree);
            };
            if (C::E1ALL) {
                mbar_expect_tx(&w1bar, C::W1RES * 4);
                bulk_load(W1s, wts, C::W1RES * 4, &w1bar);
            }
            issue_w(0);
            if (S > 1) issue_w(1);
            mbar_wait(&xfull, 0);
            if (C::E1ALL) mbar_wait(&w1bar, 0); else mbar_wait(&wbar[0], 0);
            tc_fence_after();
            mma1(0);
            for (int s = 0; s < S; ++s) {
                if (s + 1 < S && (s + 1) % NE == 0) {
                    mbar_wait(&e1free, (s / NE) & 1);                            // TMEM E of the group ending at step s has been read out
                    if ((s + 1) % C::NCHUNK == 0) mbar_wait(&xfull, ((s + 1) / C::NCHUNK) & 1);
                    if (!C::E1ALL) mbar_wait(&wbar[(s + 1) % C::NWB], ((s + 1) / C::NWB) & 1);
                    tc_fence_after();
                    TC_TRACE(s, 8);
                    mma1(s + 1);
                    TC_TRACE(s, 9);
                }
                if (s + 2 < S) {
                    if (s >= 1) mbar_wait(&dfree, (s - 1) & 1);                  // ring slot (s+2)%3 == (s-1)%3: its project MMA is complete
                    issue_w(s + 2);
                }
                mbar_wait(&dfull, s & 1);                                        // operand D of step s is written
                mbar_wait(&wbar[s % C::NWB], (s / C::NWB) & 1);
                tc_fence_after();
                TC_TRACE(s, 10);
                mma2(s);
                TC_TRACE(s, 11);
            }
        }
    } else {
        // ================= worker warps: CUDA-core phases ==========================================================
        const int quarter = warp & 3, wq = warp >> 2;        // TMEM lane quarter this warp may access; rank among the warps sharing it
        const int nwq = (NWW - quarter + 3) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int tpi = tiles_x * tiles_y;
        const uint32_t inv_tx = (65536u + (uint32_t)tiles_x - 1u) / (uint32_t)tiles_x;      // exact for tile-in-image < 65536 / tiles_x
        auto origin = [&](int ti, int& b, int& oy0, int& ox0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            b = tile / tpi;
            const int t = tile - b * tpi;
            const int ty = (int)(((uint32_t)t * inv_tx) >> 16);
            oy0 = ty * G::TH; ox0 = (t - ty * tiles_x) * G::TW;
        };
        const size_t plane = (size_t)H * W;
        float xr[C::NIT * 8];
        uint32_t xok = 0;                                      // bit i: staging item i lies inside the image
        // input staging item = (8-channel block kh, halo pixel m): item = tid + i * NTW = kh * IPIX + m; 8 channels per item.
        // The loads are unconditional (clamped addresses) and masked when consumed, so nothing waits for them here.
        auto fetch_x = [&](int b, int oy0, int ox0) {          // raw input halo tile -> registers
            const float* xb = x + (size_t)b * C::CIN * H * W;
            xok = 0;
#pragma unroll
            for (int i = 0; i < C::NIT; ++i) {
                const int item = min(tid + i * NTW, C::NITEM_X - 1);
                const int kh = item / G::IPIX, m = item - kh * G::IPIX;
                const int r = m / G::IWS, j = m - r * G::IWS;
                const int gy = oy0 - 1 + r, gx = ox0 - 1 + j;
                const bool ok = j < G::IW && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                xok |= (ok ? 1u : 0u) << i;
                const float* px = xb + (size_t)(kh * 8) * plane + min(max(gy, 0), H - 1) * W + min(max(gx, 0), W - 1);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) xr[i * 8 + kk] = __ldg(px + kk * plane);
            }
        };
        auto put_x = [&]() {                                   // registers -> split operand A of the expand MMA (zero outside the image)
#pragma unroll
            for (int i = 0; i < C::NIT; ++i) {
                const int item = tid + i * NTW;
                if (item < C::NITEM_X) {
                    const int kh = item / G::IPIX, m = item - kh * G::IPIX;
                    const int ob = kh * C::KB1 + (m >> 5) * 256 + (m & 7), mc = (m & 31) >> 3;
                    const bool ok = (xok >> i) & 1u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const float v = ok ? xr[i * 8 + kk] : 0.f;
                        const float hi = tf32_hi(v);
                        const int o = ob + ((kk >> 2) & 1) * 128 + (kk & 3) * 32 + ((mc ^ (kk & 3)) << 3);
                        XAhi[o] = hi;
                        XAlo[o] = v - hi;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&xfull);
        };
        // Output epilogue, unit = (128-pixel MMA tile mt, 8 output channels): TMEM O (hi and lo column groups) + bias (+ residual,
        // yolo_fastest.py:65) -> HBM. A warp owns units wq, wq + nwq, ...; their residual values are fetched one phase early
        // (res_prefetch, unconditional clamped loads) so the epilogue itself never waits on HBM/L2.
        constexpr int NU3 = C::MT3 * (C::COUTP / 8);
        constexpr int QMAX = (NU3 + NWW / 4 - 1) / (NWW / 4);
        float rres[C::RES ? QMAX * 8 : 1];
        size_t roff[QMAX];                                     // element offset of (image, first channel of the unit, pixel) in y / x
        uint32_t rok = 0;                                      // bit q: this lane's pixel of unit q exists
        auto res_prefetch = [&](int tb, int oy0, int ox0) {
            rok = 0;
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                const int u = min(wq + q * nwq, NU3 - 1);
                const int mt = u / (C::COUTP / 8), c0 = min((u - mt * (C::COUTP / 8)) * 8, C::COUT - 8);
                const int pix = mt * 128 + quarter * 32 + lane;
                const int oy = pix / G::TW, ox = pix - oy * G::TW;
                const int gy = oy0 + oy, gx = ox0 + ox;
                rok |= ((pix < G::OPIX && gy < H && gx < W) ? 1u : 0u) << q;
                roff[q] = ((size_t)(tb * C::COUT + c0) * H + min(gy, H - 1)) * W + min(gx, W - 1);
                if (C::RES) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rres[q * 8 + i] = __ldg(x + roff[q] + i * plane);
                }
            }
        };
        auto epilogue_out = [&]() {
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                const int u = wq + q * nwq;
                const int mt = u / (C::COUTP / 8), c0 = (u - mt * (C::COUTP / 8)) * 8;
                if (u < NU3 && c0 < C::COUT) {                 // warp-uniform
                    float vh[8], vl[8];
                    tmem_ld8(tmem + lane_base + C::TM_O + mt * 2 * C::COUTP + c0, vh);
                    tmem_ld8(tmem + lane_base + C::TM_O + mt * 2 * C::COUTP + C::COUTP + c0, vl);
                    if ((rok >> q) & 1u) {
                        float* yp = y + roff[q];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float r = (vh[i] + vl[i]) + __ldg(wts + C::OFF_B2 + c0 + i);
                            if (C::RES) r += rres[q * 8 + i];
                            yp[i * plane] = r;
                        }
                    }
                }
            }
        };

        int tb = 0, oy0 = 0, ox0 = 0, pb = 0, poy0 = 0, pox0 = 0, nb = 0, noy0 = 0, nox0 = 0;     // this / previous / next tile
        if (ntile > 0) {
            origin(0, nb, noy0, nox0);
            fetch_x(nb, noy0, nox0);
            put_x();
        }
        int s = 0;
        for (int ti = 0; ti < ntile; ++ti) {
            if (tid == 0) TC_TRACE(s, 12);
            pb = tb; poy0 = oy0; pox0 = ox0;
            tb = nb; oy0 = noy0; ox0 = nox0;
            const int iy0 = oy0 - 1, ix0 = ox0 - 1;
            const bool have_next = ti + 1 < ntile;
            if (have_next) origin(ti + 1, nb, noy0, nox0);
            if (tid == 0) TC_TRACE(s, 13);
#pragma unroll 1                 // one copy of the chunk body: the kernel has to stay inside the instruction cache
            for (int c = 0; c < C::NCHUNK; ++c, ++s) {
                const float* Wc = Ws + (s % C::NWB) * C::CB;
                // the next tile's input is staged one step before the last when one expand MMA covers the whole tile (the operand X
                // is free as soon as that MMA is complete), else during the last chunk (X is read by every chunk's expand MMA)
                constexpr int CSTAGE = (C::E1ALL && C::NCHUNK >= 2) ? C::NCHUNK - 2 : C::NCHUNK - 1;
                const bool stage_next = (c == CSTAGE) && have_next;
                if (tid == 0) TC_TRACE(s, 0);
                if (stage_next) fetch_x(nb, noy0, nox0);       // lands during the TMEM read-out below
                if (c == 0 && ti > 0) res_prefetch(pb, poy0, pox0);
                if (tid == 0) TC_TRACE(s, 14);
                mbar_wait(&wbar[s % C::NWB], (s / C::NWB) & 1);
                if (tid == 0) TC_TRACE(s, 15);
                if (s % NE == 0) mbar_wait(&e1full, (s / NE) & 1);
                tc_fence_after();
                if (tid == 0) TC_TRACE(s, 1);
                // ---- TMEM -> bias + ReLU + zero outside the image -> E[ch][halo pixel] ----------------------------
                for (int u = wq; u < C::MT1 * (C::MC / 16); u += nwq) {
                    const int mt = u / (C::MC / 16), c0 = (u - mt * (C::MC / 16)) * 16;
                    if (mt * 128 + quarter * 32 >= G::IPIX) break;      // warp-uniform: this warp's 32 pixels do not exist
                    const int pix = mt * 128 + quarter * 32 + lane;
                    const int r = pix / G::IWS, j = pix - r * G::IWS;
                    const bool ok = j < G::IW && (unsigned)(iy0 + r) < (unsigned)H && (unsigned)(ix0 + j) < (unsigned)W;
                    float bb[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b4 = ld4(Wc + C::OFF_B1 + c0 + 4 * i);
                        bb[4 * i] = b4.x; bb[4 * i + 1] = b4.y; bb[4 * i + 2] = b4.z; bb[4 * i + 3] = b4.w;
                    }
                    float v[16];
                    tmem_ld16(tmem + lane_base + C::TM_E + mt * C::N1 + (s % NE) * C::MC + c0, v);
                    float* ep = Es + c0 * C::EPS + pix;
                    const float keep = ok ? 0.f : -INFINITY;          // relu(v + b - inf) = 0 outside the image
                    if (pix < G::IPIX) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) ep[i * C::EPS] = fmaxf(v[i] + (bb[i] + keep), 0.f);
                    }
                }
                if (s % NE == NE - 1) {                        // last chunk of the expand group: TMEM E may be overwritten
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e1free);
                }
                if (tid == 0) TC_TRACE(s, 2);
                if (s > 0) mbar_wait(&dfree, (s - 1) & 1);     // project MMA of step s-1 complete: D is free, O of a finished tile is final
                if (c == 0 && ti > 0) {
                    tc_fence_after();
                    epilogue_out();
                    tc_fence_before();
                }
                if (tid == 0) TC_TRACE(s, 3);
                named_bar_sync<1, NTW>();                      // E complete
                if (tid == 0) TC_TRACE(s, 4);
                if (stage_next) put_x();                       // every expand MMA of this tile is complete (e1full of its last group)
                if (tid == 0) TC_TRACE(s, 5);
                // ---- depthwise on CUDA cores, output split into the project MMA's operand --------------------------
                dw_stage_split<G, C::MC, C::RH, NTW, C::KB3, C::EPS>(tid, Es, Wc + C::OFF_WD, Wc + C::OFF_BD, DAhi, DAlo);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dfull);
                if (tid == 0) TC_TRACE(s, 6);
                named_bar_sync<2, NTW>();                      // E free
                if (tid == 0) TC_TRACE(s, 7);
            }
        }
        if (ntile > 0) {
            res_prefetch(tb, oy0, ox0);
            mbar_wait(&dfree, (S - 1) & 1);
            tc_fence_after();
            epilogue_out();
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == NWW) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TCOLS) : "memory");
    }
}

}  // namespace yf
