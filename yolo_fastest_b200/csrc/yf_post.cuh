// yf_post.cuh — the head kernel: decode of both YOLO scales, confidence filter, warp-ballot
// ordered compaction, stable per-class sort and greedy per-class NMS, one CTA (512 threads) per image.
//
//   sort   the survivors' keys (class asc, conf desc, candidate index asc — Python's stable list.sort per class, detect.py:157-167)
//          live in SHARED memory as (float64 conf, class << 21 | slot) pairs and are ordered by one bitonic network over the next
//          power of two >= n: O(n log^2 n) compare-exchanges on chip instead of an O(n^2) rank over global memory.
//   NMS    bitmask formulation: a chunk of 32 sorted boxes is a 32x32 IoU tile — every lane evaluates its box against the 32 boxes of
//          the chunk (all IoUs in parallel, off the serial chain) and packs the results into one suppression word; the greedy
//          resolution then walks the words (one shuffle + two logic ops per box). The chunk's kept boxes knock out the later boxes
//          of the class. Classes of up to 64 boxes are resolved by one warp each, all warps in parallel (the 80-class stress
//          configuration: 80 classes x ~32 boxes); larger classes are processed by the WHOLE CTA chunk by chunk — warp 0 resolves
//          the tile, all 16 warps apply its kept boxes to the rest of the class.
//
//   YF_MODE_DETECT   src/detect.py:23-84,155-169   float64 sigmoid/exp on the fp32 logits,
//                    Python round() (half-to-even) to integer boxes, strict '>' tests, IoU
//                    without +1 as exact integer areas / float64 ratio.
//   YF_MODE_VALIDATE src/model_training/loss/yolo_loss.py:98-141 + utils/general.py:29-52,87-143
//                    fp32 everywhere, no rounding, '>=' conf test, IoU with +1 and +1e-16,
//                    a box survives iff IoU < thres.
//
// Output order is the reference's: classes ascending; inside a class conf descending with ties in
// candidate order (Python's stable list.sort, detect.py:167); candidates are visited head_large
// first, then anchor, row, column (detect.py:43,54-56).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/yf.h"

namespace yf {

constexpr int POST_NT = 512;
constexpr int POST_NW = POST_NT / 32;
constexpr int POST_MAX_CLS = 2048;

constexpr int POST_SRC_ROWS = 2;   // internal third mode: validate flavour fed decoded rows [B][N][5+nc] (general.py:87)

struct PostArgs {
    const float* head[2];
    const float* pred;    // POST_SRC_ROWS only
    int A, nc, h[2], w[2];
    double anchors[2][YF_MAX_ANCHORS][2];
    double conf_thres, nms_thres;
    int input_h, input_w, mode, max_det, do_nms;
    yf_det* out;
    int32_t* counts;
    int32_t* status;
    // per-image workspaces, each [B][NC]
    int NC;
    yf_det* rec;          // survivors in candidate order
    double* conf;         // sort key (float64 sigmoid, or the fp32 sigmoid widened)
    int32_t* cls;
    int4* sbox;           // boxes in sorted order: int32 coords (detect) or fp32 bit patterns (validate)
    int32_t* order;       // sorted position -> survivor slot
    unsigned char* alive; // per sorted position
    int sort_cap;         // elements the dynamic shared memory holds for the sort (a power of two; 0: fall back to the global rank)
};

// Ordered block compaction: every thread calls with its predicate; returns the slot of this
// thread's element among all true predicates so far (candidate order), updates `total`.
__device__ __forceinline__ int block_ordered_slot(bool pred, int* s_wc, int& total) {
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_wc[wid] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < POST_NW; ++w) {
        const int c = s_wc[w];
        if (w < wid) off += c;
        tot += c;
    }
    __syncthreads();
    const int slot = total + off + wpre;
    total += tot;
    return slot;
}

__device__ __forceinline__ double sigmoid_f64(float x) { return 1.0 / (1.0 + exp(-(double)x)); }   // detect.py:23-25
__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }  // torch.sigmoid, fp32

// detect.py:27-39 on integer boxes; true iff IoU > thres (NaN for an empty union never suppresses,
// as in YOLO_ncnn.cpp:212,221-234 — the Python reference raises ZeroDivisionError there).
__device__ __forceinline__ bool suppress_i32(const int4 a, const int4 b, double thres) {
    const long long iw = (long long)min(a.z, b.z) - (long long)max(a.x, b.x);
    const long long ih = (long long)min(a.w, b.w) - (long long)max(a.y, b.y);
    long long inter = 0;
    if (iw > 0 && ih > 0) inter = iw * ih;
    const long long uni = ((long long)a.z - a.x) * ((long long)a.w - a.y) + ((long long)b.z - b.x) * ((long long)b.w - b.y) - inter;
    return ((double)inter / (double)uni) > thres;
}

// general.py:29-52 (x1y1x2y2 branch), fp32 with the reference's operation order and no FMA
// contraction; box1 = the kept box, box2 = the candidate. Returns true iff NOT (iou < thres).
__device__ __forceinline__ bool suppress_f32(const int4 kept, const int4 cand, float thres) {
    const float k_x1 = __int_as_float(kept.x), k_y1 = __int_as_float(kept.y), k_x2 = __int_as_float(kept.z), k_y2 = __int_as_float(kept.w);
    const float c_x1 = __int_as_float(cand.x), c_y1 = __int_as_float(cand.y), c_x2 = __int_as_float(cand.z), c_y2 = __int_as_float(cand.w);
    const float ix1 = fmaxf(k_x1, c_x1), iy1 = fmaxf(k_y1, c_y1);
    const float ix2 = fminf(k_x2, c_x2), iy2 = fminf(k_y2, c_y2);
    const float iw = fmaxf(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), 0.0f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(iy2, iy1), 1.0f), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float a1 = __fmul_rn(__fadd_rn(__fsub_rn(k_x2, k_x1), 1.0f), __fadd_rn(__fsub_rn(k_y2, k_y1), 1.0f));
    const float a2 = __fmul_rn(__fadd_rn(__fsub_rn(c_x2, c_x1), 1.0f), __fadd_rn(__fsub_rn(c_y2, c_y1), 1.0f));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-16f);
    const float iou = __fdiv_rn(inter, den);
    return !(iou < thres);
}

template <int MODE>
__device__ __forceinline__ bool suppress(const int4 kept, const int4 cand, double thres_d, float thres_f) {
    return MODE == YF_MODE_DETECT ? suppress_i32(cand, kept, thres_d) : suppress_f32(kept, cand, thres_f);
}

__device__ __forceinline__ int4 shfl4(const int4 v, int src) {
    int4 r;
    r.x = __shfl_sync(0xffffffffu, v.x, src);
    r.y = __shfl_sync(0xffffffffu, v.y, src);
    r.z = __shfl_sync(0xffffffffu, v.z, src);
    r.w = __shfl_sync(0xffffffffu, v.w, src);
    return r;
}

// ---- bitmask NMS ---------------------------------------------------------------------------------------------------------------
// One 32x32 tile: lane j holds box j of the chunk (`mine`, `in`/`al` = exists / still alive). Returns the chunk's kept mask and
// updates `al`. Step 1: lane j evaluates suppress(i, j) for every earlier box i of the chunk -> word w_j (bit i). Step 2: greedy
// walk over the words: box j is kept iff it is alive and no kept earlier box has its bit set in w_j.
template <int MODE>
__device__ __forceinline__ unsigned nms_tile(const int4 mine, bool& al, int cnt, double thres_d, float thres_f) {
    const int lane = threadIdx.x & 31;
    unsigned w = 0;
    for (int i = 0; i < cnt; ++i) {
        const int4 bi = shfl4(mine, i);
        if (i < lane && suppress<MODE>(bi, mine, thres_d, thres_f)) w |= 1u << i;
    }
    const unsigned alive_in = __ballot_sync(0xffffffffu, al);
    unsigned kept = 0;
    for (int j = 0; j < cnt; ++j) {
        const unsigned wj = __shfl_sync(0xffffffffu, w, j);
        if (((alive_in >> j) & 1u) && (wj & kept) == 0u) kept |= 1u << j;
    }
    al = (kept >> lane) & 1u;
    return kept;
}

// Greedy NMS of one conf-descending segment [st, en) by one warp (detect.py:69-84 / general.py:127-136): tiles of 32 in order;
// the kept boxes of a tile knock out every later box of the segment. alive[] is read and written by the same lane only.
template <int MODE>
__device__ __forceinline__ void warp_nms_segment(const int4* __restrict__ sbox, unsigned char* __restrict__ alive,
                                                 int st, int en, double thres_d, float thres_f) {
    const int lane = threadIdx.x & 31;
    for (int s = st; s < en; s += 32) {
        const int p = s + lane;
        const bool in = p < en;
        const int4 mine = in ? sbox[p] : make_int4(0, 0, 0, 0);
        bool al = in && alive[p];
        const unsigned kept = nms_tile<MODE>(mine, al, min(32, en - s), thres_d, thres_f);
        if (in) alive[p] = al ? 1 : 0;
        if (kept == 0u) continue;
        for (int qb = s + 32; qb < en; qb += 32) {
            const int q = qb + lane;
            const bool inq = q < en;
            const int4 bq = inq ? sbox[q] : make_int4(0, 0, 0, 0);
            bool aq = inq && alive[q];
            for (unsigned mask = kept; mask; mask &= mask - 1u) {
                const int i = __ffs(mask) - 1;
                const int4 bi = shfl4(mine, i);
                if (aq && suppress<MODE>(bi, bq, thres_d, thres_f)) aq = false;
            }
            if (inq) alive[q] = aq ? 1 : 0;
        }
    }
}

// The same for a LARGE segment by the whole CTA (every thread calls it with the same arguments): per tile warp 0 resolves the
// 32x32 tile and publishes its kept boxes in shared memory, then all warps apply them to the later boxes of the segment.
template <int MODE>
__device__ __forceinline__ void block_nms_segment(const int4* __restrict__ sbox, unsigned char* __restrict__ alive,
                                                  int st, int en, double thres_d, float thres_f, int4* s_kbox, int* s_nk) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int s = st; s < en; s += 32) {
        if (wid == 0) {
            const int p = s + lane;
            const bool in = p < en;
            const int4 mine = in ? sbox[p] : make_int4(0, 0, 0, 0);
            bool al = in && alive[p];
            const unsigned kept = nms_tile<MODE>(mine, al, min(32, en - s), thres_d, thres_f);
            if (in) alive[p] = al ? 1 : 0;
            if (al) s_kbox[__popc(kept & ((1u << lane) - 1u))] = mine;
            if (lane == 0) *s_nk = __popc(kept);
        }
        __syncthreads();
        const int nk = *s_nk;
        if (nk > 0) {
            for (int q = s + 32 + (int)threadIdx.x; q < en; q += POST_NT) {
                if (!alive[q]) continue;
                const int4 bq = sbox[q];
                bool aq = true;
                for (int i = 0; i < nk && aq; ++i)
                    if (suppress<MODE>(s_kbox[i], bq, thres_d, thres_f)) aq = false;
                if (!aq) alive[q] = 0;
            }
        }
        __syncthreads();
    }
}

// sort order of the survivors: class ascending, conf descending, slot (= candidate order) ascending; key2 = class << 21 | slot
__device__ __forceinline__ bool key_before(double ca, uint32_t ka, double cb, uint32_t kb) {
    const uint32_t cla = ka >> 21, clb = kb >> 21;
    if (cla != clb) return cla < clb;
    if (ca != cb) return ca > cb;
    return ka < kb;
}

template <int MODE>
__global__ void __launch_bounds__(POST_NT) post_kernel(const PostArgs a) {
    extern __shared__ __align__(16) unsigned char post_smem[];       // sort keys: double conf[sort_cap], uint32 key2[sort_cap]
    __shared__ int s_wc[POST_NW];
    __shared__ int s_start[POST_MAX_CLS + 1];
    __shared__ int s_flags;
    __shared__ int4 s_kbox[32];
    __shared__ int s_nk;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int attrs = 5 + a.nc;
    const int n_large = a.A * a.h[0] * a.w[0];
    yf_det* rec = a.rec + (size_t)b * a.NC;
    double* confs = a.conf + (size_t)b * a.NC;
    int32_t* clss = a.cls + (size_t)b * a.NC;
    int4* sbox = a.sbox + (size_t)b * a.NC;
    int32_t* order = a.order + (size_t)b * a.NC;
    unsigned char* alive = a.alive + (size_t)b * a.NC;
    if (tid == 0) s_flags = 0;
    for (int c = tid; c <= a.nc; c += POST_NT) s_start[c] = 0;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // programmatic dependent launch (yf_kernels.cuh: pdl_wait): the heads are
    asm volatile("griddepcontrol.wait;" ::: "memory");                   // the predecessor's output
    __syncthreads();

    // ---- phase 1: decode + confidence filter + ordered compaction -------------------------------
    int n = 0;
    int flags = 0;
    const float thres_f = (float)a.conf_thres;
    for (int base = 0; base < a.NC; base += POST_NT) {
        const int idx = base + tid;
        bool pass = false;
        const float* p = nullptr;
        int hw = 0, hd = 0, an = 0, i = 0, j = 0;
        double conf = 0.0;
        if (MODE == POST_SRC_ROWS) {
            if (idx < a.NC) {
                p = a.pred + ((size_t)b * a.NC + idx) * attrs;
                const float cf = __ldg(p + 4);
                if (!isfinite(cf)) flags |= 2;
                conf = (double)cf;
                pass = cf >= thres_f;                           // general.py:100
            }
        } else if (idx < a.NC) {
            hd = idx >= n_large ? 1 : 0;
            const int local = idx - (hd ? n_large : 0);
            hw = a.h[hd] * a.w[hd];
            an = local / hw;
            const int rem = local - an * hw;
            i = rem / a.w[hd];
            j = rem - i * a.w[hd];
            p = a.head[hd] + ((size_t)b * a.A + an) * attrs * hw + rem;
            const float t4 = __ldg(p + 4 * (size_t)hw);
            if (!isfinite(t4)) flags |= 2;
            if (MODE == YF_MODE_DETECT) {
                conf = sigmoid_f64(t4);
                pass = conf > a.conf_thres;                     // detect.py:58
            } else {
                const float cf = sigmoid_f32(t4);
                conf = (double)cf;
                pass = cf >= thres_f;                           // general.py:100
            }
        }
        const int slot = block_ordered_slot(pass, s_wc, n);
        if (pass && MODE == POST_SRC_ROWS) {
            yf_det d;
            const float cx = __ldg(p), cy = __ldg(p + 1), bw = __ldg(p + 2), bh = __ldg(p + 3);
            if (!(isfinite(cx) && isfinite(cy) && isfinite(bw) && isfinite(bh))) flags |= 2;
            int best = 0;
            float bestv = __ldg(p + 5);
            for (int c = 1; c < a.nc; ++c) {                     // torch.max(dim=1): first maximum (general.py:109)
                const float v = __ldg(p + 5 + c);
                if (v > bestv) { bestv = v; best = c; }
            }
            d.x1 = __fsub_rn(cx, __fdiv_rn(bw, 2.0f)); d.y1 = __fsub_rn(cy, __fdiv_rn(bh, 2.0f));   // general.py:90-94
            d.x2 = __fadd_rn(cx, __fdiv_rn(bw, 2.0f)); d.y2 = __fadd_rn(cy, __fdiv_rn(bh, 2.0f));
            d.conf = conf; d.cls_score = (double)bestv; d.cls = best; d.src = idx;
            rec[slot] = d; confs[slot] = conf; clss[slot] = best;
            if (a.do_nms) atomicAdd(&s_start[best + 1], 1);
        } else if (pass) {
            yf_det d;
            const float t0 = __ldg(p), t1 = __ldg(p + hw), t2 = __ldg(p + 2 * (size_t)hw), t3 = __ldg(p + 3 * (size_t)hw);
            if (!(isfinite(t0) && isfinite(t1) && isfinite(t2) && isfinite(t3))) flags |= 2;
            int4 bx;
            if (MODE == YF_MODE_DETECT) {
                int best = 0;
                float bestv = __ldg(p + 5 * (size_t)hw);
                for (int c = 1; c < a.nc; ++c) {                 // np.argmax: first maximum (detect.py:59)
                    const float v = __ldg(p + (size_t)(5 + c) * hw);
                    if (v > bestv) { bestv = v; best = c; }
                }
                const double scale_w = (double)a.input_w / (double)a.w[hd];
                const double scale_h = (double)a.input_h / (double)a.h[hd];
                const double x = ((double)j + sigmoid_f64(t0)) * scale_w;             // detect.py:61-64
                const double y = ((double)i + sigmoid_f64(t1)) * scale_h;
                const double bw = exp((double)t2) * a.anchors[hd][an][0];
                const double bh = exp((double)t3) * a.anchors[hd][an][1];
                d.x1 = rint(x - bw / 2); d.y1 = rint(y - bh / 2);                      // Python round(): half-to-even
                d.x2 = rint(x + bw / 2); d.y2 = rint(y + bh / 2);
                d.conf = conf;
                d.cls_score = sigmoid_f64(bestv);
                d.cls = best;
                const double lim = 33554432.0;   // 2^25: areas stay exact in fp64 / int64
                if (!(fabs(d.x1) <= lim && fabs(d.y1) <= lim && fabs(d.x2) <= lim && fabs(d.y2) <= lim)) flags |= 1;
                const double cl = 2147483647.0;
                bx = make_int4((int)fmin(fmax(d.x1, -cl), cl), (int)fmin(fmax(d.y1, -cl), cl),
                               (int)fmin(fmax(d.x2, -cl), cl), (int)fmin(fmax(d.y2, -cl), cl));
            } else {
                int best = 0;
                float bestv = sigmoid_f32(__ldg(p + 5 * (size_t)hw));
                for (int c = 1; c < a.nc; ++c) {                 // torch.max over sigmoid scores (general.py:109)
                    const float v = sigmoid_f32(__ldg(p + (size_t)(5 + c) * hw));
                    if (v > bestv) { bestv = v; best = c; }
                }
                const float stride_w = (float)((double)a.input_w / (double)a.w[hd]);
                const float stride_h = (float)((double)a.input_h / (double)a.h[hd]);
                const float aw = (float)(a.anchors[hd][an][0] / ((double)a.input_w / (double)a.w[hd]));   // yolo_loss.py:56,113
                const float ah = (float)(a.anchors[hd][an][1] / ((double)a.input_h / (double)a.h[hd]));
                const float cx = __fmul_rn(__fadd_rn(sigmoid_f32(t0), (float)j), stride_w);               // :132-139
                const float cy = __fmul_rn(__fadd_rn(sigmoid_f32(t1), (float)i), stride_h);
                const float bw = __fmul_rn(__fmul_rn(expf(t2), aw), stride_w);
                const float bh = __fmul_rn(__fmul_rn(expf(t3), ah), stride_h);
                const float x1 = __fsub_rn(cx, __fdiv_rn(bw, 2.0f)), y1 = __fsub_rn(cy, __fdiv_rn(bh, 2.0f));   // general.py:90-94
                const float x2 = __fadd_rn(cx, __fdiv_rn(bw, 2.0f)), y2 = __fadd_rn(cy, __fdiv_rn(bh, 2.0f));
                d.x1 = x1; d.y1 = y1; d.x2 = x2; d.y2 = y2;
                d.conf = conf;
                d.cls_score = (double)bestv;
                d.cls = best;
                bx = make_int4(__float_as_int(x1), __float_as_int(y1), __float_as_int(x2), __float_as_int(y2));
            }
            d.src = idx;
            rec[slot] = d;
            confs[slot] = conf;
            clss[slot] = d.cls;
            (void)bx;
            if (a.do_nms) atomicAdd(&s_start[d.cls + 1], 1);
        }
    }
    if (flags) atomicOr(&s_flags, flags);
    __syncthreads();
    if (tid == 0 && a.status) a.status[b] = s_flags;

    if (!a.do_nms) {
        // decode_box only (detect.py:41-67): survivors in candidate order
        for (int s = tid; s < n && s < a.max_det; s += POST_NT) a.out[(size_t)b * a.max_det + s] = rec[s];
        if (tid == 0) a.counts[b] = n;
        return;
    }

    // ---- phase 2: class offsets + stable sort inside the class -----------------------------------
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c <= a.nc; ++c) { run += s_start[c]; s_start[c] = run; }   // s_start[c] = first slot of class c
    }
    __syncthreads();
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    if (np2 <= a.sort_cap) {
        // bitonic network over (class, -conf, slot) keys in shared memory; padding keys sort last (class field all ones)
        double* kc = reinterpret_cast<double*>(post_smem);
        uint32_t* kk = reinterpret_cast<uint32_t*>(kc + a.sort_cap);
        for (int s = tid; s < np2; s += POST_NT) {
            kc[s] = s < n ? confs[s] : 0.0;
            kk[s] = s < n ? ((uint32_t)clss[s] << 21 | (uint32_t)s) : 0xFFFFFFFFu;
        }
        __syncthreads();
        for (int k = 2; k <= np2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < np2; i += POST_NT) {
                    const int l = i ^ j;
                    if (l > i) {
                        const double ci = kc[i], cl = kc[l];
                        const uint32_t ki = kk[i], kl = kk[l];
                        const bool up = (i & k) == 0;                    // ascending block: the earlier key goes to the lower index
                        if (key_before(cl, kl, ci, ki) == up) { kc[i] = cl; kc[l] = ci; kk[i] = kl; kk[l] = ki; }
                    }
                }
                __syncthreads();
            }
        }
        for (int pos = tid; pos < n; pos += POST_NT) order[pos] = (int)(kk[pos] & 0x1FFFFFu);
    } else {
        // more survivors than the shared-memory sort holds: rank = #{t in same class : conf_t > conf_s, or equal and earlier}
        for (int s = tid; s < n; s += POST_NT) {
            const int c = clss[s];
            const double cf = confs[s];
            int r = 0;
            for (int t = 0; t < n; ++t) {
                const double ct = confs[t];
                r += (clss[t] == c) && (ct > cf || (ct == cf && t < s));
            }
            order[s_start[c] + r] = s;
        }
    }
    __syncthreads();
    // boxes in sorted order (separate pass: sbox currently holds candidate order, read via rec instead)
    for (int pos = tid; pos < n; pos += POST_NT) {
        const yf_det d = rec[order[pos]];
        int4 bx;
        if (MODE == YF_MODE_DETECT) {
            const double cl = 2147483647.0;
            bx = make_int4((int)fmin(fmax(d.x1, -cl), cl), (int)fmin(fmax(d.y1, -cl), cl),
                           (int)fmin(fmax(d.x2, -cl), cl), (int)fmin(fmax(d.y2, -cl), cl));
        } else {
            bx = make_int4(__float_as_int((float)d.x1), __float_as_int((float)d.y1),
                           __float_as_int((float)d.x2), __float_as_int((float)d.y2));
        }
        sbox[pos] = bx;
        alive[pos] = 1;
    }
    __syncthreads();

    // ---- phase 3: per-class greedy NMS (bitmask tiles): large classes by the whole CTA, the others one warp per class ------------
    {
        const int wid = tid >> 5;
        const float nthr_f = (float)a.nms_thres;
        constexpr int BIG = 64;
        for (int c = 0; c < a.nc; ++c)                                    // s_start is shared: every thread takes the same branches
            if (s_start[c + 1] - s_start[c] > BIG)
                block_nms_segment<MODE>(sbox, alive, s_start[c], s_start[c + 1], a.nms_thres, nthr_f, s_kbox, &s_nk);
        for (int c = wid; c < a.nc; c += POST_NW)
            if (s_start[c + 1] - s_start[c] <= BIG)
                warp_nms_segment<MODE>(sbox, alive, s_start[c], s_start[c + 1], a.nms_thres, nthr_f);
    }
    __syncthreads();

    // ---- phase 4: kept boxes in order -> out ------------------------------------------------------
    int kept = 0;
    for (int base = 0; base < n; base += POST_NT) {
        const int pos = base + tid;
        const bool k = pos < n && alive[pos];
        const int slot = block_ordered_slot(k, s_wc, kept);
        if (k && slot < a.max_det) a.out[(size_t)b * a.max_det + slot] = rec[order[pos]];
    }
    if (tid == 0) a.counts[b] = kept;
}

// NMS of one already-sorted list (the bit-exactness entry): one warp.
template <int MODE>
__global__ void nms_sorted_kernel(const int4* __restrict__ boxes, unsigned char* __restrict__ alive, int n,
                                  double thres_d, float thres_f, int32_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
    const int lane = threadIdx.x;
    for (int i = lane; i < n; i += 32) alive[i] = 1;
    __syncwarp();
    warp_nms_segment<MODE>(boxes, alive, 0, n, thres_d, thres_f);
    __syncwarp();
    int total = 0;
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        const bool k = p < n && alive[p];
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (k) keep[total + __popc(bal & ((1u << lane) - 1u))] = p;
        total += __popc(bal);
    }
    if (lane == 0) *n_keep = total;
}

// YOLOLossV3.forward(input, targets=None) (yolo_loss.py:48-68,98-141): one thread per (b, anchor, row, col).
struct ValDecodeArgs {
    const float* head;
    float* out;
    int B, A, nc, h, w;
    float stride_w, stride_h;
    float aw[YF_MAX_ANCHORS], ah[YF_MAX_ANCHORS];   // anchors / stride, as fp32 (yolo_loss.py:56,113-114)
};

__global__ void val_decode_kernel(const ValDecodeArgs a) {
    const int hw = a.h * a.w;
    const long long total = (long long)a.B * a.A * hw;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int rem = (int)(gid % hw);
    const int an = (int)((gid / hw) % a.A);
    const int b = (int)(gid / ((long long)hw * a.A));
    const int i = rem / a.w, j = rem - i * a.w;
    const int attrs = 5 + a.nc;
    const float* p = a.head + ((size_t)b * a.A + an) * attrs * hw + rem;
    float* o = a.out + (size_t)gid * attrs;
    o[0] = __fmul_rn(__fadd_rn(sigmoid_f32(__ldg(p)), (float)j), a.stride_w);
    o[1] = __fmul_rn(__fadd_rn(sigmoid_f32(__ldg(p + hw)), (float)i), a.stride_h);
    o[2] = __fmul_rn(__fmul_rn(expf(__ldg(p + 2 * (size_t)hw)), a.aw[an]), a.stride_w);
    o[3] = __fmul_rn(__fmul_rn(expf(__ldg(p + 3 * (size_t)hw)), a.ah[an]), a.stride_h);
    for (int k = 4; k < attrs; ++k) o[k] = sigmoid_f32(__ldg(p + (size_t)k * hw));
}

// ---- compaction of per-image detection slabs into one contiguous record list (multi-GPU result return) -----------------------
// dets [B][max_det] + counts [B]  ->  packed = int32 header [total, B, n_0 .. n_{B-1}] (hdr_slots int32 words reserved) followed by
// the records of image 0, 1, ... back to back (n_b = min(counts[b], max_det)). Records beyond `cap` are dropped; `total` is the
// untruncated sum, so the receiver detects the overflow. One CTA: a block-wide scan over the images, then a cooperative copy.
__global__ void __launch_bounds__(256) compact_dets_kernel(const yf_det* __restrict__ dets, const int32_t* __restrict__ counts, int B, int max_det,
                                                           int32_t* __restrict__ hdr, int hdr_slots, yf_det* __restrict__ rec, int cap) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const uint2* src = reinterpret_cast<const uint2*>(dets);          // a record = 7 x 8 bytes
    uint2* dst = reinterpret_cast<uint2*>(rec);
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + tid;
        const int n = b < B ? min(max(counts[b], 0), max_det) : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_warp[w];
        const int off = s_base + wbase + incl - n;                    // first record slot of image b
        if (b < B) {
            if (2 + b < hdr_slots) hdr[2 + b] = n;
            for (int k = 0; k < n; ++k) {
                if (off + k < cap) {
#pragma unroll
                    for (int w8 = 0; w8 < 7; ++w8) dst[(size_t)(off + k) * 7 + w8] = src[((size_t)b * max_det + k) * 7 + w8];
                }
            }
        }
        __syncthreads();
        if (tid == 255) s_base = off + n;
        __syncthreads();
    }
    if (tid == 0) { hdr[0] = s_base; hdr[1] = B; }
}

}  // namespace yf
