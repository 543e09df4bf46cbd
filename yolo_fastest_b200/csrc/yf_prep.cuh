// Pre-processing of Detect_YOLO.__pre_process (reference src/detect.py:107-129) on the GPU: BGR uint8 frames as cv2.imread returns them
// -> cv2.cvtColor(BGR2GRAY) -> cv2.resize(INTER_LINEAR) to the network input size -> uint8 gray planes, which the stem kernel then
// normalises ((x - 128) / 255, detect.py:124) while it loads them. Everything here is the integer arithmetic OpenCV 4.x uses for
// 8-bit images (modules/imgproc color_rgb: 15-bit gray coefficients; resize.cpp: 11-bit interpolation coefficients, HResizeLinear /
// VResizeLinear fixed-point casts), so the result is bit-identical to the host path: oracle/preprocess_oracle.py restates it and is
// pinned against cv2 itself.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yf {

// one entry per output column / row: the two source indices and the two 11-bit coefficients
struct PrepTap { int s0, s1, c0, c1; };

// resize.cpp (cv::resize, INTER_LINEAR, 8-bit): fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx;
// columns clamp the index AND drop the fraction at both borders, rows keep the fraction and clip the two source rows
inline void prep_taps(int dst_n, int src_n, bool is_row, PrepTap* out) {
    const double inv_scale = (double)dst_n / src_n, scale = 1.0 / inv_scale;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        int s0, s1;
        if (is_row) {
            s0 = s < 0 ? 0 : (s > src_n - 1 ? src_n - 1 : s);
            s1 = s + 1 < 0 ? 0 : (s + 1 > src_n - 1 ? src_n - 1 : s + 1);
        } else {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
            s0 = s; s1 = s + 1 > src_n - 1 ? src_n - 1 : s + 1;
        }
        const int c0 = (int)lrintf((1.f - f) * 2048.f), c1 = (int)lrintf(f * 2048.f);           // cvRound: half to even
        if (c1 == 0) s1 = s0;                           // the second tap does not contribute: never fetch it
        out[d] = {s0, s1, c0, c1};
    }
}

__device__ __forceinline__ int prep_gray(const uint8_t* __restrict__ p) {      // BGR -> Y, 15-bit coefficients (B 3735, G 19235, R 9798)
    return (__ldg(p) * 3735 + __ldg(p + 1) * 19235 + __ldg(p + 2) * 9798 + (1 << 14)) >> 15;
}

// thread = 4 consecutive output pixels of one row (one 32-bit store); grid-stride over [B][H][W / 4]
__global__ void __launch_bounds__(256)
prep_bgr_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, const PrepTap* __restrict__ xt, const PrepTap* __restrict__ yt,
                int B, int Ho, int Wo, int H, int W) {
    const int W4 = W >> 2;
    const long long total = (long long)B * H * W4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % W4);
        const long long rest = i / W4;
        const int y = (int)(rest % H), b = (int)(rest / H);
        const PrepTap ty = yt[y];
        const uint8_t* r0 = bgr + ((size_t)b * Ho + ty.s0) * (size_t)Wo * 3;
        const uint8_t* r1 = bgr + ((size_t)b * Ho + ty.s1) * (size_t)Wo * 3;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const PrepTap tx = xt[q * 4 + k];
            const int h0 = prep_gray(r0 + tx.s0 * 3) * tx.c0 + prep_gray(r0 + tx.s1 * 3) * tx.c1;      // HResizeLinear: 11-bit scaled
            const int h1 = prep_gray(r1 + tx.s0 * 3) * tx.c0 + prep_gray(r1 + tx.s1 * 3) * tx.c1;
            const int v = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;        // VResizeLinear fixed-point cast
            packed |= (uint32_t)(v & 255) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(gray + ((size_t)b * H + y) * W + q * 4) = packed;
    }
}

// Row-staged form (frame width a multiple of 4, 4-byte aligned frames): a block owns PREP_R output rows at a time. Their source rows
// are read ONCE, as coalesced 32-bit words with every load of the group in flight together, converted to gray and kept in shared
// memory; the threads then interpolate 4 consecutive output pixels each from those bytes. Every frame byte is fetched from HBM exactly
// once for identity and integer down-scales. blockDim.x = k * (Wo / 4) >= W / 4; dynamic shared memory = 2 * PREP_R * Wo bytes.
constexpr int PREP_R = 8;
__device__ __forceinline__ uint32_t prep_gray4(uint32_t w0, uint32_t w1, uint32_t w2) {      // 4 BGR pixels (12 bytes) -> 4 gray bytes
    auto byte = [](uint32_t w, int k) { return (int)__byte_perm(w, 0, 0x4440 + k); };
    auto y = [](int b, int g, int r) { return (uint32_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15); };
    return y(byte(w0, 0), byte(w0, 1), byte(w0, 2)) | (y(byte(w0, 3), byte(w1, 0), byte(w1, 1)) << 8) |
           (y(byte(w1, 2), byte(w1, 3), byte(w2, 0)) << 16) | (y(byte(w2, 1), byte(w2, 2), byte(w2, 3)) << 24);
}

__global__ void __launch_bounds__(1024)
prep_bgr_rows_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, const PrepTap* __restrict__ xt, const PrepTap* __restrict__ yt,
                     int B, int Ho, int Wo, int H, int W) {
    extern __shared__ uint32_t prep_sm[];               // [2 * PREP_R rows][Wo / 4] words of gray bytes
    __shared__ int srow[2 * PREP_R];                    // source row staged in each slot, -1 = not needed
    __shared__ PrepTap tys[PREP_R];
    const int W4 = W >> 2, Wo4 = Wo >> 2, tid = (int)threadIdx.x;
    const int sr = tid / Wo4, sg = tid - sr * Wo4, sk = (int)blockDim.x / Wo4;          // staging role: rows sr, sr + sk, ..; word group sg
    const int orow = tid / W4, oq = tid - orow * W4, ok = (int)blockDim.x / W4;         // output role: rows orow, orow + ok, ..; pixels 4 oq ..
    PrepTap tx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tx[k] = xt[min(oq, W4 - 1) * 4 + k];
    const int groups = (H + PREP_R - 1) / PREP_R;       // row groups per image
    for (long long it = blockIdx.x; it < (long long)B * groups; it += gridDim.x) {
        const int b = (int)(it / groups), y0 = (int)(it - (long long)b * groups) * PREP_R;
        if (tid < PREP_R) {
            const int yy = y0 + tid;
            PrepTap t = yt[min(yy, H - 1)];
            tys[tid] = t;
            srow[2 * tid] = yy < H ? t.s0 : -1;
            srow[2 * tid + 1] = (yy < H && t.s1 != t.s0) ? t.s1 : -1;
        }
        __syncthreads();
        if (sr < sk) {
            const uint8_t* img = bgr + (size_t)b * Ho * (size_t)Wo * 3;
#pragma unroll 4
            for (int ri = sr; ri < 2 * PREP_R; ri += sk) {
                const int r = srow[ri];
                if (r >= 0) {
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(img + (size_t)r * Wo * 3) + 3 * sg;
                    prep_sm[ri * Wo4 + sg] = prep_gray4(__ldg(src), __ldg(src + 1), __ldg(src + 2));
                }
            }
        }
        __syncthreads();
        if (orow < ok) {
            for (int ro = orow; ro < PREP_R && y0 + ro < H; ro += ok) {
                const PrepTap ty = tys[ro];
                const uint8_t* g0 = reinterpret_cast<const uint8_t*>(prep_sm + (2 * ro) * Wo4);
                const uint8_t* g1 = ty.s1 == ty.s0 ? g0 : reinterpret_cast<const uint8_t*>(prep_sm + (2 * ro + 1) * Wo4);
                uint32_t packed = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int h0 = g0[tx[k].s0] * tx[k].c0 + g0[tx[k].s1] * tx[k].c1;
                    const int h1 = g1[tx[k].s0] * tx[k].c0 + g1[tx[k].s1] * tx[k].c1;
                    const int v = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
                    packed |= (uint32_t)(v & 255) << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(gray + ((size_t)b * H + y0 + ro) * W + oq * 4) = packed;
            }
        }
        __syncthreads();
    }
}

}  // namespace yf
