// yf_tma.cuh — tensor-tile TMA (cp.async.bulk.tensor, SASS UTMALDG) for halo tiles of planar [B][C][H][W] activations.
//
// A halo rectangle of a W-innermost tensor starts 1-2 elements left of / above the tile it serves.  Measured on B200 (sm_100a, driver
// 580, tools/selftest/tma_selftest.cu, output in profiles/r02_tma_selftest.txt): tiled-mode TMA accepts NEGATIVE start coordinates in
// every dimension and boxes hanging over any edge (elements outside the tensor arrive as zeros, the mbarrier still receives the full box
// byte count — exactly the zero padding of the convolutions, yolo_fastest.py:17-19), and any start coordinate in the outer
// dimensions, but the INNERMOST start coordinate times the element size must be a multiple of 16 bytes: x = 0, 4, -4, 40, 60 load
// correctly, x = 1, -1, 2, 3 raise "illegal instruction" (fp32; the rule is on bytes).  Halo boxes therefore start at the aligned
// column S*x0 - 4 (fp32) / - 16 (uint8) instead of S*x0 - 1 and are 3 (15) columns wider on the left; the consumers index from 3.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace yf {

// ---- host: encode a rank-4 tiled tensor map over [B][C][H][W] (element size esz: 4 = fp32, 1 = uint8) ---------------------------
typedef CUresult (*yf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline yf_encode_tiled_fn tma_encoder() {
    static yf_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (yf_encode_tiled_fn)p;
    }
    return fn;
}

// box = {bw, bh, bc, 1} elements; bw * esz must be a multiple of 16, W * esz a multiple of 16, base 16-byte aligned.
// Returns 0 on success, the CUresult (or -1 when the driver entry point is missing) otherwise.
inline int tma_make_map4(CUtensorMap* map, const void* base, int esz, int B, int C, int H, int W, int bw, int bh, int bc) {
    yf_encode_tiled_fn enc = tma_encoder();
    if (!enc) return -1;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * esz, (cuuint64_t)W * H * esz, (cuuint64_t)W * H * C * esz};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

// ---- device ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// box at (x, y, c, b) -> dst (128-byte aligned shared memory, dense [bc][bh][bw]); completion (full box bytes) on bar
__device__ __forceinline__ void tma_load4(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int c, int b) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(tma_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tma_smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(b)
                 : "memory");
}

}  // namespace yf
