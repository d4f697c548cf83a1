// "Packed" operand format shared by the tensor-core GEMM (umma_packed.cu) and the kernels that feed it.
//
// A matrix X[rows][K] that will be a K-major UMMA operand is stored as a grid of blocks, one per
// (row tile of ROWS rows, K chunk of 32 fp32):  block(tile, c) at byte offset (tile * KC + c) * block_bytes.
// A block is exactly the shared-memory image the tensor core reads: the TF32 "hi" half followed by the
// "lo" half (x = hi + lo, see umma_packed.cu), each in the canonical K-major no-swizzle layout of
// 8-row x 16-byte core matrices with LBO = 144 B between K-adjacent core matrices and SBO = 1152 B
// between 8-row groups (the skew keeps every store pattern used by the writers bank-conflict free).
// One cp.async.bulk per block moves it global -> shared with no per-element work in the GEMM.
#pragma once
#include <stdint.h>

constexpr uint32_t PK_LBO = 144;
constexpr uint32_t PK_SBO = 8 * PK_LBO;   // 1152
constexpr int PK_BK = 32;                  // fp32 elements of K per block

__host__ __device__ constexpr uint32_t pk_half_bytes(int rows) { return (uint32_t)(rows / 8) * PK_SBO; }
__host__ __device__ constexpr uint32_t pk_block_bytes(int rows) { return 2 * pk_half_bytes(rows); }
__host__ __device__ constexpr int pk_chunks(int K) { return (K + PK_BK - 1) / PK_BK; }
// bytes of a packed matrix with `tiles` row tiles
__host__ __device__ inline size_t pk_matrix_bytes(int64_t tiles, int K, int rows) {
    return (size_t)tiles * pk_chunks(K) * pk_block_bytes(rows);
}
// byte offset of element (r, kk) inside one half of a block (r < ROWS, kk < 32)
__host__ __device__ constexpr uint32_t pk_elem_off(int r, int kk) {
    return (uint32_t)(r >> 3) * PK_SBO + (uint32_t)(kk >> 2) * PK_LBO + (uint32_t)(r & 7) * 16 + (uint32_t)(kk & 3) * 4;
}

#ifdef __CUDACC__
// round-to-nearest (ties away) onto the 10-bit TF32 mantissa with full-rate integer ops; NaN / Inf pass through
// unchanged (the canonical NaN 0x7fffffff would otherwise carry into the sign bit and become -0: the reference
// propagates NaN rows, e.g. x_com of an all-zero w_{t-1} row, model/conv_transfer.py:94-98)
__device__ __forceinline__ float pk_tf32(float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t r = (b + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float((b & 0x7F800000u) == 0x7F800000u ? b : r);
}
__device__ __forceinline__ void pk_split(float x, float &hi, float &lo) {
    hi = pk_tf32(x);
    lo = pk_tf32(x - hi);   // x - hi is exact in fp32
}
// store one element into a packed matrix (scalar path, used by row-per-warp producers)
__device__ __forceinline__ void pk_store1(uint8_t *base, int rows_per_tile, int KC, int64_t row, int k, float x) {
    const int64_t tile = row / rows_per_tile;
    const int r = (int)(row - tile * rows_per_tile);
    uint8_t *blk = base + ((size_t)tile * KC + (k >> 5)) * pk_block_bytes(rows_per_tile) + pk_elem_off(r, k & 31);
    float hi, lo;
    pk_split(x, hi, lo);
    *reinterpret_cast<float *>(blk) = hi;
    *reinterpret_cast<float *>(blk + pk_half_bytes(rows_per_tile)) = lo;
}
#endif
