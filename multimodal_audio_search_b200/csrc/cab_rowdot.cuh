// Row layout traits + the per-lane dot-product pieces shared by the GEMV scan and the finalize
// (re-score) kernel, so both compute a similarity with the same operation order.
//
// A row is 384 elements = 96 x 16-byte chunks (fp32) or 48 chunks (bf16).  A row is owned by a
// group of G lanes; lane g of the group holds chunks g, g+G, g+2G (3 chunks for both layouts):
//   fp32: G = 32 (one row per warp step),  4 elements per chunk, 12 query elements per lane
//   bf16: G = 16 (two rows per warp step), 8 elements per chunk, 24 query elements per lane
#pragma once
#include "cab_device.cuh"

namespace cab {

template <int DT> struct RowTraits;
template <> struct RowTraits<CAB_F32> {
    static constexpr int G = 32;              // lanes per row
    static constexpr int EPC = 4;             // elements per 16-byte chunk
    static constexpr int CPR = kDim / EPC;    // chunks per row (96)
    static constexpr int RW = 32 / G;         // rows per warp step
    static constexpr int NQ = 3 * EPC;        // query elements held per lane
    static constexpr int ROW_BYTES = kDim * 4;
};
template <> struct RowTraits<CAB_BF16> {
    static constexpr int G = 16;
    static constexpr int EPC = 8;
    static constexpr int CPR = kDim / EPC;    // 48
    static constexpr int RW = 32 / G;
    static constexpr int NQ = 3 * EPC;
    static constexpr int ROW_BYTES = kDim * 2;
};

// Normalised query elements this lane multiplies with (sklearn normalize(X): x / sqrt(sum x^2),
// zero norm -> 1).  Every warp computes the norm with the same order, so all agree bit-for-bit.
// Returns false (warp-uniform) if the query holds NaN/Inf.
// `get(i)` returns raw query element i (global memory, or the kernel-argument copy).
// raw = true: keep the query unnormalised (raw dot-product scoring).
template <int DT, typename Get>
__device__ __forceinline__ bool load_query(Get get, int lane, float (&q)[RowTraits<DT>::NQ], bool raw = false) {
    using TR = RowTraits<DT>;
    float ss = 0.f;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < kDim / 32; ++i) {
        float x = get(lane + 32 * i);
        bad |= !isfinite(x);
        ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
    bad = __any_sync(kFull, bad) || !isfinite(ss);
    float norm = sqrtf(ss);
    if (norm == 0.f || raw) norm = 1.f;
    const int g = lane & (TR::G - 1);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int e = 0; e < TR::EPC; ++e) q[j * TR::EPC + e] = get((g + TR::G * j) * TR::EPC + e) / norm;
    return !bad;
}

// Empty asm that consumes and re-defines three loaded chunks: volatile asms keep their order, so
// every use of the chunks is pinned after ALL the loads issued before this point.
__device__ __forceinline__ void keep_live(uint4 (&c)[3]) {
    asm volatile("" : "+r"(c[0].x), "+r"(c[0].y), "+r"(c[0].z), "+r"(c[0].w),
                      "+r"(c[1].x), "+r"(c[1].y), "+r"(c[1].z), "+r"(c[1].w),
                      "+r"(c[2].x), "+r"(c[2].y), "+r"(c[2].z), "+r"(c[2].w));
}

// acc += <chunk j of a row, the lane's query elements for chunk j>
template <int DT>
__device__ __forceinline__ float dot_chunk(uint4 c, const float (&q)[RowTraits<DT>::NQ], int j, float acc) {
    if constexpr (DT == CAB_F32) {
        acc = fmaf(__uint_as_float(c.x), q[4 * j + 0], acc);
        acc = fmaf(__uint_as_float(c.y), q[4 * j + 1], acc);
        acc = fmaf(__uint_as_float(c.z), q[4 * j + 2], acc);
        acc = fmaf(__uint_as_float(c.w), q[4 * j + 3], acc);
    } else {
        acc = fmaf(bf16lo(c.x), q[8 * j + 0], acc); acc = fmaf(bf16hi(c.x), q[8 * j + 1], acc);
        acc = fmaf(bf16lo(c.y), q[8 * j + 2], acc); acc = fmaf(bf16hi(c.y), q[8 * j + 3], acc);
        acc = fmaf(bf16lo(c.z), q[8 * j + 4], acc); acc = fmaf(bf16hi(c.z), q[8 * j + 5], acc);
        acc = fmaf(bf16lo(c.w), q[8 * j + 6], acc); acc = fmaf(bf16hi(c.w), q[8 * j + 7], acc);
    }
    return acc;
}

// Sum a per-lane partial over the G lanes of a row group (same tree as the scan's reduction).
template <int DT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = RowTraits<DT>::G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

}  // namespace cab
