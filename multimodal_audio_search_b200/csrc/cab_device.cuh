// Device-side building blocks shared by the scan / finalize kernels (sm_100a only).
//
//  * candidate keys:  64-bit  (orderable fp32 fused score << 32) | (0xFFFFFFFF - local row)
//    so that a plain unsigned ">" is "higher score first, then lower segment index" -- the
//    order Python's stable descending sort produces (audio_search.py:685).
//  * WarpTopK: a per-warp candidate buffer in shared memory with a running k-th-best bound;
//    almost every row is rejected by one compare, survivors are appended with a ballot and the
//    buffer is compacted (bitonic sort, keep k) only when it fills.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cab.h"

namespace cab {

constexpr int kDim = CAB_DIM;
constexpr int kMaxK = CAB_MAX_K;
constexpr int kWarpCap = 256;             // keys per WarpTopK buffer (>= kMaxK + 32, power of 2)
constexpr unsigned kFull = 0xFFFFFFFFu;

// ---- keys ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return (uint64_t(orderable(score)) << 32) | uint64_t(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - uint32_t(key); }
__device__ __forceinline__ float key_score(uint64_t key) { return unorderable(uint32_t(key >> 32)); }
// Lowest key with this score: "key > bound_key(s)" <=> score strictly greater than s.
__device__ __forceinline__ uint64_t bound_key(float score) {
    return (uint64_t(orderable(score)) << 32) | 0xFFFFFFFFull;
}

// ---- loads -----------------------------------------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (each corpus byte is used once).
__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---- effective weights (audio_search.py:656-664) in fp32 for the scan ------------------------
struct ScanWeights {
    float wa, wb;          // w_asr/(w_asr+w_audio), w_audio/(w_asr+w_audio) as fp32
};
__device__ __forceinline__ float fuse32(float sa, float sb, uint32_t flags, ScanWeights w) {
    // both pipelines: (wa, wb); one: weight 1 on it -- unless its query weight is 0, then the
    // effective weights sum to 0 and the row is skipped (:659-661); none: skipped (-inf).
    // (a positive query weight is staged as a positive fp32, see stage_weight in cab_api.cu)
    float ea = (flags & 1u) ? ((flags & 2u) ? w.wa : (w.wa > 0.f ? 1.0f : 0.0f)) : 0.0f;
    float eb = (flags & 2u) ? ((flags & 1u) ? w.wb : (w.wb > 0.f ? 1.0f : 0.0f)) : 0.0f;
    float f = fmaf(ea, sa, eb * sb);
    // :654 gate (redundant for the reference's positive threshold, kept for any threshold)
    return ((ea > 0.f || eb > 0.f) && (sa > 0.f || sb > 0.f)) ? f : -INFINITY;
}

// ---- bitonic sort (descending) of a power-of-two array of u64 keys in shared memory, by one
//      warp.  Cold path: called only when a candidate buffer fills or at the end of a scan. -----
template <int N>
__device__ __noinline__ void warp_sort_desc(uint64_t *buf, int lane) {
    static_assert((N & (N - 1)) == 0 && N >= 64, "N must be a power of two >= 64");
    for (int size = 2; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < N / 2; t += 32) {
                int lo = 2 * t - (t & (stride - 1));        // index with bit `stride` clear
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = buf[lo], b = buf[hi];
                if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncwarp();
}

// Block-wide version (all threads of the CTA participate, N up to smem size).
__device__ __forceinline__ void block_sort_desc(uint64_t *buf, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n_pow2 / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = buf[lo], b = buf[hi];
                if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// ---- per-warp running top-k ------------------------------------------------------------------
struct WarpTopK {
    uint64_t *buf;       // kWarpCap keys in shared memory, private to this warp
    uint64_t bound;      // warp-uniform: only keys > bound can still enter the top-k
    uint64_t floor_;     // the threshold bound (never lowered)
    int count;           // warp-uniform number of valid keys in buf
    int k;

    __device__ __forceinline__ void init(uint64_t *smem, int k_, uint64_t floor_key) {
        buf = smem; k = k_; floor_ = floor_key; bound = floor_key; count = 0;
    }
    // Sort, keep the best k, tighten the bound.  Leaves buf[0..count) sorted descending.
    // The sort network is sized to the fill (64 / 128 / 256 keys): at the end of a scan most
    // buffers hold a handful of keys, and a 64-key sort is ~6x cheaper than the full one.
    __device__ __noinline__ void compact(int lane) {
        const int n = count <= 64 ? 64 : (count <= 128 ? 128 : kWarpCap);
        for (int i = count + lane; i < n; i += 32) buf[i] = 0ull;
        if (n == 64) warp_sort_desc<64>(buf, lane);
        else if (n == 128) warp_sort_desc<128>(buf, lane);
        else warp_sort_desc<kWarpCap>(buf, lane);
        if (count > k) count = k;
        if (count == k) bound = buf[k - 1] > floor_ ? buf[k - 1] : floor_;
        __syncwarp();
    }
    // Warp-cooperative append: every lane calls; lanes with pass==true contribute `key`.
    __device__ __forceinline__ void push(bool pass, uint64_t key, int lane) {
        unsigned m = __ballot_sync(kFull, pass);
        if (m == 0) return;
        if (pass) buf[count + __popc(m & ((1u << lane) - 1u))] = key;
        count += __popc(m);
        if (count > kWarpCap - 32) compact(lane);
    }
};

}  // namespace cab
