// Single-query path: fused dual-corpus GEMV + weighted fusion + running top-k  (HBM-bound).
//
// Replaces the per-segment loop of search_with_fusion (audio_search.py:639-682): two cosines
// (rows are unit length, so a dot product), effective-weight fusion, threshold.  Both corpora are
// streamed exactly once; nothing is written back except the (few) survivors.
//
// Algorithmic traffic: 2 x 384 x sizeof(elem) bytes per (query, segment) = 3072 B fp32 /
// 1536 B bf16 (+1 flag byte, not counted).  At 6.5 TB/s that is one fp32 row pair every
// ~70 SM-cycles chip-wide, i.e. ~133 cycles per row per SM: the kernel has ~500 issue slots per
// row and needs only ~60, so the design goal is purely bytes-in-flight:
//   persistent CTAs; every warp keeps U row-steps (U*RW rows x 2 corpora x
//     3 x 16 B per lane = 6U independent LDG.128 per lane) in flight in registers, 128-bit
//     coalesced L1-bypassing loads; a transposing shuffle reduction (12 shuffles per 4 rows
//     instead of 40) leaves each row's (s_asr, s_audio) in one lane group; fusion in registers;
//     one compare against the warp's running k-th best rejects almost every row.
#include "cab_internal.h"
#include "cab_rowdot.cuh"

namespace cab {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

// Empty asm that consumes and re-defines three loaded chunks: volatile asms keep their order, so
// every use of the chunks is pinned after ALL the loads issued before this point.
__device__ __forceinline__ void keep_live(uint4 (&c)[3]) {
    asm volatile("" : "+r"(c[0].x), "+r"(c[0].y), "+r"(c[0].z), "+r"(c[0].w),
                      "+r"(c[1].x), "+r"(c[1].y), "+r"(c[1].z), "+r"(c[1].w),
                      "+r"(c[2].x), "+r"(c[2].y), "+r"(c[2].z), "+r"(c[2].w));
}

// MB (min resident CTAs per SM) is part of the contract with ptxas: without it ptxas caps the
// kernel at 80 registers and interleaves loads with FMAs (7 loads in flight instead of 24).
template <int DT, int U, int MB>
__global__ void __launch_bounds__(kScanThreads, MB)
gemv_scan_kernel(ScanArgs a) {
    using TR = RowTraits<DT>;
    constexpr int G = TR::G, RW = TR::RW, CPR = TR::CPR;
    constexpr int kRowsPerIter = U * RW;
    constexpr int kOwnerLanes = G / U;            // lanes sharing one finished row
    static_assert(U == 1 || U == 2 || U == 4 || U == 8, "U");

    __shared__ uint64_t s_keys[kScanWarps][kWarpCap];
    __shared__ int s_count[kScanWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.y;
    const int g = lane & (G - 1), sub = lane / G;

    float q[TR::NQ];
    bool ok = load_query<DT>(a.queries + size_t(qi) * kDim, lane, q);
    if (!ok && lane == 0) *a.nonfinite = 1;
    const ScanWeights w{a.wa32[qi], a.wb32[qi]};

    WarpTopK top;
    top.init(s_keys[warp], a.k, bound_key(a.select_threshold));

    // Which of the U row-steps this lane ends up owning after the transposing reduction.
    int u_lane = 0;
    {
        int stride = G / 2, half = U / 2;
        while (half >= 1) { if (lane & stride) u_lane += half; stride >>= 1; half >>= 1; }
    }
    const bool owner = (lane & (kOwnerLanes - 1)) == 0;

    const uint4 *__restrict__ A = reinterpret_cast<const uint4 *>(a.asr);
    const uint4 *__restrict__ B = reinterpret_cast<const uint4 *>(a.audio);
    const int64_t n = a.n_rows;
    const int64_t gwarp = int64_t(blockIdx.x) * kScanWarps + warp;
    const int64_t total_warps = int64_t(gridDim.x) * kScanWarps;

    // Dynamic chunk scheduling: a warp's first chunk is static (its global warp id); every further
    // chunk comes from a per-query atomic counter, fetched one chunk ahead so the atomic's latency
    // is hidden.  Unlike a static split this tolerates SMs that are late or busy (another kernel,
    // e.g. an NCCL collective, holding an SM) and evens out SM-to-SM speed differences.
    constexpr int kChunkRows = kRowsPerIter * (64 / kRowsPerIter);      // 64 rows per chunk
    const int64_t n_chunks = (n + kChunkRows - 1) / kChunkRows;
    unsigned int *counter = a.work_counters + qi;
    int64_t chunk = gwarp;
    unsigned int ticket = 0;                                  // lane 0: result of the in-flight atomic
    if (chunk < n_chunks && lane == 0) ticket = atomicAdd(counter, 1u);

    for (; chunk < n_chunks;
         chunk = total_warps + int64_t(__shfl_sync(kFull, ticket, 0)),
         ticket = (lane == 0 && chunk < n_chunks) ? atomicAdd(counter, 1u) : 0u)
    for (int64_t base = chunk * kChunkRows, cend = base + kChunkRows; base < cend && base < n; base += kRowsPerIter) {
        uint4 ca[U][3], cb[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int64_t r = base + u * RW + sub;
            r = r < n ? r : n - 1;                        // clamp: tail rows are masked at push
            const uint4 *pa = A + r * CPR + g;
            const uint4 *pb = B + r * CPR + g;
#pragma unroll
            for (int j = 0; j < 3; ++j) { ca[u][j] = ldg_stream(pa + G * j); cb[u][j] = ldg_stream(pb + G * j); }
        }
        const int64_t my_row = base + u_lane * RW + sub;
        const uint32_t fl = a.flags[my_row < n ? my_row : n - 1];
        // Keep all 6U loads in flight: nothing below may be scheduled between the loads above.
#pragma unroll
        for (int u = 0; u < U; ++u) { keep_live(ca[u]); keep_live(cb[u]); }

        float v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { sa = dot_chunk<DT>(ca[u][j], q, j, sa); sb = dot_chunk<DT>(cb[u][j], q, j, sb); }
            v[u][0] = sa; v[u][1] = sb;
        }
        // Transposing reduction: each split halves the number of row-steps a lane carries.
        {
            int stride = G / 2;
#pragma unroll
            for (int half = U / 2; half >= 1; half >>= 1) {
                const bool upper = (lane & stride) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float send = upper ? v[i][c] : v[i + half][c];
                        float keep = upper ? v[i + half][c] : v[i][c];
                        v[i][c] = keep + __shfl_xor_sync(kFull, send, stride);
                    }
                stride >>= 1;
            }
#pragma unroll
            for (; stride > 0; stride >>= 1) {
                v[0][0] += __shfl_xor_sync(kFull, v[0][0], stride);
                v[0][1] += __shfl_xor_sync(kFull, v[0][1], stride);
            }
        }
        const float fused = fuse32(v[0][0], v[0][1], fl, w);
        const uint64_t key = make_key(fused, uint32_t(my_row));
        top.push(owner && my_row < n && key > top.bound, key, lane);
    }

    // ---- this warp's best k, then the CTA's best k ---------------------------------------------
    top.compact(lane);
    if (lane == 0) s_count[warp] = top.count;
    __syncthreads();
    if (warp == 0) {
        for (int w2 = 1; w2 < kScanWarps; ++w2) {
            const int c2 = s_count[w2];
            for (int i = 0; i < c2; i += 32) {
                const bool in = i + lane < c2;
                const uint64_t key = in ? s_keys[w2][i + lane] : 0ull;
                top.push(in && key > top.bound, key, lane);
            }
        }
        top.compact(lane);
        const size_t list = size_t(qi) * a.n_partials + blockIdx.x;
        for (int i = lane; i < a.k; i += 32) a.partial_keys[list * a.k + i] = i < top.count ? top.buf[i] : 0ull;
    }
}

// (unroll, blocks/SM) combinations that are instantiated; anything else maps to the nearest.
static void resolve(const GemvConfig &cfg, int dtype, int *u, int *mb) {
    // Defaults from the B200 sweep (profiles/r01_gemv_sweep.md): fewer, fatter warps win --
    // fp32: 1 CTA/SM x 8 warps x 24 LDG.128 per lane (98 KB in flight per SM) = 7.33 TB/s at 10M;
    // bf16: 2 CTAs/SM x 12 LDG.128 per lane = 7.25 TB/s.
    *u = cfg.unroll ? cfg.unroll : (dtype == CAB_BF16 ? 2 : 4);
    *mb = cfg.blocks_per_sm ? cfg.blocks_per_sm : (dtype == CAB_BF16 ? 2 : 1);
    if (*u == 8) *mb = 1;                     // 48 x LDG.128 per lane: 192 registers of loads
    if (*u == 4 && *mb > 2) *mb = 2;          // 24 x LDG.128 per lane needs > 96 registers
    if (*u == 2 && *mb > 4) *mb = 4;
    if (*mb > 4) *mb = 4;
}

int gemv_grid_size(const GemvConfig &cfg, int dtype, int sm_count) {
    int u, mb;
    resolve(cfg, dtype, &u, &mb);
    return sm_count * mb;
}

template <int DT>
static void launch_dt(const ScanArgs &a, int u, int mb, dim3 grid, cudaStream_t s) {
#define CAB_CASE(U_, MB_) if (u == U_ && mb == MB_) { gemv_scan_kernel<DT, U_, MB_><<<grid, kScanThreads, 0, s>>>(a); return; }
    CAB_CASE(8, 1) CAB_CASE(4, 1) CAB_CASE(4, 2)
    CAB_CASE(2, 1) CAB_CASE(2, 2) CAB_CASE(2, 3) CAB_CASE(2, 4)
    CAB_CASE(1, 1) CAB_CASE(1, 2) CAB_CASE(1, 3) CAB_CASE(1, 4)
#undef CAB_CASE
}

void launch_gemv_scan(const ScanArgs &a, const GemvConfig &cfg, int sm_count, cudaStream_t s) {
    (void)sm_count;
    int u, mb;
    resolve(cfg, a.dtype, &u, &mb);
    dim3 grid(a.n_partials, a.n_queries);
    if (a.dtype == CAB_BF16) launch_dt<CAB_BF16>(a, u, mb, grid, s);
    else launch_dt<CAB_F32>(a, u, mb, grid, s);
}

}  // namespace cab
