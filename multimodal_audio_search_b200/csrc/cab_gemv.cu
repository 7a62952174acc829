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

// MB (min resident CTAs per SM) is part of the contract with ptxas: without it ptxas caps the
// kernel at 80 registers and interleaves loads with FMAs (7 loads in flight instead of 24).
// QT = queries scored per corpus pass (register tiling): the rows are loaded once and multiplied
// with QT query vectors held in registers, so a batch of fp32 queries costs one HBM pass per QT
// queries instead of one per query (24*QT FFMA per lane per row-step stay far below the FMA pipe:
// the kernel remains HBM-bound at QT = 4).
template <int DT, int U, int MB, int QT>
__global__ void __launch_bounds__(kScanThreads, MB)
cab_scan_kernel(ScanArgs a) {
    using TR = RowTraits<DT>;
    constexpr int G = TR::G, RW = TR::RW, CPR = TR::CPR;
    constexpr int kRowsPerIter = U * RW;
    constexpr int kOwnerLanes = G / U;            // lanes sharing one finished row
    constexpr int NV = 2 * QT;                    // values per row-step: (s_asr, s_audio) per query
    static_assert(U == 1 || U == 2 || U == 4 || U == 8, "U");

    extern __shared__ __align__(16) uint8_t scan_smem[];
    uint64_t(*s_keys)[QT][kWarpCap] = reinterpret_cast<uint64_t(*)[QT][kWarpCap]>(scan_smem);
    int(*s_count)[QT] = reinterpret_cast<int(*)[QT]>(scan_smem + sizeof(uint64_t) * kScanWarps * QT * kWarpCap);
    float *s_q = reinterpret_cast<float *>(scan_smem + sizeof(uint64_t) * kScanWarps * QT * kWarpCap + sizeof(int) * kScanWarps * QT);
    // Programmatic dependent launch (see ScanArgs::wait_early): a query / weights in device memory
    // may have been written by the kernel right in front of this one on the caller's stream (the
    // MiniLM layer that produced the embedding), so nothing is read from global memory before it
    // has completed.
    if (a.wait_early) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.inl.use_query) {
        // Kernel-argument query -> shared memory once per CTA (per-lane indexed reads of the
        // constant bank serialise 32-way; done by every warp they cost ~10 us per scan).
        for (int i = threadIdx.x; i < kDim; i += kScanThreads) s_q[i] = a.inl.q[i];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.y * QT;               // first query of this CTA's group
    const int g = lane & (G - 1), sub = lane / G;

    float q[QT][TR::NQ];
    ScanWeights w[QT];
    WarpTopK top[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const bool valid = q0 + t < a.n_queries;
        const int qi = valid ? q0 + t : q0;       // pad the group with its first query, never pushed
        const float *qsrc = a.queries + size_t(qi) * kDim;
        // a NaN/Inf query scores NaN everywhere and selects nothing; the finalize kernel reports it
        load_query<DT>([&](int i) { return a.inl.use_query ? s_q[i] : qsrc[i]; }, lane, q[t], a.norm_asr != nullptr);
        w[t] = a.inl.use_weights ? ScanWeights{a.inl.wa32, a.inl.wb32} : ScanWeights{a.wa32[qi], a.wb32[qi]};
        top[t].init(s_keys[warp][t], a.k, valid ? bound_key(a.select_threshold) : ~0ull);
    }

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
    // chunk comes from a per-group atomic counter, fetched one chunk ahead so the atomic's latency
    // is hidden.  Unlike a static split this tolerates SMs that are late or busy (another kernel,
    // e.g. an NCCL collective, holding an SM) and evens out SM-to-SM speed differences.
    // Two-level schedule: big chunks for the bulk, then quarter-size chunks for the last big chunk
    // per warp, so the tail -- warps finishing at different times while the memory system drains --
    // is a few microseconds instead of one big chunk's worth (the ticket counter is one address:
    // smaller tail chunks or a longer tail region would make the atomics the bottleneck).
    // The 64-chunk groups of the bulk are visited in a scattered order (group g at position
    // (g * super_mul) % n_super, a bijection): a library whose scores ASCEND with the row number
    // -- the adversarial order for a pruning scan, every row beating the running k-th best -- then
    // looks like a random one at group granularity (the bound is high after the first few groups
    // and whole groups are rejected by one compare), while all warps still work inside the same
    // few megabytes at any time, like the plain sweep.
    const int kChunkRows = a.rows_big, kSmallRows = a.rows_small;
    const int64_t n_big = a.n_big, tail_row0 = a.tail_row0, n_chunks = a.n_chunks;
    const int64_t n_grouped = int64_t(a.n_super) << 6;
    auto chunk_row0 = [&](int64_t c) -> int64_t {
        if (c >= n_big) return tail_row0 + (c - n_big) * kSmallRows;
        if (c < n_grouped) c = (int64_t((uint64_t(c >> 6) * a.super_mul) % a.n_super) << 6) | (c & 63);
        return c * kChunkRows;
    };
    unsigned int *counter = a.work_counters + blockIdx.y;
    int64_t chunk = gwarp;
    // The ticket counter was reset by the previous search's finalize kernel before it released its
    // dependents (griddepcontrol.launch_dependents), i.e. before this grid could start.
    unsigned int ticket = 0;                                  // lane 0: result of the in-flight atomic
    if (chunk < n_chunks && lane == 0) ticket = atomicAdd(counter, 1u);

    for (; chunk < n_chunks;
         chunk = total_warps + int64_t(__shfl_sync(kFull, ticket, 0)),
         ticket = (lane == 0 && chunk < n_chunks) ? atomicAdd(counter, 1u) : 0u)
    for (int64_t base = chunk_row0(chunk), cend = chunk < n_big ? base + kChunkRows : base + kSmallRows;
         base < cend && base < n; base += kRowsPerIter) {
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
        float len_a = 1.f, len_b = 1.f;                   // raw dot-product scoring: the rows' original lengths
        if (a.norm_asr) { len_a = a.norm_asr[my_row < n ? my_row : n - 1]; len_b = a.norm_audio[my_row < n ? my_row : n - 1]; }
        // Keep all 6U loads in flight: nothing below may be scheduled between the loads above.
#pragma unroll
        for (int u = 0; u < U; ++u) { keep_live(ca[u]); keep_live(cb[u]); }

        float v[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) { sa = dot_chunk<DT>(ca[u][j], q[t], j, sa); sb = dot_chunk<DT>(cb[u][j], q[t], j, sb); }
                v[u][2 * t] = sa; v[u][2 * t + 1] = sb;
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
                    for (int c = 0; c < NV; ++c) {
                        float send = upper ? v[i][c] : v[i + half][c];
                        float keep = upper ? v[i + half][c] : v[i][c];
                        v[i][c] = keep + __shfl_xor_sync(kFull, send, stride);
                    }
                stride >>= 1;
            }
#pragma unroll
            for (; stride > 0; stride >>= 1)
#pragma unroll
                for (int c = 0; c < NV; ++c) v[0][c] += __shfl_xor_sync(kFull, v[0][c], stride);
        }
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const float fused = fuse32(v[0][2 * t] * len_a, v[0][2 * t + 1] * len_b, fl, w[t]);
            const uint64_t key = make_key(fused, uint32_t(my_row));
            top[t].push(owner && my_row < n && key > top[t].bound, key, lane);
        }
    }

    // The partial-key slots written below were read by the previous search's finalize kernel before
    // it released its dependents; waiting here for that kernel to have COMPLETED (it has, long ago:
    // the scan took hundreds of microseconds) keeps the stream's completion order transitive -- the
    // finalize kernel after this scan waits on this grid only.
    if (!a.wait_early) asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- this warp's best k per query, then the CTA's best k --------------------------------------
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        top[t].compact(lane);
        if (lane == 0) s_count[warp][t] = top[t].count;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            if (q0 + t >= a.n_queries) break;
            for (int w2 = 1; w2 < kScanWarps; ++w2) {
                const int c2 = s_count[w2][t];
                for (int i = 0; i < c2; i += 32) {
                    const bool in = i + lane < c2;
                    const uint64_t key = in ? s_keys[w2][t][i + lane] : 0ull;
                    top[t].push(in && key > top[t].bound, key, lane);
                }
            }
            top[t].compact(lane);
            const size_t list = size_t(q0 + t) * a.n_partials + blockIdx.x;
            for (int i = lane; i < a.k; i += 32) a.partial_keys[list * a.k + i] = i < top[t].count ? top[t].buf[i] : 0ull;
        }
    }
}

// (unroll, blocks/SM) combinations that are instantiated; anything else maps to the nearest.
static void resolve(const GemvConfig &cfg, int dtype, int *u, int *mb) {
    // Defaults from the B200 sweep (profiles/r01_gemv_sweep_*.md): fewer, fatter warps win --
    // fp32: 1 CTA/SM x 8 warps x 24 LDG.128 per lane (98 KB in flight per SM) = 7.5 TB/s at 10M;
    // bf16: 2 CTAs/SM x 12 LDG.128 per lane = 7.3 TB/s.
    *u = cfg.unroll ? cfg.unroll : (dtype == CAB_BF16 ? 2 : 4);
    *mb = cfg.blocks_per_sm ? cfg.blocks_per_sm : (dtype == CAB_BF16 ? 2 : 1);
    if (*u == 8) *mb = 1;                     // 48 x LDG.128 per lane: 192 registers of loads
    if (*u == 4 && *mb > 2) *mb = 2;          // 24 x LDG.128 per lane needs > 96 registers
    if (*u == 2 && *mb > 4) *mb = 4;
    if (*mb > 4) *mb = 4;
}

GemvPlan plan_gemv(const GemvConfig &cfg, int dtype, int n_queries, int sm_count) {
    GemvPlan p{};
    const int qt_max = cfg.query_tile ? cfg.query_tile : 4;
    p.qt = n_queries >= 4 && qt_max >= 4 ? 4 : (n_queries >= 2 && qt_max >= 2 ? 2 : 1);
    if (p.qt == 1) resolve(cfg, dtype, &p.u, &p.mb);
    else {                                            // multi-query tiles: QT query register sets + U row-steps of loads
        p.u = (dtype == CAB_F32 && p.qt == 4 && cfg.unroll != 2) ? 4 : 2;     // fp32 x4: 785 vs 670 q/s at 10 M with U = 4
        p.mb = p.qt == 4 ? 1 : 2;
    }
    p.grid_x = sm_count * p.mb;
    p.groups = (n_queries + p.qt - 1) / p.qt;
    return p;
}
int gemv_max_grid(int sm_count) { return sm_count * 4; }

template <int DT, int U, int MB, int QT>
static void launch_one(const ScanArgs &a, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = sizeof(uint64_t) * kScanWarps * QT * kWarpCap + sizeof(int) * kScanWarps * QT + sizeof(float) * kDim;
    if (smem > 48 * 1024) {                       // opt-in shared memory size: per kernel AND per device
        static PerDeviceOnce once;
        int dev = -1;
        cudaGetDevice(&dev);
        if (!once.done(dev) &&
            cudaFuncSetAttribute(cab_scan_kernel<DT, U, MB, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) == cudaSuccess)
            once.mark(dev);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(kScanThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, cab_scan_kernel<DT, U, MB, QT>, a);
}

template <int DT>
static void launch_dt(const ScanArgs &a, const GemvPlan &p, dim3 grid, cudaStream_t s) {
    if (p.qt == 4) {
        if constexpr (DT == CAB_F32) { if (p.u == 4) { launch_one<DT, 4, 1, 4>(a, grid, s); return; } }
        launch_one<DT, 2, 1, 4>(a, grid, s);
        return;
    }
    if (p.qt == 2) { launch_one<DT, 2, 2, 2>(a, grid, s); return; }
#define CAB_CASE(U_, MB_) if (p.u == U_ && p.mb == MB_) { launch_one<DT, U_, MB_, 1>(a, grid, s); return; }
    CAB_CASE(8, 1) CAB_CASE(4, 1) CAB_CASE(4, 2)
    CAB_CASE(2, 1) CAB_CASE(2, 2) CAB_CASE(2, 3) CAB_CASE(2, 4)
    CAB_CASE(1, 1) CAB_CASE(1, 2) CAB_CASE(1, 3) CAB_CASE(1, 4)
#undef CAB_CASE
}

static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }

// Two-level chunk schedule (see the kernel) for `n_rows` rows scanned by `total_warps` warps that
// each take `rows_per_iter` rows per step; `want_rows` = requested rows per big chunk.
void gemv_chunk_layout(ScanArgs *a, int want_rows, int rows_per_iter, int64_t total_warps) {
    a->rows_big = want_rows >= rows_per_iter ? (want_rows / rows_per_iter) * rows_per_iter : rows_per_iter;
    a->rows_small = a->rows_big / 4 >= rows_per_iter ? (a->rows_big / 4 / rows_per_iter) * rows_per_iter : rows_per_iter;
    int64_t n_big = a->n_rows / a->rows_big - total_warps;     // big chunks handed out before the tail
    if (n_big < 0) n_big = 0;
    a->n_big = n_big;
    a->tail_row0 = n_big * a->rows_big;
    a->n_chunks = n_big + (a->n_rows - a->tail_row0 + a->rows_small - 1) / a->rows_small;
    a->n_super = uint32_t(n_big >> 6);
    uint64_t mul = 1;
    if (a->n_super > 2) {                                      // ~golden-ratio stride, coprime to the group count
        mul = uint64_t(double(a->n_super) * 0.6180339887) | 1ull;
        while (gcd64(mul, a->n_super) != 1) mul += 2;
        mul %= a->n_super;
        if (mul == 0) mul = 1;
    }
    a->super_mul = uint32_t(mul);
}

void launch_gemv_scan(const ScanArgs &a_in, const GemvPlan &p, cudaStream_t s) {
    ScanArgs a = a_in;
    const int rw = a.dtype == CAB_BF16 ? RowTraits<CAB_BF16>::RW : RowTraits<CAB_F32>::RW;
    gemv_chunk_layout(&a, a.chunk_rows, p.u * rw, int64_t(p.grid_x) * kScanWarps);
    dim3 grid(p.grid_x, p.groups);
    if (a.dtype == CAB_BF16) launch_dt<CAB_BF16>(a, p, grid, s);
    else launch_dt<CAB_F32>(a, p, grid, s);
}

}  // namespace cab
