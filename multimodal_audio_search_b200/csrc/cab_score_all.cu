// Legacy scoring modes: ONE fused similarity per segment, all N of them, no threshold, no top-k.
//
// Replaces the per-item loop of the reference's earlier engine
// (previous_iterations/streamlit_app.py:173-223, `UnifiedAudioSearch.search`), which returns
// np.array(similarities) over the whole database:
//   asr_only     sim = cos(q, asr)                        (:188-193)
//   caption_only sim = cos(q, caption)                    (:195-200)
//   adaptive     sim = 0.7 cos_asr + 0.3 cos_cap   if len(transcript.strip()) > 10
//                      0.2 cos_asr + 0.8 cos_cap   otherwise                      (:202-219)
// A missing embedding contributes 0.0 (zero row in the index).  The per-row choice between weight
// pairs is a 2-bit weight class kept in bits 2-3 of the row's flag byte; the caller passes a
// 4-entry table {w_asr, w_audio} per class, so all three strategies (and any other per-row
// weighting) are one kernel.  No success gating and no renormalisation here: that is the current
// engine's rule (audio_search.py:654-670), served by the top-k scan.
//
// HBM-bound like the top-k scan: 2 x 384 x sizeof(elem) bytes read + 4 bytes written per segment.
// Same load shape (U row-steps x 2 corpora x 3 LDG.128 per lane in flight, L1-bypassing) and the
// same dynamic chunk schedule as the top-k scan; plain butterfly reduction (the issue slots are
// there: ~60 of ~500 per row are used).  A static warp-cyclic split measured 6-10 % slower
// (profiles/r01_score_all.md).
// Several queries per call are register-tiled like the top-k scan (QT = 2 or 4 queries per corpus
// pass: the rows are loaded once and multiplied with QT query vectors held in registers), so a
// 32-query legacy call costs 8 HBM passes instead of 32; bit-identical to one query at a time.
#include "cab_internal.h"
#include "cab_rowdot.cuh"

namespace cab {

constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;

// MB = resident CTAs per SM, part of the register contract with ptxas exactly as in the top-k
// scan: fp32 U=4 needs 96 registers for the loads in flight alone.
template <int DT, int U, int MB, int QT>
__global__ void __launch_bounds__(kScoreThreads, MB)
score_all_kernel(ScoreAllArgs a) {
    using TR = RowTraits<DT>;
    constexpr int G = TR::G, RW = TR::RW, CPR = TR::CPR;
    constexpr int kRowsPerIter = U * RW;

    __shared__ float s_q[kDim];
    __shared__ float s_cw[QT][4][2];                // class weights: read per row, too many for registers at QT = 4
    if (a.use_inline_query) {
        for (int i = threadIdx.x; i < kDim; i += kScoreThreads) s_q[i] = a.q[i];
    }
    if (threadIdx.x < QT * 8) s_cw[threadIdx.x >> 3][(threadIdx.x >> 1) & 3][threadIdx.x & 1] = a.class_w[threadIdx.x >> 3][(threadIdx.x >> 1) & 3][threadIdx.x & 1];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & (G - 1), sub = lane / G;

    float q[QT][TR::NQ];
    bool finite[QT];
    bool any_bad = false;
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const float *qsrc = a.query + size_t(t < a.n_q ? t : 0) * kDim;      // pad the tile with query 0, never stored
        finite[t] = load_query<DT>([&](int i) { return a.use_inline_query ? s_q[i] : qsrc[i]; }, lane, q[t]);
        any_bad |= !finite[t] && t < a.n_q;
    }
    if (any_bad && a.nonfinite && blockIdx.x == 0 && threadIdx.x == 0) *a.nonfinite = 1;

    const uint4 *__restrict__ A = reinterpret_cast<const uint4 *>(a.asr);
    const uint4 *__restrict__ B = reinterpret_cast<const uint4 *>(a.audio);
    const int64_t n = a.n_rows;
    const int64_t gwarp = int64_t(blockIdx.x) * kScoreWarps + warp;
    const int64_t total_warps = int64_t(gridDim.x) * kScoreWarps;

    // Dynamic two-level chunk schedule, the top-k scan's (cab_gemv.cu): a warp's first chunk is
    // its global warp id, every further one comes from an atomic ticket fetched one chunk ahead;
    // big chunks for the bulk, quarter-size chunks for the last big chunk per warp so the tail is
    // short.  A warp streams chunk_rows contiguous rows (96 KB over both corpora) per ticket.
    const int kChunkRows = a.chunk_rows >= kRowsPerIter ? (a.chunk_rows / kRowsPerIter) * kRowsPerIter : kRowsPerIter;
    const int kSmallRows = kChunkRows / 4 >= kRowsPerIter ? (kChunkRows / 4 / kRowsPerIter) * kRowsPerIter : kRowsPerIter;
    int64_t n_big = n / kChunkRows - total_warps;
    if (n_big < 0) n_big = 0;
    const int64_t tail_row0 = n_big * kChunkRows;
    const int64_t n_small = (n - tail_row0 + kSmallRows - 1) / kSmallRows;
    const int64_t n_chunks = n_big + n_small;
    unsigned int *counter = a.work_counters;
    int64_t chunk = gwarp;
    unsigned int ticket = 0;
    if (chunk < n_chunks && lane == 0) ticket = atomicAdd(counter, 1u);
    float keep[QT];                                         // this lane's pending scores (row cbase + 32j + lane)
#pragma unroll
    for (int t = 0; t < QT; ++t) keep[t] = 0.f;
    // Which row of a step (kRowsPerIter == 4 rows) this lane carries after the transposing
    // reduction, and the lane that carries row (lane & 3) -- see the reduction below.
    static_assert(kRowsPerIter == 4, "the score hand-over assumes 4 rows per step");
    int u_lane = 0;
    {
        int stride = G / 2, half = U / 2;
        while (half >= 1) { if (lane & stride) u_lane += half; stride >>= 1; half >>= 1; }
    }
    const int my_local = u_lane * RW + sub;
    const int want = lane & 3;                                                // row of the step this lane stores
    const int src_lane = (want % RW) * G + (want / RW) * (G / U);             // first lane carrying that row

    for (; chunk < n_chunks;
         chunk = total_warps + int64_t(__shfl_sync(kFull, ticket, 0)),
         ticket = (lane == 0 && chunk < n_chunks) ? atomicAdd(counter, 1u) : 0u)
    for (int64_t base = chunk < n_big ? chunk * kChunkRows : tail_row0 + (chunk - n_big) * kSmallRows,
                 cbase = base, cend = chunk < n_big ? base + kChunkRows : base + kSmallRows;
         base < cend && base < n; base += kRowsPerIter) {
        uint4 ca[U][3], cb[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int64_t r = base + u * RW + sub;
            r = r < n ? r : n - 1;                        // clamp: tail rows are not stored
            const uint4 *pa = A + r * CPR + g;
            const uint4 *pb = B + r * CPR + g;
#pragma unroll
            for (int j = 0; j < 3; ++j) { ca[u][j] = ldg_stream(pa + G * j); cb[u][j] = ldg_stream(pb + G * j); }
        }
        // after the transposing reduction this lane carries row `base + my_local` of the step
        const int64_t my_row = base + my_local;
        const uint32_t fl = a.flags[my_row < n ? my_row : n - 1];
        // Keep all 6U loads in flight: nothing below may be scheduled between the loads above.
#pragma unroll
        for (int u = 0; u < U; ++u) { keep_live(ca[u]); keep_live(cb[u]); }
        constexpr int NV = 2 * QT;
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
        // Transposing reduction (the top-k scan's): each split halves the row-steps a lane carries,
        // 12 shuffles per 4 fp32 rows instead of 40 -- fewer instructions, less power under the cap.
        {
            int stride = G / 2;
#pragma unroll
            for (int half = U / 2; half >= 1; half >>= 1) {
                const bool upper = (lane & stride) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i)
#pragma unroll
                    for (int c = 0; c < NV; ++c) {
                        const float send = upper ? v[i][c] : v[i + half][c];
                        const float kept = upper ? v[i + half][c] : v[i][c];
                        v[i][c] = kept + __shfl_xor_sync(kFull, send, stride);
                    }
                stride >>= 1;
            }
#pragma unroll
            for (; stride > 0; stride >>= 1)
#pragma unroll
                for (int c = 0; c < NV; ++c) v[0][c] += __shfl_xor_sync(kFull, v[0][c], stride);
        }
        const uint32_t cls = (fl >> 2) & 3u;
        const bool mine_now = (((base - cbase) >> 2) & 7) == (lane >> 2);
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            // fp32 products and sum rounded separately, as numpy evaluates `wa * s_asr + wb * s_cap`
            // on float32 scalars (no fused multiply-add)
            float f = __fadd_rn(__fmul_rn(s_cw[t][cls][0], v[0][2 * t]), __fmul_rn(s_cw[t][cls][1], v[0][2 * t + 1]));
            if (!finite[t]) f = __int_as_float(0x7fc00000);      // NaN/Inf query: every score is NaN
            // Lane (r & 31) keeps the score of the chunk's r-th row, so 32 rows leave as ONE coalesced
            // 128-byte store (full sectors): the step's row (lane & 3) sits in lane src_lane.
            const float mine = __shfl_sync(kFull, f, src_lane);
            if (mine_now) keep[t] = mine;
        }
        const int64_t done = base + kRowsPerIter;                    // rows of this chunk scored so far end here
        const int64_t lim = cend < n ? cend : n;
        if (((done - cbase) & 31) == 0 || done >= lim) {
            const int64_t sbase = cbase + ((done - 1 - cbase) & ~int64_t(31));
            const int64_t row = sbase + lane;
            if (row < done && row < lim) {
#pragma unroll
                for (int t = 0; t < QT; ++t)
                    if (t < a.n_q) a.out[int64_t(t) * a.out_stride + row] = keep[t];
            }
        }
    }
    // The last warp to run out of chunks re-arms the counters for the next launch (every warp's
    // last ticket fetch precedes its arrival here, so nobody touches the ticket counter after).
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(a.work_counters + 1, 1u) == unsigned(total_warps - 1)) {
            a.work_counters[0] = 0u;
            a.work_counters[1] = 0u;
        }
    }
}

template <int DT, int U, int MB, int QT>
static void launch(const ScoreAllArgs &a, int sm_count, cudaStream_t s) {
    using TR = RowTraits<DT>;
    const int64_t rows_per_cta = int64_t(kScoreWarps) * U * TR::RW;
    int64_t grid = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
    if (grid > int64_t(sm_count) * MB) grid = int64_t(sm_count) * MB;   // persistent: every SM full, once
    if (grid < 1) grid = 1;
    score_all_kernel<DT, U, MB, QT><<<int(grid), kScoreThreads, 0, s>>>(a);
}

// a.n_q queries per pass: 1, 2 or 3..4 (a tile of 4, the unused slots padded and not stored)
void launch_score_all(const ScoreAllArgs &a, int sm_count, cudaStream_t s) {
    const int qt = a.n_q >= 3 ? 4 : a.n_q;
    if (a.dtype == CAB_BF16) {
        if (qt == 4) launch<CAB_BF16, 2, 1, 4>(a, sm_count, s);
        else if (qt == 2) launch<CAB_BF16, 2, 1, 2>(a, sm_count, s);
        else launch<CAB_BF16, 2, 2, 1>(a, sm_count, s);
    } else {
        if (qt == 4) launch<CAB_F32, 4, 1, 4>(a, sm_count, s);
        else if (qt == 2) launch<CAB_F32, 4, 1, 2>(a, sm_count, s);
        else launch<CAB_F32, 4, 1, 1>(a, sm_count, s);
    }
}

}  // namespace cab
