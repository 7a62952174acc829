// After the scan: (1) finalize -- merge the per-CTA partial top-k lists of one query into its
// best k rows and re-score those rows (both cosines, fp32) -> cab_candidate records;
// (2) emit -- from 1..world candidate lists per query compute the reference's float64 fusion
// score (audio_search.py:656-670), apply the strict float64 threshold (:672), order by
// (score desc, global index asc) (:685) and write the first k (:699).
//
// Both kernels touch O(k) rows per query; they are latency-, not bandwidth-bound, and run as one
// CTA per query.
#include "cab_internal.h"
#include "cab_rowdot.cuh"

namespace cab {

constexpr int kFinThreads = 512;
constexpr int kFinWarps = kFinThreads / 32;

template <int DT>
__global__ void __launch_bounds__(kFinThreads) finalize_kernel(FinalizeArgs a) {
    using TR = RowTraits<DT>;
    __shared__ uint64_t s_keys[kFinWarps][kWarpCap];
    __shared__ int s_count[kFinWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x;

    WarpTopK top;
    top.init(s_keys[warp], a.k, 0ull);
    for (int p = warp; p < a.n_partials; p += kFinWarps) {
        const size_t list = size_t(qi) * a.n_partials + p;
        const int c = a.partial_count[list];
        for (int i = 0; i < c; i += 32) {
            const bool in = i + lane < c;
            const uint64_t key = in ? a.partial_keys[list * a.k + i + lane] : 0ull;
            top.push(in && key > top.bound, key, lane);
        }
    }
    top.compact(lane);
    if (lane == 0) s_count[warp] = top.count;
    __syncthreads();
    if (warp == 0) {
        for (int w2 = 1; w2 < kFinWarps; ++w2) {
            const int c2 = s_count[w2];
            for (int i = 0; i < c2; i += 32) {
                const bool in = i + lane < c2;
                const uint64_t key = in ? s_keys[w2][i + lane] : 0ull;
                top.push(in && key > top.bound, key, lane);
            }
        }
        top.compact(lane);
        if (lane == 0) s_count[0] = top.count;
    }
    __syncthreads();
    const int n_win = s_count[0];

    // ---- re-score the winners: one lane group per row, same operation order as the scan ---------
    float q[TR::NQ];
    load_query<DT>(a.queries + size_t(qi) * kDim, lane, q);
    const int g = lane & (TR::G - 1), sub = lane / TR::G;
    const uint4 *__restrict__ A = reinterpret_cast<const uint4 *>(a.asr);
    const uint4 *__restrict__ B = reinterpret_cast<const uint4 *>(a.audio);
    cab_candidate *out = a.cands + size_t(qi) * a.k;
    for (int i0 = warp * TR::RW; i0 < a.k; i0 += kFinWarps * TR::RW) {
        const int i = i0 + sub;
        const bool have = i < n_win;                       // uniform within a lane group
        const uint32_t row = have ? key_row(s_keys[0][i]) : 0u;
        float sa = 0.f, sb = 0.f;
        if (have) {
            const uint4 *pa = A + size_t(row) * TR::CPR + g;
            const uint4 *pb = B + size_t(row) * TR::CPR + g;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                sa = dot_chunk<DT>(pa[TR::G * j], q, j, sa);
                sb = dot_chunk<DT>(pb[TR::G * j], q, j, sb);
            }
        }
        sa = group_sum<DT>(sa);
        sb = group_sum<DT>(sb);
        if (g == 0 && i < a.k) {
            cab_candidate c;
            c.index = have ? a.row_base + int64_t(row) : int64_t(-1);
            c.asr_sim = sa; c.audio_sim = sb;
            c.flags = have ? uint32_t(a.flags[row]) : 0u;
            c.pad = 0u;
            out[i] = c;
        }
    }
}

void launch_finalize(const FinalizeArgs &a, cudaStream_t s) {
    if (a.dtype == CAB_BF16) finalize_kernel<CAB_BF16><<<a.n_queries, kFinThreads, 0, s>>>(a);
    else finalize_kernel<CAB_F32><<<a.n_queries, kFinThreads, 0, s>>>(a);
}

// ---- emit -------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitMax = 1024;      // >= world(8) x CAB_MAX_K(128)

__device__ __forceinline__ uint64_t orderable64(double d) {
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kEmitThreads) emit_kernel(EmitArgs a) {
    __shared__ uint64_t s_score[kEmitMax];     // orderable float64 fusion score, 0 = not a result
    __shared__ int64_t s_index[kEmitMax];
    __shared__ uint16_t s_pos[kEmitMax];
    __shared__ int s_n;

    const int qi = blockIdx.x;
    const int n_cand = a.n_lists * a.k;
    int np2 = 64;
    while (np2 < n_cand) np2 <<= 1;
    const double wa = a.w_asr[qi], wb = a.w_audio[qi];
    if (threadIdx.x == 0) s_n = 0;

    for (int t = threadIdx.x; t < np2; t += kEmitThreads) {
        uint64_t sc = 0ull;
        int64_t gi = INT64_MAX;
        if (t < n_cand) {
            const int list = t / a.k, i = t - list * a.k;
            const cab_candidate c = a.cands[(size_t(list) * a.n_queries + qi) * a.k + i];
            if (c.index >= 0) {
                // audio_search.py:654-672, float64 with separate multiply/add like CPython
                const double sa = double(c.asr_sim), sb = double(c.audio_sim);
                double ea = (c.flags & 1u) ? wa : 0.0;
                double eb = (c.flags & 2u) ? wb : 0.0;
                const double tot = __dadd_rn(ea, eb);
                if ((sa > 0.0 || sb > 0.0) && tot > 0.0) {
                    ea = __ddiv_rn(ea, tot);
                    eb = __ddiv_rn(eb, tot);
                    const double fusion = __dadd_rn(__dmul_rn(ea, sa), __dmul_rn(eb, sb));
                    if (fusion > a.threshold) { sc = orderable64(fusion); gi = c.index; }
                }
            }
        }
        s_score[t] = sc; s_index[t] = gi; s_pos[t] = uint16_t(t);
    }
    // bitonic sort, descending by (score, -index)
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < np2 / 2; t += kEmitThreads) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t s0 = s_score[lo], s1 = s_score[hi];
                const int64_t i0 = s_index[lo], i1 = s_index[hi];
                const bool lo_lt_hi = (s0 < s1) || (s0 == s1 && i0 > i1);   // "lo ranks below hi"
                if (lo_lt_hi == desc && !(s0 == s1 && i0 == i1)) {
                    s_score[lo] = s1; s_score[hi] = s0;
                    s_index[lo] = i1; s_index[hi] = i0;
                    const uint16_t p = s_pos[lo]; s_pos[lo] = s_pos[hi]; s_pos[hi] = p;
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.k; i += kEmitThreads) {
        const bool res = s_score[i] != 0ull;
        const size_t o = size_t(qi) * a.k + i;
        if (res) {
            const int t = s_pos[i];
            const int list = t / a.k, ii = t - list * a.k;
            const cab_candidate c = a.cands[(size_t(list) * a.n_queries + qi) * a.k + ii];
            const uint64_t so = s_score[i];
            const uint64_t u = (so >> 63) ? (so & 0x7FFFFFFFFFFFFFFFull) : ~so;
            a.out_index[o] = c.index;
            a.out_fusion[o] = __longlong_as_double((long long)u);
            a.out_asr[o] = c.asr_sim;
            a.out_audio[o] = c.audio_sim;
            a.out_flags[o] = uint8_t(c.flags);
            atomicAdd(&s_n, 1);
        } else {
            a.out_index[o] = -1;
            a.out_fusion[o] = 0.0;
            a.out_asr[o] = 0.f;
            a.out_audio[o] = 0.f;
            a.out_flags[o] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.out_count[qi] = s_n;
}

void launch_emit(const EmitArgs &a, cudaStream_t s) {
    emit_kernel<<<a.n_queries, kEmitThreads, 0, s>>>(a);
}

}  // namespace cab
