// After the scan: (1) finalize -- merge the per-CTA partial top-k slots of one query into its
// best k rows and re-score those rows (both cosines, fp32) -> cab_candidate records;
// (2) emit -- from 1..world candidate lists per query compute the reference's float64 fusion
// score (audio_search.py:656-670), apply the strict float64 threshold (:672), order by
// (score desc, global index asc) (:685) and write the first k (:699).
// On a single GPU (one candidate list) emit runs inside the finalize kernel.
//
// Both touch O(k) rows per query; they are latency-, not bandwidth-bound, and run as one CTA per
// query: every global load is issued in parallel (no dependent chains), all selection happens in
// shared memory.
#include "cab_internal.h"
#include "cab_rowdot.cuh"

namespace cab {

constexpr int kFinThreads = 1024;
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kSortCap = 4096;          // keys in the shared-memory selection buffer
constexpr int kFinUnroll = 2;           // slots loaded per thread per round
constexpr int kMaxHeads = 1024;         // scan CTAs (partial lists) the fast path can rank
constexpr int kRankMax = 2048;          // survivors the fast path ranks by counting

__device__ __forceinline__ uint64_t orderable64(double d) {
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unorderable64(uint64_t o) {
    uint64_t u = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
    return __longlong_as_double((long long)u);
}

// audio_search.py:654-672 in float64 with separate multiply / add like CPython.  Returns the
// orderable score, or 0 when the candidate is not a result.
__device__ __forceinline__ uint64_t reference_fusion(const cab_candidate &c, double wa, double wb,
                                                     double threshold) {
    if (c.index < 0) return 0ull;
    const double sa = double(c.asr_sim), sb = double(c.audio_sim);
    double ea = (c.flags & 1u) ? wa : 0.0;
    double eb = (c.flags & 2u) ? wb : 0.0;
    const double tot = __dadd_rn(ea, eb);
    if (!((sa > 0.0 || sb > 0.0) && tot > 0.0)) return 0ull;
    ea = __ddiv_rn(ea, tot);
    eb = __ddiv_rn(eb, tot);
    const double fusion = __dadd_rn(__dmul_rn(ea, sa), __dmul_rn(eb, sb));
    return fusion > threshold ? orderable64(fusion) : 0ull;
}

// Block-wide bitonic sort of (score desc, index asc) triples in shared memory.
__device__ __forceinline__ void block_sort_results(uint64_t *score, int64_t *index, uint16_t *pos, int np2) {
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < np2 / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t s0 = score[lo], s1 = score[hi];
                const int64_t i0 = index[lo], i1 = index[hi];
                const bool lo_below_hi = (s0 < s1) || (s0 == s1 && i0 > i1);
                if (lo_below_hi == desc && !(s0 == s1 && i0 == i1)) {
                    score[lo] = s1; score[hi] = s0;
                    index[lo] = i1; index[hi] = i0;
                    const uint16_t p = pos[lo]; pos[lo] = pos[hi]; pos[hi] = p;
                }
            }
        }
    }
    __syncthreads();
}

// ---- peer-memory exchange helpers ----------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Slot of candidate i of query q (index within the whole call) of rank `rank` in an exchange buffer.
__device__ __forceinline__ size_t peer_slot(const PeerPush &p, int q, int k, int i) {
    return (size_t(p.rank) * p.n_queries_total + q) * k + i;
}
// After every CTA of the (last) launch has stored its candidates on all ranks: raise this rank's
// epoch flag on every rank.  Called by all threads of every CTA.
__device__ __forceinline__ void peer_signal(const PeerPush &p) {
    __threadfence_system();                       // this thread's peer stores are performed system-wide
    __syncthreads();
    if (threadIdx.x == 0 && p.signal) {
        const unsigned prev = atomicAdd(p.done_counter, 1u);
        if (prev == gridDim.x - 1) {              // last CTA: everyone's stores are ordered before this
            *p.done_counter = 0u;
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(p.flags[r] + p.parity * p.world + p.rank, p.epoch);
        }
    }
}

// Host-visible completion of a launch whose outputs live in mapped pinned host memory: called by
// all threads of every CTA after their output stores.
__device__ __forceinline__ void host_signal(const EmitArgs &e) {
    if (!e.done_flag) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && e.done_epoch) {
        const unsigned prev = atomicAdd(e.done_counter, 1u);
        if (prev == gridDim.x - 1) {
            *e.done_counter = 0u;
            __threadfence_system();
            st_release_sys(e.done_flag, e.done_epoch);
        }
    }
}

__global__ void __launch_bounds__(256) peer_push_kernel(const cab_candidate *__restrict__ local, int n_queries, int k, PeerPush p) {
    const int total = n_queries * k;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const cab_candidate c = local[t];
        const size_t slot = peer_slot(p, p.q0 + t / k, k, t % k);
        for (int r = 0; r < p.world; ++r) p.bufs[r][slot] = c;
    }
    peer_signal(p);
}
void launch_peer_push(const cab_candidate *local, int n_queries, int k, const PeerPush &peer, cudaStream_t s) {
    peer_push_kernel<<<1, 256, 0, s>>>(local, n_queries, k, peer);
}

// ---- finalize ---------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kFinThreads) finalize_kernel(FinalizeArgs a, EmitArgs e) {
    using TR = RowTraits<DT>;
    __shared__ uint64_t s_sort[kSortCap];
    __shared__ cab_candidate s_cand[kMaxK];
    __shared__ uint64_t s_head[kMaxHeads];
    __shared__ float s_q[kDim];
    __shared__ int s_cnt;
    __shared__ uint64_t s_bound;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x;
    // Launched with programmatic stream serialization: the CTA is already resident when the scan
    // finishes; everything the scan wrote is visible after this wait.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) { s_cnt = 0; s_bound = 0ull; if (a.work_counters) a.work_counters[qi] = 0u; }
    if (a.inl.use_query && threadIdx.x < kDim) s_q[threadIdx.x] = a.inl.q[threadIdx.x];   // kernel-argument query
    __syncthreads();

    // Sort the buffer, keep the best k, raise the bound.  Block-uniform.
    auto trim = [&]() {
        const int c = s_cnt;
        int np2 = 64;
        while (np2 < c) np2 <<= 1;
        for (int i = c + threadIdx.x; i < np2; i += kFinThreads) s_sort[i] = 0ull;
        block_sort_desc(s_sort, np2);
        if (threadIdx.x == 0) {
            const int kept = c < a.k ? c : a.k;
            s_cnt = kept;
            if (kept == a.k) s_bound = s_sort[a.k - 1];
        }
        __syncthreads();
    };

    const uint64_t *__restrict__ slots = a.partial_keys + size_t(qi) * a.n_partials * a.slot_stride;
    const int total = a.n_partials * a.slot_stride;
    const int32_t *__restrict__ counts = a.counts ? a.counts + size_t(qi) * a.n_partials : nullptr;

    // Warp-aggregated append of passing keys into s_sort; returns false if the buffer overflowed.
    auto append = [&](bool pass, uint64_t key) {
        const unsigned m = __ballot_sync(kFull, pass);
        if (m == 0) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
        base = __shfl_sync(kFull, base, 0);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pass && pos < kSortCap) s_sort[pos] = key;
    };

    // ---- fast path (small k): each scan CTA's slots are sorted, so the k-th largest list HEAD is
    // a lower bound on the global k-th best; only keys >= it can be results.  Typically ~2k keys
    // survive; they are ranked by counting (no sort, no barriers inside).
    bool done = false;
    if (a.n_partials <= kMaxHeads && !a.force_general && !counts) {
        for (int j = threadIdx.x; j < a.n_partials; j += kFinThreads) s_head[j] = slots[size_t(j) * a.slot_stride];
        __syncthreads();
        for (int j = threadIdx.x; j < a.n_partials; j += kFinThreads) {
            const uint64_t h = s_head[j];
            int r = 0;
            for (int t = 0; t < a.n_partials; ++t) r += s_head[t] > h;
            if (h != 0ull && r == a.k - 1) s_bound = h - 1;         // keys > bound <=> keys >= H_k
        }
        __syncthreads();
        const uint64_t bound = s_bound;                              // 0 if fewer than k non-empty lists
        for (int base = 0; base < total; base += kFinThreads) {
            const int i = base + threadIdx.x;
            const uint64_t key = i < total ? slots[i] : 0ull;
            append(key > bound, key);
        }
        __syncthreads();
        const int S = s_cnt;
        if (S <= kRankMax) {
            uint64_t mine[kRankMax / kFinThreads];
            int rank[kRankMax / kFinThreads];
#pragma unroll
            for (int u = 0; u < kRankMax / kFinThreads; ++u) {
                const int i = threadIdx.x + u * kFinThreads;
                mine[u] = i < S ? s_sort[i] : 0ull;
                rank[u] = 0;
            }
            for (int t = 0; t < S; ++t) {
                const uint64_t x = s_sort[t];
#pragma unroll
                for (int u = 0; u < kRankMax / kFinThreads; ++u) rank[u] += x > mine[u];
            }
            __syncthreads();                                         // all reads of s_sort done
#pragma unroll
            for (int u = 0; u < kRankMax / kFinThreads; ++u)
                if (threadIdx.x + u * kFinThreads < S && rank[u] < a.k) s_sort[rank[u]] = mine[u];
            if (threadIdx.x == 0) s_cnt = S < a.k ? S : a.k;
            __syncthreads();
            done = true;
        } else {
            __syncthreads();
            if (threadIdx.x == 0) { s_cnt = 0; s_bound = 0ull; }
            __syncthreads();
        }
    }

    // ---- counted lists (tensor-core scan): lists are mostly empty (tens of entries in 256 slots),
    // so index the VALID entries through a prefix sum of the counts instead of walking every slot.
    if (!done && counts) {
        int *s_pref = reinterpret_cast<int *>(s_head);                 // s_head is unused on this path
        const int P = a.n_partials < 2 * kMaxHeads - 1 ? a.n_partials : 2 * kMaxHeads - 1;
        for (int j = threadIdx.x; j < P; j += kFinThreads) {
            const int c = counts[j];
            s_pref[j + 1] = c < a.slot_stride ? c : a.slot_stride;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_pref[0] = 0; for (int j = 0; j < P; ++j) s_pref[j + 1] += s_pref[j]; }
        __syncthreads();
        const int T = s_pref[P];
        for (int base = 0; base < T; base += kFinThreads * kFinUnroll) {
            uint64_t key[kFinUnroll];
#pragma unroll
            for (int u = 0; u < kFinUnroll; ++u) {
                const int t = base + u * kFinThreads + threadIdx.x;
                key[u] = 0ull;
                if (t < T) {
                    int lo = 0, hi = P;                                // largest l with s_pref[l] <= t
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_pref[mid] <= t) lo = mid; else hi = mid; }
                    key[u] = slots[size_t(lo) * a.slot_stride + (t - s_pref[lo])];
                }
            }
            const uint64_t bound = s_bound;
#pragma unroll
            for (int u = 0; u < kFinUnroll; ++u) append(key[u] > bound, key[u]);
            __syncthreads();
            const bool full = s_cnt > kSortCap - kFinThreads * kFinUnroll;
            __syncthreads();
            if (full) trim();
        }
        trim();
        done = true;
    }

    // ---- general path: stream all slots (empty slots hold 0) through the buffer, sorting and
    // trimming to the best k whenever it fills ----------------------------------------------------
    if (!done) {
        for (int base = 0; base < total; base += kFinThreads * kFinUnroll) {
            uint64_t key[kFinUnroll];
#pragma unroll
            for (int u = 0; u < kFinUnroll; ++u) {
                const int i = base + u * kFinThreads + threadIdx.x;
                key[u] = i < total ? slots[i] : 0ull;
            }
            const uint64_t bound = s_bound;
#pragma unroll
            for (int u = 0; u < kFinUnroll; ++u) append(key[u] > bound, key[u]);
            __syncthreads();
            const bool full = s_cnt > kSortCap - kFinThreads * kFinUnroll;
            __syncthreads();               // everyone has read s_cnt before anyone appends again
            if (full) trim();
        }
        trim();
    }
    const int n_win = s_cnt;

    // ---- re-score the winners: one lane group per row, same operation order as the scan ---------
    float q[TR::NQ];
    const float *qsrc = a.queries + size_t(qi) * kDim;
    load_query<DT>([&](int i) { return a.inl.use_query ? s_q[i] : qsrc[i]; }, lane, q);
    const int g = lane & (TR::G - 1), sub = lane / TR::G;
    const uint4 *__restrict__ A = reinterpret_cast<const uint4 *>(a.asr);
    const uint4 *__restrict__ B = reinterpret_cast<const uint4 *>(a.audio);
    cab_candidate *out = a.cands + size_t(qi) * a.k;
    for (int i0 = warp * TR::RW; i0 < a.k; i0 += kFinWarps * TR::RW) {
        const int i = i0 + sub;
        const bool have = i < n_win;                       // uniform within a lane group
        const uint32_t row = have ? key_row(s_sort[i]) : 0u;
        float sa = 0.f, sb = 0.f;
        if (have) {
            const uint4 *pa = A + size_t(row) * TR::CPR + g;
            const uint4 *pb = B + size_t(row) * TR::CPR + g;
            uint4 ca[3], cb[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { ca[j] = pa[TR::G * j]; cb[j] = pb[TR::G * j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { sa = dot_chunk<DT>(ca[j], q, j, sa); sb = dot_chunk<DT>(cb[j], q, j, sb); }
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
            s_cand[i] = c;
            if (a.peer.world) {                            // sharded search: store on every rank over NVLink
                const size_t slot = peer_slot(a.peer, a.peer.q0 + qi, a.k, i);
                for (int r = 0; r < a.peer.world; ++r) a.peer.bufs[r][slot] = c;
            }
        }
    }
    if (a.peer.world) peer_signal(a.peer);
    if (!e.out_index) return;          // sharded search: candidates go to the exchange

    // ---- fused emit (single candidate list): rank the <= k candidates by counting ----------------
    __syncthreads();
    uint64_t *score = s_sort;                                        // reuse the selection buffer
    const double wa = e.inline_weights ? e.w64_asr : e.w_asr[qi], wb = e.inline_weights ? e.w64_audio : e.w_audio[qi];
    const int t = threadIdx.x;
    if (t < a.k) score[t] = reference_fusion(s_cand[t], wa, wb, e.threshold);
    __syncthreads();
    const size_t o0 = size_t(qi) * e.k;
    if (t < a.k) {
        const uint64_t sc = score[t];
        const int64_t gi = s_cand[t].index;
        int rank = 0, n_res = 0;
        for (int j = 0; j < a.k; ++j) {
            const uint64_t sj = score[j];
            n_res += sj != 0ull;
            rank += (sj > sc) || (sj == sc && sj != 0ull && s_cand[j].index < gi);
        }
        if (sc != 0ull) {                                            // (score desc, index asc), :685
            const cab_candidate c = s_cand[t];
            e.out_index[o0 + rank] = c.index;
            e.out_fusion[o0 + rank] = unorderable64(sc);
            e.out_asr[o0 + rank] = c.asr_sim;
            e.out_audio[o0 + rank] = c.audio_sim;
            e.out_flags[o0 + rank] = uint8_t(c.flags);
        }
        if (t >= n_res) {                                            // pad beyond the results
            e.out_index[o0 + t] = -1;
            e.out_fusion[o0 + t] = 0.0;
            e.out_asr[o0 + t] = 0.f;
            e.out_audio[o0 + t] = 0.f;
            e.out_flags[o0 + t] = 0;
        }
        if (t == 0) {
            e.out_count[qi] = n_res;
            if (qi == 0 && e.nonfinite_out) *e.nonfinite_out = *e.nonfinite;
        }
    }
    host_signal(e);
}

void launch_finalize(const FinalizeArgs &a, const EmitArgs *fused_emit, cudaStream_t s) {
    EmitArgs e{};
    if (fused_emit) e = *fused_emit;
    // Programmatic dependent launch: the finalize grid is staged while the scan still runs and
    // starts the moment it completes (the kernel begins with griddepcontrol.wait).
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.n_queries); cfg.blockDim = dim3(kFinThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (a.dtype == CAB_BF16) cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_BF16>, a, e);
    else cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_F32>, a, e);
}

// ---- emit (merge of several candidate lists: sharded search) ---------------------------------------
constexpr int kEmitThreads = 256;
constexpr int kEmitMax = 1024;      // >= world(8) x CAB_MAX_K(128)

__global__ void __launch_bounds__(kEmitThreads) emit_kernel(EmitArgs a) {
    __shared__ uint64_t s_score[kEmitMax];     // orderable float64 fusion score, 0 = not a result
    __shared__ int64_t s_index[kEmitMax];
    __shared__ uint16_t s_pos[kEmitMax];
    __shared__ int s_n;

    const int qi = blockIdx.x;
    const int n_cand = a.n_lists * a.k;
    int np2 = 64;
    while (np2 < n_cand) np2 <<= 1;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // programmatic dependent launch (see launch_emit)
    const double wa = a.inline_weights ? a.w64_asr : a.w_asr[qi], wb = a.inline_weights ? a.w64_audio : a.w_audio[qi];
    if (threadIdx.x == 0) s_n = 0;
    if (a.wait_flags) {
        // peer exchange: every rank's candidates of this epoch must have landed in our buffer
        if (threadIdx.x < a.n_lists) {
            const long long t0 = clock64();
            while (ld_acquire_sys(a.wait_flags + threadIdx.x) != a.wait_epoch) {
                if (clock64() - t0 > 20000000000ll) {          // ~10 s: a rank never arrived
                    if (a.status) atomicExch(a.status, 7);
                    __threadfence_system();
                    asm volatile("trap;");
                }
            }
        }
        __syncthreads();
    }
    // candidate t of this query lives at list-major position:
    auto cand_at = [&](int t) -> const cab_candidate * {
        const int list = t / a.k, i = t - list * a.k;
        return a.cands + (size_t(list) * a.n_queries + qi) * a.k + i;
    };
    for (int t = threadIdx.x; t < np2; t += kEmitThreads) {
        uint64_t sc = 0ull;
        int64_t gi = INT64_MAX;
        if (t < n_cand) {
            const cab_candidate c = *cand_at(t);
            sc = reference_fusion(c, wa, wb, a.threshold);
            if (sc) gi = c.index;
        }
        s_score[t] = sc; s_index[t] = gi; s_pos[t] = uint16_t(t);
    }
    if (n_cand <= kEmitThreads) {
        // Few candidates (e.g. 8 shards x top-10): rank by counting instead of sorting -- no barriers.
        __syncthreads();
        const int t = threadIdx.x;
        int rank = 0, n_res = 0;
        const uint64_t sc = t < n_cand ? s_score[t] : 0ull;
        const int64_t gi = t < n_cand ? s_index[t] : INT64_MAX;
        for (int j = 0; j < n_cand; ++j) {
            const uint64_t sj = s_score[j];
            n_res += sj != 0ull;
            rank += (sj > sc) || (sj == sc && sj != 0ull && s_index[j] < gi);
        }
        const size_t o0 = size_t(qi) * a.k;
        if (sc != 0ull && rank < a.k) {
            const cab_candidate c = *cand_at(t);
            a.out_index[o0 + rank] = c.index;
            a.out_fusion[o0 + rank] = unorderable64(sc);
            a.out_asr[o0 + rank] = c.asr_sim;
            a.out_audio[o0 + rank] = c.audio_sim;
            a.out_flags[o0 + rank] = uint8_t(c.flags);
        }
        const int n_out = n_res < a.k ? n_res : a.k;
        for (int i = n_out + t; i < a.k; i += kEmitThreads) {
            a.out_index[o0 + i] = -1;
            a.out_fusion[o0 + i] = 0.0;
            a.out_asr[o0 + i] = 0.f;
            a.out_audio[o0 + i] = 0.f;
            a.out_flags[o0 + i] = 0;
        }
        if (t == 0) {
            a.out_count[qi] = n_out;
            if (qi == 0 && a.nonfinite_out) { *a.nonfinite_out = *a.nonfinite; }
        }
        host_signal(a);
        return;
    }
    block_sort_results(s_score, s_index, s_pos, np2);
    for (int i = threadIdx.x; i < a.k; i += kEmitThreads) {
        const size_t o = size_t(qi) * a.k + i;
        if (s_score[i] != 0ull) {
            const cab_candidate c = *cand_at(s_pos[i]);
            a.out_index[o] = c.index;
            a.out_fusion[o] = unorderable64(s_score[i]);
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
    if (threadIdx.x == 0) {
        a.out_count[qi] = s_n;
        if (qi == 0 && a.nonfinite_out) { *a.nonfinite_out = *a.nonfinite; }
    }
    host_signal(a);
}

void launch_emit(const EmitArgs &a, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.n_queries); cfg.blockDim = dim3(kEmitThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, emit_kernel, a);
}

}  // namespace cab
