// After the scan: (1) finalize -- merge the per-CTA partial top-k slots of one query into its
// best k rows and re-score those rows (both cosines, fp32) -> cab_candidate records;
// (2) emit -- from 1..world candidate lists per query compute the reference's float64 fusion
// score (audio_search.py:656-670), apply the strict float64 threshold (:672), order by
// (score desc, global index asc) (:685) and write the first k (:699).
// On a single GPU (one candidate list) emit runs inside the finalize kernel.  Sharded search
// (corpus split over ranks): the finalize kernel stores its candidates straight into every rank's
// exchange buffer over NVLink peer memory, raises its epoch flag on every rank, waits for the
// world's flags and merges -- ONE kernel per search after the scan.  Before it touches the
// exchange it releases its dependents (griddepcontrol.launch_dependents), so the next search's
// scan streams the corpus while this search's exchange and merge are still in flight.
//
// Both touch O(k) rows per query; they are latency-, not bandwidth-bound, and run as one CTA per
// query: every global load is issued in parallel (no dependent chains), all selection happens in
// shared memory.
#include "cab_internal.h"
#include "cab_rowdot.cuh"

namespace cab {

// One CTA per query.  1024 threads (one CTA per SM) when the queries fit one wave: the shortest
// latency for a single search.  512 threads (two CTAs per SM) for bigger batches: 256 queries run as
// one wave instead of two, and their latency-bound phases overlap on each SM.
constexpr int kFinThreadsMax = 1024;
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
// A candidate record read from L2 (ld.global.cg), never from L1.  Exchange buffers are written by
// peers over NVLink while the reading kernel is already resident (launched early by programmatic
// dependent launch, then spinning on the peers' flags).  The acquire load of a flag is followed
// by an L1 invalidation (CCTL.IVALL in the SASS), which by itself covers the plain loads after the
// barrier; bypassing L1 for the records makes the merge independent of that detail and keeps
// 24 KB of use-once data out of L1.  24-byte records are 8-byte aligned: three 64-bit loads.
__device__ __forceinline__ cab_candidate load_candidate_l2(const cab_candidate *p) {
    const long long *p64 = reinterpret_cast<const long long *>(p);
    const long long w0 = __ldcg(p64), w1 = __ldcg(p64 + 1), w2 = __ldcg(p64 + 2);
    cab_candidate c;
    c.index = w0;
    c.asr_sim = __int_as_float(int(uint64_t(w1) & 0xFFFFFFFFull)); c.audio_sim = __int_as_float(int(uint64_t(w1) >> 32));
    c.flags = uint32_t(uint64_t(w2) & 0xFFFFFFFFull); c.pad = uint32_t(uint64_t(w2) >> 32);
    return c;
}
// Slot of candidate i of query q (index within the whole call) of rank `rank` in an exchange buffer.
__device__ __forceinline__ size_t peer_slot(const PeerPush &p, int q, int k, int i) {
    return (size_t(p.rank) * p.n_queries_total + q) * k + i;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Store `n` candidates (shared memory) of query q into every rank's exchange buffer and, once
// every CTA of the (last) launch has done so, raise this rank's epoch flag on every rank.
// Thread t stores candidate t % k on rank t / k: all world x k records leave in parallel.  The
// stores are ordered before the flag by the CTA barrier + ONE system-scope fence + release store in
// the signalling threads (fences are cumulative over everything that happens before them), so the
// NVLink round trip is paid once, not once per storing thread and again for the flag.
// Called by all threads of the CTA.
__device__ __forceinline__ void peer_push_and_signal(const PeerPush &p, const cab_candidate *cand, int q, int k,
                                                     int *s_last) {
    for (int t = threadIdx.x; t < p.world * k; t += blockDim.x) {
        const int r = t / k, i = t - r * k;
        p.bufs[r][peer_slot(p, q, k, i)] = cand[i];
    }
    __syncthreads();
    if (!p.signal) {                              // an earlier batch of the call: its stores must be performed
        if (threadIdx.x == 0) __threadfence_system();   // system-wide before this grid completes
        return;
    }
    if (gridDim.x > 1) {                          // several queries: the last CTA to finish raises the flags
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(p.done_counter, 1u);
            *s_last = prev == gridDim.x - 1;
            if (*s_last) *p.done_counter = 0u;
            __threadfence();                      // acquire side: the other CTAs' (fenced) stores precede their count
        }
        __syncthreads();
        if (!*s_last) return;
    }
    // st.release.sys = system-scope fence (cumulative over the CTA's stores, which the barrier above
    // ordered before it) + store: one NVLink round trip
    if (threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.parity * p.world + p.rank, p.epoch);
}

// Host-visible completion of a launch whose outputs live in mapped pinned host memory: called by
// all threads of every CTA after their output stores.
__device__ __forceinline__ void host_signal(const EmitArgs &e) {
    if (!e.done_flag) return;
    __syncthreads();                              // the CTA's output stores happen before thread 0's fence
    if (threadIdx.x != 0) return;
    __threadfence_system();                       // cumulative: performs them at system scope (PCIe)
    if (!e.done_epoch) return;                    // an earlier batch of the call does not signal
    if (gridDim.x > 1) {
        const unsigned prev = atomicAdd(e.done_counter, 1u);
        if (prev != gridDim.x - 1) return;
        *e.done_counter = 0u;
        __threadfence_system();                   // acquire side of the other CTAs' counts
    }
    st_release_sys(e.done_flag, e.done_epoch);
}

// Wait until the epoch flags of all `n_lists` ranks hold `epoch` (their candidates of this search
// have landed in our exchange buffer).  Bounded: a rank that never arrives is an error, not a hang.
__device__ __forceinline__ void wait_peer_flags(const uint32_t *flags, int n_lists, uint32_t epoch, int *status) {
    if (threadIdx.x < n_lists) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) != epoch) {
            if (clock64() - t0 > 20000000000ll) {              // ~10 s
                if (status) atomicExch(status, 7);
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
    __syncthreads();
}

// Used when there was nothing to scan (empty shard): push a prepared candidate block.
__global__ void __launch_bounds__(256) peer_push_kernel(const cab_candidate *__restrict__ local, int n_queries, int k, PeerPush p) {
    const int total = n_queries * k;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const cab_candidate c = local[t];
        const size_t slot = peer_slot(p, p.q0 + t / k, k, t % k);
        for (int r = 0; r < p.world; ++r) p.bufs[r][slot] = c;
    }
    __threadfence_system();
    __syncthreads();
    if (p.signal && threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.parity * p.world + p.rank, p.epoch);
}
void launch_peer_push(const cab_candidate *local, int n_queries, int k, const PeerPush &peer, cudaStream_t s) {
    peer_push_kernel<<<1, 256, 0, s>>>(local, n_queries, k, peer);
}

// Sharded search of an fp32 library through its bf16 shadows: the shard's EXACT top-k was produced by
// the certified single-index route as ordinary result arrays (device); turn query blockIdx.x's row
// into candidate records and push them like a finalize kernel would.
__global__ void __launch_bounds__(128) pack_push_kernel(const int64_t *__restrict__ index, const float *__restrict__ asr,
                                                        const float *__restrict__ audio, const uint8_t *__restrict__ flags,
                                                        const int32_t *__restrict__ count, int k, PeerPush p) {
    __shared__ cab_candidate s_cand[kMaxK];
    __shared__ int s_last;
    const int qi = blockIdx.x;
    const int c = count[qi];                                  // -1: the query holds NaN/Inf
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool have = i < c;
        const size_t o = size_t(qi) * k + i;
        cab_candidate cd;
        cd.index = have ? index[o] : (c < 0 ? kBadQueryIndex : int64_t(-1));
        cd.asr_sim = have ? asr[o] : 0.f;
        cd.audio_sim = have ? audio[o] : 0.f;
        cd.flags = have ? uint32_t(flags[o]) : 0u;
        cd.pad = 0u;
        s_cand[i] = cd;
    }
    __syncthreads();
    peer_push_and_signal(p, s_cand, p.q0 + qi, k, &s_last);
}
void launch_pack_push(const int64_t *index, const float *asr, const float *audio, const uint8_t *flags, const int32_t *count,
                      int n_queries, int k, const PeerPush &peer, cudaStream_t s) {
    pack_push_kernel<<<n_queries, 128, 0, s>>>(index, asr, audio, flags, count, k, peer);
}

// ---- emit: rank one query's candidates and write the first k ------------------------------------------
// `cand_at(t)` returns candidate t of n_cand (<= blockDim.x, <= kEmitMax).  score/index/pos are
// shared arrays of kEmitMax entries.  Order: (float64 fusion desc, global index asc) = Python's
// stable descending sort (:685).  Few candidates are ranked by counting (no barriers), many by a
// block bitonic sort.  Called by all threads of the CTA.
constexpr int kEmitMax = 1024;      // >= world(8) x CAB_MAX_K(128)
constexpr int kEmitCountMax = 256;

template <typename CandAt>
__device__ __forceinline__ void emit_ranked(CandAt cand_at, int n_cand, int qi, const EmitArgs &e, bool bad_query,
                                            uint64_t *score, int64_t *index, uint16_t *pos) {
    const double wa = e.inline_weights ? e.w64_asr : e.w_asr[qi], wb = e.inline_weights ? e.w64_audio : e.w_audio[qi];
    const int t = threadIdx.x;
    const size_t o0 = size_t(qi) * e.k;
    auto write = [&](int slot, const cab_candidate &c, uint64_t sc) {
        e.out_index[o0 + slot] = c.index;
        e.out_fusion[o0 + slot] = unorderable64(sc);
        e.out_asr[o0 + slot] = c.asr_sim;
        e.out_audio[o0 + slot] = c.audio_sim;
        e.out_flags[o0 + slot] = uint8_t(c.flags);
    };
    cab_candidate c{};
    uint64_t sc = 0ull;
    int64_t gi = INT64_MAX;
    if (t < n_cand) {
        c = cand_at(t);
        bad_query |= (t == 0 && c.index == kBadQueryIndex);
        sc = reference_fusion(c, wa, wb, e.threshold);
        if (sc) gi = c.index;
    }
    bad_query = __syncthreads_or(bad_query);
    if (bad_query) sc = 0ull;
    int n_res;
    if (n_cand <= kEmitCountMax) {
        if (t < n_cand) { score[t] = sc; index[t] = gi; }
        n_res = __syncthreads_count(sc != 0ull);
        if (sc != 0ull) {
            int rank = 0;
            for (int j = 0; j < n_cand; ++j) {
                const uint64_t sj = score[j];
                rank += (sj > sc) || (sj == sc && index[j] < gi);
            }
            if (rank < e.k) write(rank, c, sc);
        }
        if (n_res > e.k) n_res = e.k;
    } else {
        int np2 = 512;
        while (np2 < n_cand) np2 <<= 1;
        for (int i = t; i < np2; i += blockDim.x) {
            score[i] = i == t ? sc : 0ull; index[i] = i == t ? gi : INT64_MAX; pos[i] = uint16_t(i);
        }
        block_sort_results(score, index, pos, np2);
        n_res = __syncthreads_count(t < e.k && score[t] != 0ull);
        if (t < e.k && score[t] != 0ull) write(t, cand_at(pos[t]), score[t]);
    }
    for (int i = n_res + t; i < e.k; i += blockDim.x) {                 // pad beyond the results
        e.out_index[o0 + i] = -1;
        e.out_fusion[o0 + i] = 0.0;
        e.out_asr[o0 + i] = 0.f;
        e.out_audio[o0 + i] = 0.f;
        e.out_flags[o0 + i] = 0;
    }
    if (t == 0) {
        e.out_count[qi] = bad_query ? -1 : n_res;
        if (bad_query && e.nonfinite_out) *e.nonfinite_out = 1;
    }
}

// ---- finalize ---------------------------------------------------------------------------------
template <int DT, int kFinThreads>
__global__ void __launch_bounds__(kFinThreads, kFinThreadsMax / kFinThreads) finalize_kernel(FinalizeArgs a, EmitArgs e) {
    using TR = RowTraits<DT>;
    constexpr int kFinWarps = kFinThreads / 32;
    __shared__ uint64_t s_sort[kSortCap];
    __shared__ cab_candidate s_cand[kMaxK];
    __shared__ uint64_t s_head[kMaxHeads];
    __shared__ float s_q[kDim];
    __shared__ int s_cnt;
    __shared__ int s_last;
    __shared__ uint64_t s_bound;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x;
    // Launched with programmatic stream serialization: the CTA is already resident when the scan
    // finishes; everything the scan wrote is visible after this wait.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool stamp = e.stamps && qi == 0 && threadIdx.x == 0;
    if (stamp) e.stamps[0] = globaltimer_ns();
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

    // Warp-aggregated append of passing keys into s_sort (keys beyond the buffer are dropped; every
    // caller either bounds the number of survivors beforehand or checks s_cnt afterwards).
    auto append = [&](bool pass, uint64_t key) {
        const unsigned m = __ballot_sync(kFull, pass);
        if (m == 0) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
        base = __shfl_sync(kFull, base, 0);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pass && pos < kSortCap) s_sort[pos] = key;
    };
    // Rank the S <= kRankMax keys in s_sort by counting and leave the best k sorted in s_sort[0..k).
    auto rank_by_counting = [&](int S) {
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
        for (int base = 0; base < total; base += kFinThreads * 4) {  // 4 independent loads in flight per thread
            uint64_t key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * kFinThreads + threadIdx.x;
                key[u] = i < total ? slots[i] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) append(key[u] > bound, key[u]);
        }
        __syncthreads();
        const int S = s_cnt;
        if (S <= kRankMax) {
            rank_by_counting(S);
            done = true;
        } else {
            __syncthreads();
            if (threadIdx.x == 0) { s_cnt = 0; s_bound = 0ull; }
            __syncthreads();
        }
    }

    // ---- counted lists (tensor-core scan), fast path: the scan's per-query score-level histogram
    // gives a bound -- the lower edge of the highest level L at or above which >= k candidates were
    // pushed -- that at most suffix(L) keys exceed.  If that fits the ranking buffer, each warp
    // walks whole lists (coalesced, only the valid entries) and the survivors are ranked by
    // counting: no sort, two barriers.
    if (!done && counts && a.levels && !a.force_general) {
        int *s_lv = reinterpret_cast<int *>(s_head);                   // s_head is unused on this path
        if (threadIdx.x < 64) s_lv[threadIdx.x] = a.levels[size_t(qi) * 64 + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            int suffix = 0, lvl = 0, above = 0;
            for (int j = 63; j >= 0; --j) { suffix += s_lv[j]; if (suffix >= a.k) { lvl = j; break; } }
            above = suffix;                                            // candidates ever pushed at levels >= lvl
            if (suffix < a.k) { lvl = -1; }                            // fewer than k candidates at all: keep everything
            s_lv[64] = above;
            s_bound = lvl > 0 ? bound_key(a.select_threshold + float(lvl) * a.level_step - 2e-6f) : 0ull;
        }
        __syncthreads();
        if (s_lv[64] <= kRankMax) {
            const uint64_t bound = s_bound;
            for (int l = warp; l < a.n_partials; l += kFinWarps) {
                int c = counts[l];
                c = c < a.slot_stride ? c : a.slot_stride;
                const uint64_t *lp = slots + size_t(l) * a.slot_stride;
                for (int i = lane; i < ((c + 31) & ~31); i += 32) {
                    const uint64_t key = i < c ? lp[i] : 0ull;
                    append(key > bound, key);
                }
            }
            __syncthreads();
            const int S = s_cnt;
            if (S <= kRankMax) {                                       // always (S <= suffix(L)); checked anyway
                rank_by_counting(S);
                done = true;
            } else {
                __syncthreads();
                if (threadIdx.x == 0) { s_cnt = 0; s_bound = 0ull; }
                __syncthreads();
            }
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_bound = 0ull;
            __syncthreads();
        }
    }

    // ---- counted lists, general path: index the VALID entries through a prefix sum of the counts,
    // stream them through the buffer, sort + trim whenever it fills ---------------------------------
    if (!done && counts) {
        int *s_pref = reinterpret_cast<int *>(s_head);
        const int P = a.n_partials < 2 * kMaxHeads - 1 ? a.n_partials : 2 * kMaxHeads - 1;
        for (int j = threadIdx.x; j < P; j += kFinThreads) {
            const int c = counts[j];
            s_pref[j + 1] = c < a.slot_stride ? c : a.slot_stride;
        }
        __syncthreads();
        if (warp == 0) {                                               // inclusive scan, 32 lists per step
            int carry = 0;
            for (int j0 = 0; j0 < P; j0 += 32) {
                int v = j0 + lane < P ? s_pref[j0 + lane + 1] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(kFull, v, o); if (lane >= o) v += n; }
                if (j0 + lane < P) s_pref[j0 + lane + 1] = v + carry;
                carry += __shfl_sync(kFull, v, 31);
            }
            if (lane == 0) s_pref[0] = 0;
        }
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
    int n_win = s_cnt;
    // scan score below which rows were NOT selected (certificate of the shadow path, see FinalizeArgs)
    const float cert_cutoff = (a.cert_out && n_win == a.k) ? key_score(s_sort[a.k - 1]) : a.select_threshold;
    if (stamp) e.stamps[1] = globaltimer_ns();

    // Everything the next scan on this stream touches -- its ticket counters (reset above), the
    // partial-key slots, counts and levels (all consumed into shared memory by now) -- is released:
    // let the dependent grid start while this one re-scores, exchanges and merges.
    __threadfence();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- re-score the winners: one lane group per row, same operation order as the scan ---------
    float q[TR::NQ];
    const float *qsrc = a.queries + size_t(qi) * kDim;
    const bool finite = load_query<DT>([&](int i) { return a.inl.use_query ? s_q[i] : qsrc[i]; }, lane, q, a.norm_asr != nullptr);
    if (!finite) n_win = 0;                                // NaN/Inf query: no results, reported below
    const int g = lane & (TR::G - 1), sub = lane / TR::G;
    const uint4 *__restrict__ A = reinterpret_cast<const uint4 *>(a.asr);
    const uint4 *__restrict__ B = reinterpret_cast<const uint4 *>(a.audio);
    cab_candidate *out = a.cands ? a.cands + size_t(qi) * a.k : nullptr;
    for (int i0 = warp * TR::RW; i0 < a.k; i0 += kFinWarps * TR::RW) {
        const int i = i0 + sub;
        const bool have = i < n_win;                       // uniform within a lane group
        const uint32_t row = have ? key_row(s_sort[i]) : 0u;
        float sa = 0.f, sb = 0.f, len_a = 1.f, len_b = 1.f;
        uint32_t row_flags = 0u;
        if (have) {
            const uint4 *pa = A + size_t(row) * TR::CPR + g;
            const uint4 *pb = B + size_t(row) * TR::CPR + g;
            uint4 ca[3], cb[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { ca[j] = pa[TR::G * j]; cb[j] = pb[TR::G * j]; }
            row_flags = a.flags[row];                      // in flight together with the row (one round trip, not two)
            if (a.norm_asr) { len_a = a.norm_asr[row]; len_b = a.norm_audio[row]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { sa = dot_chunk<DT>(ca[j], q, j, sa); sb = dot_chunk<DT>(cb[j], q, j, sb); }
        }
        sa = group_sum<DT>(sa);
        sb = group_sum<DT>(sb);
        sa *= len_a; sb *= len_b;                          // raw dot products (option raw_dot), else x 1
        if (g == 0 && i < a.k) {
            cab_candidate c;
            c.index = have ? a.row_base + int64_t(row) : (finite ? int64_t(-1) : kBadQueryIndex);
            c.asr_sim = sa; c.audio_sim = sb;
            c.flags = row_flags;
            c.pad = 0u;
            if (out) out[i] = c;
            s_cand[i] = c;
        }
    }
    __syncthreads();
    if (stamp) e.stamps[2] = globaltimer_ns();
    uint64_t *score = s_sort;                                        // the selection buffer is free now
    int64_t *index = reinterpret_cast<int64_t *>(s_sort + kEmitMax);
    uint16_t *pos = reinterpret_cast<uint16_t *>(s_sort + 2 * kEmitMax);
    if (a.peer.world) {
        // sharded search: this shard's candidates go to every rank over NVLink peer memory
        peer_push_and_signal(a.peer, s_cand, a.peer.q0 + qi, a.k, &s_last);
        if (stamp) e.stamps[3] = globaltimer_ns();
        if (!e.out_index) return;                                    // merged by a separate emit launch
        wait_peer_flags(e.wait_flags, e.n_lists, e.wait_epoch, e.status);
        if (stamp) e.stamps[4] = globaltimer_ns();
        const cab_candidate *lists = e.cands;                        // [n_lists][n_queries][k], this rank's buffer
        const int nq = e.n_queries, k = e.k;
        emit_ranked([&](int t) {
            const int list = t / k, i = t - list * k;
            return load_candidate_l2(lists + (size_t(list) * nq + qi) * k + i);
        }, e.n_lists * k, qi, e, !finite, score, index, pos);
        host_signal(e);
        if (stamp) e.stamps[5] = globaltimer_ns();
        return;
    }
    if (!e.out_index) return;          // candidates only (cab_search_candidates)

    // ---- fused emit (single candidate list) ------------------------------------------------------------
    emit_ranked([&](int t) { return s_cand[t]; }, a.k, qi, e, !finite, score, index, pos);
    if (a.cert_out) {                  // a.k <= 256: emit_ranked left score[t] = orderable float64 fusion of candidate t
        const double line = double(cert_cutoff) + double(a.cert_eps);
        const int t = threadIdx.x;
        const int n_safe = __syncthreads_count(t < a.k && score[t] != 0ull && unorderable64(score[t]) > line);
        if (t == 0) a.cert_out[qi] = (n_safe >= e.k || line <= e.threshold || !finite) ? 1 : 0;
    }
    host_signal(e);
}

void launch_finalize(const FinalizeArgs &a, const EmitArgs *fused_emit, int sm_count, cudaStream_t s) {
    EmitArgs e{};
    if (fused_emit) e = *fused_emit;
    // Programmatic dependent launch: the finalize grid is staged while the scan still runs and
    // starts the moment it completes (the kernel begins with griddepcontrol.wait).
    cudaLaunchConfig_t cfg{};
    // a sharded search whose merge rides in this kernel spins on the peers' flags: keep its CTAs at
    // one per SM (all co-resident, no second CTA starved behind a spinning one)
    const bool wide = a.n_queries <= sm_count || (a.peer.world && e.out_index);
    cfg.gridDim = dim3(a.n_queries); cfg.blockDim = dim3(wide ? kFinThreadsMax : kFinThreadsMax / 2); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (a.dtype == CAB_BF16) {
        if (wide) cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_BF16, kFinThreadsMax>, a, e);
        else cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_BF16, kFinThreadsMax / 2>, a, e);
    } else {
        if (wide) cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_F32, kFinThreadsMax>, a, e);
        else cudaLaunchKernelEx(&cfg, finalize_kernel<CAB_F32, kFinThreadsMax / 2>, a, e);
    }
}

// ---- emit as its own launch: merge of several candidate lists (cab_merge_candidates, and sharded
// searches whose merge cannot ride in the finalize kernel) -----------------------------------------------
__global__ void __launch_bounds__(kEmitMax) emit_kernel(EmitArgs a) {
    __shared__ uint64_t s_score[kEmitMax];     // orderable float64 fusion score, 0 = not a result
    __shared__ int64_t s_index[kEmitMax];
    __shared__ uint16_t s_pos[kEmitMax];

    const int qi = blockIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // programmatic dependent launch (see launch_emit)
    if (a.wait_flags) wait_peer_flags(a.wait_flags, a.n_lists, a.wait_epoch, a.status);
    const int nq = a.n_queries, k = a.k;
    const cab_candidate *lists = a.cands;
    emit_ranked([&](int t) {
        const int list = t / k, i = t - list * k;
        return load_candidate_l2(lists + (size_t(list) * nq + qi) * k + i);
    }, a.n_lists * k, qi, a, false, s_score, s_index, s_pos);
    host_signal(a);
}

void launch_emit(const EmitArgs &a, cudaStream_t s) {
    const int n_cand = a.n_lists * a.k;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.n_queries); cfg.blockDim = dim3(n_cand <= 256 ? 256 : kEmitMax); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, emit_kernel, a);
}

}  // namespace cab
