// Host-side internal API between the C-ABI (cab_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/cab.h"

namespace cab {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a property of (kernel, DEVICE): one of these
// per kernel remembers on which device ordinals it has been set (a process may hold indices on
// several GPUs, cab.h "Conventions").  Thread-safe; ordinals >= 64 simply set it on every launch.
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool done(int dev) const { return dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
    void mark(int dev) { if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

// ---- ingest (cab_ingest.cu) --------------------------------------------------------------------
// Normalise n raw fp32 rows (sklearn `normalize` semantics) into the store at row offset `dst_row`.
// src == nullptr writes zero rows.  *nonfinite (device int) is set to 1 if any value is NaN/Inf.
// norms[dst_row + i] receives the length of raw row i (0 for a zero / missing row).
// shadow (fp32 stores only, may be null): a bf16 copy of the normalised rows, same row offsets.
void launch_normalize_rows(const float *src, void *dst, int dtype, int64_t dst_row, int64_t n,
                           int *nonfinite, float *norms, void *shadow, cudaStream_t s);
// bf16 shadow of rows [r0, r0 + n) of an fp32 store that is already normalised.
void launch_shadow_rows(const float *src, void *dst, int64_t r0, int64_t n, cudaStream_t s);
// Deterministic synthetic rows (synth.py) for global rows [r0, r1) of stream 0/1, raw fp32.
struct SynthParams {
    uint32_t seed;
    uint64_t n_total;
    uint32_t n_plants_total;   // n_queries * plants
    uint32_t plants;           // per query
    uint64_t base, inv_stride;
    int mode;                  // 0 planted, 1 ascending, 2 clustered (synth.MODES)
};
void launch_synth_rows(const SynthParams &p, int stream_id, int partial, int64_t r0, int64_t n,
                       float *out, cudaStream_t s);
void launch_synth_flags(uint32_t seed, int partial, int64_t r0, int64_t n, uint8_t *out,
                        cudaStream_t s);
void launch_widen_rows(const void *src, int dtype, int64_t r0, int64_t n, float *out,
                       cudaStream_t s);

// Single-query searches carry their parameters in the kernel arguments (no H2D copy ahead of the
// scan): the weights always, the raw query vector too when it comes from host memory.
struct InlineParams {
    int use_weights;           // wa32/wb32/w64 below are valid (n_queries == 1)
    int use_query;             // q below is valid
    float wa32, wb32;
    double w64_asr, w64_audio;
    float q[CAB_DIM];
};

// ---- scan (cab_gemv.cu) ------------------------------------------------------------------------
struct ScanArgs {
    const void *asr;           // [n_rows x 384] normalised rows, fp32 or bf16
    const void *audio;
    const uint8_t *flags;      // [n_rows]
    // Raw dot-product scoring (option "raw_dot"; clean_audio_search.py:306 ranks by np.dot, not by
    // cosine): non-null = the stored rows' original lengths; a similarity is then
    // <q_raw, row_hat> * |row| = <q_raw, row> and the query is NOT normalised.  Null = cosine.
    const float *norm_asr, *norm_audio;
    int64_t n_rows;
    int dtype;
    const float *queries;      // device, raw fp32 [n_queries x 384]
    const float *wa32;         // device [n_queries] w_asr/(w_asr+w_audio)
    const float *wb32;
    int n_queries;
    int k;
    float select_threshold;    // fp32 bound used while scanning (slightly below the fp64 one)
    // outputs: per (query, scan CTA) the CTA's best keys, unused slots = 0
    uint64_t *partial_keys;    // [n_queries][n_partials][k]            (GEMV)
                               // [n_queries][n_partials][kGemmListCap] (GEMM, with counts)
    int32_t *partial_counts;   // GEMM only: [n_queries][n_partials]
    int n_partials;            // == grid size of the scan (GEMV) / number of CTA pairs (GEMM)
    unsigned int *work_counters;   // GEMV: [n_queries] chunk tickets, zero on entry (finalize resets them)
    int chunk_rows;                // GEMV: rows per dynamically scheduled chunk (request; 0 = default)
    // GEMV chunk schedule, filled in by launch_gemv_scan (see ChunkLayout in cab_gemv.cu)
    int rows_big, rows_small;      // rows per big / tail chunk
    int64_t n_big, n_chunks, tail_row0;
    uint32_t n_super, super_mul;   // visiting order of the 64-chunk groups: group g is read at (g * super_mul) % n_super
    // Where the scan executes griddepcontrol.wait (it is launched with programmatic stream
    // serialization).  1: before its first global read -- the device query / staged weights may have
    // been written by the kernel right in front of it on the caller's stream.  0: only before it
    // publishes its partial lists: everything it touches earlier (ticket counters, its own shared
    // memory) was released by the preceding finalize kernel BEFORE that kernel's
    // griddepcontrol.launch_dependents, so the scan of search i+1 streams the corpus while search
    // i's finalize / exchange / merge is still in flight.
    int wait_early;
    InlineParams inl;
};
struct GemvConfig {
    int variant;               // 0 = LDG register pipeline (the only one built)
    int blocks_per_sm;         // 0 = default
    int unroll;                // 0 = default
    int query_tile;            // max queries scored per corpus pass: 0 = default (fp32 4, bf16 2), 1, 2, 4
};
// How one GEMV launch over `n_queries` queries is shaped.
struct GemvPlan {
    int u, mb, qt;             // row-steps in flight, CTAs per SM, queries per pass
    int grid_x;                // CTAs per query group == partial lists per query
    int groups;                // ceil(n_queries / qt)
};
GemvPlan plan_gemv(const GemvConfig &cfg, int dtype, int n_queries, int sm_count);
int gemv_max_grid(int sm_count);       // upper bound of grid_x over all plans (buffer sizing)
void launch_gemv_scan(const ScanArgs &a, const GemvPlan &plan, cudaStream_t s);

// ---- legacy modes: all N fused scores, per-row weight class (cab_score_all.cu) --------------------
struct ScoreAllArgs {
    const void *asr;
    const void *audio;
    const uint8_t *flags;      // bits 2-3 of a row's byte: its weight class
    int64_t n_rows;
    int dtype;
    const float *query;        // device, raw fp32 [n_q x 384] (unused when use_inline_query)
    int n_q;                   // queries scored by this launch (1..4): the corpus is read once for all of them
    int use_inline_query;      // n_q == 1 only
    float q[CAB_DIM];          // host query, carried in the kernel arguments
    float class_w[4][4][2];    // per query of the launch: {w_asr, w_audio} per weight class
    float *out;                // device: query t's scores at out + t * out_stride; all NaN if it holds NaN/Inf
    int64_t out_stride;
    int *nonfinite;            // also set in that case (may be null)
    unsigned int *work_counters;   // [2] chunk tickets + finished warps; zero on entry, re-armed by the kernel
    int chunk_rows;                // rows per dynamically scheduled chunk
};
void launch_score_all(const ScoreAllArgs &a, int sm_count, cudaStream_t s);

// ---- tensor-core scan (cab_gemm_tc.cu) -----------------------------------------------------------
constexpr int kGemmListCap = 256;       // slots per (CTA pair, query) candidate list
constexpr int kGemmQueriesPerPass = 256;

// ---- finalize + emit (cab_finalize.cu) ----------------------------------------------------------
// Peer-memory exchange (sharded search): where the finalize kernel also stores this rank's
// candidates, and the epoch flag it raises on every rank when all of them are written.
struct PeerPush {
    int world;                       // 0 = no peer exchange
    int rank;
    int parity;                      // exchange buffers are double-buffered by epoch parity
    int signal;                      // raise the flags at the end of this launch
    int q0, n_queries_total;         // this launch handles queries [q0, q0 + n_queries) of the call
    uint32_t epoch;
    cab_candidate *bufs[CAB_MAX_WORLD];   // rank r's buffer: [2][world][n_queries_total][k]
    uint32_t *flags[CAB_MAX_WORLD];       // rank r's flags:  [2][world]
    unsigned int *done_counter;      // local: CTAs of this launch that finished pushing
};
struct FinalizeArgs {
    const void *asr;
    const void *audio;
    const uint8_t *flags;
    const float *norm_asr, *norm_audio;   // raw dot-product scoring, see ScanArgs
    int dtype;
    int64_t row_base;          // global index of local row 0
    const float *queries;
    int n_queries;
    int k;
    // Partial lists of query q: n_partials lists of slot_stride slots at
    //   partial_keys + (q * n_partials + list) * slot_stride.
    // counts == nullptr: lists are sorted descending and unused slots hold 0 (GEMV scan);
    // counts != nullptr: list holds counts[q * n_partials + list] unsorted keys (GEMM scan).
    const uint64_t *partial_keys;
    const int32_t *counts;
    int n_partials;
    int slot_stride;
    cab_candidate *cands;      // out [n_queries][k], best-first by scan score, index -1 = empty
    int force_general;         // test hook: skip the head-bound fast path
    unsigned int *work_counters;   // [n_queries] reset to 0 for the next scan (may be null)
    // tensor-core scan only: the per-query score-level histogram it left behind ([n_queries][64],
    // level j = fused score in [select_threshold + j*step, +step)); gives the finalize a bound that
    // at most (k + one level's population) keys exceed, so they are ranked by counting, not sorted
    const int32_t *levels;
    float select_threshold;
    float level_step;
    // fp32 library searched through its bf16 shadow (tensor-core PRE-selection of k > e.k rows, then
    // this kernel's exact fp32 re-score): cert_out[q] = 1 iff the emitted top-e.k is provably the
    // exact one -- at least e.k re-scored candidates lie above (score of the worst selected row in
    // the scan + cert_eps), the line no unselected row can reach -- else 0 and the host re-runs that
    // query on the exact scan.  Null = plain search.
    uint8_t *cert_out;
    float cert_eps;                // bound on |bf16 scan score - exact score|
    PeerPush peer;
    InlineParams inl;
};
struct EmitArgs;
// fused_emit != nullptr (single candidate list): the emit stage runs inside the same kernel.
void launch_finalize(const FinalizeArgs &a, const EmitArgs *fused_emit, int sm_count, cudaStream_t s);

struct EmitArgs {
    const cab_candidate *cands;   // [n_lists][n_queries][k]
    int n_lists;
    int n_queries;
    int k;
    const double *w_asr;          // device [n_queries]
    const double *w_audio;
    int inline_weights;           // n_queries == 1: use w64_asr / w64_audio instead of the arrays
    double w64_asr, w64_audio;
    double threshold;
    int64_t *out_index;           // device [n_queries][k] (all required here)
    double *out_fusion;
    float *out_asr;
    float *out_audio;
    uint8_t *out_flags;
    int32_t *out_count;
    // A NaN/Inf query (sklearn raises ValueError for the reference) has no results: out_count = -1,
    // and *nonfinite_out (host-visible block, zeroed by the host before the launch) is set to 1.
    int *nonfinite_out;           // may be null (device outputs: out_count = -1 is the signal)
    // zero-copy host results: outputs point into mapped pinned host memory; when every CTA of the
    // (last) launch has written, done_flag (also in that block) is set to done_epoch for the host
    uint32_t *done_flag;          // null = outputs are ordinary device memory
    uint32_t done_epoch;
    unsigned int *done_counter;   // device scratch, zero between launches
    // peer exchange: wait until wait_flags[0..n_lists) all hold wait_epoch before reading cands
    const uint32_t *wait_flags;   // null = no wait
    uint32_t wait_epoch;
    int *status;                  // set to 7 (and the kernel traps) if the wait times out
    // sharded search, diagnostic: %globaltimer (ns) at {scan complete, best k selected (dependents
    // released), winners re-scored, own flag raised, all flags seen, results written} of query 0's
    // CTA, 8 slots per search (2 unused); null = not recorded
    unsigned long long *stamps;
};
// cab_candidate.index of every slot of a query that holds NaN/Inf (travels through the exchange so
// that a merge without the query -- cab_merge_candidates -- reports it too)
constexpr int64_t kBadQueryIndex = -2;
// Push a local candidate block [n_queries x k] to every peer and raise the flags (used when there
// was nothing to scan; otherwise the finalize kernel pushes).
void launch_peer_push(const cab_candidate *local, int n_queries, int k, const PeerPush &peer, cudaStream_t s);
void launch_emit(const EmitArgs &a, cudaStream_t s);
// Result arrays [n_queries x k] (device) of this shard -> candidate records, pushed to every rank;
// raises this rank's flags (peer.signal must be 1, peer.q0 the first query).
void launch_pack_push(const int64_t *index, const float *asr, const float *audio, const uint8_t *flags, const int32_t *count,
                      int n_queries, int k, const PeerPush &peer, cudaStream_t s);

}  // namespace cab
