/*
 * cab.h -- C-ABI of the B200-native fused dual-corpus retrieval engine ("cab" = ClipABit).
 *
 * Drop-in boundary for ONE path of ClipABit/Multimodal-Audio-Search:
 *   DualPipelineAudioSearch.search_with_fusion        /root/reference/audio_search.py:624-699
 * minus the query-embedding call (:635) and the keyword weight rule (:457-622), which stay on
 * the host.  The reference has no FFI/plugin interface (it is one Python file); each entry point
 * below names the reference statement(s) it replaces.  The Python side that mirrors the
 * reference's call surface is multimodal_audio_search_b200/engine.py; the binding a maintainer
 * would add to the reference is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every function returns a cab_status (0 = ok) and never throws;
 *   - cab_last_error(idx) returns a static/handle-owned message for the last failure;
 *   - a handle is externally synchronised (one caller at a time); distinct handles may be used
 *     from distinct threads (the reference keeps one engine per Streamlit session, :708-711);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the handle's own non-blocking stream;
 *     to run on the default stream pass cudaStreamLegacy / cudaStreamPerThread explicitly);
 *   - `*_loc` says where a caller buffer lives: CAB_HOST or CAB_DEVICE;
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with
 *     CAB_ERR_NO_DEVICE.
 */
#ifndef CAB_H_
#define CAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CAB_API __attribute__((visibility("default")))
#else
#define CAB_API
#endif

#define CAB_VERSION 100          /* 0.1.0 */
#define CAB_DIM 384              /* all-MiniLM-L6-v2, audio_search.py:92 */
#define CAB_MAX_K 128            /* reference uses 10 (:699); BASELINE configs use up to 100 */
#define CAB_MAX_QUERIES 4096     /* per cab_search call */

typedef struct cab_index cab_index;

typedef enum cab_status {
    CAB_OK = 0,
    CAB_ERR_INVALID = 1,      /* bad argument */
    CAB_ERR_CUDA = 2,         /* a CUDA call failed (sticky on the handle) */
    CAB_ERR_NONFINITE = 3,    /* NaN/Inf in rows or query: sklearn raises ValueError here */
    CAB_ERR_NO_DEVICE = 4,    /* no CUDA device / not an sm_100 part */
    CAB_ERR_NOMEM = 5
} cab_status;

typedef enum cab_dtype { CAB_F32 = 0, CAB_BF16 = 1 } cab_dtype;      /* storage type of rows */
typedef enum cab_loc { CAB_HOST = 0, CAB_DEVICE = 1 } cab_loc;
typedef enum cab_path {                                              /* which scan kernel */
    CAB_PATH_AUTO = 0,       /* GEMV for few queries; tensor-core GEMM from option "gemm_min_queries" on
                                (default: 4 queries on libraries of >= 65 536 segments, else 64) for a bf16
                                index or an fp32 index with bf16 shadows */
    CAB_PATH_GEMV = 1,       /* HBM-bound fused GEMV + top-k (CUDA cores, fp32 accumulate) */
    CAB_PATH_GEMM = 2        /* tcgen05/TMEM GEMM with fused weighting/top-k epilogue (bf16) */
} cab_path;

/* Row flags: the reference's per-segment `asr_success` / `audio_success` (:284, :289), which
 * select the effective weights (:656-664).  A missing embedding (`None`, :640-651) is passed
 * as an all-zero row: its cosine is exactly 0.0 like the reference's. */
#define CAB_FLAG_ASR 1u
#define CAB_FLAG_AUDIO 2u
/* Bits 2-3 of the flag byte: the row's weight class (0..3), read only by cab_score_all (the
 * earlier engine's per-row weighting, previous_iterations/streamlit_app.py:214-219); the top-k
 * search ignores them and hands them back untouched in out_flags. */
#define CAB_FLAG_CLASS_SHIFT 2
#define CAB_FLAG_CLASS_MASK 0x0Cu

CAB_API int cab_version(void);
CAB_API const char *cab_status_string(int status);
/* Last error text of `idx` (or of the calling thread's last failed create when idx == NULL). */
CAB_API const char *cab_last_error(const cab_index *idx);
/* Number of visible CUDA devices (0 if none / driver missing). Never fails. */
CAB_API int cab_device_count(void);

/* ---- segment store: replaces the `audio_segments` list as the thing that is scanned --------
 * (audio_search.py:115; rows appended at :797).  Two row-major [rows x 384] matrices (ASR,
 * audio), L2-normalised on ingest, 128-byte-aligned rows, plus one flag byte per row. */
CAB_API int cab_index_create(int dim, int dtype, int64_t capacity_rows, int device, cab_index **out);
CAB_API int cab_index_destroy(cab_index *idx);
CAB_API int cab_index_reserve(cab_index *idx, int64_t capacity_rows);
CAB_API int64_t cab_index_size(const cab_index *idx);
CAB_API int64_t cab_index_capacity(const cab_index *idx);
CAB_API int cab_index_dtype(const cab_index *idx);
CAB_API int cab_index_device(const cab_index *idx);
/* Global segment index of local row 0 (corpus sharded by segment over ranks, default 0). */
CAB_API int cab_index_set_row_base(cab_index *idx, int64_t row_base);
CAB_API int64_t cab_index_row_base(const cab_index *idx);
CAB_API int cab_index_clear(cab_index *idx);

/* Append n_rows segments.  asr_rows / audio_rows: fp32 [n_rows x dim], raw (any length; the
 * engine normalises like sklearn `normalize`, zero rows stay zero) or NULL (= all rows missing);
 * flags: n_rows bytes of CAB_FLAG_* or NULL (= both pipelines succeeded).
 * Fails with CAB_ERR_NONFINITE (nothing appended) if any value is NaN/Inf.
 * Replaces: audio_segments.extend(...) at :797 + the per-call normalize(Y) of :646/:651. */
CAB_API int cab_index_append(cab_index *idx, const float *asr_rows, const float *audio_rows,
                     const uint8_t *flags, int64_t n_rows, int rows_loc, void *stream);

/* Append global rows [r0, r1) of the deterministic synthetic library (seed, n_total rows,
 * planted neighbours, optional partial flags) generated ON DEVICE; bit-identical to
 * multimodal_audio_search_b200/synth.py.  Benchmark/test data source (BASELINE.json configs).
 * `partial`: bit 0 = ~10 % ASR-only / ~10 % audio-only rows; bits 8-15 = distribution: 0 planted
 * neighbours over isotropic noise, 1 "ascending" scores (defeats top-k pruning), 2 "clustered". */
CAB_API int cab_index_append_synth(cab_index *idx, uint32_t seed, int64_t n_total, int64_t r0, int64_t r1,
                           int n_queries, int plants, int partial, void *stream);
/* Raw synthetic query vectors [q0, q1) as fp32 [n x dim] into `out` (host or device). */
CAB_API int cab_synth_queries(int device, uint32_t seed, int q0, int q1, float *out, int out_loc);

/* Copy normalised rows [r0, r1) back as fp32 (bf16 storage is widened) -- test/debug hook. */
CAB_API int cab_index_read_rows(cab_index *idx, int corpus /*0 asr, 1 audio*/, int64_t r0, int64_t r1,
                        float *out, int out_loc);

/* Overwrite / read back the flag bytes of rows [r0, r1) (host or device buffer of r1 - r0 bytes):
 * success bits and weight class of segments that are already in the index (e.g. assigning the
 * legacy weight classes to a library that was ingested without them). */
CAB_API int cab_index_write_flags(cab_index *idx, int64_t r0, int64_t r1, const uint8_t *flags, int flags_loc);
CAB_API int cab_index_read_flags(cab_index *idx, int64_t r0, int64_t r1, uint8_t *out, int out_loc);

/* ---- persistent index file (the reference keeps its library only
 * in the Streamlit session, audio_search.py:708-711, and loses it when the session ends) ---------
 * Layout (little endian): 4096-byte header {magic "CABIDX01", version, dim, dtype, n_rows,
 * row_base, section offsets}, then the ASR rows, the audio rows (exactly as resident in HBM:
 * L2-normalised fp32/bf16, row-major), the flag bytes and (version 2) the fp32 original lengths
 * of the ASR and audio rows, each section 4096-byte aligned so the file can be mmap-ed.
 * Version-1 files (no lengths) still load; their rows count as unit length.  A reload is bit-identical: rows are not re-normalised.  Segment metadata
 * (texts, times, audio) stays with the caller. */
CAB_API int cab_index_save(cab_index *idx, const char *path);
/* Load rows [r0, r1) of a file into a new index on `device` (r1 < 0: to the end).  The new
 * index's row_base is the file's row_base + r0, so shards of one file merge correctly. */
CAB_API int cab_index_load(const char *path, int device, int64_t r0, int64_t r1, cab_index **out);
/* Header fields without touching CUDA (any may be NULL). */
CAB_API int cab_index_file_info(const char *path, int *dim, int *dtype, int64_t *n_rows, int64_t *row_base);

/* ---- search: replaces the per-segment loop, threshold, sort and [:k]  (:639-685, :699) -------
 * queries : fp32 [n_queries x dim], raw (re-normalised like normalize(X) at :646);
 * w_asr, w_audio : host arrays [n_queries] of the query weights from
 *                  _analyze_query_for_weights (:632), float64 like the reference's floats;
 *                  finite, >= 0, w_asr + w_audio > 0.  A weight of 0 makes the search
 *                  single-corpus: rows whose only successful pipeline carries weight 0 have
 *                  effective weights summing to 0 and are skipped (:659-661);
 * k <= CAB_MAX_K; threshold: the reference's 0.1 (:672), strict `>` evaluated in float64.
 * Outputs (each [n_queries x k], may be NULL if not wanted; host or device per out_loc):
 *   out_index  global segment index, best first (score desc, index asc = Python's stable
 *              sort, :685); -1 beyond out_count[q]
 *   out_fusion float64 fusion_score  = eff_w_asr*asr_sim + eff_w_audio*audio_sim  (:667-670)
 *   out_asr / out_audio  float32 cosine similarities (:646, :651)
 *   out_flags  the row's CAB_FLAG_* byte (gives the effective weights of :656-664)
 *   out_count  [n_queries] number of results (< k when fewer rows pass the threshold).
 * A query holding NaN/Inf (sklearn raises ValueError for the reference): with host outputs the
 * call fails with CAB_ERR_NONFINITE; with device outputs (nothing is read back, the call does
 * not wait) that query gets out_count = -1 and no results, the other queries are unaffected.
 * Stream contract: calls on one handle share its workspace and are ordered by the stream they
 * run on; a call that arrives on a different stream than the previous one first waits for the
 * device (so a device-tensor call on the caller's stream may be followed by a host call on the
 * handle's own stream).  On one stream a search is launched so that it may START before the
 * previous search has finished (programmatic dependent launch) but never reads a device query
 * before the kernel in front of it has completed -- unless option "queries_settled" is set. */
CAB_API int cab_search(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
               const double *w_audio, int n_queries, int k, double threshold, int path,
               int64_t *out_index, double *out_fusion, float *out_asr, float *out_audio,
               uint8_t *out_flags, int32_t *out_count, int out_loc, void *stream);

/* ---- sharded search (corpus split by segment over ranks) ----------------------------------
 * Step 1, per rank: local top-k as packed candidates (cab_candidate[n_queries x k], device).
 * Step 2, host plumbing: all-gather the candidate blocks of all ranks (NCCL over NVLink).
 * Step 3, per rank: merge world x k candidates per query into the final top-k
 *         (w_asr = w_audio = NULL reuses the weights staged by step 1 on the same handle). */
typedef struct cab_candidate {
    int64_t index;       /* global segment index, -1 = empty slot */
    float asr_sim;
    float audio_sim;
    uint32_t flags;
    uint32_t pad;
} cab_candidate;         /* 24 bytes */

CAB_API int cab_search_candidates(cab_index *idx, const float *queries, int queries_loc,
                          const double *w_asr, const double *w_audio, int n_queries, int k,
                          double threshold, int path, cab_candidate *out_device, void *stream);
CAB_API int cab_merge_candidates(cab_index *idx, const cab_candidate *cands_device, int n_lists,
                         int n_queries, int k, const double *w_asr, const double *w_audio,
                         double threshold, int64_t *out_index, double *out_fusion, float *out_asr,
                         float *out_audio, uint8_t *out_flags, int32_t *out_count, int out_loc,
                         void *stream);

/* ---- legacy scoring modes: every segment's fused similarity, no threshold, no top-k ------------
 * Replaces the per-item loop of the reference's earlier engine, `UnifiedAudioSearch.search`
 * (previous_iterations/streamlit_app.py:173-223), which returns np.array(similarities) over the
 * whole database for strategy "asr_only" / "caption_only" / "adaptive".
 *   out[q][r] = class_weights[q][c][0] * cos(query q, asr row r)
 *             + class_weights[q][c][1] * cos(query q, audio row r),    c = weight class of row r
 * in fp32 with the products and the sum rounded separately (numpy float32 scalar arithmetic); a
 * missing embedding (zero row) contributes 0.0; no success gating, no renormalisation.
 * class_weights: HOST float [n_queries][4][2].  asr_only = {1,0} for every class, caption_only =
 * {0,1}, adaptive = class 0 {0.2,0.8}, class 1 {0.7,0.3} with class 1 = "transcript longer than
 * 10 characters" (:216).  out: [n_queries][cab_index_size] floats, host or device.
 * A NaN/Inf query: CAB_ERR_NONFINITE with host output; all-NaN scores with device output. */
CAB_API int cab_score_all(cab_index *idx, const float *queries, int queries_loc, int n_queries,
                  const float *class_weights, float *out, int out_loc, void *stream);

/* ---- sharded search with the exchange fused into the kernels (NVLink peer memory) ---------------
 * Instead of steps 2+3 above through NCCL: the finalize kernel of every rank stores its k
 * candidates per query straight into EVERY rank's exchange buffer (peer stores over NVLink /
 * NVSwitch) and then raises a per-rank epoch flag; the merge kernel waits for the world's flags
 * and merges.  One process per GPU; buffers are shared through CUDA IPC handles that the host
 * side exchanges once (any transport; sharded.py uses torch.distributed).
 *   cab_peer_init   : allocate this rank's exchange buffer, return its 64-byte IPC handle
 *   cab_peer_attach : open all ranks' handles (world x 64 bytes, rank order)
 *   cab_search_sharded : collective -- every rank calls it with the same n_queries / k / path.
 *                     Outputs as cab_search; every rank receives the same merged result. */
#define CAB_IPC_HANDLE_BYTES 64
#define CAB_MAX_WORLD 8
CAB_API int cab_peer_init(cab_index *idx, int rank, int world, int max_queries, int max_k, void *ipc_handle_out);
CAB_API int cab_peer_attach(cab_index *idx, const void *all_handles);
CAB_API int cab_search_sharded(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
                               const double *w_audio, int n_queries, int k, double threshold, int path,
                               int64_t *out_index, double *out_fusion, float *out_asr, float *out_audio,
                               uint8_t *out_flags, int32_t *out_count, int out_loc, void *stream);

/* Diagnostic: copy this rank's whole exchange buffer to the host after a device synchronise --
 * [256 bytes: epoch flags 2 x world u32][2 halves x world x max_queries x max_k cab_candidate].
 * `needed` (may be NULL) receives its size; out == NULL only queries the size. */
CAB_API int cab_peer_snapshot(cab_index *idx, void *out, size_t out_bytes, size_t *needed);

/* Diagnostic (option "stamp_exchange" = 1): %globaltimer stamps, in ns, of the last sharded
 * searches whose merge ran inside the finalize kernel -- per search a row of 8 uint64:
 * {scan complete, best k selected (dependents released), winners re-scored, own epoch flag raised
 * on every rank, all ranks' flags seen, results written, 0, 0}.  Copies up to max_rows rows
 * (oldest first) into `out` (host) and returns the number of rows copied (0 on error, with
 * cab_last_error set).  Synchronises the device.  NOTE: returns a row count, not a cab_status. */
CAB_API int cab_index_exchange_stamps(cab_index *idx, uint64_t *out, int max_rows);

/* ---- tuning / introspection -------------------------------------------------------------- */
/* Options: "gemv_unroll" (row-steps in flight per warp: 1/2/4/8, 0 = default),
 * "gemv_blocks_per_sm" (resident CTAs per SM, 0 = default), "gemv_query_tile" (queries scored per
 * corpus pass, 1/2/4, 0 = default 4), "gemv_batch", "gemm_min_queries",
 * "time_kernels", "sync_after_search", "stamp_exchange";
 * "tensor_core_shadow" = 1 (fp32 index): keep bf16 shadow copies of both corpora (+50 % memory);
 * cab_search batches of >= "gemm_min_queries" queries (default 4, k <= 112) are then PRE-selected on the
 * tensor cores from the shadows (max(96, k + k/2 + 32) rows per query), re-scored exactly from the fp32
 * rows, and certified per query: the top-k is provably the exact one when k re-scored candidates
 * lie above (scan score of the worst selected row + 4e-3, the bound on the bf16 error of a cosine
 * of unit vectors).  Uncertified queries are re-run on the exact GEMV scan before the call
 * returns (get_option "last_uncertified" / "total_uncertified" / "total_shadow_queries"), so the
 * results are the fp32 path's; with device outputs the call synchronises the stream once.
 * "raw_dot" = 1: similarities are raw dot products <query, row> of the vectors AS APPENDED instead
 * of cosines (the index keeps every row's original length; the query is not normalised) -- the
 * ranking rule of the reference's earlier engine, previous_iterations/clean_audio_search.py:306-310
 * (`float(np.dot(query_embedding, embedding))`), for embeddings that are not unit length.  Served
 * by the GEMV path; threshold, fusion and order are unchanged.  Ignored by cab_score_all.
 * "queries_settled" = 1: the caller promises that DEVICE query buffers passed to a search were
 * completely written before the previous search on this handle was issued (pre-computed query
 * batches); the scan of search i+1 then streams the corpus while search i's finalize / exchange /
 * merge is still in flight.  Default 0: a search waits for the kernel in front of it before it
 * reads its query.  Unknown keys fail. */
CAB_API int cab_index_set_option(cab_index *idx, const char *key, int64_t value);
CAB_API int64_t cab_index_get_option(const cab_index *idx, const char *key);
/* Kernels launched by this handle since creation (for bench.py's gpu_launches). */
CAB_API int64_t cab_index_launch_count(const cab_index *idx);
/* Device time (ms, CUDA events on the handle's stream) of the scan kernel of the last search
 * when option "time_kernels" is 1; < 0 if not recorded. */
CAB_API double cab_index_last_scan_ms(const cab_index *idx);

#ifdef __cplusplus
}
#endif
#endif /* CAB_H_ */
