// C-ABI of the engine (include/cab.h): device-resident segment store + search orchestration.
// No CPU fallback anywhere: if CUDA is unusable every entry point reports CAB_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "cab_internal.h"

using namespace cab;

namespace cab {
// tensor-core path (cab_gemm_tc.cu)
struct GemmScanArgs;
int gemm_partials_per_query(int sm_count);
bool gemm_path_available();
void launch_gemm_scan(const ScanArgs &a, int sm_count, void *workspace, size_t workspace_bytes,
                      cudaStream_t s, std::string *err);
size_t gemm_workspace_bytes(int n_queries, int k, int sm_count);
const int32_t *gemm_levels(const void *workspace, int sm_count);
float gemm_level_step();
}  // namespace cab

static thread_local std::string g_last_error;

struct cab_index {
    int device = 0;
    int dtype = CAB_F32;
    int sm_count = 148;
    int64_t capacity = 0, size = 0, row_base = 0;
    void *asr = nullptr, *audio = nullptr;
    float *norm_asr = nullptr, *norm_audio = nullptr;   // original row lengths (raw dot-product scoring)
    // fp32 index, option tensor_core_shadow: bf16 copies of both corpora -- batches are PRE-selected on
    // the tensor cores from the shadows, re-scored exactly from the fp32 rows and certified per query
    void *shadow_asr = nullptr, *shadow_audio = nullptr;
    uint8_t *d_cert = nullptr;     size_t sz_cert = 0;    // per query of the last shadow batch: 1 = provably exact
    int cert_pending = 0;                                // queries of the batch whose certificates are to be checked
    int64_t last_uncertified = 0, total_uncertified = 0, total_shadow_queries = 0;
    uint8_t *flags = nullptr;
    cudaStream_t own_stream = nullptr;
    // search workspace (device)
    // per-search parameter block, one H2D copy: [w64 asr|w64 audio|w32 a|w32 b|queries (if host)]
    uint8_t *d_params = nullptr;   size_t sz_params = 0;
    double *d_w64 = nullptr;       // views into d_params, set by stage_params
    float *d_w32 = nullptr;
    int staged_nq = 0;             // number of queries whose weights are currently staged
    bool inline_w_valid = false;   // the last search carried one query's weights inline
    double inline_w64[2] = {0, 0};
    uint64_t *d_partial_keys = nullptr;  size_t sz_pkeys = 0;
    cab_candidate *d_cands = nullptr;    size_t sz_cands = 0;
    uint8_t *d_out = nullptr;      size_t d_out_bytes = 0;  // packed outputs
    uint8_t *d_gemm_ws = nullptr;  size_t d_gemm_ws_bytes = 0;
    int *d_nonfinite = nullptr;              // set by the ingest / score_all kernels (rows, or a score_all query)
    unsigned int *d_counters = nullptr;     // [0..63] GEMV chunk tickets, [64..65] score_all tickets (zero between launches)
    // pinned host staging
    uint8_t *h_in = nullptr;  size_t h_in_bytes = 0;
    uint8_t *h_out = nullptr; size_t h_out_bytes = 0;
    uint8_t *h_rows = nullptr; size_t h_rows_bytes = 0;
    float *d_rows = nullptr;  size_t d_rows_bytes = 0;   // raw-row device staging (append)
    uint8_t *d_scores = nullptr; size_t d_scores_bytes = 0;   // cab_score_all with host output
    cudaEvent_t ev_in = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};     // append staging slots
    bool ev_in_pending = false, timed = false;
    // options
    GemvConfig gemv{0, 0, 0, 0};
    int64_t opt_time_kernels = 0, opt_sync = 0, opt_gemm_min_queries = 0 /* auto, see gemm_min_queries() */, opt_gemv_batch = 32;
    int64_t opt_queries_settled = 0, opt_stamp_exchange = 0, opt_raw_dot = 0;
    int64_t opt_finalize_general = 0, opt_chunk_rows = 0;    // 0 = auto: 96 KB of corpus per chunk (fp32 32 rows, bf16 64), profiles/r01_gemv_chunk_sweep.md
    // peer-memory exchange (sharded search)
    int peer_world = 0, peer_rank = 0, peer_qcap = 0, peer_kcap = 0;
    bool peer_attached = false;
    uint8_t *d_peer = nullptr;                       // [flags 2 x world u32 | pad to 256][cands 2 x world x qcap x kcap]
    uint8_t *peer_ptr[CAB_MAX_WORLD] = {};           // every rank's buffer as mapped in this process
    uint32_t peer_epoch = 0;
    unsigned int *d_done = nullptr;
    int *d_status = nullptr;
    // Stream of the previous search / scoring call.  The workspace (partial lists, ticket counters,
    // parameter block, exchange buffers) is shared by consecutive calls, which are ordered by the
    // stream they run on; a call that arrives on ANOTHER stream first waits for the device.
    cudaStream_t last_stream = nullptr;
    bool last_stream_set = false;
    unsigned long long *d_stamps = nullptr;          // [kStampRows][4] %globaltimer stamps of the last sharded searches
    uint32_t stamp_calls = 0;
    unsigned int *d_host_done = nullptr;             // CTA counter for host-visible completion
    uint32_t host_epoch = 0;
    int64_t launches = 0;
    std::string err;
    int sticky = CAB_OK;
};

static int fail(cab_index *idx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (idx) { idx->err = buf; if (code == CAB_ERR_CUDA) idx->sticky = code; }
    g_last_error = buf;
    return code;
}

#define CU(idx, expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(idx, e_ == cudaErrorMemoryAllocation ? CAB_ERR_NOMEM : CAB_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CHECK_HANDLE(idx)                                                        \
    do {                                                                         \
        if (!(idx)) return fail(nullptr, CAB_ERR_INVALID, "null index handle");  \
        if ((idx)->sticky != CAB_OK) return (idx)->sticky;                       \
    } while (0)

// Calls on one handle share its workspace and are ordered by their stream.  A caller that moves to
// another stream (e.g. a device-tensor call on torch's stream followed by a host call on the
// handle's own stream) gets the ordering it implicitly expects: the device is drained once.
static int enter_stream(cab_index *idx, cudaStream_t s) {
    if (idx->last_stream_set && idx->last_stream != s) {
        CU(idx, cudaSetDevice(idx->device));
        CU(idx, cudaDeviceSynchronize());
    }
    idx->last_stream = s; idx->last_stream_set = true;
    return CAB_OK;
}

static size_t elem_size(int dtype) { return dtype == CAB_BF16 ? 2 : 4; }
static size_t align_up(size_t x, size_t a);
constexpr int kStampRows = 64, kStampCols = 8;
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int cab_version(void) { return CAB_VERSION; }

const char *cab_status_string(int s) {
    switch (s) {
        case CAB_OK: return "ok";
        case CAB_ERR_INVALID: return "invalid argument";
        case CAB_ERR_CUDA: return "CUDA error";
        case CAB_ERR_NONFINITE: return "input contains NaN or infinity";
        case CAB_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
        case CAB_ERR_NOMEM: return "out of device memory";
        default: return "unknown status";
    }
}

const char *cab_last_error(const cab_index *idx) { return idx ? idx->err.c_str() : g_last_error.c_str(); }

int cab_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int grow(cab_index *idx, int64_t new_cap) {
    if (new_cap <= idx->capacity) return CAB_OK;
    const size_t row_bytes = CAB_DIM * elem_size(idx->dtype);
    void *na = nullptr, *nb = nullptr;
    float *la = nullptr, *lb = nullptr;
    void *sa_ = nullptr, *sb_ = nullptr;
    const bool shadow = idx->shadow_asr != nullptr;
    uint8_t *nf = nullptr;
    CU(idx, cudaSetDevice(idx->device));
    cudaError_t e = cudaMalloc(&na, size_t(new_cap) * row_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&nb, size_t(new_cap) * row_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&la, size_t(new_cap) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&lb, size_t(new_cap) * sizeof(float));
    if (e == cudaSuccess && shadow) e = cudaMalloc(&sa_, size_t(new_cap) * CAB_DIM * 2);
    if (e == cudaSuccess && shadow) e = cudaMalloc(&sb_, size_t(new_cap) * CAB_DIM * 2);
    if (e == cudaSuccess) e = cudaMalloc(&nf, align_up(size_t(new_cap), 16));     // the tensor-core epilogue reads flag WORDS
    if (e != cudaSuccess) {
        cudaFree(na); cudaFree(nb); cudaFree(nf); cudaFree(la); cudaFree(lb); cudaFree(sa_); cudaFree(sb_); cudaGetLastError();
        return fail(idx, CAB_ERR_NOMEM, "cannot allocate %lld rows (%s)", (long long)new_cap, cudaGetErrorString(e));
    }
    if (idx->size > 0) {
        CU(idx, cudaMemcpyAsync(na, idx->asr, size_t(idx->size) * row_bytes, cudaMemcpyDeviceToDevice, idx->own_stream));
        CU(idx, cudaMemcpyAsync(nb, idx->audio, size_t(idx->size) * row_bytes, cudaMemcpyDeviceToDevice, idx->own_stream));
        CU(idx, cudaMemcpyAsync(nf, idx->flags, size_t(idx->size), cudaMemcpyDeviceToDevice, idx->own_stream));
        CU(idx, cudaMemcpyAsync(la, idx->norm_asr, size_t(idx->size) * sizeof(float), cudaMemcpyDeviceToDevice, idx->own_stream));
        CU(idx, cudaMemcpyAsync(lb, idx->norm_audio, size_t(idx->size) * sizeof(float), cudaMemcpyDeviceToDevice, idx->own_stream));
        if (shadow) {
            CU(idx, cudaMemcpyAsync(sa_, idx->shadow_asr, size_t(idx->size) * CAB_DIM * 2, cudaMemcpyDeviceToDevice, idx->own_stream));
            CU(idx, cudaMemcpyAsync(sb_, idx->shadow_audio, size_t(idx->size) * CAB_DIM * 2, cudaMemcpyDeviceToDevice, idx->own_stream));
        }
    }
    CU(idx, cudaStreamSynchronize(idx->own_stream));
    cudaFree(idx->asr); cudaFree(idx->audio); cudaFree(idx->flags); cudaFree(idx->norm_asr); cudaFree(idx->norm_audio);
    cudaFree(idx->shadow_asr); cudaFree(idx->shadow_audio);
    idx->asr = na; idx->audio = nb; idx->flags = nf; idx->norm_asr = la; idx->norm_audio = lb; idx->capacity = new_cap;
    idx->shadow_asr = sa_; idx->shadow_audio = sb_;
    return CAB_OK;
}

int cab_index_create(int dim, int dtype, int64_t capacity_rows, int device, cab_index **out) {
    if (!out) return fail(nullptr, CAB_ERR_INVALID, "out is null");
    *out = nullptr;
    if (dim != CAB_DIM) return fail(nullptr, CAB_ERR_INVALID, "dim must be %d (all-MiniLM-L6-v2), got %d", CAB_DIM, dim);
    if (dtype != CAB_F32 && dtype != CAB_BF16) return fail(nullptr, CAB_ERR_INVALID, "dtype must be CAB_F32 or CAB_BF16");
    if (capacity_rows < 0 || capacity_rows > 0xFFFFFFF0ll) return fail(nullptr, CAB_ERR_INVALID, "capacity_rows out of range");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, CAB_ERR_NO_DEVICE, "no CUDA device visible; this engine has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, CAB_ERR_INVALID, "device %d out of range (%d visible)", device, n);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, CAB_ERR_NO_DEVICE, "cannot query device %d", device); }
    if (prop.major != 10) return fail(nullptr, CAB_ERR_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    cab_index *idx = new (std::nothrow) cab_index();
    if (!idx) return fail(nullptr, CAB_ERR_NOMEM, "host allocation failed");
    idx->device = device; idx->dtype = dtype; idx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&idx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_nonfinite, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(idx->d_nonfinite, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_host_done, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(idx->d_host_done, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_counters, 128 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(idx->d_counters, 0, 128 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&idx->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&idx->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&idx->ev_t1);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, CAB_ERR_CUDA, "index setup failed: %s", cudaGetErrorString(e));
        cab_index_destroy(idx);
        return rc;
    }
    if (capacity_rows > 0) {
        int rc = grow(idx, capacity_rows);
        if (rc != CAB_OK) { g_last_error = idx->err; cab_index_destroy(idx); return rc; }
    }
    *out = idx;
    return CAB_OK;
}

int cab_index_destroy(cab_index *idx) {
    if (!idx) return CAB_OK;
    cudaSetDevice(idx->device);
    if (idx->own_stream) cudaStreamSynchronize(idx->own_stream);
    cudaFree(idx->asr); cudaFree(idx->audio); cudaFree(idx->flags); cudaFree(idx->norm_asr); cudaFree(idx->norm_audio);
    cudaFree(idx->shadow_asr); cudaFree(idx->shadow_audio); cudaFree(idx->d_cert);
    cudaFree(idx->d_params);
    cudaFree(idx->d_partial_keys); cudaFree(idx->d_cands);
    cudaFree(idx->d_out); cudaFree(idx->d_gemm_ws); cudaFree(idx->d_nonfinite); cudaFree(idx->d_rows); cudaFree(idx->d_counters); cudaFree(idx->d_scores);
    for (int r = 0; r < idx->peer_world; ++r)
        if (idx->peer_attached && r != idx->peer_rank && idx->peer_ptr[r]) cudaIpcCloseMemHandle(idx->peer_ptr[r]);
    cudaFree(idx->d_peer); cudaFree(idx->d_done); cudaFree(idx->d_status); cudaFree(idx->d_host_done); cudaFree(idx->d_stamps);

    cudaFreeHost(idx->h_in); cudaFreeHost(idx->h_out); cudaFreeHost(idx->h_rows);
    if (idx->ev_in) cudaEventDestroy(idx->ev_in);
    if (idx->ev_t0) cudaEventDestroy(idx->ev_t0);
    if (idx->ev_t1) cudaEventDestroy(idx->ev_t1);
    for (cudaEvent_t e : idx->ev_slot) if (e) cudaEventDestroy(e);
    if (idx->own_stream) cudaStreamDestroy(idx->own_stream);
    cudaGetLastError();
    delete idx;
    return CAB_OK;
}

int cab_index_reserve(cab_index *idx, int64_t capacity_rows) {
    CHECK_HANDLE(idx);
    if (capacity_rows > 0xFFFFFFF0ll) return fail(idx, CAB_ERR_INVALID, "capacity_rows out of range");
    return grow(idx, capacity_rows);
}
int64_t cab_index_size(const cab_index *idx) { return idx ? idx->size : -1; }
int64_t cab_index_capacity(const cab_index *idx) { return idx ? idx->capacity : -1; }
int cab_index_dtype(const cab_index *idx) { return idx ? idx->dtype : -1; }
int cab_index_device(const cab_index *idx) { return idx ? idx->device : -1; }
int cab_index_set_row_base(cab_index *idx, int64_t row_base) {
    CHECK_HANDLE(idx);
    if (row_base < 0) return fail(idx, CAB_ERR_INVALID, "row_base must be >= 0");
    idx->row_base = row_base;
    return CAB_OK;
}
int64_t cab_index_row_base(const cab_index *idx) { return idx ? idx->row_base : -1; }
int cab_index_clear(cab_index *idx) { CHECK_HANDLE(idx); idx->size = 0; return CAB_OK; }

static int ensure_pinned(cab_index *idx, uint8_t **p, size_t *have, size_t want) {
    if (*have >= want) return CAB_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *have = 0; }
    want = align_up(want, 4096);
    CU(idx, cudaMallocHost((void **)p, want));
    *have = want;
    return CAB_OK;
}

static int check_nonfinite(cab_index *idx, cudaStream_t s, bool *bad) {
    int h = 0;
    CU(idx, cudaMemcpyAsync(&h, idx->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(idx, cudaStreamSynchronize(s));
    *bad = h != 0;
    if (h) { CU(idx, cudaMemsetAsync(idx->d_nonfinite, 0, sizeof(int), s)); CU(idx, cudaStreamSynchronize(s)); }
    return CAB_OK;
}

int cab_index_append(cab_index *idx, const float *asr_rows, const float *audio_rows,
                     const uint8_t *flags, int64_t n_rows, int rows_loc, void *stream) {
    CHECK_HANDLE(idx);
    if (n_rows < 0) return fail(idx, CAB_ERR_INVALID, "n_rows < 0");
    if (n_rows == 0) return CAB_OK;
    if (rows_loc != CAB_HOST && rows_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "rows_loc");
    if (idx->size + n_rows > 0xFFFFFFF0ll) return fail(idx, CAB_ERR_INVALID, "more than 2^32 rows per index");
    CU(idx, cudaSetDevice(idx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    if (idx->size + n_rows > idx->capacity) {
        int rc = grow(idx, std::max(idx->size + n_rows, idx->capacity * 2));        // geometric growth
        if (rc == CAB_ERR_NOMEM && idx->capacity * 2 > idx->size + n_rows) rc = grow(idx, idx->size + n_rows);   // ... or exact fit
        if (rc != CAB_OK) return rc;
    }
    const int64_t dst0 = idx->size;
    if (rows_loc == CAB_DEVICE) {
        launch_normalize_rows(asr_rows, idx->asr, idx->dtype, dst0, n_rows, idx->d_nonfinite, idx->norm_asr, idx->shadow_asr, s);
        launch_normalize_rows(audio_rows, idx->audio, idx->dtype, dst0, n_rows, idx->d_nonfinite, idx->norm_audio, idx->shadow_audio, s);
        idx->launches += 2;
        if (flags) CU(idx, cudaMemcpyAsync(idx->flags + dst0, flags, size_t(n_rows), cudaMemcpyDeviceToDevice, s));
        else CU(idx, cudaMemsetAsync(idx->flags + dst0, 3, size_t(n_rows), s));
    } else {
        // Two staging slots (pinned host + device), used alternately: the CPU copies chunk i+1 into
        // its pinned slot while the copy engine and the normalise kernels work on chunk i.
        const int64_t chunk = 8192;                                    // rows per slot (2 corpora x 12 MB)
        const size_t row_bytes = CAB_DIM * sizeof(float);
        const size_t slot_bytes = 2 * size_t(chunk) * row_bytes;
        int rc = ensure_pinned(idx, &idx->h_rows, &idx->h_rows_bytes, 2 * slot_bytes);
        if (rc != CAB_OK) return rc;
        if (idx->d_rows_bytes < 2 * slot_bytes) {
            cudaFree(idx->d_rows); idx->d_rows = nullptr; idx->d_rows_bytes = 0;
            CU(idx, cudaMalloc((void **)&idx->d_rows, 2 * slot_bytes));
            idx->d_rows_bytes = 2 * slot_bytes;
        }
        for (int i = 0; i < 2; ++i)
            if (!idx->ev_slot[i]) CU(idx, cudaEventCreateWithFlags(&idx->ev_slot[i], cudaEventDisableTiming));
        int64_t step = 0;
        for (int64_t r = 0; r < n_rows; r += chunk, ++step) {
            const int slot = int(step & 1);
            const int64_t m = std::min(chunk, n_rows - r);
            float *ha = reinterpret_cast<float *>(idx->h_rows + slot * slot_bytes), *hb = ha + chunk * CAB_DIM;
            float *da = idx->d_rows + slot * (slot_bytes / sizeof(float)), *db = da + chunk * CAB_DIM;
            if (step >= 2) CU(idx, cudaEventSynchronize(idx->ev_slot[slot]));     // the slot's previous chunk is consumed
            if (asr_rows) { memcpy(ha, asr_rows + r * CAB_DIM, size_t(m) * row_bytes); CU(idx, cudaMemcpyAsync(da, ha, size_t(m) * row_bytes, cudaMemcpyHostToDevice, s)); }
            if (audio_rows) { memcpy(hb, audio_rows + r * CAB_DIM, size_t(m) * row_bytes); CU(idx, cudaMemcpyAsync(db, hb, size_t(m) * row_bytes, cudaMemcpyHostToDevice, s)); }
            launch_normalize_rows(asr_rows ? da : nullptr, idx->asr, idx->dtype, dst0 + r, m, idx->d_nonfinite, idx->norm_asr, idx->shadow_asr, s);
            launch_normalize_rows(audio_rows ? db : nullptr, idx->audio, idx->dtype, dst0 + r, m, idx->d_nonfinite, idx->norm_audio, idx->shadow_audio, s);
            idx->launches += 2;
            CU(idx, cudaEventRecord(idx->ev_slot[slot], s));
        }
        if (flags) CU(idx, cudaMemcpyAsync(idx->flags + dst0, flags, size_t(n_rows), cudaMemcpyHostToDevice, s));
        else CU(idx, cudaMemsetAsync(idx->flags + dst0, 3, size_t(n_rows), s));
    }
    CU(idx, cudaGetLastError());
    bool bad = false;
    int rc = check_nonfinite(idx, s, &bad);
    if (rc != CAB_OK) return rc;
    if (bad) return fail(idx, CAB_ERR_NONFINITE, "Input contains NaN or infinity (rows not appended)");
    idx->size += n_rows;
    return CAB_OK;
}

static uint32_t h_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
static uint64_t modinv(uint64_t a, uint64_t m) {      // a^-1 mod m, gcd(a, m) == 1, m <= 2^32
    __int128 t = 0, nt = 1, r = m, nr = a % m;
    while (nr != 0) { __int128 q = r / nr; __int128 x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    if (t < 0) t += m;
    return (uint64_t)t;
}
// Mirror of synth.plant_spec()
static SynthParams make_synth(uint32_t seed, int64_t n_total, int n_queries, int plants) {
    SynthParams p{};
    p.seed = seed; p.n_total = uint64_t(n_total); p.plants = uint32_t(plants > 0 ? plants : 1);
    p.n_plants_total = 0; p.base = 0; p.inv_stride = 1;
    if (n_total <= 1 || int64_t(n_queries) * plants == 0) return p;
    uint64_t n = uint64_t(n_total);
    uint64_t stride = uint64_t(double(n_total) * 0.6180339887) | 1ull;
    while (gcd64(stride, n) != 1) stride += 2;
    stride %= n;
    if (stride == 0) stride = 1;
    p.base = h_mix32(seed ^ 0xABCD1234u) % n;
    p.inv_stride = modinv(stride, n);
    p.n_plants_total = uint32_t(int64_t(n_queries) * plants);
    return p;
}

int cab_index_append_synth(cab_index *idx, uint32_t seed, int64_t n_total, int64_t r0, int64_t r1,
                           int n_queries, int plants, int partial, void *stream) {
    CHECK_HANDLE(idx);
    if (n_total <= 0 || n_total > 0xFFFFFFF0ll || r0 < 0 || r1 < r0 || r1 > n_total || n_queries < 0 || plants < 0)
        return fail(idx, CAB_ERR_INVALID, "bad synthetic library range");
    if (int64_t(n_queries) * plants > n_total) return fail(idx, CAB_ERR_INVALID, "more planted rows than library rows");
    const int64_t n_rows = r1 - r0;
    if (n_rows == 0) return CAB_OK;
    CU(idx, cudaSetDevice(idx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    if (idx->size + n_rows > idx->capacity) {
        int rc = grow(idx, std::max(idx->size + n_rows, idx->capacity * 2));        // geometric growth
        if (rc == CAB_ERR_NOMEM && idx->capacity * 2 > idx->size + n_rows) rc = grow(idx, idx->size + n_rows);   // ... or exact fit
        if (rc != CAB_OK) return rc;
    }
    const int mode = (partial >> 8) & 0xFF;
    partial &= 1;
    if (mode > 2) return fail(idx, CAB_ERR_INVALID, "unknown synthetic distribution %d", mode);
    SynthParams p = make_synth(seed, n_total, n_queries, mode == 0 ? plants : 0);
    p.mode = mode;
    const int64_t chunk = 65536;
    const size_t need = size_t(chunk) * CAB_DIM * sizeof(float);
    if (idx->d_rows_bytes < need) {
        cudaFree(idx->d_rows); idx->d_rows = nullptr; idx->d_rows_bytes = 0;
        CU(idx, cudaMalloc((void **)&idx->d_rows, need));
        idx->d_rows_bytes = need;
    }
    for (int st = 0; st < 2; ++st) {
        void *dst = st == 0 ? idx->asr : idx->audio;
        for (int64_t r = 0; r < n_rows; r += chunk) {
            const int64_t m = std::min(chunk, n_rows - r);
            launch_synth_rows(p, st, partial ? 1 : 0, r0 + r, m, idx->d_rows, s);
            launch_normalize_rows(idx->d_rows, dst, idx->dtype, idx->size + r, m, idx->d_nonfinite, st == 0 ? idx->norm_asr : idx->norm_audio,
                                  st == 0 ? idx->shadow_asr : idx->shadow_audio, s);
            idx->launches += 2;
        }
    }
    launch_synth_flags(seed, partial, r0, n_rows, idx->flags + idx->size, s);
    idx->launches += 1;
    CU(idx, cudaGetLastError());
    CU(idx, cudaStreamSynchronize(s));
    idx->size += n_rows;
    return CAB_OK;
}

int cab_synth_queries(int device, uint32_t seed, int q0, int q1, float *out, int out_loc) {
    if (!out || q0 < 0 || q1 < q0) return fail(nullptr, CAB_ERR_INVALID, "bad query range");
    if (q1 == q0) return CAB_OK;
    if (cab_device_count() == 0) return fail(nullptr, CAB_ERR_NO_DEVICE, "no CUDA device visible");
    CU(nullptr, cudaSetDevice(device));
    SynthParams p{};
    p.seed = seed; p.n_total = 0xFFFFFFFFull; p.n_plants_total = 0; p.plants = 1;
    const size_t bytes = size_t(q1 - q0) * CAB_DIM * sizeof(float);
    float *d = out;
    if (out_loc == CAB_HOST) CU(nullptr, cudaMalloc((void **)&d, bytes));
    launch_synth_rows(p, 2, 0, q0, q1 - q0, d, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && out_loc == CAB_HOST) e = cudaMemcpy(out, d, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (out_loc == CAB_HOST) cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, CAB_ERR_CUDA, "synthetic queries: %s", cudaGetErrorString(e));
    return CAB_OK;
}

int cab_index_read_rows(cab_index *idx, int corpus, int64_t r0, int64_t r1, float *out, int out_loc) {
    CHECK_HANDLE(idx);
    if (!out || r0 < 0 || r1 < r0 || r1 > idx->size || (corpus != 0 && corpus != 1))
        return fail(idx, CAB_ERR_INVALID, "bad row range");
    if (r1 == r0) return CAB_OK;
    CU(idx, cudaSetDevice(idx->device));
    const size_t bytes = size_t(r1 - r0) * CAB_DIM * sizeof(float);
    float *d = out;
    if (out_loc == CAB_HOST) CU(idx, cudaMalloc((void **)&d, bytes));
    launch_widen_rows(corpus == 0 ? idx->asr : idx->audio, idx->dtype, r0, r1 - r0, d, idx->own_stream);
    idx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && out_loc == CAB_HOST) e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, idx->own_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->own_stream);
    if (out_loc == CAB_HOST) cudaFree(d);
    if (e != cudaSuccess) return fail(idx, CAB_ERR_CUDA, "read_rows: %s", cudaGetErrorString(e));
    return CAB_OK;
}

int cab_index_write_flags(cab_index *idx, int64_t r0, int64_t r1, const uint8_t *flags, int flags_loc) {
    CHECK_HANDLE(idx);
    if (!flags || r0 < 0 || r1 < r0 || r1 > idx->size) return fail(idx, CAB_ERR_INVALID, "bad flag range");
    if (flags_loc != CAB_HOST && flags_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "flags_loc");
    if (r1 == r0) return CAB_OK;
    CU(idx, cudaSetDevice(idx->device));
    CU(idx, cudaMemcpyAsync(idx->flags + r0, flags, size_t(r1 - r0),
                            flags_loc == CAB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, idx->own_stream));
    CU(idx, cudaStreamSynchronize(idx->own_stream));
    return CAB_OK;
}

int cab_index_read_flags(cab_index *idx, int64_t r0, int64_t r1, uint8_t *out, int out_loc) {
    CHECK_HANDLE(idx);
    if (!out || r0 < 0 || r1 < r0 || r1 > idx->size) return fail(idx, CAB_ERR_INVALID, "bad flag range");
    if (out_loc != CAB_HOST && out_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "out_loc");
    if (r1 == r0) return CAB_OK;
    CU(idx, cudaSetDevice(idx->device));
    CU(idx, cudaMemcpyAsync(out, idx->flags + r0, size_t(r1 - r0),
                            out_loc == CAB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, idx->own_stream));
    CU(idx, cudaStreamSynchronize(idx->own_stream));
    return CAB_OK;
}

// ---- persistent index file ---------------------------------------------------------------------
namespace {
struct FileHeader {
    char magic[8];
    uint32_t version, dim, dtype, reserved;
    uint64_t n_rows, row_base, off_asr, off_audio, off_flags, file_bytes;
    uint64_t off_norm_asr, off_norm_audio;      // version 2: fp32 original row lengths (raw dot-product scoring)
};
constexpr char kMagic[8] = {'C', 'A', 'B', 'I', 'D', 'X', '0', '1'};
constexpr size_t kFileAlign = 4096;
constexpr size_t kIoChunk = size_t(32) << 20;

int read_header(const char *path, FILE **fp, FileHeader *h, cab_index *idx) {
    if (!path) return fail(idx, CAB_ERR_INVALID, "null path");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(idx, CAB_ERR_INVALID, "cannot open '%s' for reading", path);
    uint8_t raw[kFileAlign];
    if (fread(raw, 1, kFileAlign, f) != kFileAlign) { fclose(f); return fail(idx, CAB_ERR_INVALID, "'%s': truncated header", path); }
    memcpy(h, raw, sizeof(FileHeader));
    const size_t eb = h->dtype == CAB_BF16 ? 2 : 4;
    bool ok = memcmp(h->magic, kMagic, 8) == 0 && (h->version == 1 || h->version == 2) && h->dim == CAB_DIM &&
              (h->dtype == CAB_F32 || h->dtype == CAB_BF16) && h->n_rows <= 0xFFFFFFF0ull &&
              h->off_asr == kFileAlign && h->off_audio >= h->off_asr + h->n_rows * CAB_DIM * eb &&
              h->off_flags >= h->off_audio + h->n_rows * CAB_DIM * eb && h->file_bytes >= h->off_flags + h->n_rows;
    if (ok && h->version == 2)
        ok = h->off_norm_asr >= h->off_flags + h->n_rows && h->off_norm_audio >= h->off_norm_asr + h->n_rows * 4 &&
             h->file_bytes >= h->off_norm_audio + h->n_rows * 4;
    if (ok && h->version == 1) h->off_norm_asr = h->off_norm_audio = 0;
    if (ok) {
        fseek(f, 0, SEEK_END);
        ok = uint64_t(ftell(f)) >= h->file_bytes;
    }
    if (!ok) { fclose(f); return fail(idx, CAB_ERR_INVALID, "'%s' is not a valid cab index file", path); }
    *fp = f;
    return CAB_OK;
}
}  // namespace

int cab_index_file_info(const char *path, int *dim, int *dtype, int64_t *n_rows, int64_t *row_base) {
    FILE *f = nullptr;
    FileHeader h;
    int rc = read_header(path, &f, &h, nullptr);
    if (rc != CAB_OK) return rc;
    fclose(f);
    if (dim) *dim = int(h.dim);
    if (dtype) *dtype = int(h.dtype);
    if (n_rows) *n_rows = int64_t(h.n_rows);
    if (row_base) *row_base = int64_t(h.row_base);
    return CAB_OK;
}

int cab_index_save(cab_index *idx, const char *path) {
    CHECK_HANDLE(idx);
    if (!path) return fail(idx, CAB_ERR_INVALID, "null path");
    CU(idx, cudaSetDevice(idx->device));
    const size_t row_bytes = CAB_DIM * elem_size(idx->dtype);
    const uint64_t n = uint64_t(idx->size);
    FileHeader h{};
    memcpy(h.magic, kMagic, 8);
    h.version = 2; h.dim = CAB_DIM; h.dtype = uint32_t(idx->dtype); h.n_rows = n; h.row_base = uint64_t(idx->row_base);
    h.off_asr = kFileAlign;
    h.off_audio = align_up(h.off_asr + n * row_bytes, kFileAlign);
    h.off_flags = align_up(h.off_audio + n * row_bytes, kFileAlign);
    h.off_norm_asr = align_up(h.off_flags + n, kFileAlign);
    h.off_norm_audio = align_up(h.off_norm_asr + n * 4, kFileAlign);
    h.file_bytes = align_up(h.off_norm_audio + n * 4, kFileAlign);
    int rc = ensure_pinned(idx, &idx->h_rows, &idx->h_rows_bytes, kIoChunk);
    if (rc != CAB_OK) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return fail(idx, CAB_ERR_INVALID, "cannot open '%s' for writing", path);
    std::vector<uint8_t> zeros(kFileAlign, 0);
    memcpy(zeros.data(), &h, sizeof h);
    bool ok = fwrite(zeros.data(), 1, kFileAlign, f) == kFileAlign;
    memset(zeros.data(), 0, kFileAlign);
    const struct { const void *src; uint64_t off, bytes; } sect[5] = {
        {idx->asr, h.off_asr, n * row_bytes}, {idx->audio, h.off_audio, n * row_bytes}, {idx->flags, h.off_flags, n},
        {idx->norm_asr, h.off_norm_asr, n * 4}, {idx->norm_audio, h.off_norm_audio, n * 4}};
    uint64_t pos = kFileAlign;
    for (int s3 = 0; s3 < 5 && ok; ++s3) {
        for (; pos < sect[s3].off && ok; ) { size_t m = std::min<uint64_t>(kFileAlign, sect[s3].off - pos); ok = fwrite(zeros.data(), 1, m, f) == m; pos += m; }
        for (uint64_t done = 0; done < sect[s3].bytes && ok; ) {
            const size_t m = size_t(std::min<uint64_t>(kIoChunk, sect[s3].bytes - done));
            cudaError_t e = cudaMemcpyAsync(idx->h_rows, static_cast<const uint8_t *>(sect[s3].src) + done, m, cudaMemcpyDeviceToHost, idx->own_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(idx->own_stream);
            if (e != cudaSuccess) { fclose(f); return fail(idx, CAB_ERR_CUDA, "save: %s", cudaGetErrorString(e)); }
            ok = fwrite(idx->h_rows, 1, m, f) == m;
            done += m; pos += m;
        }
    }
    for (; pos < h.file_bytes && ok; ) { size_t m = std::min<uint64_t>(kFileAlign, h.file_bytes - pos); ok = fwrite(zeros.data(), 1, m, f) == m; pos += m; }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(idx, CAB_ERR_INVALID, "write to '%s' failed", path);
    return CAB_OK;
}

int cab_index_load(const char *path, int device, int64_t r0, int64_t r1, cab_index **out) {
    if (!out) return fail(nullptr, CAB_ERR_INVALID, "out is null");
    *out = nullptr;
    FILE *f = nullptr;
    FileHeader h;
    int rc = read_header(path, &f, &h, nullptr);
    if (rc != CAB_OK) return rc;
    if (r1 < 0) r1 = int64_t(h.n_rows);
    if (r0 < 0 || r0 > r1 || uint64_t(r1) > h.n_rows) { fclose(f); return fail(nullptr, CAB_ERR_INVALID, "row range [%lld, %lld) outside the file's %llu rows", (long long)r0, (long long)r1, (unsigned long long)h.n_rows); }
    cab_index *idx = nullptr;
    rc = cab_index_create(int(h.dim), int(h.dtype), r1 - r0, device, &idx);
    if (rc != CAB_OK) { fclose(f); return rc; }
    const size_t row_bytes = CAB_DIM * elem_size(idx->dtype);
    const uint64_t n = uint64_t(r1 - r0);
    rc = ensure_pinned(idx, &idx->h_rows, &idx->h_rows_bytes, kIoChunk);
    const struct { void *dst; uint64_t off, bytes; } sect[5] = {
        {idx->asr, h.off_asr + uint64_t(r0) * row_bytes, n * row_bytes},
        {idx->audio, h.off_audio + uint64_t(r0) * row_bytes, n * row_bytes},
        {idx->flags, h.off_flags + uint64_t(r0), n},
        {idx->norm_asr, h.off_norm_asr + uint64_t(r0) * 4, n * 4},
        {idx->norm_audio, h.off_norm_audio + uint64_t(r0) * 4, n * 4}};
    const int n_sect = h.version >= 2 ? 5 : 3;
    if (rc == CAB_OK && h.version < 2 && n > 0) {
        // version-1 file: the original row lengths were not recorded; its rows are served as unit length
        std::vector<float> ones(size_t(std::min<uint64_t>(n, 1u << 20)), 1.0f);
        for (uint64_t done = 0; done < n && rc == CAB_OK; done += ones.size()) {
            const size_t m = size_t(std::min<uint64_t>(ones.size(), n - done));
            cudaError_t e = cudaMemcpy(idx->norm_asr + done, ones.data(), m * 4, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(idx->norm_audio + done, ones.data(), m * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) rc = fail(nullptr, CAB_ERR_CUDA, "load: %s", cudaGetErrorString(e));
        }
    }
    for (int s3 = 0; s3 < n_sect && rc == CAB_OK; ++s3) {
        if (fseek(f, long(sect[s3].off), SEEK_SET) != 0) { rc = fail(nullptr, CAB_ERR_INVALID, "seek in '%s' failed", path); break; }
        for (uint64_t done = 0; done < sect[s3].bytes; ) {
            const size_t m = size_t(std::min<uint64_t>(kIoChunk, sect[s3].bytes - done));
            if (fread(idx->h_rows, 1, m, f) != m) { rc = fail(nullptr, CAB_ERR_INVALID, "'%s': short read", path); break; }
            cudaError_t e = cudaMemcpyAsync(static_cast<uint8_t *>(sect[s3].dst) + done, idx->h_rows, m, cudaMemcpyHostToDevice, idx->own_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(idx->own_stream);
            if (e != cudaSuccess) { rc = fail(nullptr, CAB_ERR_CUDA, "load: %s", cudaGetErrorString(e)); break; }
            done += m;
        }
    }
    fclose(f);
    if (rc != CAB_OK) { cab_index_destroy(idx); return rc; }
    idx->size = int64_t(n);
    idx->row_base = int64_t(h.row_base) + r0;
    *out = idx;
    return CAB_OK;
}

// ---- search ------------------------------------------------------------------------------------
struct OutLayout {
    size_t index, fusion, asr, audio, flags, count, nonfinite, done, total;
};
static OutLayout out_layout(int nq, int k) {
    OutLayout L;
    size_t o = 0;
    L.index = o; o += align_up(size_t(nq) * k * 8, 256);
    L.fusion = o; o += align_up(size_t(nq) * k * 8, 256);
    L.asr = o; o += align_up(size_t(nq) * k * 4, 256);
    L.audio = o; o += align_up(size_t(nq) * k * 4, 256);
    L.flags = o; o += align_up(size_t(nq) * k, 256);
    L.count = o; o += align_up(size_t(nq) * 4, 256);
    L.nonfinite = o; o += 128;
    L.done = o; o += 128;
    L.total = o;
    return L;
}

template <typename T>
static int ensure_dev(cab_index *idx, T **p, size_t *have, size_t want) {
    if (*have >= want) return CAB_OK;
    cudaFree(*p); *p = nullptr; *have = 0;
    want = align_up(want, 256);
    CU(idx, cudaMalloc((void **)p, want));
    *have = want;
    return CAB_OK;
}

static int ensure_workspace(cab_index *idx, int nq, int k, int n_partials, int scan_batch, int slots, size_t gemm_ws) {
    int rc;
    if ((rc = ensure_dev(idx, &idx->d_params, &idx->sz_params, size_t(nq) * (24 + CAB_DIM * 4)))) return rc;
    if ((rc = ensure_dev(idx, &idx->d_partial_keys, &idx->sz_pkeys, size_t(scan_batch) * n_partials * slots * 8))) return rc;
    if ((rc = ensure_dev(idx, &idx->d_cands, &idx->sz_cands, size_t(nq) * k * sizeof(cab_candidate)))) return rc;
    if ((rc = ensure_dev(idx, &idx->d_out, &idx->d_out_bytes, out_layout(nq, k).total))) return rc;
    if (gemm_ws && (rc = ensure_dev(idx, &idx->d_gemm_ws, &idx->d_gemm_ws_bytes, gemm_ws))) return rc;
    return CAB_OK;
}

// Resolve output pointers for the emit stage: user device pointers; or, for host results, the
// handle's MAPPED PINNED host block (the kernel writes the results straight into host memory over
// PCIe and raises a completion flag there -- no D2H copy, no stream synchronisation); device
// scratch for outputs the caller did not ask for.
static int make_emit(cab_index *idx, int n_lists, int nq, int k, double threshold,
                     int64_t *out_index, double *out_fusion, float *out_asr, float *out_audio,
                     uint8_t *out_flags, int32_t *out_count, int out_loc, EmitArgs *out) {
    const OutLayout L = out_layout(nq, k);
    EmitArgs ea{};
    ea.n_lists = n_lists; ea.n_queries = nq; ea.k = k;
    ea.w_asr = idx->d_w64; ea.w_audio = idx->d_w64 + nq; ea.threshold = threshold;
    const bool dev = out_loc == CAB_DEVICE;
    uint8_t *d = idx->d_out;
    if (!dev) {
        int rc = ensure_pinned(idx, &idx->h_out, &idx->h_out_bytes, L.total);
        if (rc != CAB_OK) return rc;
        d = idx->h_out;                                   // UVA: pinned host memory is device-addressable
        ea.done_flag = reinterpret_cast<uint32_t *>(d + L.done);
        ea.done_epoch = ++idx->host_epoch ? idx->host_epoch : ++idx->host_epoch;
        ea.done_counter = idx->d_host_done;
        *reinterpret_cast<volatile uint32_t *>(d + L.done) = 0u;
        *reinterpret_cast<volatile int *>(d + L.nonfinite) = 0;          // a CTA whose query holds NaN/Inf stores 1
        ea.nonfinite_out = reinterpret_cast<int *>(d + L.nonfinite);
    }
    ea.out_index = dev && out_index ? out_index : reinterpret_cast<int64_t *>(d + L.index);
    ea.out_fusion = dev && out_fusion ? out_fusion : reinterpret_cast<double *>(d + L.fusion);
    ea.out_asr = dev && out_asr ? out_asr : reinterpret_cast<float *>(d + L.asr);
    ea.out_audio = dev && out_audio ? out_audio : reinterpret_cast<float *>(d + L.audio);
    ea.out_flags = dev && out_flags ? out_flags : d + L.flags;
    ea.out_count = dev && out_count ? out_count : reinterpret_cast<int32_t *>(d + L.count);
    *out = ea;
    return CAB_OK;
}

// fp32 scan weight w / (w_asr + w_audio).  The scan tells "this pipeline's query weight is 0"
// (rows that only have this pipeline are skipped, :659-661) from "positive" by the sign of the
// staged value, so a positive weight never rounds to 0.
static float stage_weight(double w, double tot) {
    if (!(w > 0) || !(tot > 0)) return 0.f;
    return std::max(float(w / tot), FLT_MIN);
}

// One H2D copy of the per-search parameter block (weights, and the queries if they are on the host).
static int stage_params(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
                        const double *w_audio, int nq, cudaStream_t s, const float **dq) {
    const size_t wbytes = size_t(nq) * 24;
    const size_t qbytes = queries && queries_loc == CAB_HOST ? size_t(nq) * CAB_DIM * 4 : 0;
    int rc = ensure_pinned(idx, &idx->h_in, &idx->h_in_bytes, wbytes + qbytes);
    if (rc != CAB_OK) return rc;
    if (idx->ev_in_pending) { CU(idx, cudaEventSynchronize(idx->ev_in)); idx->ev_in_pending = false; }
    double *h64 = reinterpret_cast<double *>(idx->h_in);
    float *h32 = reinterpret_cast<float *>(idx->h_in + size_t(nq) * 16);
    for (int i = 0; i < nq; ++i) {
        h64[i] = w_asr[i]; h64[nq + i] = w_audio[i];
        const double tot = w_asr[i] + w_audio[i];
        h32[i] = stage_weight(w_asr[i], tot);
        h32[nq + i] = stage_weight(w_audio[i], tot);
    }
    if (qbytes) memcpy(idx->h_in + wbytes, queries, qbytes);
    CU(idx, cudaMemcpyAsync(idx->d_params, idx->h_in, wbytes + qbytes, cudaMemcpyHostToDevice, s));
    CU(idx, cudaEventRecord(idx->ev_in, s));
    idx->ev_in_pending = true;
    idx->d_w64 = reinterpret_cast<double *>(idx->d_params);
    idx->d_w32 = reinterpret_cast<float *>(idx->d_params + size_t(nq) * 16);
    idx->staged_nq = nq;
    idx->inline_w_valid = false;
    if (dq) *dq = qbytes ? reinterpret_cast<const float *>(idx->d_params + wbytes) : queries;
    return CAB_OK;
}

// Smallest batch that PATH_AUTO sends to the tensor cores.  Measured crossover (tools/crossover_probe.py,
// profiles/r02_crossover.md): the tensor-core scan costs about the same for 2 as for 256 queries and
// beats the register-tiled GEMV (4 queries per corpus pass) from 4 queries on, at every library size
// from 100 K segments; below ~64 K segments its fixed cost (cluster launch, TMEM, tensor maps) loses.
static int64_t gemm_min_queries(const cab_index *idx) {
    if (idx->opt_gemm_min_queries > 0) return idx->opt_gemm_min_queries;
    return idx->size >= 65536 ? 4 : 64;
}

// fp32 library with bf16 shadows: rows preselected per query by the tensor-core scan (wide margin:
// re-scoring 96 rows costs microseconds, a failed certificate a corpus pass), and whether the route
// applies to this index / k at all.
static int shadow_k_sel(int k) { return std::min(CAB_MAX_K, std::max(96, k + k / 2 + 32)); }
static bool shadow_possible(const cab_index *idx, int k) {
    return idx->dtype == CAB_F32 && idx->shadow_asr && !idx->opt_raw_dot && shadow_k_sel(k) >= k + 16 && idx->size > 0;
}

// Stage parameters; run scan + finalize -> idx->d_cands[nq x k]; with `out` != null (single GPU)
// the finalize kernel also emits the final results.
struct UserOut {
    int64_t *index; double *fusion; float *asr; float *audio; uint8_t *flags; int32_t *count; int loc;
};
// Sharded search: where the finalize kernel pushes this shard's candidates; when the whole call is
// one scan pass of few queries the merge of all shards' candidates rides in the same kernel.
struct Sharded {
    PeerPush pp;
    bool fused;              // out: the finalize kernel waited for the world's flags and emitted the results
};
constexpr int kFusedMergeMaxQueries = 64;      // CTAs that spin on peer flags must all be co-resident (one per SM)

static int run_local(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
                     const double *w_audio, int nq, int k, double threshold, int path,
                     const UserOut *out, cab_candidate *cand_dst, Sharded *sh, cudaStream_t s) {
    if (!queries || !w_asr || !w_audio) return fail(idx, CAB_ERR_INVALID, "queries / weights are null");
    if (queries_loc != CAB_HOST && queries_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "queries_loc");
    if (nq <= 0 || nq > CAB_MAX_QUERIES) return fail(idx, CAB_ERR_INVALID, "n_queries must be in 1..%d", CAB_MAX_QUERIES);
    if (k <= 0 || k > CAB_MAX_K) return fail(idx, CAB_ERR_INVALID, "k must be in 1..%d", CAB_MAX_K);
    if (!std::isfinite(threshold)) return fail(idx, CAB_ERR_INVALID, "threshold must be finite");
    if (path != CAB_PATH_AUTO && path != CAB_PATH_GEMV && path != CAB_PATH_GEMM) return fail(idx, CAB_ERR_INVALID, "unknown path %d", path);
    if (out && out->loc != CAB_HOST && out->loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "out_loc");
    for (int i = 0; i < nq; ++i)
        if (!(std::isfinite(w_asr[i]) && std::isfinite(w_audio[i]) && w_asr[i] >= 0 && w_audio[i] >= 0))
            return fail(idx, CAB_ERR_INVALID, "weights must be finite and >= 0");
    for (int i = 0; i < nq; ++i)
        if (!(w_asr[i] + w_audio[i] > 0)) return fail(idx, CAB_ERR_INVALID, "w_asr + w_audio must be > 0 for every query");
    // fp32 library with bf16 shadows: the tensor cores PRE-select k_sel > k rows per query from the
    // shadows, the finalize kernel re-scores them exactly from the fp32 rows and certifies each query
    // (FinalizeArgs::cert_out); only a plain cab_search takes this route (the caller of run_local
    // re-runs uncertified queries on the exact scan).
    const int k_sel_shadow = shadow_k_sel(k);
    const bool shadow_ok = shadow_possible(idx, k) && out && !sh && !cand_dst;
    bool use_gemm = false;
    if (path == CAB_PATH_GEMM) {
        if (idx->dtype != CAB_BF16 && !shadow_ok)
            return fail(idx, CAB_ERR_INVALID, "the tensor-core path needs a bf16 index (or an fp32 index with option tensor_core_shadow, "
                                              "k <= 112, through cab_search)");
        if (idx->opt_raw_dot) return fail(idx, CAB_ERR_INVALID, "raw dot-product scoring (option raw_dot) is served by the GEMV path only");
        if (!gemm_path_available()) return fail(idx, CAB_ERR_INVALID, "tensor-core path not built");
        use_gemm = true;
    } else if (path == CAB_PATH_AUTO) {
        use_gemm = (idx->dtype == CAB_BF16 || shadow_ok) && nq >= gemm_min_queries(idx) && gemm_path_available() && !idx->opt_raw_dot;
    }
    const bool certified = use_gemm && idx->dtype == CAB_F32;
    const int k_sel = certified ? k_sel_shadow : k;                // rows selected by the scan per query
    constexpr float kShadowEps = 4.0e-3f;                          // >= 2^-8 + 2^-18: |bf16 x bf16 cosine - fp32 cosine| of unit vectors
    CU(idx, cudaSetDevice(idx->device));
    const int n_partials = use_gemm ? gemm_partials_per_query(idx->sm_count) : gemv_max_grid(idx->sm_count);
    const int batch = use_gemm ? std::min(nq, kGemmQueriesPerPass) : int(std::min<int64_t>(nq, idx->opt_gemv_batch));
    int rc = ensure_workspace(idx, nq, k_sel, n_partials, use_gemm ? kGemmQueriesPerPass : batch,
                              use_gemm ? kGemmListCap : k, use_gemm ? gemm_workspace_bytes(nq, k_sel, idx->sm_count) : 0);
    if (rc != CAB_OK) return rc;
    if (certified && (rc = ensure_dev(idx, &idx->d_cert, &idx->sz_cert, size_t(nq)))) return rc;
    idx->cert_pending = certified ? nq : 0;
    // One query: parameters travel in the kernel arguments, no H2D copy ahead of the scan.
    InlineParams inl{};
    const float *dq = nullptr;
    if (nq == 1 && !use_gemm) {
        const double tot = w_asr[0] + w_audio[0];
        inl.use_weights = 1;
        inl.wa32 = stage_weight(w_asr[0], tot);
        inl.wb32 = stage_weight(w_audio[0], tot);
        inl.w64_asr = w_asr[0]; inl.w64_audio = w_audio[0];
        if (queries_loc == CAB_HOST) { inl.use_query = 1; memcpy(inl.q, queries, sizeof inl.q); dq = reinterpret_cast<const float *>(idx->d_params); }
        else dq = queries;
        idx->d_w64 = reinterpret_cast<double *>(idx->d_params);          // not read when inline_weights is set
        idx->d_w32 = reinterpret_cast<float *>(idx->d_params + 16);
        idx->staged_nq = 0;
        idx->inline_w_valid = true; idx->inline_w64[0] = w_asr[0]; idx->inline_w64[1] = w_audio[0];
    } else if ((rc = stage_params(idx, queries, queries_loc, w_asr, w_audio, nq, s, &dq))) return rc;

    const PeerPush *peer = sh ? &sh->pp : nullptr;
    const bool fuse_merge = sh && out && nq <= batch && nq <= kFusedMergeMaxQueries && idx->size > 0;
    const bool emit_here = out && (!sh || fuse_merge);             // the finalize kernel writes the results
    if (sh) sh->fused = fuse_merge;
    EmitArgs ea{};
    if (emit_here && (rc = make_emit(idx, fuse_merge ? peer->world : 1, nq, k, threshold, out->index, out->fusion,
                                     out->asr, out->audio, out->flags, out->count, out->loc, &ea))) return rc;
    if (inl.use_weights) { ea.inline_weights = 1; ea.w64_asr = inl.w64_asr; ea.w64_audio = inl.w64_audio; }
    if (fuse_merge) {
        ea.cands = peer->bufs[peer->rank];
        ea.wait_flags = peer->flags[peer->rank] + peer->parity * peer->world;
        ea.wait_epoch = peer->epoch;
        ea.status = idx->d_status;
        if (idx->opt_stamp_exchange && idx->d_stamps) ea.stamps = idx->d_stamps + size_t(idx->stamp_calls++ % kStampRows) * kStampCols;
    }
    idx->timed = false;
    cab_candidate *cands = cand_dst ? cand_dst : idx->d_cands;      // cab_search_candidates: straight into the caller's block
    if (idx->size == 0) {
        // nothing to scan: every candidate slot is empty
        CU(idx, cudaMemsetAsync(cands, 0xFF, size_t(nq) * k * sizeof(cab_candidate), s));
        if (emit_here) { ea.cands = cands; launch_emit(ea, s); idx->launches += 1; CU(idx, cudaGetLastError()); }
        if (peer) { PeerPush pp = *peer; pp.q0 = 0; pp.signal = 1; launch_peer_push(cands, nq, k, pp, s); idx->launches += 1; CU(idx, cudaGetLastError()); }
        return CAB_OK;
    }

    ScanArgs sa{};
    sa.asr = idx->asr; sa.audio = idx->audio; sa.flags = idx->flags; sa.n_rows = idx->size;
    if (idx->opt_raw_dot) { sa.norm_asr = idx->norm_asr; sa.norm_audio = idx->norm_audio; }
    sa.dtype = idx->dtype; sa.k = k_sel;
    sa.select_threshold = float(threshold) - 1e-6f;
    if (certified) {                                   // scan the shadows; nothing above (threshold - eps) may be missed
        sa.asr = idx->shadow_asr; sa.audio = idx->shadow_audio; sa.dtype = CAB_BF16;
        sa.select_threshold = float(threshold) - kShadowEps - 1e-6f;
    }
    sa.partial_keys = idx->d_partial_keys;
    sa.inl = inl;
    // A device query may have been written by the kernel right in front of the scan on the caller's
    // stream: the scan then waits for its predecessor before its first global read.  A host query
    // travels in the kernel arguments / through an ordinary memcpy, and "queries_settled" is the
    // caller's promise that device queries were complete before the previous search was issued.
    sa.wait_early = (queries_loc == CAB_DEVICE && !idx->opt_queries_settled) ? 1 : 0;
    sa.n_partials = n_partials; sa.work_counters = idx->d_counters; sa.chunk_rows = idx->opt_chunk_rows ? int(idx->opt_chunk_rows) : (idx->dtype == CAB_BF16 ? 64 : 32);
    FinalizeArgs fa{};
    fa.inl = inl;
    fa.asr = idx->asr; fa.audio = idx->audio; fa.flags = idx->flags; fa.dtype = idx->dtype;
    fa.norm_asr = sa.norm_asr; fa.norm_audio = sa.norm_audio;
    fa.row_base = idx->row_base; fa.k = k_sel; fa.partial_keys = idx->d_partial_keys;
    if (certified) fa.cert_eps = kShadowEps;
    fa.n_partials = n_partials; fa.force_general = int(idx->opt_finalize_general);
    fa.slot_stride = use_gemm ? kGemmListCap : k;
    fa.work_counters = use_gemm ? nullptr : idx->d_counters;
    fa.counts = use_gemm ? reinterpret_cast<const int32_t *>(idx->d_gemm_ws) : nullptr;
    fa.levels = use_gemm ? gemm_levels(idx->d_gemm_ws, idx->sm_count) : nullptr;
    fa.select_threshold = sa.select_threshold; fa.level_step = gemm_level_step();

    if (idx->opt_time_kernels) CU(idx, cudaEventRecord(idx->ev_t0, s));
    for (int q0 = 0; q0 < nq; q0 += batch) {
        const int m = std::min(batch, nq - q0);
        sa.queries = dq + size_t(q0) * CAB_DIM; sa.wa32 = idx->d_w32 + q0; sa.wb32 = idx->d_w32 + nq + q0;
        sa.n_queries = m;
        if (use_gemm) {
            std::string err;
            launch_gemm_scan(sa, idx->sm_count, idx->d_gemm_ws, idx->d_gemm_ws_bytes, s, &err);
            if (!err.empty()) return fail(idx, CAB_ERR_CUDA, "%s", err.c_str());
        } else {
            const GemvPlan plan = plan_gemv(idx->gemv, idx->dtype, m, idx->sm_count);
            sa.n_partials = fa.n_partials = plan.grid_x;
            launch_gemv_scan(sa, plan, s);
        }
        if (idx->opt_time_kernels && q0 + batch >= nq) CU(idx, cudaEventRecord(idx->ev_t1, s));
        fa.queries = sa.queries; fa.n_queries = m; fa.cands = peer ? nullptr : cands + size_t(q0) * k_sel;
        if (certified) fa.cert_out = idx->d_cert + q0;
        if (peer) { fa.peer = *peer; fa.peer.q0 = q0; fa.peer.signal = q0 + batch >= nq ? 1 : 0; }
        if (emit_here) {
            EmitArgs eb = ea;                           // this batch's slice of the outputs
            eb.n_queries = m;
            eb.w_asr = ea.w_asr + q0; eb.w_audio = ea.w_audio + q0;
            eb.out_index = ea.out_index + size_t(q0) * k; eb.out_fusion = ea.out_fusion + size_t(q0) * k;
            eb.out_asr = ea.out_asr + size_t(q0) * k; eb.out_audio = ea.out_audio + size_t(q0) * k;
            eb.out_flags = ea.out_flags + size_t(q0) * k; eb.out_count = ea.out_count + q0;
            if (q0 + batch < nq) eb.done_epoch = 0;                                   // only the last batch signals completion
            launch_finalize(fa, &eb, idx->sm_count, s);
        } else {
            launch_finalize(fa, nullptr, idx->sm_count, s);
        }
        idx->launches += use_gemm ? 3 : 2;          // (prologue +) scan + finalize
    }
    if (idx->opt_time_kernels) idx->timed = true;
    CU(idx, cudaGetLastError());
    return CAB_OK;
}

// Host results: wait for the kernel's completion flag in the mapped pinned block, then hand the
// values to the caller's arrays (nothing to do for device outputs).
static int finish_outputs(cab_index *idx, int nq, int k, const UserOut &o, cudaStream_t s) {
    if (o.loc == CAB_DEVICE) {
        if (idx->opt_sync) CU(idx, cudaStreamSynchronize(s));
        return CAB_OK;
    }
    const OutLayout L = out_layout(nq, k);
    const uint8_t *h = idx->h_out;
    const volatile uint32_t *flag = reinterpret_cast<const volatile uint32_t *>(h + L.done);
    const uint32_t want = idx->host_epoch;
    bool seen = false;
    for (long spins = 0; spins < 400000000L; ++spins) {               // bounded: falls back to a stream sync
        if (*flag == want) { seen = true; break; }
        if ((spins & 0xFFF) == 0xFFF && cudaStreamQuery(s) != cudaErrorNotReady) break;   // finished or failed
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (!seen) {
        CU(idx, cudaStreamSynchronize(s));
        if (*flag != want) return fail(idx, CAB_ERR_CUDA, "search finished without raising its completion flag");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    idx->ev_in_pending = false;
    if (*reinterpret_cast<const volatile int *>(h + L.nonfinite))
        return fail(idx, CAB_ERR_NONFINITE, "Input contains NaN or infinity (query)");
    const size_t n = size_t(nq) * k;
    if (o.index) memcpy(o.index, h + L.index, n * 8);
    if (o.fusion) memcpy(o.fusion, h + L.fusion, n * 8);
    if (o.asr) memcpy(o.asr, h + L.asr, n * 4);
    if (o.audio) memcpy(o.audio, h + L.audio, n * 4);
    if (o.flags) memcpy(o.flags, h + L.flags, n);
    if (o.count) memcpy(o.count, h + L.count, size_t(nq) * 4);
    return CAB_OK;
}

// Shadow path (fp32 library, tensor-core preselection): read the per-query certificates of the batch
// that just ran and re-run every query that is not provably exact on the exact scan, into the same
// output slots.  Most of the batch uncertified: the whole batch goes to the exact scan.  Otherwise
// the uncertified queries are gathered into one contiguous batch (register-tiled exact scan, 4
// queries per corpus pass), their results come back through host memory and are scattered.
static int rerun_uncertified(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
                             const double *w_audio, int k, double threshold, const UserOut &o, cudaStream_t s) {
    const int nq = idx->cert_pending;
    idx->cert_pending = 0;
    std::vector<uint8_t> cert(size_t(nq), 1);
    CU(idx, cudaMemcpyAsync(cert.data(), idx->d_cert, size_t(nq), cudaMemcpyDeviceToHost, s));
    CU(idx, cudaStreamSynchronize(s));
    std::vector<int> todo;
    for (int i = 0; i < nq; ++i) if (!cert[size_t(i)]) todo.push_back(i);
    const int bad = int(todo.size());
    idx->last_uncertified = bad; idx->total_uncertified += bad; idx->total_shadow_queries += nq;
    if (bad == 0) return CAB_OK;
    int rc;
    if (bad * 4 > nq * 3) {
        if ((rc = run_local(idx, queries, queries_loc, w_asr, w_audio, nq, k, threshold, CAB_PATH_GEMV, &o, nullptr, nullptr, s))) return rc;
        return finish_outputs(idx, nq, k, o, s);
    }
    // gather the queries (device buffer) and their weights
    const size_t qbytes = size_t(CAB_DIM) * sizeof(float);
    float *d_q = nullptr;
    CU(idx, cudaMalloc((void **)&d_q, size_t(bad) * qbytes));
    const size_t nb = size_t(bad);
    std::vector<double> wa(nb), wb(nb);
    cudaError_t e = cudaSuccess;
    for (int j = 0; j < bad && e == cudaSuccess; ++j) {
        wa[size_t(j)] = w_asr[todo[size_t(j)]]; wb[size_t(j)] = w_audio[todo[size_t(j)]];
        e = cudaMemcpyAsync(d_q + size_t(j) * CAB_DIM, queries + size_t(todo[size_t(j)]) * CAB_DIM, qbytes,
                            queries_loc == CAB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s);
    }
    if (e != cudaSuccess) { cudaFree(d_q); return fail(idx, CAB_ERR_CUDA, "gathering uncertified queries: %s", cudaGetErrorString(e)); }
    const size_t n = size_t(bad) * k;
    std::vector<int64_t> t_index(n); std::vector<double> t_fusion(n); std::vector<float> t_asr(n), t_audio(n);
    std::vector<uint8_t> t_flags(n); std::vector<int32_t> t_count(nb);
    const UserOut tmp{t_index.data(), t_fusion.data(), t_asr.data(), t_audio.data(), t_flags.data(), t_count.data(), CAB_HOST};
    rc = run_local(idx, d_q, CAB_DEVICE, wa.data(), wb.data(), bad, k, threshold, CAB_PATH_GEMV, &tmp, nullptr, nullptr, s);
    if (rc == CAB_OK) rc = finish_outputs(idx, bad, k, tmp, s);
    cudaFree(d_q);                                       // synchronises; the scan has finished (host outputs)
    if (rc != CAB_OK) return rc;
    const bool dev = o.loc == CAB_DEVICE;
    auto put = [&](void *dst_base, const void *src_base, size_t elem, size_t count, int j) -> cudaError_t {
        if (!dst_base) return cudaSuccess;
        uint8_t *dst = static_cast<uint8_t *>(dst_base) + size_t(todo[size_t(j)]) * count * elem;
        const uint8_t *src = static_cast<const uint8_t *>(src_base) + size_t(j) * count * elem;
        if (!dev) { memcpy(dst, src, count * elem); return cudaSuccess; }
        return cudaMemcpyAsync(dst, src, count * elem, cudaMemcpyHostToDevice, s);
    };
    for (int j = 0; j < bad && e == cudaSuccess; ++j) {
        e = put(o.index, t_index.data(), 8, size_t(k), j);
        if (e == cudaSuccess) e = put(o.fusion, t_fusion.data(), 8, size_t(k), j);
        if (e == cudaSuccess) e = put(o.asr, t_asr.data(), 4, size_t(k), j);
        if (e == cudaSuccess) e = put(o.audio, t_audio.data(), 4, size_t(k), j);
        if (e == cudaSuccess) e = put(o.flags, t_flags.data(), 1, size_t(k), j);
        if (e == cudaSuccess) e = put(o.count, t_count.data(), 4, 1, j);
    }
    if (e == cudaSuccess && dev) e = cudaStreamSynchronize(s);       // the temporaries die with this frame
    if (e != cudaSuccess) return fail(idx, CAB_ERR_CUDA, "scattering re-run results: %s", cudaGetErrorString(e));
    return CAB_OK;
}

int cab_search(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
               const double *w_audio, int n_queries, int k, double threshold, int path,
               int64_t *out_index, double *out_fusion, float *out_asr, float *out_audio,
               uint8_t *out_flags, int32_t *out_count, int out_loc, void *stream) {
    CHECK_HANDLE(idx);
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    const UserOut o{out_index, out_fusion, out_asr, out_audio, out_flags, out_count, out_loc};
    int rc = enter_stream(idx, s);
    if (rc != CAB_OK) return rc;
    rc = run_local(idx, queries, queries_loc, w_asr, w_audio, n_queries, k, threshold, path, &o, nullptr, nullptr, s);
    if (rc != CAB_OK) return rc;
    if ((rc = finish_outputs(idx, n_queries, k, o, s)) != CAB_OK) { idx->cert_pending = 0; return rc; }
    if (idx->cert_pending) return rerun_uncertified(idx, queries, queries_loc, w_asr, w_audio, k, threshold, o, s);
    return CAB_OK;
}

int cab_search_candidates(cab_index *idx, const float *queries, int queries_loc,
                          const double *w_asr, const double *w_audio, int n_queries, int k,
                          double threshold, int path, cab_candidate *out_device, void *stream) {
    CHECK_HANDLE(idx);
    if (!out_device) return fail(idx, CAB_ERR_INVALID, "out_device is null");
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    int rc = enter_stream(idx, s);
    if (rc != CAB_OK) return rc;
    rc = run_local(idx, queries, queries_loc, w_asr, w_audio, n_queries, k, threshold, path, nullptr, out_device, nullptr, s);
    if (rc != CAB_OK) return rc;
    if (idx->opt_sync) CU(idx, cudaStreamSynchronize(s));
    return CAB_OK;
}

int cab_merge_candidates(cab_index *idx, const cab_candidate *cands_device, int n_lists,
                         int n_queries, int k, const double *w_asr, const double *w_audio,
                         double threshold, int64_t *out_index, double *out_fusion, float *out_asr,
                         float *out_audio, uint8_t *out_flags, int32_t *out_count, int out_loc,
                         void *stream) {
    CHECK_HANDLE(idx);
    if (!cands_device || n_lists <= 0 || (!w_asr) != (!w_audio)) return fail(idx, CAB_ERR_INVALID, "bad merge arguments");
    // w_asr == w_audio == NULL: reuse the weights staged by the preceding cab_search_candidates
    const bool reuse_inline = !w_asr && n_queries == 1 && idx->inline_w_valid;
    if (!w_asr && !reuse_inline && idx->staged_nq != n_queries) return fail(idx, CAB_ERR_INVALID, "no staged weights for %d queries", n_queries);
    if (n_queries <= 0 || n_queries > CAB_MAX_QUERIES || k <= 0 || k > CAB_MAX_K) return fail(idx, CAB_ERR_INVALID, "bad n_queries / k");
    if (size_t(n_lists) * k > 1024) return fail(idx, CAB_ERR_INVALID, "n_lists * k must be <= 1024");
    if (out_loc != CAB_HOST && out_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "out_loc");
    CU(idx, cudaSetDevice(idx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    int rc = enter_stream(idx, s);
    if (rc != CAB_OK) return rc;
    if ((rc = ensure_workspace(idx, n_queries, k, 1, 1, k, 0))) return rc;
    if (w_asr && (rc = stage_params(idx, nullptr, CAB_DEVICE, w_asr, w_audio, n_queries, s, nullptr))) return rc;
    const UserOut o{out_index, out_fusion, out_asr, out_audio, out_flags, out_count, out_loc};
    EmitArgs ea{};
    if ((rc = make_emit(idx, n_lists, n_queries, k, threshold, out_index, out_fusion, out_asr,
                        out_audio, out_flags, out_count, out_loc, &ea))) return rc;
    ea.cands = cands_device;
    if (reuse_inline) { ea.inline_weights = 1; ea.w64_asr = idx->inline_w64[0]; ea.w64_audio = idx->inline_w64[1]; }
    launch_emit(ea, s);
    idx->launches += 1;
    CU(idx, cudaGetLastError());
    return finish_outputs(idx, n_queries, k, o, s);
}

// ---- legacy modes: every segment's fused score ---------------------------------------------------
int cab_score_all(cab_index *idx, const float *queries, int queries_loc, int n_queries,
                  const float *class_weights, float *out, int out_loc, void *stream) {
    CHECK_HANDLE(idx);
    if (!queries || !class_weights || !out) return fail(idx, CAB_ERR_INVALID, "queries / class_weights / out are null");
    if (queries_loc != CAB_HOST && queries_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "queries_loc");
    if (out_loc != CAB_HOST && out_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "out_loc");
    if (n_queries <= 0 || n_queries > CAB_MAX_QUERIES) return fail(idx, CAB_ERR_INVALID, "n_queries must be in 1..%d", CAB_MAX_QUERIES);
    for (int i = 0; i < n_queries * 8; ++i)
        if (!std::isfinite(class_weights[i])) return fail(idx, CAB_ERR_INVALID, "class weights must be finite");
    if (idx->size == 0) return CAB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    CU(idx, cudaSetDevice(idx->device));
    {
        int rc = enter_stream(idx, s);
        if (rc != CAB_OK) return rc;
    }
    const bool host_out = out_loc == CAB_HOST;
    const size_t per_query = size_t(idx->size) * sizeof(float);
    float *d_out = out;
    if (host_out) {
        int rc = ensure_dev(idx, &idx->d_scores, &idx->d_scores_bytes, per_query * n_queries);
        if (rc != CAB_OK) return rc;
        d_out = reinterpret_cast<float *>(idx->d_scores);
    }
    ScoreAllArgs a{};
    a.asr = idx->asr; a.audio = idx->audio; a.flags = idx->flags; a.n_rows = idx->size; a.dtype = idx->dtype;
    a.nonfinite = host_out ? idx->d_nonfinite : nullptr;     // device results: a bad query gives NaN scores
    a.work_counters = idx->d_counters + 64;
    a.chunk_rows = idx->opt_chunk_rows ? int(idx->opt_chunk_rows) : (idx->dtype == CAB_BF16 ? 64 : 32);
    idx->timed = false;
    if (idx->opt_time_kernels) CU(idx, cudaEventRecord(idx->ev_t0, s));
    // Several queries: register-tiled, 4 (then 2, then 1) per corpus pass; host queries are staged once.
    const float *dq = queries;
    if (queries_loc == CAB_HOST && n_queries > 1) {
        const size_t qbytes = size_t(n_queries) * CAB_DIM * sizeof(float);
        int rc = ensure_dev(idx, &idx->d_params, &idx->sz_params, qbytes);
        if (rc != CAB_OK) return rc;
        idx->staged_nq = 0; idx->inline_w_valid = false;            // the parameter block no longer holds search weights
        CU(idx, cudaMemcpyAsync(idx->d_params, queries, qbytes, cudaMemcpyHostToDevice, s));
        dq = reinterpret_cast<const float *>(idx->d_params);
    }
    a.out_stride = idx->size;
    for (int q = 0; q < n_queries; ) {
        const int m = n_queries - q >= 3 ? std::min(4, n_queries - q) : n_queries - q;
        a.n_q = m;
        if (queries_loc == CAB_HOST && n_queries == 1) { a.use_inline_query = 1; memcpy(a.q, queries, sizeof a.q); a.query = nullptr; }
        else { a.use_inline_query = 0; a.query = dq + size_t(q) * CAB_DIM; }
        memset(a.class_w, 0, sizeof a.class_w);
        memcpy(a.class_w, class_weights + size_t(q) * 8, size_t(m) * 8 * sizeof(float));
        a.out = d_out + size_t(q) * idx->size;
        launch_score_all(a, idx->sm_count, s);
        idx->launches += 1;
        q += m;
    }
    if (idx->opt_time_kernels) { CU(idx, cudaEventRecord(idx->ev_t1, s)); idx->timed = true; }
    CU(idx, cudaGetLastError());
    if (!host_out) {
        if (idx->opt_sync) CU(idx, cudaStreamSynchronize(s));
        return CAB_OK;
    }
    int bad = 0;
    CU(idx, cudaMemcpyAsync(out, d_out, per_query * n_queries, cudaMemcpyDeviceToHost, s));
    CU(idx, cudaMemcpyAsync(&bad, idx->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(idx, cudaStreamSynchronize(s));
    if (bad) {
        CU(idx, cudaMemsetAsync(idx->d_nonfinite, 0, sizeof(int), s));
        CU(idx, cudaStreamSynchronize(s));
        return fail(idx, CAB_ERR_NONFINITE, "Input contains NaN or infinity (query)");
    }
    return CAB_OK;
}

// ---- peer-memory exchange ----------------------------------------------------------------------
static size_t peer_flags_bytes() { return 256; }
static size_t peer_half_elems(const cab_index *idx) { return size_t(idx->peer_world) * idx->peer_qcap * idx->peer_kcap; }

int cab_peer_init(cab_index *idx, int rank, int world, int max_queries, int max_k, void *ipc_handle_out) {
    CHECK_HANDLE(idx);
    if (!ipc_handle_out || world < 1 || world > CAB_MAX_WORLD || rank < 0 || rank >= world ||
        max_queries < 1 || max_queries > CAB_MAX_QUERIES || max_k < 1 || max_k > CAB_MAX_K)
        return fail(idx, CAB_ERR_INVALID, "bad peer_init arguments");
    if (idx->d_peer) return fail(idx, CAB_ERR_INVALID, "peer exchange already initialised on this index");
    static_assert(sizeof(cudaIpcMemHandle_t) == CAB_IPC_HANDLE_BYTES, "IPC handle size");
    CU(idx, cudaSetDevice(idx->device));
    idx->peer_world = world; idx->peer_rank = rank; idx->peer_qcap = max_queries; idx->peer_kcap = max_k;
    const size_t bytes = peer_flags_bytes() + 2 * peer_half_elems(idx) * sizeof(cab_candidate);
    CU(idx, cudaMalloc((void **)&idx->d_peer, bytes));
    CU(idx, cudaMemset(idx->d_peer, 0, bytes));
    CU(idx, cudaMalloc((void **)&idx->d_done, sizeof(unsigned int)));
    CU(idx, cudaMemset(idx->d_done, 0, sizeof(unsigned int)));
    CU(idx, cudaMalloc((void **)&idx->d_status, sizeof(int)));
    CU(idx, cudaMemset(idx->d_status, 0, sizeof(int)));
    CU(idx, cudaMalloc((void **)&idx->d_stamps, size_t(kStampRows) * kStampCols * sizeof(unsigned long long)));
    CU(idx, cudaMemset(idx->d_stamps, 0, size_t(kStampRows) * kStampCols * sizeof(unsigned long long)));
    CU(idx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(idx, cudaIpcGetMemHandle(&h, idx->d_peer));
    memcpy(ipc_handle_out, &h, sizeof h);
    return CAB_OK;
}

int cab_peer_attach(cab_index *idx, const void *all_handles) {
    CHECK_HANDLE(idx);
    if (!idx->d_peer || !all_handles) return fail(idx, CAB_ERR_INVALID, "cab_peer_init first");
    if (idx->peer_attached) return fail(idx, CAB_ERR_INVALID, "peers already attached");
    CU(idx, cudaSetDevice(idx->device));
    for (int r = 0; r < idx->peer_world; ++r) {
        if (r == idx->peer_rank) { idx->peer_ptr[r] = idx->d_peer; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t *>(all_handles) + size_t(r) * CAB_IPC_HANDLE_BYTES, sizeof h);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(idx, CAB_ERR_CUDA, "cannot map rank %d's exchange buffer: %s", r, cudaGetErrorString(e)); }
        idx->peer_ptr[r] = static_cast<uint8_t *>(ptr);
    }
    idx->peer_attached = true;
    return CAB_OK;
}

int cab_search_sharded(cab_index *idx, const float *queries, int queries_loc, const double *w_asr,
                       const double *w_audio, int n_queries, int k, double threshold, int path,
                       int64_t *out_index, double *out_fusion, float *out_asr, float *out_audio,
                       uint8_t *out_flags, int32_t *out_count, int out_loc, void *stream) {
    CHECK_HANDLE(idx);
    if (!idx->peer_attached) return fail(idx, CAB_ERR_INVALID, "cab_peer_init / cab_peer_attach first");
    if (n_queries > idx->peer_qcap || k > idx->peer_kcap)
        return fail(idx, CAB_ERR_INVALID, "n_queries / k exceed the exchange buffer (%d x %d)", idx->peer_qcap, idx->peer_kcap);
    if (out_loc != CAB_HOST && out_loc != CAB_DEVICE) return fail(idx, CAB_ERR_INVALID, "out_loc");
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->own_stream;
    const uint32_t epoch = ++idx->peer_epoch;
    Sharded sh{};
    PeerPush &pp = sh.pp;
    pp.world = idx->peer_world; pp.rank = idx->peer_rank; pp.parity = int(epoch & 1u); pp.epoch = epoch;
    pp.n_queries_total = n_queries; pp.done_counter = idx->d_done;
    const size_t half = peer_half_elems(idx);
    for (int r = 0; r < idx->peer_world; ++r) {
        pp.flags[r] = reinterpret_cast<uint32_t *>(idx->peer_ptr[r]);
        pp.bufs[r] = reinterpret_cast<cab_candidate *>(idx->peer_ptr[r] + peer_flags_bytes()) + size_t(pp.parity) * half;
    }
    const UserOut o{out_index, out_fusion, out_asr, out_audio, out_flags, out_count, out_loc};
    int rc = enter_stream(idx, s);
    if (rc != CAB_OK) return rc;
    if (path != CAB_PATH_GEMV && shadow_possible(idx, k) && n_queries >= gemm_min_queries(idx) && gemm_path_available()) {
        // fp32 shard with bf16 shadows, a batch: the shard's EXACT top-k comes from the certified
        // single-index route (tensor-core preselection, exact re-score, local re-runs of uncertified
        // queries -- a purely local matter, no rank needs to agree), lands in device scratch, is packed
        // into candidate records and pushed; the merge below is the ordinary one.
        CU(idx, cudaSetDevice(idx->device));
        if ((rc = ensure_dev(idx, &idx->d_out, &idx->d_out_bytes, out_layout(n_queries, CAB_MAX_K).total))) return rc;
        const OutLayout L = out_layout(n_queries, k);
        uint8_t *d = idx->d_out;
        const UserOut lo{reinterpret_cast<int64_t *>(d + L.index), reinterpret_cast<double *>(d + L.fusion),
                         reinterpret_cast<float *>(d + L.asr), reinterpret_cast<float *>(d + L.audio), d + L.flags,
                         reinterpret_cast<int32_t *>(d + L.count), CAB_DEVICE};
        if ((rc = run_local(idx, queries, queries_loc, w_asr, w_audio, n_queries, k, threshold, CAB_PATH_AUTO, &lo, nullptr, nullptr, s))) return rc;
        if (idx->cert_pending && (rc = rerun_uncertified(idx, queries, queries_loc, w_asr, w_audio, k, threshold, lo, s))) return rc;
        if ((rc = stage_params(idx, nullptr, CAB_DEVICE, w_asr, w_audio, n_queries, s, nullptr))) return rc;   // the merge's weights
        pp.q0 = 0; pp.signal = 1;
        launch_pack_push(lo.index, lo.asr, lo.audio, lo.flags, lo.count, n_queries, k, pp, s);
        idx->launches += 1;
        CU(idx, cudaGetLastError());
        sh.fused = false;
    } else {
        rc = run_local(idx, queries, queries_loc, w_asr, w_audio, n_queries, k, threshold, path, &o, nullptr, &sh, s);
        if (rc != CAB_OK) return rc;
    }
    if (!sh.fused) {
        // many queries / several scan passes / an empty shard: the merge is its own launch
        EmitArgs ea{};
        if ((rc = make_emit(idx, idx->peer_world, n_queries, k, threshold, out_index, out_fusion, out_asr,
                            out_audio, out_flags, out_count, out_loc, &ea))) return rc;
        ea.cands = pp.bufs[idx->peer_rank];
        if (n_queries == 1 && idx->inline_w_valid) { ea.inline_weights = 1; ea.w64_asr = idx->inline_w64[0]; ea.w64_audio = idx->inline_w64[1]; }
        ea.wait_flags = pp.flags[idx->peer_rank] + pp.parity * idx->peer_world;
        ea.wait_epoch = epoch;
        ea.status = idx->d_status;
        launch_emit(ea, s);
        idx->launches += 1;
        CU(idx, cudaGetLastError());
    }
    return finish_outputs(idx, n_queries, k, o, s);
}

int cab_peer_snapshot(cab_index *idx, void *out, size_t out_bytes, size_t *needed) {
    CHECK_HANDLE(idx);
    if (!idx->d_peer) return fail(idx, CAB_ERR_INVALID, "cab_peer_init first");
    const size_t bytes = peer_flags_bytes() + 2 * peer_half_elems(idx) * sizeof(cab_candidate);
    if (needed) *needed = bytes;
    if (!out) return CAB_OK;
    if (out_bytes < bytes) return fail(idx, CAB_ERR_INVALID, "snapshot buffer too small (%zu < %zu)", out_bytes, bytes);
    CU(idx, cudaSetDevice(idx->device));
    CU(idx, cudaDeviceSynchronize());
    CU(idx, cudaMemcpy(out, idx->d_peer, bytes, cudaMemcpyDeviceToHost));
    return CAB_OK;
}

int cab_index_exchange_stamps(cab_index *idx, uint64_t *out, int max_rows) {
    CHECK_HANDLE(idx);
    if (!out || max_rows < 0) return fail(idx, CAB_ERR_INVALID, "bad stamp buffer");
    if (!idx->d_stamps) return 0;
    CU(idx, cudaSetDevice(idx->device));
    const int have = int(std::min<uint32_t>(idx->stamp_calls, kStampRows));
    const int n = std::min(have, max_rows);
    std::vector<unsigned long long> all(size_t(kStampRows) * kStampCols);
    CU(idx, cudaDeviceSynchronize());
    CU(idx, cudaMemcpy(all.data(), idx->d_stamps, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {                    // oldest first among the last n
        const uint32_t call = idx->stamp_calls - uint32_t(n) + uint32_t(i);
        memcpy(out + size_t(i) * kStampCols, all.data() + size_t(call % kStampRows) * kStampCols, kStampCols * sizeof(uint64_t));
    }
    return n;
}

int cab_index_set_option(cab_index *idx, const char *key, int64_t value) {
    CHECK_HANDLE(idx);
    if (!key) return fail(idx, CAB_ERR_INVALID, "null option key");
    std::string k(key);
    if (k == "gemv_variant") { if (value != 0) return fail(idx, CAB_ERR_INVALID, "gemv_variant: only 0 (LDG register pipeline) is built"); idx->gemv.variant = int(value); }
    else if (k == "gemv_blocks_per_sm") { if (value < 0 || value > 8) return fail(idx, CAB_ERR_INVALID, "gemv_blocks_per_sm in 0..8"); idx->gemv.blocks_per_sm = int(value); }
    else if (k == "gemv_unroll") { if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return fail(idx, CAB_ERR_INVALID, "gemv_unroll in {0,1,2,4,8}"); idx->gemv.unroll = int(value); }
    else if (k == "gemv_query_tile") { if (value != 0 && value != 1 && value != 2 && value != 4) return fail(idx, CAB_ERR_INVALID, "gemv_query_tile in {0,1,2,4}"); idx->gemv.query_tile = int(value); }
    else if (k == "time_kernels") idx->opt_time_kernels = value != 0;
    else if (k == "sync_after_search") idx->opt_sync = value != 0;
    else if (k == "finalize_general") idx->opt_finalize_general = value != 0;
    else if (k == "queries_settled") idx->opt_queries_settled = value != 0;
    else if (k == "raw_dot") idx->opt_raw_dot = value != 0;
    else if (k == "tensor_core_shadow") {
        if (idx->dtype != CAB_F32) return fail(idx, CAB_ERR_INVALID, "tensor_core_shadow is for fp32 indices (a bf16 index is its own tensor-core operand)");
        CU(idx, cudaSetDevice(idx->device));
        CU(idx, cudaDeviceSynchronize());
        if (value && !idx->shadow_asr) {
            const size_t bytes = size_t(std::max<int64_t>(idx->capacity, 1)) * CAB_DIM * 2;
            cudaError_t e = cudaMalloc(&idx->shadow_asr, bytes);
            if (e == cudaSuccess) e = cudaMalloc(&idx->shadow_audio, bytes);
            if (e != cudaSuccess) {
                cudaFree(idx->shadow_asr); idx->shadow_asr = idx->shadow_audio = nullptr; cudaGetLastError();
                return fail(idx, CAB_ERR_NOMEM, "cannot allocate the bf16 shadows (%s)", cudaGetErrorString(e));
            }
            launch_shadow_rows(static_cast<const float *>(idx->asr), idx->shadow_asr, 0, idx->size, idx->own_stream);
            launch_shadow_rows(static_cast<const float *>(idx->audio), idx->shadow_audio, 0, idx->size, idx->own_stream);
            idx->launches += 2;
            CU(idx, cudaGetLastError());
            CU(idx, cudaStreamSynchronize(idx->own_stream));
        } else if (!value && idx->shadow_asr) {
            cudaFree(idx->shadow_asr); cudaFree(idx->shadow_audio);
            idx->shadow_asr = idx->shadow_audio = nullptr;
        }
    }
    else if (k == "stamp_exchange") { idx->opt_stamp_exchange = value != 0; idx->stamp_calls = 0; }
    else if (k == "gemv_chunk_rows") { if (value < 0 || value > 4096) return fail(idx, CAB_ERR_INVALID, "gemv_chunk_rows in 0..4096 (0 = auto)"); idx->opt_chunk_rows = value; }
    else if (k == "gemm_min_queries") { if (value < 0) return fail(idx, CAB_ERR_INVALID, "gemm_min_queries >= 0 (0 = auto)"); idx->opt_gemm_min_queries = value; }
    else if (k == "gemv_batch") { if (value < 1 || value > 64) return fail(idx, CAB_ERR_INVALID, "gemv_batch in 1..64"); idx->opt_gemv_batch = value; }
    else return fail(idx, CAB_ERR_INVALID, "unknown option '%s'", key);
    return CAB_OK;
}

int64_t cab_index_get_option(const cab_index *idx, const char *key) {
    if (!idx || !key) return -1;
    std::string k(key);
    if (k == "gemv_variant") return idx->gemv.variant;
    if (k == "gemv_blocks_per_sm") return idx->gemv.blocks_per_sm;
    if (k == "gemv_unroll") return idx->gemv.unroll;
    if (k == "time_kernels") return idx->opt_time_kernels;
    if (k == "sync_after_search") return idx->opt_sync;
    if (k == "finalize_general") return idx->opt_finalize_general;
    if (k == "queries_settled") return idx->opt_queries_settled;
    if (k == "raw_dot") return idx->opt_raw_dot;
    if (k == "tensor_core_shadow") return idx->shadow_asr != nullptr;
    if (k == "last_uncertified") return idx->last_uncertified;
    if (k == "total_uncertified") return idx->total_uncertified;
    if (k == "total_shadow_queries") return idx->total_shadow_queries;
    if (k == "stamp_exchange") return idx->opt_stamp_exchange;
    if (k == "gemv_chunk_rows") return idx->opt_chunk_rows;
    if (k == "gemm_min_queries") return gemm_min_queries(idx);
    if (k == "gemv_batch") return idx->opt_gemv_batch;
    if (k == "sm_count") return idx->sm_count;
    if (k == "gemv_query_tile") return idx->gemv.query_tile;
    if (k == "gemv_grid") return plan_gemv(idx->gemv, idx->dtype, 1, idx->sm_count).grid_x;
    return -1;
}

int64_t cab_index_launch_count(const cab_index *idx) { return idx ? idx->launches : -1; }

double cab_index_last_scan_ms(const cab_index *idx) {
    if (!idx || !idx->timed) return -1.0;
    if (cudaEventSynchronize(idx->ev_t1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, idx->ev_t0, idx->ev_t1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    return double(ms);
}

