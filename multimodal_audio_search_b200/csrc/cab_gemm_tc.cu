// Batched path (>= 64 queries, bf16 storage): dense contraction on the 5th-gen tensor cores.
//
//   scores_asr  [256 queries x 128 rows] = Q[256 x 384] . A_tile[128 x 384]^T      (bf16 x bf16 -> fp32)
//   scores_audio[256 queries x 128 rows] = Q[256 x 384] . B_tile[128 x 384]^T
//
// replaces, for a batch of queries, the per-segment loop of search_with_fusion
// (audio_search.py:639-682); weighting, threshold pruning and the running top-k are fused into
// the epilogue so no score is ever written to memory (256 x 10M scores would be 10 GB).
//
// Mapping to sm_100a:
//   * one persistent CTA PAIR per two SMs (cluster 2x1x1, 74 pairs), tcgen05.mma.cta_group::2 with
//     M = 256 (128 queries per CTA) and N = 256 = BOTH corpora of a 128-segment tile in one
//     instruction: each CTA stages [64 ASR rows | 64 audio rows] of its half of the tile as one
//     128-row B operand, so one accumulator holds s_asr and s_audio side by side and the query
//     operand is read once for both.  K = 384 as 24 MMAs of K = 16 per tile.  The pair reads each
//     corpus row exactly once: HBM traffic is the algorithmic 1536 B per segment per pass.
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) -- the 128 queries of
//     each CTA stay resident in shared memory (96 KB), corpus k-blocks (2 x 64 rows x 128 B =
//     16 KB) flow through a 7-stage mbarrier ring (112 KB in flight per SM);
//   * accumulators in TMEM: 2 stages x 256 columns = all 512 columns, so the MMAs of tile i+1
//     overlap the epilogue of tile i.  Column map of a stage: [0,64) s_asr and [64,128) s_audio of
//     the leader CTA's 64 segments, [128,192) / [192,256) the same for the peer CTA's 64 segments;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane of the pair's
//     leader CTA) + TMEM allocator, warps 2..5 = epilogue.  An epilogue thread owns one TMEM lane
//     = one query: it reads the two score rows with tcgen05.ld, fuses them with the row flags,
//     compares with its private running bound (one compare rejects almost every segment) and
//     appends survivors to its private candidate list in global memory (L2 resident); a full
//     list is compacted warp-cooperatively (bitonic sort in shared memory, keep k).
//   * cross-CTA bound sharing: every appended candidate bumps a per-query score-level histogram in
//     global memory (one RED); each epilogue warp refreshes one query's bound per tile from the
//     histogram's suffix sums (>= k candidates at or above a level => the global k-th best is at
//     least that level).  Without it pruning stays at the 0.1 threshold for the whole scan.
//   Roofline: 2 x 2 x 384 flop per (query, segment); at 256 queries the kernel needs 256 flop per
//   corpus byte, i.e. it sits on the HBM/tensor ridge.
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "cab_device.cuh"
#include "cab_internal.h"

namespace cab {

// ---- shapes --------------------------------------------------------------------------------------
constexpr int kTileRows = 128;                 // corpus rows per tile (N of the MMA), per CTA pair
constexpr int kHalfRows = kTileRows / 2;       // rows staged by each CTA
constexpr int kQPerCta = 128;                  // queries (M rows) per CTA
constexpr int kKBlock = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int kNumKBlocks = kDim / kKBlock;    // 6
constexpr int kUmmaK = 16;
constexpr int kStages = 7;                     // ring of 16 KB k-blocks ([ASR half | audio half])
constexpr uint32_t kCorpusBytes = kHalfRows * kKBlock * 2;         // 8192: one corpus, one k-block, this CTA's rows
constexpr uint32_t kStageBytes = 2 * kCorpusBytes;                 // 16384
constexpr uint32_t kQBlockBytes = kQPerCta * kKBlock * 2;          // 16384
constexpr int kTmemCols = 512;
constexpr int kGemmThreads = 192;              // 6 warps
constexpr int kEpiWarp0 = 2;
constexpr int kLevels = 64;                    // score levels of the shared per-query histogram
constexpr float kLevelStep = 0.004f;

// smem layout (dynamic, 1024-byte aligned)
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffRing = kOffQ + kNumKBlocks * kQBlockBytes;                  // 98304
constexpr uint32_t kOffScratch = kOffRing + kStages * kStageBytes;                 // 98304 + 114688 = 212992
constexpr uint32_t kOffBars = kOffScratch + 4 * kWarpCap * 8;                      // + 8 KB sort scratch
constexpr uint32_t kNumBars = 2 * kStages + 4 + 1;
constexpr uint32_t kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr uint32_t kGemmSmemBytes = kOffTmemPtr + 16 + 1024;                       // + alignment slack

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // L2 cache hints (createpolicy encodings)
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {      // same smem offset in CTA rank 0 of the cluster
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0)); return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_bar) : "memory");
}
// Bounded wait: a protocol bug must surface as an error (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *status, int code) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) {           // ~2 s
            if (status) atomicExch(status, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar,
                                                 int c0, int c1, uint64_t hint) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(hint) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {       // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Packed fp32 pairs (FMUL2 / FFMA2) and the 3-input maximum (FMNMX3) of sm_100: the epilogue's
// per-score work in a third of the issue slots -- under the 1 kW cap issue slots are clock.
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

// K-major, 128-byte-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 | LBO (16 B) | SBO = 8 rows x 128 B = 1024 | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;            // stride byte offset
    d |= uint64_t(1) << 46;                    // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                    // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 256 (pair), N = 256.
constexpr uint32_t kUmmaN = 2 * kTileRows;        // 256: [ASR | audio] x [leader | peer] halves
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kUmmaN >> 3) << 17) | (uint32_t(256 >> 4) << 24);

// ---- per-pass prologue (one launch): normalise the queries like sklearn normalize(X), round to
// bf16, zero-pad to 256 rows; clear the level histogram and the protocol status word --------------------
__global__ void __launch_bounds__(256) gemm_prologue_kernel(const float *__restrict__ q_raw, int n_queries,
                                                            __nv_bfloat16 *__restrict__ out, int32_t *__restrict__ levels,
                                                            int *__restrict__ status) {
    // programmatic dependent launch: the previous pass's finalize kernel has released the histogram
    // (it consumed it before its launch_dependents); waiting for its completion here keeps the
    // stream's completion order transitive for the kernels behind this one
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = gtid >> 5;
    for (int i = gtid; i < kGemmQueriesPerPass * kLevels; i += gridDim.x * blockDim.x) levels[i] = 0;
    if (gtid == 0) *status = 0;
    if (row >= kGemmQueriesPerPass) return;
    float v[kDim / 32];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kDim / 32; ++i) {
        v[i] = row < n_queries ? q_raw[size_t(row) * kDim + lane + 32 * i] : 0.f;
        ss = fmaf(v[i], v[i], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
    // a NaN/Inf query becomes a NaN operand row: every score of it is NaN, nothing is selected,
    // and the finalize kernel (which re-reads the raw query) reports it
    float norm = sqrtf(ss);
    if (norm == 0.f) norm = 1.f;
#pragma unroll
    for (int i = 0; i < kDim / 32; ++i) out[size_t(row) * kDim + lane + 32 * i] = __float2bfloat16_rn(v[i] / norm);
}

// ---- the scan ----------------------------------------------------------------------------------------
struct GemmParams {
    const uint8_t *flags;
    int64_t n_rows;
    const float *wa32, *wb32;      // [n_queries]
    int n_queries;
    int k;
    float select_threshold;
    uint64_t *lists;               // [256 queries][n_pairs][kGemmListCap]
    int32_t *counts;               // [256 queries][n_pairs]
    int n_pairs;
    int *status;                   // != 0: protocol timeout code
    // Cross-CTA bound sharing: levels[q][j] counts the candidates of query q pushed (by any CTA)
    // with fused score in [thr + j*kLevelStep, thr + (j+1)*kLevelStep).  If the counts of levels
    // >= j sum to k or more, the global k-th best score is >= thr + j*kLevelStep: a safe bound.
    int32_t *levels;               // [256 queries][kLevels]
    // Visiting order of the tile rounds (round r = the n_pairs tiles [r * n_pairs, (r+1) * n_pairs)):
    // the it-th round a pair works on is (it * round_mul) % n_rounds_full, a bijection.  A library
    // whose scores ascend with the row number -- every row beating the running k-th best, the
    // adversarial order for the pruning epilogue -- then looks like a random one at round
    // granularity; the pairs still sweep n_pairs consecutive tiles at a time.
    uint32_t n_rounds_full;        // rounds in which every pair has a tile
    uint32_t round_mul;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_scan_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_asr,
                 const __grid_constant__ CUtensorMap map_audio, GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms); identical offset in both CTAs of the pair.
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1;

    const uint32_t bar0 = smem_base + kOffBars;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int t) { return bar0 + 8u * (2 * kStages + t); };
    auto tmem_empty_bar = [&](int t) { return bar0 + 8u * (2 * kStages + 2 + t); };
    const uint32_t q_bar = bar0 + 8u * (2 * kStages + 4);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(tmem_full_bar(t), 1); mbar_init(tmem_empty_bar(t), 2 * 4); }
        mbar_init(q_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                   // TMEM: all 512 columns of both SMs of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_holder)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // Programmatic dependent launch: barrier init, TMEM allocation and the cluster handshake above
    // overlapped the prologue kernel; its bf16 queries and the cleared histogram are read below.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int64_t n_tiles = (p.n_rows + kTileRows - 1) / kTileRows;
    // tile of this pair's it-th round (>= n_tiles: no more work)
    auto tile_of = [&](uint32_t it) -> int64_t {
        const uint32_t r = it < p.n_rounds_full ? uint32_t((uint64_t(it) * p.round_mul) % p.n_rounds_full) : it;
        return int64_t(r) * p.n_pairs + pair;
    };

    if (warp == 0) {
        // ===== TMA producer (one lane in each CTA): own 128 queries once, then own half of every tile
        if (elect_one()) {
            const uint32_t q_bar_leader = mapa_rank0(q_bar);
            if (leader) mbar_expect_tx(q_bar, 2 * kNumKBlocks * kQBlockBytes);
            for (int kb = 0; kb < kNumKBlocks; ++kb)
                tma_load_2d_pair(smem_base + kOffQ + kb * kQBlockBytes, &map_q, q_bar_leader, kb * kKBlock,
                                 int(cta_rank) * kQPerCta, kEvictLast);
            uint32_t n = 0, rnd = 0;
            for (int64_t tile = tile_of(0); tile < n_tiles; tile = tile_of(++rnd)) {
                const int row0 = int(tile * kTileRows) + int(cta_rank) * kHalfRows;
                for (int kb = 0; kb < kNumKBlocks; ++kb, ++n) {
                    const int s = n % kStages;
                    const uint32_t ph = (n / kStages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u, p.status, 1);
                    if (leader) mbar_expect_tx(full_bar(s), 2 * kStageBytes);
                    const uint32_t dst = smem_base + kOffRing + s * kStageBytes;
                    const uint32_t bar = mapa_rank0(full_bar(s));
                    tma_load_2d_pair(dst, &map_asr, bar, kb * kKBlock, row0, kEvictFirst);
                    tma_load_2d_pair(dst + kCorpusBytes, &map_audio, bar, kb * kKBlock, row0, kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one lane of the leader CTA drives the tensor cores of both SMs
        if (leader && elect_one()) {
            mbar_wait(q_bar, 0, p.status, 2);
            tc_fence_after();
            uint32_t n = 0, it = 0;
            for (int64_t tile = tile_of(0); tile < n_tiles; tile = tile_of(++it)) {
                const uint32_t t = it & 1u, tph = (it >> 1) & 1u;
                mbar_wait(tmem_empty_bar(t), tph ^ 1u, p.status, 3);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + t * 256u;
                for (int kb = 0; kb < kNumKBlocks; ++kb, ++n) {
                    const int s = n % kStages;
                    const uint32_t ph = (n / kStages) & 1u;
                    mbar_wait(full_bar(s), ph, p.status, 4);
                    tc_fence_after();
                    const uint64_t a0 = umma_desc(smem_base + kOffQ + kb * kQBlockBytes);
                    const uint64_t b0 = umma_desc(smem_base + kOffRing + s * kStageBytes);
#pragma unroll
                    for (int k = 0; k < kKBlock / kUmmaK; ++k)              // +32 bytes per K step
                        tc_mma_pair(d_tmem, a0 + uint64_t(2 * k), b0 + uint64_t(2 * k), kInstrDesc,
                                    (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(empty_bar(s));                           // frees the stage in both CTAs
                }
                tc_commit_pair(tmem_full_bar(t));                           // accumulators ready, both CTAs
            }
        }
    } else {
        // ===== epilogue: thread <-> TMEM lane <-> query.  The four warps never synchronise with each
        // other: each keeps the tile's 128 row flags in registers (lane l holds rows 4l..4l+3).
        const int ew = warp & 3;                              // TMEM lane quarter this warp may access
        const int qloc = ew * 32 + lane;
        const int q = int(cta_rank) * kQPerCta + qloc;        // query index within this pass
        const bool q_valid = q < p.n_queries;
        ScanWeights w{0.f, 0.f};
        if (q_valid) { w.wa = p.wa32[q]; w.wb = p.wb32[q]; }
        uint64_t bound = q_valid ? bound_key(p.select_threshold) : ~0ull;
        int cnt = 0;
        uint64_t *list = p.lists + (size_t(q) * p.n_pairs + pair) * kGemmListCap;
        uint64_t *scratch = reinterpret_cast<uint64_t *>(smem + kOffScratch) + (warp - kEpiWarp0) * kWarpCap;
        const uint32_t lane_addr = tmem_base + (uint32_t(ew * 32) << 16);
        const uint32_t tmem_empty_leader0 = mapa_rank0(tmem_empty_bar(0));
        const uint32_t tmem_empty_leader1 = mapa_rank0(tmem_empty_bar(1));
        const bool thr_nonneg = p.select_threshold >= 0.f;    // then the :654 gate is redundant

        // flags of rows 4*lane .. 4*lane+3 of a tile, 0 for rows beyond the library
        auto load_flags = [&](int64_t tile) -> uint32_t {
            const int64_t r = tile * kTileRows + 4 * lane;
            if (tile >= n_tiles || r >= p.n_rows) return 0u;
            uint32_t wd = *reinterpret_cast<const uint32_t *>(p.flags + r);   // capacity is a multiple of 4 rows
            const int64_t left = p.n_rows - r;
            if (left < 4) wd &= (1u << (8 * int(left))) - 1u;
            return wd;
        };
        // Push one (already fused) candidate; called with warp-divergent predicates.
        int32_t *my_levels = p.levels + size_t(q_valid ? q : 0) * kLevels;
        auto push = [&](float fused, uint32_t row) {
            const uint64_t key = make_key(fused, row);
            if (key > bound) {
                list[cnt++] = key;
                int lvl = int((fused - p.select_threshold) * (1.0f / kLevelStep));
                lvl = lvl < 0 ? 0 : (lvl > kLevels - 1 ? kLevels - 1 : lvl);
                atomicAdd(my_levels + lvl, 1);                                 // RED, no return value
            }
        };
        // Levels of the query owned by lane (it & 31): lane l holds levels l and 32 + l.
        const int q_warp0 = int(cta_rank) * kQPerCta + ew * 32;
        auto load_levels = [&](uint32_t it_, int &lo, int &hi) {
            const int qr = q_warp0 + int(it_ & 31u);
            const int32_t *lp = p.levels + size_t(qr < p.n_queries ? qr : 0) * kLevels;
            lo = __ldcg(lp + lane);
            hi = __ldcg(lp + 32 + lane);
        };
        int lv_lo, lv_hi;
        load_levels(0, lv_lo, lv_hi);

        uint32_t it = 0;
        uint32_t flag_word = load_flags(tile_of(0));
        for (int64_t tile = tile_of(0); tile < n_tiles; tile = tile_of(++it)) {
            const uint32_t t = it & 1u, tph = (it >> 1) & 1u;
            const uint32_t row0 = uint32_t(tile * kTileRows);
            const uint32_t fw = flag_word;
            flag_word = load_flags(tile_of(it + 1));                           // prefetch the next tile's flags
            const bool fast = thr_nonneg && __all_sync(kFull, (fw & 0x03030303u) == 0x03030303u);   // bits 2-3: weight class, not used here
            {   // refresh the bound of lane (it & 31)'s query from the shared level histogram
                int s_hi = lv_hi, s_lo = lv_lo;                                // inclusive suffix sums over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int a = __shfl_down_sync(kFull, s_hi, o), b = __shfl_down_sync(kFull, s_lo, o);
                    if (lane + o < 32) { s_hi += a; s_lo += b; }
                }
                s_lo += __shfl_sync(kFull, s_hi, 0);
                const unsigned b_hi = __ballot_sync(kFull, s_hi >= p.k), b_lo = __ballot_sync(kFull, s_lo >= p.k);
                const int lvl = b_hi ? 32 + (31 - __clz(b_hi)) : (b_lo ? 31 - __clz(b_lo) : -1);
                if (lvl >= 0 && lane == int(it & 31u)) {
                    const uint64_t nb = bound_key(p.select_threshold + float(lvl) * kLevelStep - 2e-6f);
                    if (nb > bound) bound = nb;
                }
                load_levels(it + 1, lv_lo, lv_hi);                             // prefetch for the next tile
            }
            mbar_wait(tmem_full_bar(t), tph, p.status, 5);
            tc_fence_after();
#pragma unroll 1
            for (int chunk = 0; chunk < kTileRows / 32; ++chunk) {
                // segments [32*chunk, +32) of the tile: s_asr at column 128*(chunk/2) + 32*(chunk%2), s_audio 64 further
                const uint32_t ta = lane_addr + t * 256u + uint32_t((chunk >> 1) * 128 + (chunk & 1) * 32);
                uint32_t va[32], vb[32];
                tc_ld32(ta, va);
                tc_ld32(ta + 64u, vb);
                tc_wait_ld();
                const uint32_t col0 = row0 + uint32_t(chunk * 32);
                const float bound_f = key_score(bound);
                uint32_t m = 0;                                                // bit j: column j may enter the list
                float f[32];
                bool any = true;
                if (fast) {
                    // branch-free, packed: f = wa * s_asr + wb * s_audio for two columns per instruction
                    // (FMUL2 + FFMA2, bit-identical to the scalar form), then ONE maximum over the 32
                    // columns (FMNMX3 tree) and one compare: almost every chunk ends here
                    const uint64_t wa2 = pack2(__float_as_uint(w.wa), __float_as_uint(w.wa));
                    const uint64_t wb2 = pack2(__float_as_uint(w.wb), __float_as_uint(w.wb));
#pragma unroll
                    for (int j = 0; j < 32; j += 2)
                        unpack2(fma2(wa2, pack2(va[j], va[j + 1]), mul2(wb2, pack2(vb[j], vb[j + 1]))), f[j], f[j + 1]);
                    float mx[11];
#pragma unroll
                    for (int j = 0; j < 10; ++j) mx[j] = max3(f[3 * j], f[3 * j + 1], f[3 * j + 2]);
                    mx[10] = fmaxf(f[30], f[31]);
                    const float m0 = max3(mx[0], mx[1], mx[2]), m1 = max3(mx[3], mx[4], mx[5]), m2 = max3(mx[6], mx[7], mx[8]);
                    const float top = max3(max3(m0, m1, m2), mx[9], mx[10]);
                    any = __any_sync(kFull, top >= bound_f);
                    if (any) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) m |= f[j] >= bound_f ? (1u << j) : 0u;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t wd = __shfl_sync(kFull, fw, (chunk * 32 + j) >> 2);
                        f[j] = fuse32(__uint_as_float(va[j]), __uint_as_float(vb[j]), (wd >> (8 * (j & 3))) & 0xFFu, w);
                        m |= f[j] >= bound_f ? (1u << j) : 0u;
                    }
                }
                // Rare survivors.  The loop runs over the union of the lanes' masks, so the column index
                // is warp-uniform and the register pair is picked by a jump table (no divergence, no
                // TMEM re-read); only the lanes that flagged the column push.
                uint32_t um = any ? __reduce_or_sync(kFull, m) : 0u;
                while (um) {
                    const int j = __ffs(um) - 1;
                    um &= um - 1;
                    float fj = 0.f;
                    switch (j) {
#define CAB_PICK(J) case J: fj = f[J]; break;
                        CAB_PICK(0) CAB_PICK(1) CAB_PICK(2) CAB_PICK(3) CAB_PICK(4) CAB_PICK(5) CAB_PICK(6) CAB_PICK(7)
                        CAB_PICK(8) CAB_PICK(9) CAB_PICK(10) CAB_PICK(11) CAB_PICK(12) CAB_PICK(13) CAB_PICK(14) CAB_PICK(15)
                        CAB_PICK(16) CAB_PICK(17) CAB_PICK(18) CAB_PICK(19) CAB_PICK(20) CAB_PICK(21) CAB_PICK(22) CAB_PICK(23)
                        CAB_PICK(24) CAB_PICK(25) CAB_PICK(26) CAB_PICK(27) CAB_PICK(28) CAB_PICK(29) CAB_PICK(30) CAB_PICK(31)
#undef CAB_PICK
                    }
                    if ((m >> j) & 1u) push(fj, col0 + uint32_t(j));
                }
                // A list that could overflow during the next chunk is compacted by the whole warp.
                unsigned need = __ballot_sync(kFull, cnt > kGemmListCap - 32);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const uint64_t *lp = reinterpret_cast<const uint64_t *>(
                        __shfl_sync(kFull, reinterpret_cast<unsigned long long>(list), src));
                    const int c = __shfl_sync(kFull, cnt, src);
                    __syncwarp();
                    for (int i = lane; i < kWarpCap; i += 32) scratch[i] = i < c ? lp[i] : 0ull;
                    warp_sort_desc<kWarpCap>(scratch, lane);
                    uint64_t *lw = const_cast<uint64_t *>(lp);
                    for (int i = lane; i < p.k; i += 32) lw[i] = scratch[i];
                    const uint64_t kth = scratch[p.k - 1];
                    __syncwarp();
                    if (lane == src) { cnt = p.k; if (kth > bound) bound = kth; }
                }
            }
            // this warp's TMEM reads of stage t are complete: one arrive per warp on the leader's barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(t == 0 ? tmem_empty_leader0 : tmem_empty_leader1);
        }
        if (q < kGemmQueriesPerPass) p.counts[size_t(q) * p.n_pairs + pair] = q_valid ? cnt : 0;
    }

    // ---- teardown: everyone (both CTAs) done before TMEM is released -------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// [rows x 384] bf16 row-major, box = 64 elements (128 B) x box_rows, 128-byte swizzle, zero OOB fill.
static bool make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t box_rows, std::string *err) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { *err = "cuTensorMapEncodeTiled is not available from the driver"; return false; }
    cuuint64_t dims[2] = {cuuint64_t(kDim), rows};
    cuuint64_t strides[1] = {cuuint64_t(kDim) * 2};
    cuuint32_t box[2] = {cuuint32_t(kKBlock), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)); return false; }
    return true;
}

bool gemm_path_available() { return true; }
int gemm_partials_per_query(int sm_count) { return sm_count / 2; }

// workspace: [counts 256 x n_pairs int32][bf16 queries 256 x 384][levels 256 x 64 int32][status int]
static size_t ws_counts_bytes(int sm_count) { return ((size_t(kGemmQueriesPerPass) * (sm_count / 2) * 4 + 255) / 256) * 256; }
static size_t ws_queries_bytes() { return size_t(kGemmQueriesPerPass) * kDim * 2; }
static size_t ws_levels_bytes() { return size_t(kGemmQueriesPerPass) * kLevels * 4; }
size_t gemm_workspace_bytes(int, int, int sm_count) { return ws_counts_bytes(sm_count) + ws_queries_bytes() + ws_levels_bytes() + 256; }
const int32_t *gemm_levels(const void *workspace, int sm_count) {
    return reinterpret_cast<const int32_t *>(static_cast<const uint8_t *>(workspace) + ws_counts_bytes(sm_count) + ws_queries_bytes());
}
float gemm_level_step() { return kLevelStep; }

void launch_gemm_scan(const ScanArgs &a, int sm_count, void *workspace, size_t workspace_bytes,
                      cudaStream_t s, std::string *err) {
    if (a.dtype != CAB_BF16) { *err = "tensor-core scan needs bf16 rows"; return; }
    if (a.n_queries < 1 || a.n_queries > kGemmQueriesPerPass) { *err = "tensor-core scan takes 1..256 queries per pass"; return; }
    if (workspace_bytes < gemm_workspace_bytes(a.n_queries, a.k, sm_count)) { *err = "tensor-core workspace too small"; return; }
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    int32_t *counts = reinterpret_cast<int32_t *>(ws);
    __nv_bfloat16 *qb = reinterpret_cast<__nv_bfloat16 *>(ws + ws_counts_bytes(sm_count));
    int32_t *levels = reinterpret_cast<int32_t *>(ws + ws_counts_bytes(sm_count) + ws_queries_bytes());
    int *status = reinterpret_cast<int *>(ws + ws_counts_bytes(sm_count) + ws_queries_bytes() + ws_levels_bytes());

    {   // opt-in shared memory size: per kernel AND per device
        static PerDeviceOnce once;
        int dev = -1;
        cudaGetDevice(&dev);
        if (!once.done(dev)) {
            cudaError_t e = cudaFuncSetAttribute(gemm_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemmSmemBytes));
            if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return; }
            once.mark(dev);
        }
    }
    CUtensorMap mq, ma, mb;
    if (!make_map(&mq, qb, kGemmQueriesPerPass, kQPerCta, err)) return;
    if (!make_map(&ma, a.asr, uint64_t(a.n_rows), kHalfRows, err)) return;
    if (!make_map(&mb, a.audio, uint64_t(a.n_rows), kHalfRows, err)) return;

    GemmParams p{};
    p.flags = a.flags; p.n_rows = a.n_rows; p.wa32 = a.wa32; p.wb32 = a.wb32; p.n_queries = a.n_queries;
    p.k = a.k; p.select_threshold = a.select_threshold; p.lists = a.partial_keys; p.counts = counts;
    p.n_pairs = sm_count / 2; p.status = status; p.levels = levels;
    {
        const int64_t n_tiles = (a.n_rows + kTileRows - 1) / kTileRows;
        const uint64_t full = uint64_t(n_tiles / p.n_pairs);
        uint64_t mul = 1;
        if (full > 2) {                                    // ~golden-ratio stride, coprime to the round count
            mul = uint64_t(double(full) * 0.6180339887) | 1ull;
            auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
            while (gcd(mul, full) != 1) mul += 2;
            mul %= full;
            if (mul == 0) mul = 1;
        }
        p.n_rounds_full = uint32_t(full); p.round_mul = uint32_t(mul);
    }

    // prologue -> scan, both with programmatic stream serialization: the prologue's launch overlaps
    // the previous pass's finalize, the scan's setup (barriers, TMEM, cluster sync) the prologue.
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t c0{};
    c0.gridDim = dim3(kGemmQueriesPerPass / 8); c0.blockDim = dim3(256); c0.dynamicSmemBytes = 0; c0.stream = s;
    c0.attrs = attr; c0.numAttrs = 1;
    cudaLaunchKernelEx(&c0, gemm_prologue_kernel, a.queries, a.n_queries, qb, levels, status);
    cudaLaunchConfig_t c1{};
    c1.gridDim = dim3(2 * p.n_pairs); c1.blockDim = dim3(kGemmThreads); c1.dynamicSmemBytes = kGemmSmemBytes; c1.stream = s;
    c1.attrs = attr; c1.numAttrs = 1;
    cudaLaunchKernelEx(&c1, gemm_scan_kernel, mq, ma, mb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) *err = std::string("gemm_scan_kernel launch: ") + cudaGetErrorString(e);
}

}  // namespace cab
