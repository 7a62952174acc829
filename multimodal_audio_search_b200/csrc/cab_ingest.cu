// Ingest-side kernels: normalise-on-append, the deterministic synthetic library generator and a
// row read-back.  All are simple streaming kernels (one warp per 384-element row, 128-bit
// accesses); none is on the search hot path.
#include "cab_device.cuh"
#include "cab_internal.h"

namespace cab {

// ---- normalise-on-append -----------------------------------------------------------------------
// Replaces the per-call `normalize(Y)` inside sklearn's cosine_similarity (pairwise.py:1744-1748,
// called from audio_search.py:646/:651): norm = sqrt(sum x^2) in fp32, zero norm -> divide by 1.
template <bool kBf16>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float *__restrict__ src,
                                                             void *__restrict__ dst,
                                                             int64_t dst_row, int64_t n,
                                                             int *__restrict__ nonfinite,
                                                             float *__restrict__ norms,
                                                             __nv_bfloat16 *__restrict__ shadow) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += n_warps) {
        float4 v[3];
        float ss = 0.f;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            v[j] = src ? reinterpret_cast<const float4 *>(src + r * kDim)[lane + 32 * j]
                       : make_float4(0.f, 0.f, 0.f, 0.f);
            bad |= !(isfinite(v[j].x) && isfinite(v[j].y) && isfinite(v[j].z) && isfinite(v[j].w));
            ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss);
            ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
        if (__any_sync(kFull, bad) || !isfinite(ss)) { if (lane == 0) *nonfinite = 1; }
        float norm = sqrtf(ss);
        const int64_t out_row = dst_row + r;
        if (lane == 0) norms[out_row] = norm;         // the row's length, for raw dot-product scoring (0 for a missing row)
        if (norm == 0.f) norm = 1.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float4 o4 = make_float4(v[j].x / norm, v[j].y / norm, v[j].z / norm, v[j].w / norm);
            if constexpr (kBf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o4.x, o4.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn(o4.z, o4.w);
                uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
                reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(dst) + out_row * kDim)[lane + 32 * j] = pk;
            } else {
                reinterpret_cast<float4 *>(reinterpret_cast<float *>(dst) + out_row * kDim)[lane + 32 * j] = o4;
                if (shadow) {                          // bf16 copy of an fp32 library for the tensor-core preselection
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o4.x, o4.y);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o4.z, o4.w);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
                    reinterpret_cast<uint2 *>(shadow + out_row * kDim)[lane + 32 * j] = pk;
                }
            }
        }
    }
}

void launch_normalize_rows(const float *src, void *dst, int dtype, int64_t dst_row, int64_t n,
                           int *nonfinite, float *norms, void *shadow, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype == CAB_BF16)
        normalize_rows_kernel<true><<<int(blocks), 256, 0, s>>>(src, dst, dst_row, n, nonfinite, norms, nullptr);
    else
        normalize_rows_kernel<false><<<int(blocks), 256, 0, s>>>(src, dst, dst_row, n, nonfinite, norms,
                                                                  static_cast<__nv_bfloat16 *>(shadow));
}

// ---- bf16 shadow of stored fp32 rows (option tensor_core_shadow switched on after the rows) --------
__global__ void shadow_rows_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int64_t n_elems) {
    int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    for (; i < n_elems; i += int64_t(gridDim.x) * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4 *>(src + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(dst + i) = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
    }
}
void launch_shadow_rows(const float *src, void *dst, int64_t r0, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n * kDim / 4 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    shadow_rows_kernel<<<int(blocks), 256, 0, s>>>(src + r0 * kDim, static_cast<__nv_bfloat16 *>(dst) + r0 * kDim, n * kDim);
}

// ---- read back ---------------------------------------------------------------------------------
__global__ void widen_rows_kernel(const void *__restrict__ src, int dtype, int64_t r0, int64_t n,
                                  float *__restrict__ out) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t total = n * kDim;
    for (; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        if (dtype == CAB_BF16)
            out[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(src)[r0 * kDim + i]);
        else
            out[i] = reinterpret_cast<const float *>(src)[r0 * kDim + i];
    }
}
void launch_widen_rows(const void *src, int dtype, int64_t r0, int64_t n, float *out, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n * kDim + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    widen_rows_kernel<<<int(blocks), 256, 0, s>>>(src, dtype, r0, n, out);
}

// ---- synthetic library (must stay bit-identical to multimodal_audio_search_b200/synth.py) ------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
constexpr uint32_t kGold = 0x9E3779B9u, kC1 = 0x85EBCA6Bu, kColMul = 0x9E3779B1u;
constexpr uint32_t kPlantMixSalt = 0x51ED270Bu;

__host__ __device__ __forceinline__ uint32_t stream_key(uint32_t seed, uint32_t stream) {
    return mix32(seed ^ (stream * kGold + kC1));
}
__device__ __forceinline__ float synth_elem(uint32_t row_key, int j) {
    uint32_t h = mix32(row_key + uint32_t(j + 1) * kColMul);
    uint32_t s = (h & 0xFFu) + ((h >> 8) & 0xFFu) + ((h >> 16) & 0xFFu) + (h >> 24);
    return float(int(s) - 510);
}

// `partial`: rows whose pipeline "failed" (flag bit clear, see synth.row_flags) have no embedding
// in the reference (None, audio_search.py:344/350) -> an all-zero row here.
// p.mode: 0 planted neighbours over isotropic noise; 1 "ascending" (every row scores a hair above
// its predecessor: the adversarial order for a pruning scan); 2 "clustered" (1024 shared centres).
// Twin of synth.corpus_rows_at -- bit-identical, so every fp32 operation of mode 1 is spelled out
// with its rounding (no FMA contraction).
__global__ void __launch_bounds__(256) synth_rows_kernel(SynthParams p, int stream_id, int partial,
                                                         int64_t r0, int64_t n, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t skey = stream_key(p.seed, uint32_t(stream_id));
    const uint32_t qkey = stream_key(p.seed, 2u);
    const uint32_t fkey = stream_key(p.seed, 3u);
    const uint32_t ckey = stream_key(p.seed, 4u);                          // cluster of a row
    const uint32_t d_rk = mix32(qkey);                                     // raw query 0: the ascent direction
    const uint32_t e_rk = mix32(stream_key(p.seed, 5u + uint32_t(stream_id)));   // second axis of this corpus's plane
    const uint32_t centre_key = stream_key(p.seed, 8u);
    const float asc_step = __fdiv_rn(0.8f, __ull2float_rn(p.n_total > 1 ? p.n_total - 1 : 1ull));
    for (int64_t i = warp; i < n; i += n_warps) {
        const uint64_t r = uint64_t(r0 + i);
        const uint32_t rk = mix32(skey ^ uint32_t(r));
        if (partial) {
            const uint32_t hv = mix32(mix32(fkey ^ uint32_t(r))) % 10u;
            const uint32_t fl = hv == 0u ? 1u : (hv == 1u ? 2u : 3u);
            if (!(fl & (1u << stream_id))) {
#pragma unroll
                for (int c = 0; c < kDim / 32; ++c) out[i * kDim + lane + 32 * c] = 0.f;
                continue;
            }
        }
        if (p.mode == 1) {
            const float alpha = __fadd_rn(0.15f, __fmul_rn(__ull2float_rn(r), asc_step));
            const float beta = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(alpha, alpha)));
#pragma unroll
            for (int c = 0; c < kDim / 32; ++c) {
                const int j = lane + 32 * c;
                out[i * kDim + j] = __fadd_rn(__fmul_rn(alpha, synth_elem(d_rk, j)), __fmul_rn(beta, synth_elem(e_rk, j)));
            }
            continue;
        }
        if (p.mode == 2) {
            const uint32_t crk = mix32(centre_key ^ (mix32(ckey ^ uint32_t(r)) % 1024u));
#pragma unroll
            for (int c = 0; c < kDim / 32; ++c) {
                const int j = lane + 32 * c;
                out[i * kDim + j] = 3.0f * synth_elem(crk, j) + 2.0f * synth_elem(rk, j);   // exact: small integers
            }
            continue;
        }
        float m = 0.f, nn = 1.f;
        uint32_t qrk = 0;
        if (p.n_plants_total) {
            uint64_t d = (r + p.n_total - p.base) % p.n_total;
            uint64_t g = (d * p.inv_stride) % p.n_total;
            if (g < p.n_plants_total) {
                uint32_t kind = uint32_t(g % 3);
                if (kind == 2u || kind == uint32_t(stream_id)) {
                    uint32_t hm = mix32((p.seed ^ kPlantMixSalt) ^ (uint32_t(g) * kColMul));
                    m = float(1u + ((hm >> (8 * stream_id)) & 7u));
                    nn = float(1u + ((hm >> (8 * stream_id + 4)) & 7u));
                    qrk = mix32(qkey ^ uint32_t(g / p.plants));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kDim / 32; ++c) {
            int j = lane + 32 * c;
            float x = synth_elem(rk, j);
            if (m != 0.f) x = m * synth_elem(qrk, j) + nn * x;       // exact: small integers
            out[i * kDim + j] = x;
        }
    }
}
void launch_synth_rows(const SynthParams &p, int stream_id, int partial, int64_t r0, int64_t n,
                       float *out, cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_rows_kernel<<<int(blocks), 256, 0, s>>>(p, stream_id, partial, r0, n, out);
}

__global__ void synth_flags_kernel(uint32_t seed, int partial, int64_t r0, int64_t n,
                                   uint8_t *__restrict__ out) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t fkey = stream_key(seed, 3u);
    for (; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        uint8_t f = 3;
        if (partial) {
            uint32_t hv = mix32(mix32(fkey ^ uint32_t(r0 + i))) % 10u;
            f = hv == 0u ? 1 : (hv == 1u ? 2 : 3);
        }
        out[i] = f;
    }
}
void launch_synth_flags(uint32_t seed, int partial, int64_t r0, int64_t n, uint8_t *out,
                        cudaStream_t s) {
    if (n <= 0) return;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_flags_kernel<<<int(blocks), 256, 0, s>>>(seed, partial, r0, n, out);
}

}  // namespace cab
