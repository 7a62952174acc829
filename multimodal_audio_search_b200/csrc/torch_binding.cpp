// PyTorch extension over the C-ABI (include/cab.h): tensors in, tensors out, on torch's current
// CUDA stream.  PyTorch is plumbing only (device memory + streams); all arithmetic is in libcab.so.
//   torch.ops.cab.append(handle, asr, audio, flags)
//   torch.ops.cab.search(handle, queries, w_asr, w_audio, k, threshold, path)
//       -> (index i64 [Q,k], fusion f64 [Q,k], asr_sim f32, audio_sim f32, flags u8, count i32 [Q])
//   torch.ops.cab.search_candidates(...) -> u8 [Q,k,24];  torch.ops.cab.merge_candidates(...)
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include "../../include/cab.h"

namespace {

cab_index *handle_of(int64_t h) { return reinterpret_cast<cab_index *>(static_cast<intptr_t>(h)); }

// torch's current stream as the C-ABI wants it.  The C-ABI reads NULL as "the handle's own stream",
// but torch's default stream IS the NULL handle: name it explicitly so the work is ordered with
// the caller's stream.
cudaStream_t current_stream() {
    cudaStream_t s = at::cuda::getCurrentCUDAStream();
    return s ? s : cudaStreamLegacy;
}

void check(int status, cab_index *idx) {
    if (status == CAB_OK) return;
    const char *msg = cab_last_error(idx);
    std::string text = (msg && *msg) ? msg : cab_status_string(status);
    if (status == CAB_ERR_NONFINITE) TORCH_CHECK_VALUE(false, text);
    TORCH_CHECK(false, "cab status ", status, ": ", text);
}

void check_rows(const at::Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.dim() == 2 && t.size(1) == CAB_DIM && t.is_contiguous(),
                name, " must be a contiguous CUDA float32 [n x 384] tensor");
}

std::vector<double> host_weights(const at::Tensor &w, int64_t nq, const char *name) {
    at::Tensor c = w.to(at::kCPU, at::kDouble).contiguous();
    TORCH_CHECK(c.numel() == nq, name, " must hold one weight per query");
    return std::vector<double>(c.data_ptr<double>(), c.data_ptr<double>() + nq);
}

void append(int64_t h, const c10::optional<at::Tensor> &asr, const c10::optional<at::Tensor> &audio,
            const c10::optional<at::Tensor> &flags) {
    TORCH_CHECK(asr.has_value() || audio.has_value(), "at least one corpus must be given");
    const at::Tensor &first = asr.has_value() ? *asr : *audio;
    if (asr.has_value()) check_rows(*asr, "asr");
    if (audio.has_value()) check_rows(*audio, "audio");
    const int64_t n = first.size(0);
    if (asr.has_value() && audio.has_value()) TORCH_CHECK(asr->size(0) == audio->size(0), "row counts differ");
    if (flags.has_value())
        TORCH_CHECK(flags->is_cuda() && flags->scalar_type() == at::kByte && flags->numel() == n && flags->is_contiguous(),
                    "flags must be a contiguous CUDA uint8 [n] tensor");
    c10::cuda::CUDAGuard guard(first.device());
    cudaStream_t s = current_stream();
    check(cab_index_append(handle_of(h), asr.has_value() ? asr->data_ptr<float>() : nullptr,
                           audio.has_value() ? audio->data_ptr<float>() : nullptr,
                           flags.has_value() ? flags->data_ptr<uint8_t>() : nullptr, n, CAB_DEVICE, s),
          handle_of(h));
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
search(int64_t h, const at::Tensor &queries, const at::Tensor &w_asr, const at::Tensor &w_audio, int64_t k,
       double threshold, int64_t path) {
    check_rows(queries, "queries");
    const int64_t nq = queries.size(0);
    const auto wa = host_weights(w_asr, nq, "w_asr"), wb = host_weights(w_audio, nq, "w_audio");
    c10::cuda::CUDAGuard guard(queries.device());
    auto opt = queries.options();
    at::Tensor oi = at::empty({nq, k}, opt.dtype(at::kLong)), of = at::empty({nq, k}, opt.dtype(at::kDouble));
    at::Tensor oa = at::empty({nq, k}, opt.dtype(at::kFloat)), ob = at::empty({nq, k}, opt.dtype(at::kFloat));
    at::Tensor ofl = at::empty({nq, k}, opt.dtype(at::kByte)), oc = at::empty({nq}, opt.dtype(at::kInt));
    cudaStream_t s = current_stream();
    check(cab_search(handle_of(h), queries.data_ptr<float>(), CAB_DEVICE, wa.data(), wb.data(), int(nq), int(k),
                     threshold, int(path), oi.data_ptr<int64_t>(), of.data_ptr<double>(), oa.data_ptr<float>(),
                     ob.data_ptr<float>(), ofl.data_ptr<uint8_t>(), oc.data_ptr<int32_t>(), CAB_DEVICE, s),
          handle_of(h));
    return {oi, of, oa, ob, ofl, oc};
}

at::Tensor search_candidates(int64_t h, const at::Tensor &queries, const at::Tensor &w_asr, const at::Tensor &w_audio,
                             int64_t k, double threshold, int64_t path) {
    check_rows(queries, "queries");
    const int64_t nq = queries.size(0);
    const auto wa = host_weights(w_asr, nq, "w_asr"), wb = host_weights(w_audio, nq, "w_audio");
    c10::cuda::CUDAGuard guard(queries.device());
    at::Tensor out = at::empty({nq, k, int64_t(sizeof(cab_candidate))}, queries.options().dtype(at::kByte));
    cudaStream_t s = current_stream();
    check(cab_search_candidates(handle_of(h), queries.data_ptr<float>(), CAB_DEVICE, wa.data(), wb.data(), int(nq),
                                int(k), threshold, int(path), reinterpret_cast<cab_candidate *>(out.data_ptr<uint8_t>()), s),
          handle_of(h));
    return out;
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
merge_candidates(int64_t h, const at::Tensor &gathered, const at::Tensor &w_asr, const at::Tensor &w_audio, int64_t k,
                 double threshold) {
    TORCH_CHECK(gathered.is_cuda() && gathered.scalar_type() == at::kByte && gathered.dim() == 4 &&
                    gathered.size(2) == k && gathered.size(3) == int64_t(sizeof(cab_candidate)) && gathered.is_contiguous(),
                "gathered must be a contiguous CUDA uint8 [world, Q, k, 24] tensor");
    const int64_t world = gathered.size(0), nq = gathered.size(1);
    const auto wa = host_weights(w_asr, nq, "w_asr"), wb = host_weights(w_audio, nq, "w_audio");
    c10::cuda::CUDAGuard guard(gathered.device());
    auto opt = gathered.options();
    at::Tensor oi = at::empty({nq, k}, opt.dtype(at::kLong)), of = at::empty({nq, k}, opt.dtype(at::kDouble));
    at::Tensor oa = at::empty({nq, k}, opt.dtype(at::kFloat)), ob = at::empty({nq, k}, opt.dtype(at::kFloat));
    at::Tensor ofl = at::empty({nq, k}, opt.dtype(at::kByte)), oc = at::empty({nq}, opt.dtype(at::kInt));
    cudaStream_t s = current_stream();
    check(cab_merge_candidates(handle_of(h), reinterpret_cast<const cab_candidate *>(gathered.data_ptr<uint8_t>()),
                               int(world), int(nq), int(k), wa.data(), wb.data(), threshold, oi.data_ptr<int64_t>(),
                               of.data_ptr<double>(), oa.data_ptr<float>(), ob.data_ptr<float>(),
                               ofl.data_ptr<uint8_t>(), oc.data_ptr<int32_t>(), CAB_DEVICE, s),
          handle_of(h));
    return {oi, of, oa, ob, ofl, oc};
}

}  // namespace

TORCH_LIBRARY(cab, m) {
    m.def("append(int handle, Tensor? asr, Tensor? audio, Tensor? flags) -> ()", &append);
    m.def("search(int handle, Tensor queries, Tensor w_asr, Tensor w_audio, int k, float threshold, int path) -> "
          "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", &search);
    m.def("search_candidates(int handle, Tensor queries, Tensor w_asr, Tensor w_audio, int k, float threshold, int path) -> Tensor",
          &search_candidates);
    m.def("merge_candidates(int handle, Tensor gathered, Tensor w_asr, Tensor w_audio, int k, float threshold) -> "
          "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", &merge_candidates);
}
