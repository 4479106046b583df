// Host-buffer entry of the DPA-loss path (include/halva_dpa.h: dpa_plan_host_spans, dpa_stage_host_batch).
//
// When the hidden states live in pinned host memory the step is bounded by the host->device link long before the
// kernels are (1.07 GB per cfg-2 batch; eight GPUs share the host's memory bandwidth).  Labels and phrase ids are
// known on the host, and a hidden row (s, t) is only ever read by the kernels when it predicts a contributing label
// (llava/train/halva_trainer.py:403-407,550-562: label[s,t+1] != -100 and, for the pair batch, sign[s,t+1] not in
// {0, -100}).  Those rows form a few contiguous spans per sequence, so one asynchronous copy per span is issued; every
// other row of the device tensors is never read by the library (row compaction).  Planning is a single pass over the
// host labels (microseconds), done here per batch and inside the caller's timed region.
#include <chrono>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/halva_dpa.h"

namespace {

constexpr long long kIgnore = -100;

template <class I>
int plan_spans(const I* labels, const I* signs, int S, int T, int merge_gap, dpa_span_t* out, int cap, long long* rows) {
  int n = 0;
  long long nrows = 0;
  for (int s = 0; s < S; ++s) {
    const I* y = labels + (size_t)s * T;
    const I* g = signs ? signs + (size_t)s * T : nullptr;
    int open = -1, last = -1;          // current span [open, last]
    for (int t = 0; t + 1 < T; ++t) {
      bool on = (long long)y[t + 1] != kIgnore;
      if (on && g) { const long long v = (long long)g[t + 1]; on = v != 0 && v != kIgnore; }
      if (!on) continue;
      if (open >= 0 && t - last - 1 <= merge_gap) { last = t; continue; }
      if (open >= 0) {
        if (n >= cap) return -1;
        out[n++] = dpa_span_t{s, open, last + 1};
        nrows += last + 1 - open;
      }
      open = last = t;
    }
    if (open >= 0) {
      if (n >= cap) return -1;
      out[n++] = dpa_span_t{s, open, last + 1};
      nrows += last + 1 - open;
    }
  }
  if (rows) *rows = nrows;
  return n;
}

int plan_any(const void* labels, const void* signs, int S, int T, int index_dtype, int merge_gap, dpa_span_t* out, int cap, long long* rows) {
  if (index_dtype == DPA_I64)
    return plan_spans(static_cast<const long long*>(labels), static_cast<const long long*>(signs), S, T, merge_gap, out, cap, rows);
  return plan_spans(static_cast<const int*>(labels), static_cast<const int*>(signs), S, T, merge_gap, out, cap, rows);
}

}  // namespace

extern "C" {

int dpa_plan_host_spans(const void* labels, const void* signs, int S, int T, int index_dtype, int merge_gap, dpa_span_t* out, int cap) {
  if (!labels || !out || S < 0 || T < 2 || cap < 0 || merge_gap < 0) return DPA_ERR_ARG;
  if (index_dtype != DPA_I64 && index_dtype != DPA_I32) return DPA_ERR_DTYPE;
  const int n = plan_any(labels, signs, S, T, index_dtype, merge_gap, out, cap, nullptr);
  return n < 0 ? DPA_ERR_WORKSPACE : n;
}

int dpa_stage_host_batch(const dpa_host_batch_t* b, void* stream, dpa_stage_info_t* info) {
  if (!b) return DPA_ERR_ARG;
  if (b->B < 1 || b->T < 2 || b->B_ref < 1 || b->T_ref < 2 || b->H < 1) return DPA_ERR_BAD_SHAPE;
  if (b->index_dtype != DPA_I64 && b->index_dtype != DPA_I32) return DPA_ERR_DTYPE;
  if (b->dtype != DPA_BF16 && b->dtype != DPA_F32) return DPA_ERR_DTYPE;
  if (!b->h_pair || !b->labels_pair || !b->signs_pair || !b->h_ref_pol || !b->labels_ref || !b->d_h_pair || !b->d_labels_pair ||
      !b->d_signs_pair || !b->d_h_ref_pol || !b->d_labels_ref)
    return DPA_ERR_ARG;
  if ((b->h_ref_frz == nullptr) != (b->d_h_ref_frz == nullptr)) return DPA_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto t0 = std::chrono::steady_clock::now();
  const size_t isz = b->index_dtype == DPA_I64 ? 8 : 4;
  const size_t row_bytes = (size_t)b->H * (b->dtype == DPA_F32 ? 4 : 2);
  const int S = 2 * b->B;
  long long bytes = 0, rows_pair = 0, rows_ref = 0;
  int n_pair = 0, n_ref = 0;
  thread_local std::vector<dpa_span_t> spans;
  if (!b->dense) {
    // a sequence holds at most T / 2 spans
    const size_t cap = (size_t)S * (b->T / 2 + 1) + (size_t)b->B_ref * (b->T_ref / 2 + 1);
    if (spans.size() < cap) spans.resize(cap);
    n_pair = plan_any(b->labels_pair, b->signs_pair, S, b->T, b->index_dtype, b->merge_gap, spans.data(), (int)cap, &rows_pair);
    if (n_pair < 0) return DPA_ERR_WORKSPACE;
    n_ref = plan_any(b->labels_ref, nullptr, b->B_ref, b->T_ref, b->index_dtype, b->merge_gap, spans.data() + n_pair, (int)(cap - n_pair), &rows_ref);
    if (n_ref < 0) return DPA_ERR_WORKSPACE;
  }
  const auto t1 = std::chrono::steady_clock::now();
  auto copy = [&](void* dst, const void* src, size_t n) -> cudaError_t {
    bytes += (long long)n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
  };
  if (copy(b->d_labels_pair, b->labels_pair, (size_t)S * b->T * isz) != cudaSuccess) return DPA_ERR_CUDA;
  if (copy(b->d_signs_pair, b->signs_pair, (size_t)S * b->T * isz) != cudaSuccess) return DPA_ERR_CUDA;
  if (copy(b->d_labels_ref, b->labels_ref, (size_t)b->B_ref * b->T_ref * isz) != cudaSuccess) return DPA_ERR_CUDA;
  if (b->dense) {
    if (copy(b->d_h_pair, b->h_pair, (size_t)S * b->T * row_bytes) != cudaSuccess) return DPA_ERR_CUDA;
    if (copy(b->d_h_ref_pol, b->h_ref_pol, (size_t)b->B_ref * b->T_ref * row_bytes) != cudaSuccess) return DPA_ERR_CUDA;
    if (b->h_ref_frz && copy(b->d_h_ref_frz, b->h_ref_frz, (size_t)b->B_ref * b->T_ref * row_bytes) != cudaSuccess) return DPA_ERR_CUDA;
    rows_pair = (long long)S * b->T; rows_ref = (long long)b->B_ref * b->T_ref;
  } else {
    for (int i = 0; i < n_pair; ++i) {
      const dpa_span_t& sp = spans[i];
      const size_t off = ((size_t)sp.seq * b->T + sp.t0) * row_bytes, n = (size_t)(sp.t1 - sp.t0) * row_bytes;
      if (copy(static_cast<char*>(b->d_h_pair) + off, static_cast<const char*>(b->h_pair) + off, n) != cudaSuccess) return DPA_ERR_CUDA;
    }
    for (int i = n_pair; i < n_pair + n_ref; ++i) {
      const dpa_span_t& sp = spans[i];
      const size_t off = ((size_t)sp.seq * b->T_ref + sp.t0) * row_bytes, n = (size_t)(sp.t1 - sp.t0) * row_bytes;
      if (copy(static_cast<char*>(b->d_h_ref_pol) + off, static_cast<const char*>(b->h_ref_pol) + off, n) != cudaSuccess) return DPA_ERR_CUDA;
      if (b->h_ref_frz && copy(static_cast<char*>(b->d_h_ref_frz) + off, static_cast<const char*>(b->h_ref_frz) + off, n) != cudaSuccess)
        return DPA_ERR_CUDA;
    }
  }
  const auto t2 = std::chrono::steady_clock::now();
  if (info) {
    info->bytes = bytes; info->spans_pair = n_pair; info->spans_ref = n_ref;
    info->rows_pair = (int32_t)rows_pair; info->rows_ref = (int32_t)rows_ref;
    info->plan_us = std::chrono::duration<float, std::micro>(t1 - t0).count();
    info->issue_us = std::chrono::duration<float, std::micro>(t2 - t1).count();
  }
  return DPA_OK;
}

}  // extern "C"
