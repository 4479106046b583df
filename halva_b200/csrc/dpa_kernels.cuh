// Integer / bookkeeping / reduction kernels around the GEMMs of the DPA-loss path.
// HBM-bound byte and index work: coalesced grid-stride loops, no tensor cores.
#pragma once
#include "dpa_gemm.cuh"
#include "../../include/halva_dpa.h"

namespace dpa {

constexpr int IGNORE_INDEX = -100;          // /root/reference/llava/constants.py:7
constexpr int kMaxSign = 255;               // phrase ids are small: mask_tag counts <MASK> spans (train_halva.py:275-332)
constexpr int kScanBlock = 1024;            // positions per scan block
constexpr int kNoCol = 0xFFFF;

template <bool I64>
__device__ __forceinline__ int load_index(const void* p, size_t i) {
  if constexpr (I64) return (int)reinterpret_cast<const long long*>(p)[i];
  else return reinterpret_cast<const int*>(p)[i];
}

struct MetaArgs {
  const void* labels_pair; const void* signs_pair; const void* labels_ref;
  int B, T, B_ref, T_ref, V, dense;
  int n_pos_pair;          // 2B * (T-1)
  int n_pos_ref;           // B_ref * (T_ref-1)
  int nblk_pair, nblk_ref; // scan blocks
  int cap_pair, cap_ref;
  // state
  DpaMeta* meta;
  unsigned* bitmap;        // [2][8]  presence of phrase ids per half (pos / neg)
  int* block_counts;       // [nblk_pair + nblk_ref]
  int* src_pair; int* lab_pair; int* col_pair;   // col_pair: (sequence << 16) | column (kNoCol = none)
  int* src_ref; int* act_ref;
  int* status;             // user-visible [8]
};

// ---- pass 0: reset bookkeeping, then presence bitmap of phrase ids over signs[:, 1:] of each half
//      (halva_trainer.py:560 maps -100 -> 0; :412 takes the sorted unique over the whole half batch)
template <bool I64>
__global__ void meta_bitmap_kernel(MetaArgs a) {
  __shared__ unsigned bm[16];
  __shared__ int bad;
  if (threadIdx.x < 16) bm[threadIdx.x] = 0;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const int Tm = a.T - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_pos_pair; i += gridDim.x * blockDim.x) {
    const int s = i / Tm, t = i - s * Tm;
    int g = load_index<I64>(a.signs_pair, (size_t)s * a.T + t + 1);
    if (g == IGNORE_INDEX) g = 0;
    if (g < 0 || g > kMaxSign) { bad = 1; continue; }
    const int half = s >= a.B;
    atomicOr(&bm[half * 8 + (g >> 5)], 1u << (g & 31));
  }
  __syncthreads();
  if (threadIdx.x < 16 && bm[threadIdx.x]) atomicOr(&a.bitmap[threadIdx.x], bm[threadIdx.x]);
  if (threadIdx.x == 0 && bad) atomicOr(&a.meta->status, DPA_STATUS_BAD_SIGN);
}

// rank of phrase id g among the ids present in its half; rank 0 (the smallest id, normally 0) is dropped
__device__ __forceinline__ int sign_rank(const unsigned* bm8, int g) {
  int r = 0;
  const int w = g >> 5;
  for (int i = 0; i < w; ++i) r += __popc(bm8[i]);
  r += __popc(bm8[w] & ((1u << (g & 31)) - 1u));
  return r;
}

template <bool I64>
__device__ __forceinline__ bool pair_position(const MetaArgs& a, const unsigned* bm, int i, int& src, int& lab, int& col) {
  const int Tm = a.T - 1;
  const int s = i / Tm, t = i - s * Tm;
  src = s * a.T + t;                                   // hidden row (s, t) predicts label (s, t+1)
  const int y = load_index<I64>(a.labels_pair, (size_t)s * a.T + t + 1);
  int g = load_index<I64>(a.signs_pair, (size_t)s * a.T + t + 1);
  if (g == IGNORE_INDEX) g = 0;
  const int half = s >= a.B;
  int rank = 0;
  if (g >= 0 && g <= kMaxSign) rank = sign_rank(bm + half * 8, g);
  const bool active = (y != IGNORE_INDEX) && rank >= 1;
  lab = (y == IGNORE_INDEX) ? 0 : y;                   // -100 -> index 0 (halva_trainer.py:406)
  col = (s << 16) | (active ? (rank - 1) : kNoCol);
  return active;
}

// ---- pass 1: per-block counts of active rows
template <bool I64>
__global__ void meta_count_kernel(MetaArgs a) {
  __shared__ unsigned bm[16];
  __shared__ int cnt;
  if (threadIdx.x < 16) bm[threadIdx.x] = a.bitmap[threadIdx.x];
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int blk = blockIdx.x;
  int local = 0;
  if (blk < a.nblk_pair) {
    const int i = blk * kScanBlock + threadIdx.x;
    if (i < a.n_pos_pair) {
      int src, lab, col;
      const bool act = pair_position<I64>(a, bm, i, src, lab, col);
      local = (act || a.dense) ? 1 : 0;
    }
  } else {
    const int i = (blk - a.nblk_pair) * kScanBlock + threadIdx.x;
    if (i < a.n_pos_ref) {
      const int Tm = a.T_ref - 1;
      const int s = i / Tm, t = i - s * Tm;
      const int y = load_index<I64>(a.labels_ref, (size_t)s * a.T_ref + t + 1);
      local = (y != IGNORE_INDEX || a.dense) ? 1 : 0;
    }
  }
  const unsigned ball = __ballot_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && ball) atomicAdd(&cnt, __popc(ball));
  __syncthreads();
  if (threadIdx.x == 0) a.block_counts[blk] = cnt;
}

// ---- pass 2: stable compaction (block offset = sum of earlier counts, warp-scan inside the block) + DpaMeta
template <bool I64>
__global__ void meta_fill_kernel(MetaArgs a) {
  __shared__ unsigned bm[16];
  __shared__ int warp_sums[kScanBlock / 32];
  __shared__ int base_s, total_pair_s, total_ref_s;
  if (threadIdx.x < 16) bm[threadIdx.x] = a.bitmap[threadIdx.x];
  const int blk = blockIdx.x;
  const bool is_pair = blk < a.nblk_pair;
  // block offset inside its segment and segment totals (block counts are few: <= ~100)
  if (threadIdx.x < 32) {
    int before = 0, tp = 0, trf = 0;
    for (int j = threadIdx.x; j < a.nblk_pair + a.nblk_ref; j += 32) {
      const int c = a.block_counts[j];
      if (j < a.nblk_pair) { tp += c; if (is_pair && j < blk) before += c; }
      else { trf += c; if (!is_pair && j < blk) before += c; }
    }
    for (int o = 16; o; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      tp += __shfl_xor_sync(0xffffffffu, tp, o);
      trf += __shfl_xor_sync(0xffffffffu, trf, o);
    }
    if (threadIdx.x == 0) { base_s = before; total_pair_s = tp; total_ref_s = trf; }
  }
  __syncthreads();
  int src = 0, lab = 0, col = 0, flag = 0, act = 0;
  if (is_pair) {
    const int i = blk * kScanBlock + threadIdx.x;
    if (i < a.n_pos_pair) {
      act = pair_position<I64>(a, bm, i, src, lab, col);
      flag = (act || a.dense) ? 1 : 0;
    }
  } else {
    const int i = (blk - a.nblk_pair) * kScanBlock + threadIdx.x;
    if (i < a.n_pos_ref) {
      const int Tm = a.T_ref - 1;
      const int s = i / Tm, t = i - s * Tm;
      src = s * a.T_ref + t;
      act = load_index<I64>(a.labels_ref, (size_t)s * a.T_ref + t + 1) != IGNORE_INDEX;
      flag = (act || a.dense) ? 1 : 0;
    }
  }
  const unsigned ball = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int in_warp = __popc(ball & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[w] = __popc(ball);
  __syncthreads();
  if (w == 0) {
    int v = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    warp_sums[lane] = v - warp_sums[lane];   // exclusive
  }
  __syncthreads();
  if (flag) {
    const int dst = base_s + warp_sums[w] + in_warp;
    if (is_pair) {
      if (act && (lab < 0 || lab >= a.V)) { atomicOr(&a.meta->status, DPA_STATUS_BAD_LABEL); lab = 0; }
      a.src_pair[dst] = src; a.lab_pair[dst] = lab; a.col_pair[dst] = col;
    } else {
      a.src_ref[dst] = src; a.act_ref[dst] = act;
    }
  }
  if (blk == 0 && threadIdx.x == 0) {
    DpaMeta* m = a.meta;
    const int np = total_pair_s, nr = total_ref_s;
    m->n_pair = np; m->n_ref = nr;
    m->p0 = (np + kRowPad - 1) / kRowPad * kRowPad;
    m->n_tot = m->p0 + (nr + kRowPad - 1) / kRowPad * kRowPad;
    int kp = -1, kn = -1;
    for (int i = 0; i < 8; ++i) { kp += __popc(bm[i]); kn += __popc(bm[8 + i]); }
    m->K_pos = kp; m->K_neg = kn; m->K = kp;
    int st = atomicOr(&m->status, 0);
    if (kp != kn) st |= DPA_STATUS_K_MISMATCH;
    else if (kp <= 0) st |= DPA_STATUS_K_ZERO;
    atomicOr(&m->status, st);
  }
}

// pad rows of the packed lists up to the tile boundary (labels 0, no column): one small block
__global__ void meta_pad_kernel(MetaArgs a) {
  const DpaMeta* m = a.meta;
  for (int i = m->n_pair + threadIdx.x; i < m->p0; i += blockDim.x) { a.src_pair[i] = -1; a.lab_pair[i] = 0; a.col_pair[i] = kNoCol; }
  const int nr_pad = m->n_tot - m->p0;
  for (int i = m->n_ref + threadIdx.x; i < nr_pad; i += blockDim.x) { a.src_ref[i] = -1; a.act_ref[i] = 0; }
  if (threadIdx.x == 0) {
    a.status[0] = m->status; a.status[1] = m->K; a.status[2] = m->n_pair; a.status[3] = m->n_ref;
    a.status[4] = m->K_pos; a.status[5] = m->K_neg; a.status[6] = m->p0; a.status[7] = m->n_tot;
  }
}

// ---- gather active hidden rows into the packed, tile-padded operand buffers (one warp per row, 16 B vectors).
//      NORM: the rows are the pre-norm residual stream and the decoder's final RMSNorm is applied on the way
//      (modelling_llama.py:64-69: variance in fp32, x * rsqrt(var + eps) cast back to bf16, then weight * (.) in bf16).
struct GatherArgs {
  const DpaMeta* meta;
  const int* src_pair; const int* src_ref;
  const __nv_bfloat16* h_pair; const __nv_bfloat16* h_ref_pol; const __nv_bfloat16* h_ref_frz;
  __nv_bfloat16* Hp;   // [cap_pair + cap_ref][H]  policy rows: pair segment [0,p0), reference segment [p0, n_tot)
  __nv_bfloat16* Hf;   // [cap_ref][H]             frozen-model rows of the reference segment
  int H;
  const __nv_bfloat16* norm_w; const __nv_bfloat16* norm_w_ref; float norm_eps;
  float* rinv;         // [cap_pair + cap_ref]  rsqrt(var + eps) of the policy rows (kept for the backward)
};
__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
template <bool NORM>
__global__ void gather_rows_kernel(GatherArgs a) {
  const int p0 = a.meta->p0, n_tot = a.meta->n_tot;
  const int nr_pad = n_tot - p0;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int vec = a.H / 8;                                 // uint4 per row
  const int r_end = a.h_ref_frz ? n_tot + nr_pad : n_tot;    // (ref_logits given: no frozen hidden rows)
  for (int r = wid; r < r_end; r += warps) {
    const uint4* src = nullptr; uint4* dst;
    const __nv_bfloat16* w = a.norm_w;
    if (r < p0) {
      const int s = a.src_pair[r];
      if (s >= 0) src = reinterpret_cast<const uint4*>(a.h_pair + (size_t)s * a.H);
      dst = reinterpret_cast<uint4*>(a.Hp + (size_t)r * a.H);
    } else if (r < n_tot) {
      const int s = a.src_ref[r - p0];
      if (s >= 0) src = reinterpret_cast<const uint4*>(a.h_ref_pol + (size_t)s * a.H);
      dst = reinterpret_cast<uint4*>(a.Hp + (size_t)r * a.H);
    } else {
      const int j = r - n_tot;
      const int s = a.src_ref[j];
      if (s >= 0) src = reinterpret_cast<const uint4*>(a.h_ref_frz + (size_t)s * a.H);
      dst = reinterpret_cast<uint4*>(a.Hf + (size_t)j * a.H);
      w = a.norm_w_ref;
    }
    if (!src) {
      for (int i = lane; i < vec; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
      if (NORM && lane == 0 && r < n_tot) a.rinv[r] = 0.f;
      continue;
    }
    if constexpr (!NORM) {
      for (int i = lane; i < vec; i += 32) dst[i] = __ldg(src + i);
    } else {
      float ss = 0.f;
      for (int i = lane; i < vec; i += 32) {
        const uint4 v = __ldg(src + i);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = bf2_to_f2(q[k]); ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); }
      }
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float ri = rsqrtf(ss / (float)a.H + a.norm_eps);
      if (lane == 0 && r < n_tot) a.rinv[r] = ri;
      const uint4* wv = reinterpret_cast<const uint4*>(w);
      for (int i = lane; i < vec; i += 32) {
        const uint4 v = __ldg(src + i), g = __ldg(wv + i);     // second pass over the 8 KB row hits L1
        const uint32_t q[4] = {v.x, v.y, v.z, v.w}, gq[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = bf2_to_f2(q[k]), gw = bf2_to_f2(gq[k]);
          o[k] = pack_bf16x2(gw.x * bf16_round_f32(f.x * ri), gw.y * bf16_round_f32(f.y * ri));
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---- fp32 inputs (dpa_shape_t::dtype == DPA_F32): the tensor-core operands are split-bf16 triples, x = hi + mid + lo with
//      hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)  (the residuals are exact in fp32; 24 mantissa bits in all).
//      8 fp32 -> 3 x 8 bf16; the three parts of a buffer are `part` elements apart.
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4 (&out)[3]) {
  float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    out[p] = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] -= bf16_round_f32(r[k]);
  }
}
// gather of the active fp32 hidden rows into packed split-bf16 operand buffers (one warp per row)
struct GatherSplitArgs {
  const DpaMeta* meta;
  const int* src_pair; const int* src_ref;
  const float* h_pair; const float* h_ref_pol; const float* h_ref_frz;
  __nv_bfloat16* Hp; size_t part_p;   // 3 x [cap_pair + cap_ref][H]
  __nv_bfloat16* Hf; size_t part_f;   // 3 x [cap_ref][H]
  int H;
};
__global__ void gather_rows_split_kernel(GatherSplitArgs a) {
  const int p0 = a.meta->p0, n_tot = a.meta->n_tot;
  const int nr_pad = n_tot - p0;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int vec = a.H / 8;
  for (int r = wid; r < n_tot + nr_pad; r += warps) {
    const float4* src = nullptr; uint4* dst; size_t part;
    if (r < n_tot) {
      const int sr = r < p0 ? a.src_pair[r] : a.src_ref[r - p0];
      if (sr >= 0) src = reinterpret_cast<const float4*>((r < p0 ? a.h_pair : a.h_ref_pol) + (size_t)sr * a.H);
      dst = reinterpret_cast<uint4*>(a.Hp + (size_t)r * a.H); part = a.part_p / 8;
    } else {
      const int j = r - n_tot, sr = a.src_ref[j];
      if (sr >= 0) src = reinterpret_cast<const float4*>(a.h_ref_frz + (size_t)sr * a.H);
      dst = reinterpret_cast<uint4*>(a.Hf + (size_t)j * a.H); part = a.part_f / 8;
    }
    for (int i = lane; i < vec; i += 32) {
      uint4 o[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      if (src) split8(__ldg(src + 2 * i), __ldg(src + 2 * i + 1), o);
      dst[i] = o[0]; dst[i + part] = o[1]; dst[i + 2 * part] = o[2];
    }
  }
}
// lm_head weights fp32 [V][H] -> split-bf16 triple, parts n elements apart (per call: the weights change every optimiser step)
__global__ void split_f32_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t nv = n >> 3, stride = (size_t)gridDim.x * blockDim.x;
  const float4* p = reinterpret_cast<const float4*>(in);
  uint4* o = reinterpret_cast<uint4*>(out);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    uint4 t[3];
    split8(__ldg(p + 2 * i), __ldg(p + 2 * i + 1), t);
    o[i] = t[0]; o[i + nv] = t[1]; o[i + 2 * nv] = t[2];
  }
}

// ---- backward of the folded final RMSNorm, in place over the dH rows of the active policy rows:
//      dn = bf16(dy * w), dx = r dn - x r^3 mean(dn x), d_norm_w += dy * bf16(x r)   (autograd of modelling_llama.py:64-69)
struct NormBwdArgs {
  const DpaMeta* meta;
  const int* src_pair; const int* src_ref;
  const __nv_bfloat16* h_pair; const __nv_bfloat16* h_ref_pol;     // pre-norm rows
  __nv_bfloat16* dH_pair; __nv_bfloat16* dH_ref;                    // in: dy, out: dx
  const __nv_bfloat16* norm_w; const float* rinv;
  float* dnw32;        // [H] fp32 accumulator or nullptr
  int H;
};
__global__ void __launch_bounds__(256) norm_backward_kernel(NormBwdArgs a) {
  extern __shared__ float sm_dnw[];                                 // [H] when dnw32
  const int p0 = a.meta->p0, n_pair = a.meta->n_pair, n_ref = a.meta->n_ref;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int vec = a.H / 8;
  if (a.dnw32) { for (int i = threadIdx.x; i < a.H; i += blockDim.x) sm_dnw[i] = 0.f; __syncthreads(); }
  const uint4* wv = reinterpret_cast<const uint4*>(a.norm_w);
  const int total = n_pair + n_ref;
  for (int q = blockIdx.x * wpb + wib; q < total; q += gridDim.x * wpb) {
    const bool is_pair = q < n_pair;
    const int r = is_pair ? q : p0 + (q - n_pair);
    const int s = is_pair ? a.src_pair[q] : a.src_ref[q - n_pair];
    const uint4* x = reinterpret_cast<const uint4*>((is_pair ? a.h_pair : a.h_ref_pol) + (size_t)s * a.H);
    uint4* dy = reinterpret_cast<uint4*>((is_pair ? a.dH_pair : a.dH_ref) + (size_t)s * a.H);
    const float ri = a.rinv[r];
    float acc = 0.f;
    for (int i = lane; i < vec; i += 32) {
      const uint4 dv = dy[i], xv = __ldg(x + i), g = __ldg(wv + i);
      const uint32_t dq[4] = {dv.x, dv.y, dv.z, dv.w}, xq[4] = {xv.x, xv.y, xv.z, xv.w}, gq[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = bf2_to_f2(dq[k]), xx = bf2_to_f2(xq[k]), gw = bf2_to_f2(gq[k]);
        acc = fmaf(bf16_round_f32(d.x * gw.x), xx.x, acc);
        acc = fmaf(bf16_round_f32(d.y * gw.y), xx.y, acc);
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float c = acc / (float)a.H * ri * ri * ri;
    for (int i = lane; i < vec; i += 32) {
      const uint4 dv = dy[i], xv = __ldg(x + i), g = __ldg(wv + i);
      const uint32_t dq[4] = {dv.x, dv.y, dv.z, dv.w}, xq[4] = {xv.x, xv.y, xv.z, xv.w}, gq[4] = {g.x, g.y, g.z, g.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = bf2_to_f2(dq[k]), xx = bf2_to_f2(xq[k]), gw = bf2_to_f2(gq[k]);
        o[k] = pack_bf16x2(ri * bf16_round_f32(d.x * gw.x) - xx.x * c, ri * bf16_round_f32(d.y * gw.y) - xx.y * c);
        if (a.dnw32) {
          atomicAdd(&sm_dnw[i * 8 + 2 * k], d.x * bf16_round_f32(xx.x * ri));
          atomicAdd(&sm_dnw[i * 8 + 2 * k + 1], d.y * bf16_round_f32(xx.y * ri));
        }
      }
      dy[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (a.dnw32) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.H; i += blockDim.x) if (sm_dnw[i] != 0.f) atomicAdd(&a.dnw32[i], sm_dnw[i]);
  }
}

// ---- packed-row ranges the row-wise kernels can be restricted to (the fused-gradient forward finishes the pair
//      segment first and the reference-sample segment chunk by chunk)
enum : int { SEG_ALL = 0, SEG_PAIR = 1, SEG_REF_CHUNK = 2 };
struct RowRange { int seg, chunk, chunk_rows; };
__device__ __forceinline__ void row_range(const DpaMeta* m, const RowRange& rr, int& lo, int& hi) {
  const int p0 = m->p0, n_tot = m->n_tot;
  if (rr.seg == SEG_PAIR) { lo = 0; hi = p0; }
  else if (rr.seg == SEG_REF_CHUNK) { lo = max(rr.chunk * rr.chunk_rows, p0); hi = min((rr.chunk + 1) * rr.chunk_rows, n_tot); }
  else { lo = 0; hi = n_tot; }
}

// ---- per-row results: log-prob of the target (pair), KL row (reference sample)
struct FinishArgs {
  const DpaMeta* meta;
  RowRange rr;
  const float4* stats_pair; const float4* stats_ref; const float* stats_u;
  const int* act_ref;
  float* lse_pair; float* logp_pair;
  float* lse_ref_p; float* lse_ref_r; float* kl_row;
};
__global__ void finish_rows_kernel(FinishArgs a) {
  const int p0 = a.meta->p0;
  int lo, hi; row_range(a.meta, a.rr, lo, hi);
  for (int r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x) {
    if (r < p0) {
      const float4 s = a.stats_pair[r];
      const float lse = s.x + logf(s.y);
      a.lse_pair[r] = lse; a.logp_pair[r] = s.z - lse;          // z[y] - LSE (halva_trainer.py:407)
    } else {
      const int j = r - p0;
      const float4 s = a.stats_ref[j];
      const float lp = s.x + logf(s.y), lr = s.z + logf(s.w);
      a.lse_ref_p[j] = lp; a.lse_ref_r[j] = lr;
      // sum_v q (log q - log p) = sum_v q (z_r - z_p) - (LSE_r - LSE_p)   (halva_trainer.py:583-586).
      // LSE_r - LSE_p is formed as (m_r - m_p) + log(s_r / s_p): the two maxima are close (exact difference) and
      // the ratio is near 1, so the ~1e-6 absolute rounding of each LSE does not swamp a KL row of ~1e-3.
      const float dlse = (s.z - s.x) + logf(s.w / s.y);
      if (a.stats_u) a.kl_row[j] = a.act_ref[j] ? (a.stats_u[j] / s.w - dlse) : 0.f;
      else a.kl_row[j] = dlse;       // fused-gradient forward: dz_kl_elem_kernel finishes the row from the kept logits
    }
  }
}

// ---- P[b,k], N[b,k] and the loss assembly in ONE launch: one block per sequence computes its row of P / N (one warp per
//      phrase column round-robin, fixed summation order); the last block to finish (ticket in DpaMeta) assembles
//      align = mean log(1+exp(N-P)), sigma for the backward and / or div = sum(kl)/B_ref and the total.
enum : int { LOSS_ALIGN = 1, LOSS_TOTAL = 2 };
struct LossArgs {
  DpaMeta* meta;
  const int* src_pair; const int* col_pair; const float* logp_pair;
  float* PN;        // [2B][kMaxSign+1]
  const float* kl_row;
  float* sigma;     // [B][kMaxSign+1]  d align / d x
  float* losses;    // [4]
  int B, B_ref, T; float alpha;
  int phases;       // bit 0: P / N sums, alignment loss + sigma;  bit 1: divergence + total (needs every kl_row)
  int ticket;       // which DpaMeta::pad slot counts finished blocks (one per launch of a call)
};
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}
__global__ void __launch_bounds__(256) loss_kernel(LossArgs a) {
  __shared__ float sh[32];
  __shared__ int is_last;
  const int K = a.meta->K, status = a.meta->status;
  if (a.phases & LOSS_ALIGN) {
    const int s = blockIdx.x;                         // grid = 2B sequences
    const int n = a.meta->n_pair;
    // packed rows are sequence-major: binary-search this sequence's range by source row
    const int lo_key = s * a.T, hi_key = (s + 1) * a.T;
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.src_pair[mid] < lo_key) lo = mid + 1; else hi = mid; }
    const int r0 = lo;
    hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.src_pair[mid] < hi_key) lo = mid + 1; else hi = mid; }
    const int r1 = lo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = warp; k < K && k <= kMaxSign; k += nw) {
      float acc = 0.f;
      for (int r = r0 + lane; r < r1; r += 32)
        if ((a.col_pair[r] & 0xFFFF) == k) acc += a.logp_pair[r];
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) a.PN[(size_t)s * (kMaxSign + 1) + k] = acc;
    }
    if (gridDim.x > 1) {                              // the last block to get here assembles the loss
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) is_last = atomicAdd(&a.meta->pad[a.ticket], 1) == (int)gridDim.x - 1;
      __syncthreads();
      if (!is_last) return;
      __threadfence();
    } else {
      __syncthreads();
    }
    const int BK_ = a.B * max(K, 0);
    float acc = 0.f;
    for (int i = threadIdx.x; i < BK_; i += blockDim.x) {
      const int b = i / K, k = i - b * K;
      const float P = __ldcg(&a.PN[(size_t)b * (kMaxSign + 1) + k]);
      const float N = __ldcg(&a.PN[(size_t)(a.B + b) * (kMaxSign + 1) + k]);
      const float e = expf(N - P);                       // literal log(1 + exp(x)) (halva_trainer.py:567)
      acc += logf(1.f + e);
      a.sigma[(size_t)b * (kMaxSign + 1) + k] = (e / (1.f + e)) / (float)BK_;
    }
    const float align_sum = block_sum(acc, sh);
    if (threadIdx.x == 0) {
      float align = BK_ > 0 ? align_sum / (float)BK_ : nanf("");   // mean of empty -> NaN, like the reference
      if (status & (DPA_STATUS_BAD_SIGN | DPA_STATUS_K_MISMATCH | DPA_STATUS_K_ZERO | DPA_STATUS_BAD_LABEL)) align = nanf("");
      a.losses[1] = align;
    }
  } else if (blockIdx.x != 0) {
    return;
  }
  if (a.phases & LOSS_TOTAL) {
    float kacc = 0.f;
    const int nr = a.meta->n_ref;
    for (int i = threadIdx.x; i < nr; i += blockDim.x) kacc += a.kl_row[i];
    const float kl_sum = block_sum(kacc, sh);
    if (threadIdx.x == 0) {
      const float align = a.losses[1];                            // written by this thread or by the LOSS_ALIGN launch
      const float div = kl_sum / (float)a.B_ref;                  // / batch size (halva_trainer.py:587-588)
      a.losses[0] = align + a.alpha * div; a.losses[2] = div; a.losses[3] = 0.f;
    }
  }
}

// ---- backward row coefficients (incl. the upstream gradient): dz = coef (onehot - p) / coef (p - q)
struct CoefArgs {
  const DpaMeta* meta;
  const int* col_pair; const int* act_ref; const float* sigma; const float* d_loss;
  float* coef_pair; float* coef_ref;
  int B, B_ref; float alpha;
  float* formed_for;     // fused-gradient forward: remembers the upstream gradient the gradients are being formed for
};
__global__ void coef_kernel(CoefArgs a) {
  const int p0 = a.meta->p0, nr_pad = a.meta->n_tot - p0;
  float go = a.d_loss ? *a.d_loss : 1.f;
  if (a.formed_for && blockIdx.x == 0 && threadIdx.x == 0) *a.formed_for = go;
  // Ill-formed phrase ids / labels: the reference raises (halva_trainer.py:567) or returns NaN; the device cannot raise,
  // so the gradients are poisoned the same way loss_kernel poisons the loss -- finite-but-wrong dH / dW must never reach
  // the optimiser.
  if (a.meta->status & (DPA_STATUS_BAD_SIGN | DPA_STATUS_K_MISMATCH | DPA_STATUS_K_ZERO | DPA_STATUS_BAD_LABEL)) go = nanf("");
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p0 + nr_pad; r += gridDim.x * blockDim.x) {
    if (r < p0) {
      const int c = a.col_pair[r], k = c & 0xFFFF, s = c >> 16;
      float v = 0.f;
      if (k != kNoCol && r < a.meta->n_pair) {
        const int b = s >= a.B ? s - a.B : s;
        const float sg = a.sigma[(size_t)b * (kMaxSign + 1) + k] * go;
        v = s >= a.B ? sg : -sg;                       // d loss / d logp: -sigma (pos), +sigma (neg)
      }
      a.coef_pair[r] = v;
    } else {
      const int j = r - p0;
      a.coef_ref[j] = (j < a.meta->n_ref && a.act_ref[j]) ? go * a.alpha / (float)a.B_ref : 0.f;
    }
  }
}

// ---- fused-gradient forward: from the chunk's kept bf16 logits, dz = coef (softmax(z_pol) - softmax(z_frozen)) in place
//      over the policy logits, and the KL row sum_v q_v (z_r - z_p) - (LSE_r - LSE_p)  (halva_trainer.py:583-586).
//      HBM-bound elementwise pass: 4 B read + 2 B written per logit; one block per row at a time.
struct DzElemArgs {
  const DpaMeta* meta; int chunk, chunk_rows;
  __nv_bfloat16* dz; const __nv_bfloat16* zr; int ldz, v_local;
  const float* lse_p; const float* lse_r; const float* coef; const int* act;   // indexed by reference row j
  float* kl_row;                                                               // in: LSE_r - LSE_p, out: KL of the row
  int sub_dlse;     // vocab-parallel: the row's KL is the SUM over the shards of sum_v q (z_r - z_p) minus (LSE_r - LSE_p) once --
};                  // only one shard (rank 0) subtracts it; 1 without vocabulary sharding
__device__ __forceinline__ uint32_t dz_pack2(uint32_t zp, uint32_t zr, float c, float nlp, float nlr, float& t) {
  const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zp));
  const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zr));
  const float qx = fast_exp2(fmaf(r.x, kLog2e, nlr)), qy = fast_exp2(fmaf(r.y, kLog2e, nlr));
  t = fmaf(qx, r.x - p.x, t);
  t = fmaf(qy, r.y - p.y, t);
  return pack_bf16x2(c * (fast_exp2(fmaf(p.x, kLog2e, nlp)) - qx), c * (fast_exp2(fmaf(p.y, kLog2e, nlp)) - qy));
}
__global__ void __launch_bounds__(256) dz_kl_elem_kernel(DzElemArgs a) {
  __shared__ float sh[32];
  const int p0 = a.meta->p0, n_tot = a.meta->n_tot;
  const int c_lo = a.chunk * a.chunk_rows;
  const int lo = max(c_lo, p0), hi = min(c_lo + a.chunk_rows, n_tot);
  const int nvec = (a.v_local + 7) >> 3;             // columns past v_local hold logits of zero-filled weight rows (0 - 0)
  constexpr int U = 4;
  for (int r = lo + blockIdx.x; r < hi; r += gridDim.x) {
    const int j = r - p0;
    uint4* pz = reinterpret_cast<uint4*>(a.dz + (size_t)(r - c_lo) * a.ldz);
    const uint4* pr = reinterpret_cast<const uint4*>(a.zr + (size_t)(r - c_lo) * a.ldz);
    if (!a.act[j]) {                                  // padding / ignored rows: no loads at all
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) pz[v] = make_uint4(0, 0, 0, 0);
      if (threadIdx.x == 0) a.kl_row[j] = 0.f;
      continue;
    }
    const float c = a.coef[j];
    const float nlp = -a.lse_p[j] * kLog2e, nlr = -a.lse_r[j] * kLog2e;
    float t = 0.f;
    for (int v0 = threadIdx.x; v0 < nvec; v0 += U * blockDim.x) {
      uint4 zp[U], zr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * blockDim.x;
        if (v < nvec) { zp[u] = pz[v]; zr[u] = __ldg(pr + v); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * blockDim.x;
        if (v < nvec)
          pz[v] = make_uint4(dz_pack2(zp[u].x, zr[u].x, c, nlp, nlr, t), dz_pack2(zp[u].y, zr[u].y, c, nlp, nlr, t),
                             dz_pack2(zp[u].z, zr[u].z, c, nlp, nlr, t), dz_pack2(zp[u].w, zr[u].w, c, nlp, nlr, t));
      }
    }
    const float tot = block_sum(t, sh);
    if (threadIdx.x == 0) a.kl_row[j] = a.sub_dlse ? tot - a.kl_row[j] : tot;   // sum_v q (z_r - z_p) - (LSE_r - LSE_p)
  }
}

// ---- fused-gradient forward with the frozen model's logits GIVEN (dpa_args_t::ref_logits) instead of computed: what the
//      stash_frz GEMM would have left for the chunk's reference-sample rows -- bf16 logits in zr and the per (row, column
//      group) {max, sum-exp} partials -- read from the caller's [rows, V] tensor.  HBM-bound: 4 (fp32) or 2 (bf16) bytes
//      read + 2 written per logit of an ACTIVE row; one block per row, one warp per column group (8 columns per lane and
//      step, 256 contiguous columns per warp and step).
struct StashGivenArgs {
  const DpaMeta* meta; int chunk, chunk_rows;
  const void* logits; long long ld;                 // [B_ref * T_ref][ld], column c = vocabulary id v_offset + c
  const int* src_ref; const int* act_ref;           // indexed by reference row j
  __nv_bfloat16* zr; int ldz, v_local;
  float2* part; int row_cap, n_groups, group_cols;  // [n_groups][row_cap][2]: slot 1 = frozen model
};
template <bool F32>
__global__ void __launch_bounds__(256) stash_given_kernel(StashGivenArgs a) {
  const int p0 = a.meta->p0, n_tot = a.meta->n_tot;
  const int c_lo = a.chunk * a.chunk_rows;
  const int lo = max(c_lo, p0), hi = min(c_lo + a.chunk_rows, n_tot);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int kSteps = 5;                          // up to 5 x 256 columns = one column group of the statistics GEMMs
  const int steps = a.group_cols / 256;              // (1 for a small vocabulary shard: single-tile units)
  for (int r = lo + blockIdx.x; r < hi; r += gridDim.x) {
    const int j = r - p0;
    if (!a.act_ref[j]) {                             // padding / ignored rows: finite dummies, never used (coef == 0)
      for (int g = threadIdx.x; g < a.n_groups; g += blockDim.x) a.part[((size_t)g * a.row_cap + j) * 2 + 1] = make_float2(0.f, 1.f);
      continue;
    }
    const size_t src = (size_t)a.src_ref[j] * (size_t)a.ld;
    __nv_bfloat16* out = a.zr + (size_t)(r - c_lo) * a.ldz;
    for (int g = warp; g < a.n_groups; g += nw) {
      float v[kSteps * 8];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < kSteps; ++i) {
        const int col = g * a.group_cols + i * 256 + lane * 8;
        float* x = v + i * 8;
        if (i >= steps) {                              // past this (narrower) group
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = -INFINITY;
          continue;
        }
        if (col + 8 <= a.v_local) {
          if constexpr (F32) {
            const float4* q = reinterpret_cast<const float4*>(static_cast<const float*>(a.logits) + src + col);
            const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
            x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w; x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w;
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = bf16_round_f32(x[k]);
          } else {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(a.logits) + src + col));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = bf2_to_f2(w[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float t = -INFINITY;
            if (col + k < a.v_local) {
              if constexpr (F32) t = bf16_round_f32(static_cast<const float*>(a.logits)[src + col + k]);
              else t = __bfloat162float(static_cast<const __nv_bfloat16*>(a.logits)[src + col + k]);
            }
            x[k] = t;
          }
        }
        if (col < a.v_local) {
          float z[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { z[k] = (col + k < a.v_local) ? x[k] : 0.f; m = fmaxf(m, x[k]); }
          *reinterpret_cast<uint4*>(out + col) = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]),
                                                           pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
        }
      }
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float mb = m * kLog2e;
      float ssum = 0.f;
#pragma unroll
      for (int i = 0; i < kSteps * 8; ++i) ssum += fast_exp2(fmaf(v[i], kLog2e, -mb));    // -inf (past v_local) -> 0
      for (int o = 16; o; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      if (lane == 0) a.part[((size_t)g * a.row_cap + j) * 2 + 1] = make_float2(m, ssum);
    }
  }
}

// ---- fused-gradient forward, pair rows of row chunk 0: dz = coef (onehot(y) - softmax(z)) in place over the logits the
//      statistics GEMM kept (2 B read + 2 B written per logit) -- the recompute GEMM only runs for pair rows beyond chunk 0
struct DzPairElemArgs {
  const DpaMeta* meta; int row_lo, row_hi;          // packed pair rows [row_lo, min(row_hi, p0)) -- the chunk's kept logits
  __nv_bfloat16* dz; int ldz, v_local, v_offset;
  const int* labels; const float* lse; const float* coef;        // indexed by packed row
};
__device__ __forceinline__ uint32_t dzp_pack2(uint32_t z, float c, float nl, int col, int y) {
  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&z));
  float a = -c * fast_exp2(fmaf(f.x, kLog2e, nl)), b = -c * fast_exp2(fmaf(f.y, kLog2e, nl));
  a += (col == y) ? c : 0.f;
  b += (col + 1 == y) ? c : 0.f;
  return pack_bf16x2(a, b);
}
__global__ void __launch_bounds__(256) dz_pair_elem_kernel(DzPairElemArgs a) {
  const int hi = min(a.meta->p0, a.row_hi);
  const int nvec = (a.v_local + 7) >> 3;
  constexpr int U = 4;
  for (int r = a.row_lo + blockIdx.x; r < hi; r += gridDim.x) {
    uint4* pz = reinterpret_cast<uint4*>(a.dz + (size_t)r * a.ldz);
    const float c = a.coef[r];
    if (c == 0.f) {                                   // padding rows (and ignored rows in dense mode)
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) pz[v] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const float nl = -a.lse[r] * kLog2e;
    const int y = a.labels[r] - a.v_offset;
    for (int v0 = threadIdx.x; v0 < nvec; v0 += U * blockDim.x) {
      uint4 z[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int v = v0 + u * blockDim.x; if (v < nvec) z[u] = pz[v]; }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * blockDim.x;
        if (v < nvec) {
          const int col = v * 8;
          pz[v] = make_uint4(dzp_pack2(z[u].x, c, nl, col, y), dzp_pack2(z[u].y, c, nl, col + 2, y),
                             dzp_pack2(z[u].z, c, nl, col + 4, y), dzp_pack2(z[u].w, c, nl, col + 6, y));
        }
      }
    }
  }
}

// ---- fp32 inputs, fused-gradient forward: the kept logits are fp32 (the promoted sums of the split-bf16 GEMMs), the dLogits
//      leave as split-bf16 triples (hi, mid, lo `part` elements apart) for the split gradient GEMMs: 4-8 B read + 6 B written
//      per logit.  Same formulas as the bf16 kernels above and as the recompute epilogues (EpiDzPair / EpiDzRef, SPLIT).
struct DzPairElemF32Args {
  const DpaMeta* meta; int row_lo, row_hi;
  const float* z; __nv_bfloat16* dz; size_t part; int ldz, v_local, v_offset;
  const int* labels; const float* lse; const float* coef;
};
__global__ void __launch_bounds__(256) dz_pair_elem_f32_kernel(DzPairElemF32Args a) {
  const int hi = min(a.meta->p0, a.row_hi);
  const int nvec = (a.v_local + 7) >> 3;
  for (int r = a.row_lo + blockIdx.x; r < hi; r += gridDim.x) {
    const float4* pz = reinterpret_cast<const float4*>(a.z + (size_t)(r - a.row_lo) * a.ldz);
    uint4* po = reinterpret_cast<uint4*>(a.dz + (size_t)(r - a.row_lo) * a.ldz);
    const size_t pv = a.part >> 3;
    const float c = a.coef[r];
    if (c == 0.f) {
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) { po[v] = make_uint4(0, 0, 0, 0); po[v + pv] = make_uint4(0, 0, 0, 0); po[v + 2 * pv] = make_uint4(0, 0, 0, 0); }
      continue;
    }
    const float nl = -a.lse[r] * kLog2e;
    const int y = a.labels[r] - a.v_offset;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const float4 z0 = pz[2 * v], z1 = pz[2 * v + 1];
      float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      const int col = v * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) { z[k] = -c * fast_exp2(fmaf(z[k], kLog2e, nl)); z[k] += (col + k == y) ? c : 0.f; }
      uint4 o[3];
      split8(make_float4(z[0], z[1], z[2], z[3]), make_float4(z[4], z[5], z[6], z[7]), o);
      po[v] = o[0]; po[v + pv] = o[1]; po[v + 2 * pv] = o[2];
    }
  }
}
struct DzElemF32Args {
  const DpaMeta* meta; int chunk, chunk_rows;
  const float* zp; const float* zr; __nv_bfloat16* dz; size_t part; int ldz, v_local;
  const float* lse_p; const float* lse_r; const float* coef; const int* act;
  float* kl_row;
};
__global__ void __launch_bounds__(256) dz_kl_elem_f32_kernel(DzElemF32Args a) {
  __shared__ float sh[32];
  const int p0 = a.meta->p0, n_tot = a.meta->n_tot;
  const int c_lo = a.chunk * a.chunk_rows;
  const int lo = max(c_lo, p0), hi = min(c_lo + a.chunk_rows, n_tot);
  const int nvec = (a.v_local + 7) >> 3;
  const size_t pv = a.part >> 3;
  for (int r = lo + blockIdx.x; r < hi; r += gridDim.x) {
    const int j = r - p0;
    const float4* pp = reinterpret_cast<const float4*>(a.zp + (size_t)(r - c_lo) * a.ldz);
    const float4* pr = reinterpret_cast<const float4*>(a.zr + (size_t)(r - c_lo) * a.ldz);
    uint4* po = reinterpret_cast<uint4*>(a.dz + (size_t)(r - c_lo) * a.ldz);
    if (!a.act[j]) {
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) { po[v] = make_uint4(0, 0, 0, 0); po[v + pv] = make_uint4(0, 0, 0, 0); po[v + 2 * pv] = make_uint4(0, 0, 0, 0); }
      if (threadIdx.x == 0) a.kl_row[j] = 0.f;
      continue;
    }
    const float c = a.coef[j];
    const float nlp = -a.lse_p[j] * kLog2e, nlr = -a.lse_r[j] * kLog2e;
    float t = 0.f;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const float4 a0 = pp[2 * v], a1 = pp[2 * v + 1], b0 = __ldg(pr + 2 * v), b1 = __ldg(pr + 2 * v + 1);
      const float zp[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, zr[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float d[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool in = v * 8 + k < a.v_local;          // columns past v_local hold logits of zero-filled weight rows
        const float q = in ? fast_exp2(fmaf(zr[k], kLog2e, nlr)) : 0.f;
        t = fmaf(q, zr[k] - zp[k], t);
        d[k] = in ? c * (fast_exp2(fmaf(zp[k], kLog2e, nlp)) - q) : 0.f;
      }
      uint4 o[3];
      split8(make_float4(d[0], d[1], d[2], d[3]), make_float4(d[4], d[5], d[6], d[7]), o);
      po[v] = o[0]; po[v + pv] = o[1]; po[v + 2 * pv] = o[2];
    }
    const float tot = block_sum(t, sh);
    if (threadIdx.x == 0) a.kl_row[j] = tot - a.kl_row[j];
  }
}
// ---- fused-gradient backward: the gradients were formed during the forward for d_loss = 1 (or the caller's hint); scale them by the
//      upstream gradient (a device scalar) unless they were formed for exactly that value (1, or the caller's hint)
struct ScaleSegs { void* p[4]; size_t n[4]; int count; };       // up to four gradient tensors, one launch
template <class T>
__global__ void scale_grads_kernel(ScaleSegs sg, const float* __restrict__ d_loss, const float* __restrict__ formed_for) {
  const float want = *d_loss, have = *formed_for;
  if (want == have) return;
  const float go = want / have;          // (have == 0 with want != 0: the zero gradients cannot be rescaled -> NaN, loudly)
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = 0; k < sg.count; ++k) {
    uint4* p = reinterpret_cast<uint4*>(sg.p[k]);
    const size_t nv = sg.n[k] * sizeof(T) / 16;
    for (size_t i = t0; i < nv; i += stride) {
      uint4 v = p[i];
      uint32_t* w = reinterpret_cast<uint32_t*>(&v);
      if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
          w[j] = pack_bf16x2(f.x * go, f.y * go);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = __float_as_uint(__uint_as_float(w[j]) * go);
      }
      p[i] = v;
    }
  }
}

// dW when no row at all contributed (no chunk ran): zero it; otherwise the last chunk's GEMM epilogue already wrote it
__global__ void zero_if_no_rows_kernel(const DpaMeta* meta, void* out, size_t bytes) {
  if (meta->n_tot != 0) return;
  uint4* p = reinterpret_cast<uint4*>(out);
  const size_t nv = bytes >> 4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) p[i] = make_uint4(0, 0, 0, 0);
}

// dense gradient outputs -> 0 before the dH epilogues scatter the contributing rows into them (two tensors, one launch;
// 16-byte streaming stores, four in flight per thread)
struct ZeroSegs { void* p[2]; size_t bytes[2]; };
__global__ void __launch_bounds__(512) zero_fill_kernel(ZeroSegs z) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint4* p = reinterpret_cast<uint4*>(z.p[k]);
    const size_t nv = z.bytes[k] >> 4;
    size_t i = t0;
    for (; i + 3 * stride < nv; i += 4 * stride) {
      __stcs(p + i, zero); __stcs(p + i + stride, zero); __stcs(p + i + 2 * stride, zero); __stcs(p + i + 3 * stride, zero);
    }
    for (; i < nv; i += stride) __stcs(p + i, zero);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(out + i) = o;
  }
}

}  // namespace dpa
