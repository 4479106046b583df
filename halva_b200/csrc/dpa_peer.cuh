// Peer-memory exchange of the vocab-parallel lm_head mode: the dH partial sums of all vocabulary shards are reduced
// over NVLink / NVSwitch by this library's own kernels, reading and writing the peers' buffers directly.
//
// Why not ncclAllReduce: the number of contributing rows lives in DEVICE memory (DpaMeta::n_tot, written by the metadata
// kernels of the same call) and a collective with a host-side count would need a device->host read-back and a stream
// synchronisation in the middle of the step.  Here every kernel takes its row range from DpaMeta, so the whole exchange is
// stream-ordered, touches only the rows that exist (a quarter of the dense tensors at 10 % phrase density), sums the shards
// in a fixed order in fp32 (every rank ends up with bit-identical rows), and runs beside the chunk's dW GEMM.
//
// Layout of one rank's symmetric allocation (cudaMalloc by dpa_comm's first use, mapped into every peer with CUDA IPC):
//   [0, 4096)   control: uint32 flag[kPeerSlots][8] (written by the peers), uint32 epoch[kPeerSlots] (local), float kl_part[8]
//   [4096, ..)  statistics gather region [G][record]: rank r pushes its per-row statistics record rows into slot r of
//               every peer (the forward's exchange: 36 B/row, only the rows of the segment / chunk being finished)
//   [.., ..)    two chunk buffers [chunk_rows][H] bf16: the dH GEMM of chunk c writes its partial sums into buffer c & 1
// Protocol per chunk, on the communicator's stream (forked from the caller's stream after the chunk's dH GEMM):
//   barrier(slot A)  -- every rank's partial sums of the chunk are complete
//   reduce           -- rank r sums rows slice r of all G buffers (peer loads) and writes the result back into slice r of
//                       all G buffers (peer stores): an in-place all-reduce, each slice touched by exactly one rank
//   barrier(slot B)  -- every rank's slice has landed everywhere
//   scatter          -- local: packed rows -> their [s, t] positions in dH_pair / dH_ref
#pragma once
#include "dpa_kernels.cuh"

namespace dpa {

constexpr int kPeerMax = 8;                 // ranks of one NVSwitch box
constexpr int kPeerSlots = 16;              // barrier slots (flag rows) in the control block
constexpr size_t kPeerCtrlBytes = 4096;
constexpr int kPeerFlagWords = kPeerSlots * kPeerMax;      // epoch[] follows
constexpr size_t kPeerLossOff = 1024;       // float kl_part[8] inside the control block: every shard's part of the divergence

struct PeerView {
  char* base[kPeerMax];                     // base[r]: rank r's symmetric allocation as mapped in THIS process
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bulk data moves with weak accesses (a little faster than system-scope relaxed ones: 0.84 vs 0.91 ms per cfg-3 exchange on
// two GPUs), ordered by the fences / flag barriers around them: loads bypass L1 (.cg; L1 is invalidated at kernel start anyway
// and peer memory is never cached in this GPU's L2), stores are followed by __threadfence_system() before the next barrier.
__device__ __forceinline__ uint4 ld_peer_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v4(uint4* p, const uint4& v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One warp: lane r signals rank r (epoch of this slot, kept in the local control block, so a CUDA-graph replay advances it
// too) and waits for rank r's signal.  Completes only when every rank's stream has reached the same barrier.
__global__ void peer_barrier_kernel(PeerView v, int slot) {
  unsigned* ctl = reinterpret_cast<unsigned*>(v.base[v.rank]);
  const int lane = threadIdx.x;
  const unsigned e = ctl[kPeerFlagWords + slot] + 1u;
  __syncwarp();
  if (lane == 0) ctl[kPeerFlagWords + slot] = e;
  __threadfence_system();                   // everything this GPU wrote before (own kernels, stream order) is visible first
  if (lane < v.world) {
    st_release_sys(reinterpret_cast<unsigned*>(v.base[lane]) + slot * kPeerMax + v.rank, e);
    const unsigned* mine = ctl + slot * kPeerMax + lane;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      __nanosleep(64);
      if (++spins > (1u << 26)) __trap();   // seconds: a peer died or the ranks disagree on the call sequence
    }
  }
}

struct PeerReduceArgs {
  PeerView v;
  const DpaMeta* meta; int chunk, chunk_rows, H;
  size_t buf_off;                           // byte offset of this chunk's buffer inside every rank's allocation
};
// In-place all-reduce of the chunk's packed dH rows [c_lo, min(c_hi, n_tot)): this rank owns slice `rank` of the rows.
// G = ranks (compile-time: the G peer loads of a vector are independent and U = 16 / G vectors are in flight per thread).
template <int G>
__global__ void __launch_bounds__(256, 2) peer_reduce_rows_kernel(PeerReduceArgs a) {
  const int c_lo = a.chunk * a.chunk_rows;
  const int rows = max(min(c_lo + a.chunk_rows, a.meta->n_tot) - c_lo, 0);
  const long long vec_row = a.H / 8;                                    // uint4 per row
  const long long v0 = (long long)rows * a.v.rank / G * vec_row, v1 = (long long)rows * (a.v.rank + 1) / G * vec_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = G >= 16 ? 1 : 16 / G;
  for (long long i0 = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < v1; i0 += U * stride) {
    uint4 in[U][G];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
#pragma unroll
      for (int r = 0; r < G; ++r)
        if (i < v1) in[u][r] = ld_peer_v4(reinterpret_cast<const uint4*>(a.v.base[r] + a.buf_off) + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < v1) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < G; ++r) {                                   // fixed order: identical result on every rank
          const uint32_t w[4] = {in[u][r].x, in[u][r].y, in[u][r].z, in[u][r].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 f = bf2_to_f2(w[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
        }
        const uint4 out = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
#pragma unroll
        for (int r = 0; r < G; ++r) st_peer_v4(reinterpret_cast<uint4*>(a.v.base[r] + a.buf_off) + i, out);
      }
    }
  }
  __threadfence_system();                   // this thread's peer stores are ordered before the barrier that follows
}
inline cudaError_t launch_peer_reduce(const PeerReduceArgs& a, int blocks, cudaStream_t st) {
  switch (a.v.world) {
    case 2: peer_reduce_rows_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: peer_reduce_rows_kernel<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: peer_reduce_rows_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    case 5: peer_reduce_rows_kernel<5><<<blocks, 256, 0, st>>>(a); break;
    case 6: peer_reduce_rows_kernel<6><<<blocks, 256, 0, st>>>(a); break;
    case 7: peer_reduce_rows_kernel<7><<<blocks, 256, 0, st>>>(a); break;
    case 8: peer_reduce_rows_kernel<8><<<blocks, 256, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// packed rows of one chunk (already reduced, local buffer) -> their [s, t] positions; one warp per row
struct ScatterChunkArgs {
  const DpaMeta* meta; int chunk, chunk_rows;
  const int* src_pair; const int* src_ref;
  const __nv_bfloat16* packed;              // [chunk_rows][H]: row i = packed row c_lo + i
  __nv_bfloat16* dH_pair; __nv_bfloat16* dH_ref; int H;
};
__global__ void scatter_chunk_rows_kernel(ScatterChunkArgs a) {
  const int p0 = a.meta->p0, n_pair = a.meta->n_pair, n_ref = a.meta->n_ref;
  const int c_lo = a.chunk * a.chunk_rows, c_hi = min(c_lo + a.chunk_rows, a.meta->n_tot);
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int vec = a.H / 8;
  for (int r = c_lo + wid; r < c_hi; r += warps) {
    int s = -1; bool is_pair = r < p0;
    if (is_pair) { if (r < n_pair) s = a.src_pair[r]; }
    else { const int j = r - p0; if (j < n_ref) s = a.src_ref[j]; }
    if (s < 0) continue;                    // padding rows
    const uint4* src = reinterpret_cast<const uint4*>(a.packed + (size_t)(r - c_lo) * a.H);
    uint4* dst = reinterpret_cast<uint4*>((is_pair ? a.dH_pair : a.dH_ref) + (size_t)s * a.H);
    for (int i = lane; i < vec; i += 32) dst[i] = __ldcg(src + i);
  }
}

}  // namespace dpa
