// Persistent, warp-specialised tcgen05 GEMM for sm_100a with pluggable fused epilogues.
//
//   D[128*CG x BN] (+)= A[128*CG x K] * B[BN x K]^T      bf16 operands, fp32 accumulators in TMEM
//   CG = 2 (default): a cluster of two CTAs on one TPC shares the tile -- tcgen05.mma.cta_group::2, M = 256, each CTA
//   stages its own 128 rows of A and half of B; CG = 1: one CTA per 128-row tile.
//
//   warps 0..3         epilogue: tcgen05.ld (thread == accumulator row) -> fused maths -> HBM
//   warp 4 (one lane)  TMA producer: cp.async.bulk.tensor -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 5 (one lane)  tcgen05.mma issuer; tcgen05.commit frees smem slots and publishes accumulators
//   warp 6             TMEM allocator (all 512 columns: two accumulator stages)
//
// The DPA loss uses it through its epilogues (see dpa_api.cu, struct Plan): forward row statistics of pair rows
// (optionally keeping their bf16 logits), statistics + kept logits of reference-sample rows (fused-gradient forward),
// the DUAL variants that hold the policy and frozen-model logits of the same tile side by side in TMEM (statistics
// incl. the KL numerator; recompute of dLogits), the recompute of pair-row dLogits, and the two gradient GEMMs
// dH = dZ*W, dW = dZ^T*H.  fp32 inputs run the same classes through dpa_gemm_split_kernel: split-bf16 operand triples
// as six K segments, K-chunk promotion of the accumulators through registers (GemmCfg::NSEG / PROMOTE).
// A work "unit" is (row tile, group of consecutive column tiles); row counts live in device memory
// (DpaMeta) so that nothing on the path needs a host synchronisation.
#pragma once
#include "dpa_ptx.cuh"

namespace dpa {

constexpr int BM = 128;      // rows per CTA tile == TMEM lanes == epilogue threads
constexpr int kRowPad = 256; // packed row segments and row chunks are padded to this (one cta_group::2 tile pair)
constexpr int BK = 64;       // K per smem stage: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;   // K per tcgen05.mma (bf16)
constexpr int kNumThreads = 256;
// warp roles: 0-3 epilogue (TMEM lane quadrant = warp id), 4 TMA producer, 5 MMA issuer, 6 TMEM allocator.  The SM's
// schedulers favour the higher warp id of a sub-partition, so the two latency-critical single-lane roles sit above the
// epilogue warps they share a sub-partition with.
#ifndef DPA_ROLES_LOW
constexpr int kProducerWarp = 4, kMmaWarp = 5, kAllocWarp = 6, kEpiWarp0 = 0;
#else   // experiment: single-lane roles below the epilogue warps
constexpr int kProducerWarp = 0, kMmaWarp = 1, kAllocWarp = 2, kEpiWarp0 = 4;
#endif
#ifndef DPA_SUPER_M
#define DPA_SUPER_M 16
#endif
constexpr int kSuperM = DPA_SUPER_M;  // row-tile units per L2 super-block of the unit order

enum : int { KMAJOR = 0, MNMAJOR = 1 };

// Device-resident bookkeeping written by the metadata kernels (dpa_kernels.cu). All int32.
struct DpaMeta {
  int n_pair;    // active pair rows (packed order: sequence-major, pos sequences first)
  int n_ref;     // active reference-sample rows
  int p0;        // round_up(n_pair, kRowPad): first packed row of the reference-sample segment
  int n_tot;     // p0 + round_up(n_ref, kRowPad): padded rows of the packed policy space
  int K;         // phrase columns (K_pos == K_neg)
  int K_pos, K_neg;
  int status;    // 0 = ok, else DPA_STATUS_* bits
  int pad[8];
};

enum : int {
  KIND_STATIC = 0,   // debug GEMM: static M
  KIND_FWD_PAIR,     // rows [0, p0)
  KIND_FWD_REF,      // rows [p0, n_tot) of the policy buffer + rows [0, ..) of the frozen buffer
  KIND_DZ_PAIR,      // chunk ∩ [0, p0)
  KIND_DZ_REF,       // chunk ∩ [p0, n_tot)
  KIND_DH,           // chunk rows x H, K = V
  KIND_DW,           // V x H, K = chunk rows
  KIND_STASH_POL,    // chunk ∩ [p0, n_tot), policy rows: statistics + bf16 logits kept for the chunk
  KIND_STASH_FRZ,    // same rows of the frozen-model buffer
};

struct GemmWork {
  const DpaMeta* meta;
  int kind;
  int chunk;           // backward chunk index
  int chunk_rows;      // rows per backward chunk (multiple of 128)
  int m_static;        // KIND_STATIC / KIND_DW: M extent
  int n_cols;          // N extent
  int k_static;        // K extent unless KIND_DW
  int tiles_per_unit;  // consecutive column tiles per work unit
  int super_m;         // row-tile units per L2 super-block of the unit order (0 = kSuperM)
  int hint_a, hint_b;  // L2 eviction priority of the A / B operand loads: 0 = none, 1 = evict_first, 2 = evict_last
  int dz_row0;         // KIND_DH / KIND_DW: first row of this chunk's slot inside the kept-logits / dz buffer the dz map spans
  // K-lockstep of the CTA groups of a launch (0 = off): before the producer loads k-block kb with kb % sync_every == 0 it
  // waits until every group that runs a unit in the same round has reached that k-block of that tile.  The groups of a
  // round share operand panels; held within sync_every (+ the smem ring) k-blocks of each other, a panel's k-slices stay
  // L2-resident between their first and last consumer instead of being re-read from DRAM -- measured on the long-K gradient
  // GEMMs: 13.4 -> 5.6 GB read per launch, and the whole (power-bound) step 6 % faster (profiles/r02e_klockstep.md).
  int sync_every;
  unsigned* sync_ctr;  // one zero-initialised counter per launch
};

struct TileRange {
  int m_tiles;   // row tiles of this launch
  int a_row0;    // A  row coordinate of tile 0
  int a2_row0;   // A2 row coordinate of tile 0 (DUAL)
  int e_row0;    // epilogue's logical row of tile 0
  int k_blocks;  // BK-blocks along K
  int a_k0, b_k0;
};

__device__ __forceinline__ TileRange compute_range(const GemmWork& w) {
  TileRange r;
  r.a2_row0 = 0; r.a_k0 = 0; r.b_k0 = 0;
  r.k_blocks = (w.k_static + BK - 1) / BK;
  if (w.kind == KIND_STATIC) {
    r.m_tiles = (w.m_static + BM - 1) / BM; r.a_row0 = 0; r.e_row0 = 0;
    return r;
  }
  const int p0 = w.meta->p0, n_tot = w.meta->n_tot;
  const int c_lo = w.chunk * w.chunk_rows, c_hi = c_lo + w.chunk_rows;
  int lo = 0, hi = 0;
  switch (w.kind) {
    case KIND_FWD_PAIR: lo = 0; hi = p0; r.a_row0 = 0; r.e_row0 = 0; break;
    case KIND_FWD_REF: lo = p0; hi = n_tot; r.a_row0 = p0; r.a2_row0 = 0; r.e_row0 = 0; break;
    case KIND_DZ_PAIR: lo = c_lo; hi = min(c_hi, p0); r.a_row0 = lo; r.e_row0 = lo; break;
    case KIND_DZ_REF:
    case KIND_STASH_POL: lo = max(c_lo, p0); hi = min(c_hi, n_tot); r.a_row0 = lo; r.a2_row0 = lo - p0; r.e_row0 = lo; break;
    case KIND_STASH_FRZ: lo = max(c_lo, p0); hi = min(c_hi, n_tot); r.a_row0 = lo - p0; r.e_row0 = lo; break;
    case KIND_DH: lo = c_lo; hi = min(c_hi, n_tot); r.a_row0 = w.dz_row0; r.e_row0 = lo; break;
    case KIND_DW: {
      lo = c_lo; hi = min(c_hi, n_tot);
      const int rows = max(hi - lo, 0);
      r.m_tiles = rows > 0 ? (w.m_static + BM - 1) / BM : 0;
      r.a_row0 = 0; r.e_row0 = 0; r.k_blocks = rows / BK; r.a_k0 = w.dz_row0; r.b_k0 = lo;
      return r;
    }
  }
  r.m_tiles = max(hi - lo, 0) / BM;
  return r;
}

// NSEG = 6: split-bf16 operands for fp32 inputs ("bf16x6").  x = hi + mid + lo, each part the bf16 rounding of what the
// previous parts left (the residuals are exact in fp32; three parts carry 24 mantissa bits).  The product keeps the six
// terms down to 2^-16 of the leading one,  hi*hi + hi*mid + mid*hi + hi*lo + lo*hi + mid*mid  (the dropped ones are
// <= 2^-24 relative, the fp32 rounding level), accumulated in the same TMEM accumulator as six K segments.
struct SplitMaps { CUtensorMap A[2], B[2], A2[2], B2[2]; };   // tensor maps of the (mid, lo) parts; A2/B2: DUAL's second pair
// Segment order: smallest terms first, hi*hi last.  The tensor core adds each K=16 step into the fp32 accumulator with
// truncation, an error proportional to the accumulator's magnitude at that step: with the large term last, only its own
// K/16 steps run on a large accumulator (measured: the error of an H=4096 logit against fp64 halves vs hi*hi first).
//   segment:       0       1      2      3       4      5
//   (A, B) part: (mid,mid) (hi,lo) (lo,hi) (hi,mid) (mid,hi) (hi,hi)
constexpr uint32_t kSegPartA = 0x010201u, kSegPartB = 0x001021u;   // nibble s = part (0 hi, 1 mid, 2 lo) of A / B in segment s

template <int BN_, bool DUAL_, int AMAJ_, int BMAJ_, int CG_ = 1, int NSEG_ = 1>
struct GemmCfg {
  static constexpr int NSEG = NSEG_;
  static_assert(NSEG == 1 || NSEG == 6, "NSEG");
  static constexpr int BN = BN_;                 // N of one tcgen05.mma / columns of one accumulator
  static constexpr bool DUAL = DUAL_;
  static constexpr int AMAJ = AMAJ_, BMAJ = BMAJ_;
  static constexpr int CG = CG_;                 // cta_group: 2 = a CTA pair shares one M=256 MMA, each staging half of B
  static constexpr int BN_CTA = BN / CG;         // B rows staged by this CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (DUAL ? 2 : 1);
  static constexpr int STAGES_RAW = (227 * 1024 - 1280) / STAGE_BYTES;   // 7 x 32 KiB (pair), 4 x 48 KiB (single CTA / dual pair), 3 x 64 KiB (dual single CTA)
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = DUAL ? 2 * BN : BN;             // TMEM columns per accumulator stage
  // BN = 512 (the long-K gradient GEMMs): one accumulator stage of all 512 TMEM columns, two N = 256 MMAs per K step.
  // A 256 x 512 tile per CTA pair moves 25 % fewer operand bytes per FLOP than 256 x 256, a wave of 74 pairs covers every
  // column tile of its row units (so each dz panel is read once), and the CTAs of a wave re-align at every tile boundary
  // (there is no second stage to run ahead into); the epilogue is no longer hidden, but K is 250-500 blocks long.
  static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;
  static constexpr int UMMA_N = BN > 256 ? 256 : BN;              // N of one tcgen05.mma
  static constexpr int NMMA = BN / UMMA_N;
  static constexpr uint32_t B_MMA_STEP = ((UMMA_N / CG) * 128) >> 4;   // smem distance (>>4) between the B rows of two MMAs
  // PROMOTE (the split-operand / fp32 path): the tensor core adds each K = 16 step into its fp32 accumulator with
  // TRUNCATION (tools/tc_accum_probe.py: a same-sign K = 4096 dot product comes out 9e-6 low; the bias grows with the
  // accumulator's magnitude and the number of steps).  So the K loop is cut into chunks of PROMOTE_KB k-blocks, each chunk
  // accumulates from zero in one of the two TMEM stages, and the epilogue warps add the chunk results in registers with
  // round-to-nearest (the other stage is being filled meanwhile); the sum is parked in a third TMEM region for the
  // unchanged fused epilogue.  Inside a chunk the accumulator stays 1 / #chunks of the result, and so does the bias.
  static constexpr bool PROMOTE = NSEG > 1;
  static constexpr int PROMOTE_KB = 4;                             // 4 k-blocks = K 256 = 16 tensor-core steps per chunk
  static constexpr int MAIN_COL = 2 * ACC_COLS;                    // TMEM column of the promoted sum
  static_assert(!PROMOTE || (3 * ACC_COLS <= 512 && ACC_COLS <= 128), "promotion keeps ACC_COLS fp32 registers per thread");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(ACC_STAGES * ACC_COLS <= 512, "the accumulator stages must fit TMEM");
  static_assert(BN % 64 == 0 && BN <= 512 && (BMAJ_ == 0 /* K-major: one box of BN_CTA rows */ ? (BN_CTA % 16 == 0 && BN_CTA <= 256) : BN_CTA % 64 == 0), "BN");
  static_assert(BN <= 256 || (!DUAL && NSEG == 1), "512-column tiles: single accumulator, plain bf16 operands");
  static_assert(CG == 1 || CG == 2, "cta_group");
};

// issue the TMA loads of one operand tile for one K block
template <int CG>
__device__ __forceinline__ void load_box(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t policy) {
  if (policy) {
    if constexpr (CG == 1) tma_load_2d_hint(dst, tm, bar, c0, c1, policy); else tma_load_2d_pair_hint(dst, tm, bar, c0, c1, policy);
  } else {
    if constexpr (CG == 1) tma_load_2d(dst, tm, bar, c0, c1); else tma_load_2d_pair(dst, tm, bar, c0, c1);
  }
}
template <int MAJ, int ROWS, int CG>
__device__ __forceinline__ void load_operand(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int row, int k, uint64_t policy = 0) {
  if constexpr (MAJ == KMAJOR) {                               // box {64 k, ROWS}
    load_box<CG>(dst, tm, bar, k, row, policy);
  } else {
#pragma unroll
    for (int j = 0; j < ROWS / 64; ++j)                        // box {64 mn, 64 k}: [atom][k][64 mn]
      load_box<CG>(dst + j * (BK * 128), tm, bar, row + j * 64, k, policy);
  }
}
__device__ __forceinline__ uint64_t l2_policy(int hint) { return hint == 1 ? kPolicyEvictFirst : hint == 2 ? kPolicyEvictLast : 0ull; }
// K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart, K advances inside the row (UMMA_K * 2 bytes per MMA).
// MN-major: atoms [64 mn x 8 k] 1024 B apart along K, BK * 128 B apart along MN (UMMA_K * 128 bytes per MMA).
template <int MAJ> struct OperandDesc {
  static constexpr uint64_t HI = MAJ == KMAJOR ? smem_desc_hi_sw128(0, 1024) : smem_desc_hi_sw128(BK * 128, 1024);
  static constexpr uint32_t KSTEP = (MAJ == KMAJOR ? UMMA_K * 2 : UMMA_K * 128) >> 4;   // in the >>4 address field
};

__device__ __forceinline__ void load_chunk(uint32_t taddr, float (&v)[32]);
__device__ __forceinline__ void store_chunk(uint32_t taddr, const float* v) {    // registers -> 32 fp32 TMEM columns
  __syncwarp();                      // tcgen05.st is warp-collective (.sync.aligned)
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
  tmem_st_32x32(taddr, r);
}

template <class Cfg, class Epi>
__device__ __forceinline__ void dpa_gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                                              const CUtensorMap& tmB2, const SplitMaps* sm, const GemmWork& work,
                                              const typename Epi::Params& ep) {
  constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES, CG = Cfg::CG, NSEG = Cfg::NSEG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
  uint8_t* sA2 = sB + STAGES * Cfg::B_BYTES;
  uint8_t* sB2 = sA2 + (Cfg::DUAL ? STAGES * Cfg::A_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                      // CG == 2: only the leader CTA's are used
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;   // CG == 2: only the leader CTA's are used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CG == 2: the two CTAs of a cluster work on one 256-row tile pair; rank 0 ("leader") issues the MMAs for both
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int group_id = blockIdx.x / CG, n_cta_groups = gridDim.x / CG;

  const TileRange tr = compute_range(work);
  const int n_tiles = (work.n_cols + BN - 1) / BN;
  const int tpu = work.tiles_per_unit;
  const int n_groups = (n_tiles + tpu - 1) / tpu;
  const int m_units = (tr.m_tiles + CG - 1) / CG;             // row tiles of BM * CG rows
  const int n_units = tr.k_blocks > 0 ? m_units * n_groups : 0;

  if constexpr (CG == 2) cluster_sync_all();                   // both CTAs resident before the paired TMEM allocation
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], CG); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG * BM); }   // (ACC_STAGES == 1 uses [0] only)
    fence_barrier_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (Cfg::DUAL) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == kAllocWarp) { tmem_alloc<CG>(tmem_slot, 512); tmem_relinquish<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (row-tile unit, column group): super-blocks of super_m row units walk the column groups together
  const int super_m = work.super_m > 0 ? work.super_m : kSuperM;
  auto decode = [&](int u, int& mu, int& ng) {
    const int per_sb = super_m * n_groups;
    const int sb = u / per_sb, r = u - sb * per_sb;
    const int sb_m = min(super_m, m_units - sb * super_m);
    ng = r / sb_m;
    mu = sb * super_m + (r - ng * sb_m);
  };

  // The producer and MMA warps run their loops with all 32 lanes converged (lane 0 issues): lanes parked at the
  // final cluster barrier while lane 0 loops starve it (measured: 3x slower mainloop with cta_group::2).
  if (warp == kProducerWarp) {
    // ------------------------------------------------------------- TMA producer (one per CTA)
    int st = 0; uint32_t ph = 0;
    const uint64_t pol_a = l2_policy(work.hint_a), pol_b = l2_policy(work.hint_b);
    for (int u = group_id; u < n_units; u += n_cta_groups) {
      int mu, ng; decode(u, mu, ng);
      const int mt = mu * CG + cta_rank;
      // K-lockstep (GemmWork::sync_every): every unit passes tiles_per_unit x spt sync points -- a unit of the last column
      // group with fewer tiles still visits the points of its missing tiles -- so a round's barrier arithmetic is uniform.
      // (sync_every < 0: one sync point per UNIT, at its first k-block -- the cheap form for the short-K statistics GEMMs)
      // (split-bf16 operands: the NSEG K segments of a tile are NSEG passes over the k-blocks, each with its own sync points)
      const bool usync = work.sync_every < 0;
      const bool ksync = work.sync_every > 0;
      const int spt = ksync ? NSEG * ((tr.k_blocks + work.sync_every - 1) / work.sync_every) : 0;   // sync points per tile
      const int round = (u - group_id) / n_cta_groups;
      const long long sync_base = (long long)round * n_cta_groups * (usync ? 1 : tpu * spt);   // every earlier round was full
      const int round_groups = min(n_cta_groups, n_units - round * n_cta_groups);
      auto k_sync = [&](int point) {
        // cumulative-counter barrier over the groups of this round: each leader adds 1 per point on arrival, nobody passes
        // point j before all have arrived at it
        const long long target = sync_base + (long long)(point + 1) * round_groups;
        if (elect_one()) {
          if (leader) atomicAdd(work.sync_ctr, 1u);
          unsigned spins = 0;
          while ((long long)*reinterpret_cast<volatile unsigned*>(work.sync_ctr) < target) {
            __nanosleep(32);
            if (++spins > (1u << 26)) __trap();
          }
        }
        __syncwarp();
      };
      if (usync) k_sync(0);
      const int nt_end = ksync ? (ng + 1) * tpu : min((ng + 1) * tpu, n_tiles);
      for (int nt = ng * tpu; nt < nt_end; ++nt) {
       if (nt >= n_tiles) {                         // (ksync only) a missing tile of the last column group: its sync points
         for (int j = 0; j < spt; ++j) k_sync((nt - ng * tpu) * spt + j);
         continue;
       }
#pragma unroll 1
       for (int seg = 0; seg < NSEG; ++seg) {
        const CUtensorMap* mA = &tmA; const CUtensorMap* mB = &tmB; const CUtensorMap* mA2 = &tmA2; const CUtensorMap* mB2 = &tmB2;
        if constexpr (NSEG > 1) {                     // split-bf16: which part of A and of B this K segment streams
          const int pa = (kSegPartA >> (4 * seg)) & 15, pb = (kSegPartB >> (4 * seg)) & 15;
          if (pa) { mA = &sm->A[pa - 1]; mA2 = &sm->A2[pa - 1]; }
          if (pb) { mB = &sm->B[pb - 1]; mB2 = &sm->B2[pb - 1]; }
        }
        for (int kb = 0; kb < tr.k_blocks; ++kb) {
          if (ksync && kb % work.sync_every == 0) k_sync((nt - ng * tpu) * spt + seg * (spt / NSEG) + kb / work.sync_every);
          mbar_wait(&empty[st], ph ^ 1);
          if (elect_one()) {
            if constexpr (CG == 1) mbar_arrive_expect_tx(&full[st], Cfg::STAGE_BYTES);
            load_operand<Cfg::AMAJ, BM, CG>(sA + st * Cfg::A_BYTES, mA, &full[st], tr.a_row0 + mt * BM, tr.a_k0 + kb * BK, pol_a);
            if constexpr (Cfg::NMMA == 1) {
              load_operand<Cfg::BMAJ, Cfg::BN_CTA, CG>(sB + st * Cfg::B_BYTES, mB, &full[st], nt * BN + cta_rank * Cfg::BN_CTA,
                                                       tr.b_k0 + kb * BK, pol_b);
            } else {
              // 512-column tile: MMA h covers columns [h * 256, h * 256 + 256); this CTA stages its half of each
#pragma unroll
              for (int h = 0; h < Cfg::NMMA; ++h)
                load_operand<Cfg::BMAJ, Cfg::UMMA_N / CG, CG>(sB + st * Cfg::B_BYTES + h * (Cfg::UMMA_N / CG) * 128, mB, &full[st],
                                                              nt * BN + h * Cfg::UMMA_N + cta_rank * (Cfg::UMMA_N / CG), tr.b_k0 + kb * BK, pol_b);
            }
            if constexpr (Cfg::DUAL) {
              load_operand<Cfg::AMAJ, BM, CG>(sA2 + st * Cfg::A_BYTES, mA2, &full[st], tr.a2_row0 + mt * BM, tr.a_k0 + kb * BK);
              load_operand<Cfg::BMAJ, Cfg::BN_CTA, CG>(sB2 + st * Cfg::B_BYTES, mB2, &full[st], nt * BN + cta_rank * Cfg::BN_CTA,
                                                       tr.b_k0 + kb * BK);
            }
            if constexpr (CG == 2) {
              // both CTAs' bytes are accounted on the leader's barrier (the .cta_group::2 loads signal it)
              if (leader) mbar_arrive_expect_tx(&full[st], CG * Cfg::STAGE_BYTES);
              else mbar_arrive_cluster(&full[st], 0);
            }
          }
          __syncwarp();
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
       }
      }
    }
  } else if (warp == kMmaWarp && leader) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(BM * CG, Cfg::UMMA_N, Cfg::AMAJ, Cfg::BMAJ);
    constexpr uint32_t AS = Cfg::ACC_STAGES;
    using DA = OperandDesc<Cfg::AMAJ>;
    using DB = OperandDesc<Cfg::BMAJ>;
    const uint32_t a_base = smem_u32(sA) >> 4, b_base = smem_u32(sB) >> 4;
    const uint32_t a2_base = smem_u32(sA2) >> 4, b2_base = smem_u32(sB2) >> 4;
    int st = 0; uint32_t ph = 0; uint32_t it = 0;
    for (int u = group_id; u < n_units; u += n_cta_groups) {
      int mu, ng; decode(u, mu, ng);
      const int nt_end = min((ng + 1) * tpu, n_tiles);
      for (int nt = ng * tpu; nt < nt_end; ++nt, it += Cfg::PROMOTE ? 0 : 1) {
        uint32_t as = it % AS, aph = (it / AS) & 1;
        if constexpr (!Cfg::PROMOTE) { mbar_wait(&tempty[as], aph ^ 1); tc_fence_after(); }
        uint32_t acc = tmem_base + as * Cfg::ACC_COLS;
        const int kb_total = tr.k_blocks * NSEG;       // the K segments of split operands accumulate into the same tile
        int kc = 0;                                    // PROMOTE: k-block inside the current chunk
        for (int kb = 0; kb < kb_total; ++kb) {
          if constexpr (Cfg::PROMOTE) {
            if (kc == 0) {                             // a new chunk starts from zero in the next accumulator stage
              as = it % AS; aph = (it / AS) & 1;
              mbar_wait(&tempty[as], aph ^ 1);
              tc_fence_after();
              acc = tmem_base + as * Cfg::ACC_COLS;
            }
          }
          const int kacc = Cfg::PROMOTE ? kc : kb;     // 0 = first k-block of this accumulator: overwrite
          mbar_wait(&full[st], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_base + st * (Cfg::A_BYTES >> 4), b_lo = b_base + st * (Cfg::B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#pragma unroll
              for (int h = 0; h < Cfg::NMMA; ++h)
                umma_bf16<CG>(acc + h * Cfg::UMMA_N, DA::HI | (uint64_t)(a_lo + k * DA::KSTEP),
                              DB::HI | (uint64_t)(b_lo + h * Cfg::B_MMA_STEP + k * DB::KSTEP), idesc, (kacc | k) != 0);
            }
            if constexpr (Cfg::DUAL) {
              const uint32_t a2_lo = a2_base + st * (Cfg::A_BYTES >> 4), b2_lo = b2_base + st * (Cfg::B_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16<CG>(acc + BN, DA::HI | (uint64_t)(a2_lo + k * DA::KSTEP), DB::HI | (uint64_t)(b2_lo + k * DB::KSTEP), idesc,
                              (kacc | k) != 0);
            }
            umma_commit<CG>(&empty[st]);              // smem slot reusable (in both CTAs) once these MMAs retire
            if constexpr (Cfg::PROMOTE) {
              if (kc == Cfg::PROMOTE_KB - 1 || kb == kb_total - 1) umma_commit<CG>(&tfull[as]);
            } else {
              if (kb == kb_total - 1) umma_commit<CG>(&tfull[as]);
            }
          }
          __syncwarp();
          if constexpr (Cfg::PROMOTE) {
            if (++kc == Cfg::PROMOTE_KB || kb == kb_total - 1) { kc = 0; ++it; }
          }
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // ------------------------------------------------------------- epilogue (thread == accumulator row)
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int trow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t it = 0;
    for (int u = group_id; u < n_units; u += n_cta_groups) {
      int mu, ng; decode(u, mu, ng);
      const int mt = mu * CG + cta_rank;
      typename Epi::State s;
      Epi::unit_begin(ep, s, tr.e_row0 + mt * BM + trow, mt * BM + trow);
      const int nt_end = min((ng + 1) * tpu, n_tiles);
      if constexpr (Cfg::PROMOTE) {
        const int n_chunks = (tr.k_blocks * NSEG + Cfg::PROMOTE_KB - 1) / Cfg::PROMOTE_KB;
        for (int nt = ng * tpu; nt < nt_end; ++nt) {
          float sum[Cfg::ACC_COLS];
#pragma unroll
          for (int i = 0; i < Cfg::ACC_COLS; ++i) sum[i] = 0.f;
          for (int ch = 0; ch < n_chunks; ++ch, ++it) {
            const uint32_t as = it % Cfg::ACC_STAGES, aph = (it / Cfg::ACC_STAGES) & 1;
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < Cfg::ACC_COLS / 32; ++c) {
              float v[32];
              load_chunk(lane_base + as * Cfg::ACC_COLS + c * 32, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[c * 32 + i] += v[i];          // fp32 round-to-nearest
            }
            tc_fence_before();
            if (leader) mbar_arrive(&tempty[as]); else mbar_arrive_cluster(&tempty[as], 0);
          }
          // park the promoted sum in TMEM (this warp's own lanes) and run the fused epilogue on it unchanged
#pragma unroll
          for (int c = 0; c < Cfg::ACC_COLS / 32; ++c) store_chunk(lane_base + Cfg::MAIN_COL + c * 32, &sum[c * 32]);
          tmem_st_wait();
          Epi::tile(ep, s, lane_base + Cfg::MAIN_COL, nt * BN);
        }
      } else {
      for (int nt = ng * tpu; nt < nt_end; ++nt, ++it) {
        const uint32_t as = it % Cfg::ACC_STAGES, aph = (it / Cfg::ACC_STAGES) & 1;
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        Epi::tile(ep, s, lane_base + as * Cfg::ACC_COLS, nt * BN);
        tc_fence_before();
        if (leader) mbar_arrive(&tempty[as]); else mbar_arrive_cluster(&tempty[as], 0);
      }
      }
      Epi::unit_end(ep, s, ng);
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // remote arrives have landed before either CTA exits
  if (warp == kAllocWarp) tmem_dealloc<CG>(tmem_base, 512);
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(kNumThreads, 1)
dpa_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                const GemmWork work, const typename Epi::Params ep) {
  static_assert(Cfg::NSEG == 1, "split operands go through dpa_gemm_split_kernel");
  dpa_gemm_body<Cfg, Epi>(tmA, tmB, tmA2, tmB2, nullptr, work, ep);
}
// split-bf16 operands (fp32 inputs): the same mainloop over six K segments
template <class Cfg, class Epi>
__global__ void __launch_bounds__(kNumThreads, 1)
dpa_gemm_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                      const __grid_constant__ SplitMaps sm, const GemmWork work, const typename Epi::Params ep) {
  static_assert(Cfg::NSEG == 6, "NSEG");
  dpa_gemm_body<Cfg, Epi>(tmA, tmB, tmA2, tmB2, &sm, work, ep);
}

// =====================================================================================================
// Epilogues. `tile(ep, s, taddr, col0)`: taddr = TMEM address of this thread's lane, accumulator column 0.
// =====================================================================================================
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void load_chunk(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  __syncwarp();                      // tcgen05.ld is warp-collective (.sync.aligned)
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <bool ROUND>
__device__ __forceinline__ void round_chunk(float (&v)[32]) {
  if constexpr (ROUND) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round_f32(v[i]);   // nn.Linear emits bf16 logits (modeling_llama.py:1028)
  }
}
// online (max, sum-exp) update over one 32-column chunk; returns the rescale factor applied to the old sum
__device__ __forceinline__ float online_update(const float (&v)[32], int nvalid, float& m, float& ssum, float (&e)[32]) {
  float cmax = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, i < nvalid ? v[i] : -INFINITY);
  const float m_new = fmaxf(m, cmax);
  const float scale = fast_exp2((m - m_new) * kLog2e);            // m = -inf on first use -> 0
  const float mb = m_new * kLog2e;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    e[i] = i < nvalid ? fast_exp2(fmaf(v[i], kLog2e, -mb)) : 0.f;
    acc += e[i];
  }
  ssum = fmaf(ssum, scale, acc);
  m = m_new;
  return scale;
}

__device__ __forceinline__ void store_bf16_chunk(__nv_bfloat16* dst, const float (&d)[32]);
__device__ __forceinline__ void store_out_chunk(__nv_bfloat16* dst, const float (&v)[32]);
__device__ __forceinline__ void store_out_chunk(float* dst, const float (&v)[32]);

// ---- forward statistics, pair rows: per (row, column group) partial {max, sum-exp, target logit}.
//      Fused-gradient forward: rows below `store_rows` (the pair rows of row chunk 0) also keep their bf16 logits in the
//      chunk's dz buffer, so that dz = c (onehot - p) is an elementwise pass instead of a recompute GEMM.
template <int BN, bool ROUND, class ZT = __nv_bfloat16>      // ZT = float: the fp32 path keeps fp32 logits
struct EpiStatsPair {
  struct Params {
    const int* labels;     // packed row -> target id (global vocab id)
    float4* part;          // [n_groups][row_cap]
    int row_cap, v_local, v_offset;
    ZT* z; int ldz, store_rows;                // z == nullptr: statistics only
  };
  struct State { float m, s, zy; int yloc; int row; ZT* out; };
  __device__ static void unit_begin(const Params& p, State& s, int erow, int) {
    s.m = -INFINITY; s.s = 0.f; s.zy = 0.f; s.row = erow;
    s.yloc = p.labels[erow] - p.v_offset;
    s.out = (p.z && erow < p.store_rows) ? p.z + (size_t)erow * p.ldz : nullptr;
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.v_local) break;
      float v[32], e[32];
      load_chunk(taddr + c * 32, v);
      round_chunk<ROUND>(v);
      if (s.out) store_out_chunk(s.out + cbase, v);
      online_update(v, p.v_local - cbase, s.m, s.s, e);
      if ((unsigned)(s.yloc - cbase) < 32u) {
        const int j = s.yloc - cbase;
#pragma unroll
        for (int i = 0; i < 32; ++i) s.zy = (i == j) ? v[i] : s.zy;
      }
    }
  }
  __device__ static void unit_end(const Params& p, State& s, int ng) {
    p.part[(size_t)ng * p.row_cap + s.row] = make_float4(s.m, s.s, s.zy, 0.f);
  }
};

// ---- forward statistics, reference-sample rows (DUAL): policy logits in columns [0,BN), frozen in [BN,2BN)
//      partial {m_p, s_p, m_r, s_r} and u = sum_v e^{z_r - m_r} (z_r - z_p)
template <int BN, bool ROUND>
struct EpiStatsRef {
  struct Params {
    float4* part;          // [n_groups][row_cap]
    float* part_u;         // [n_groups][row_cap]
    int row_cap, v_local;
  };
  struct State { float mp, sp, mr, sr, u; int row; };
  __device__ static void unit_begin(const Params&, State& s, int erow, int) {
    s.mp = -INFINITY; s.sp = 0.f; s.mr = -INFINITY; s.sr = 0.f; s.u = 0.f; s.row = erow;
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.v_local) break;
      const int nvalid = p.v_local - cbase;
      float vp[32], vr[32], e[32];
      load_chunk(taddr + c * 32, vp);
      round_chunk<ROUND>(vp);
      online_update(vp, nvalid, s.mp, s.sp, e);
      load_chunk(taddr + BN + c * 32, vr);
      round_chunk<ROUND>(vr);
      const float scale = online_update(vr, nvalid, s.mr, s.sr, e);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(e[i], vr[i] - vp[i], acc);   // e[i] == 0 beyond nvalid
      s.u = fmaf(s.u, scale, acc);
    }
  }
  __device__ static void unit_end(const Params& p, State& s, int ng) {
    p.part[(size_t)ng * p.row_cap + s.row] = make_float4(s.mp, s.sp, s.mr, s.sr);
    p.part_u[(size_t)ng * p.row_cap + s.row] = s.u;
  }
};

__device__ __forceinline__ void store_bf16_chunk(__nv_bfloat16* dst, const float (&d)[32]) {
  uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    o[i] = make_uint4(pack_bf16x2(d[8 * i + 0], d[8 * i + 1]), pack_bf16x2(d[8 * i + 2], d[8 * i + 3]),
                      pack_bf16x2(d[8 * i + 4], d[8 * i + 5]), pack_bf16x2(d[8 * i + 6], d[8 * i + 7]));
}

// split-bf16 store of a dLogits chunk for the fp32 path: hi = bf16(d), mid = bf16(d - hi), lo = bf16(d - hi - mid);
// the three parts are `part_stride` elements apart
__device__ __forceinline__ void store_split_chunk(__nv_bfloat16* hi, size_t part_stride, const float (&d)[32]) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = d[i] - bf16_round_f32(d[i]);
  store_bf16_chunk(hi, d);
  store_bf16_chunk(hi + part_stride, r);
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] -= bf16_round_f32(r[i]);
  store_bf16_chunk(hi + 2 * part_stride, r);
}

// ---- fused-gradient forward, reference-sample rows: per (row, column group) partial {max, sum-exp} of one model's
//      logits, and the bf16-rounded logits of the row chunk are kept (policy -> the chunk's dz buffer, frozen -> zr) so
//      that dz = c (p - q) and the KL row are one elementwise pass once the chunk's log-sum-exps are known, instead of a
//      second pair of GEMMs.  Only valid in reference-bf16 numerics (the logits ARE bf16 there,
//      modeling_llama.py:1028), hence always rounded.
template <int BN, class ZT = __nv_bfloat16>      // ZT = float (fp32 inputs): the kept logits are the promoted fp32 sums, unrounded
struct EpiStatsStore {
  struct Params {
    const DpaMeta* meta;
    float2* part;            // [n_groups][row_cap][2]: {m, s} of the policy (which = 0) / frozen (which = 1) logits
    ZT* z;                   // [chunk_rows][ldz]
    int row_cap, v_local, ldz, chunk_row0, which;
  };
  struct State { float m, s; int row; ZT* out; };
  __device__ static void unit_begin(const Params& p, State& s, int erow, int) {
    s.m = -INFINITY; s.s = 0.f;
    s.row = erow - p.meta->p0;                                   // reference-sample row id
    s.out = p.z + (size_t)(erow - p.chunk_row0) * p.ldz;
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.v_local) break;
      float v[32], e[32];
      load_chunk(taddr + c * 32, v);
      round_chunk<sizeof(ZT) == 2>(v);
      store_out_chunk(s.out + cbase, v);                         // a partial last chunk still stores 32 columns (ldz % 64 == 0)
      online_update(v, p.v_local - cbase, s.m, s.s, e);
    }
  }
  __device__ static void unit_end(const Params& p, State& s, int ng) {
    p.part[((size_t)ng * p.row_cap + s.row) * 2 + p.which] = make_float2(s.m, s.s);
  }
};

// ---- backward recompute, pair rows: dz = coef * (onehot(y) - softmax(z))  -> bf16 scratch [chunk_rows][ldz]
template <int BN, bool ROUND, bool SPLIT = false>
struct EpiDzPair {
  struct Params {
    const int* labels; const float* lse; const float* coef;   // indexed by packed row
    __nv_bfloat16* dz; int ldz; int chunk_row0; int v_local, v_offset;
    size_t part_stride;                                       // SPLIT: elements between the hi / mid / lo copies of dz
  };
  struct State { float nl, c; int yloc; __nv_bfloat16* out; };
  __device__ static void unit_begin(const Params& p, State& s, int erow, int) {
    s.nl = -p.lse[erow] * kLog2e; s.c = p.coef[erow];
    s.yloc = p.labels[erow] - p.v_offset;
    s.out = p.dz + (size_t)(erow - p.chunk_row0) * p.ldz;
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.ldz) break;
      float v[32], d[32];
      load_chunk(taddr + c * 32, v);
      round_chunk<ROUND>(v);
#pragma unroll
      for (int i = 0; i < 32; ++i) d[i] = -s.c * fast_exp2(fmaf(v[i], kLog2e, s.nl));
      if ((unsigned)(s.yloc - cbase) < 32u) {
        const int j = s.yloc - cbase;
#pragma unroll
        for (int i = 0; i < 32; ++i) d[i] += (i == j) ? s.c : 0.f;
      }
      if constexpr (SPLIT) store_split_chunk(s.out + cbase, p.part_stride, d);
      else store_bf16_chunk(s.out + cbase, d);
    }
  }
  __device__ static void unit_end(const Params&, State&, int) {}
};

// ---- backward recompute, reference-sample rows (DUAL): dz = coef * (softmax(z_pol) - softmax(z_frozen))
template <int BN, bool ROUND, bool SPLIT = false>
struct EpiDzRef {
  struct Params {
    const DpaMeta* meta;
    const float* lse_p; const float* lse_r; const float* coef;   // indexed by reference row j
    __nv_bfloat16* dz; int ldz; int chunk_row0; int v_local;
    size_t part_stride;                                          // SPLIT: elements between the hi / mid / lo copies of dz
  };
  struct State { float nlp, nlr, c; __nv_bfloat16* out; };
  __device__ static void unit_begin(const Params& p, State& s, int erow, int) {
    // erow is the packed policy row; the reference-sample arrays are indexed from the segment start p0
    const int j = erow - p.meta->p0;
    s.out = p.dz + (size_t)(erow - p.chunk_row0) * p.ldz;
    s.nlp = -p.lse_p[j] * kLog2e; s.nlr = -p.lse_r[j] * kLog2e; s.c = p.coef[j];
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.ldz) break;
      float vp[32], vr[32], d[32];
      load_chunk(taddr + c * 32, vp);
      round_chunk<ROUND>(vp);
      load_chunk(taddr + BN + c * 32, vr);
      round_chunk<ROUND>(vr);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        d[i] = s.c * (fast_exp2(fmaf(vp[i], kLog2e, s.nlp)) - fast_exp2(fmaf(vr[i], kLog2e, s.nlr)));
      if constexpr (SPLIT) store_split_chunk(s.out + cbase, p.part_stride, d);
      else store_bf16_chunk(s.out + cbase, d);
    }
  }
  __device__ static void unit_end(const Params&, State&, int) {}
};

// ---- dH = dZ * W: fp32 accumulators -> bf16 rows scattered to their source positions
__device__ __forceinline__ void store_out_chunk(__nv_bfloat16* dst, const float (&v)[32]) { store_bf16_chunk(dst, v); }
__device__ __forceinline__ void store_out_chunk(float* dst, const float (&v)[32]) {
  float4* o = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <class OutT> struct StoreRowsParams {
  const DpaMeta* meta;
  const int* src_pair; const int* src_ref;       // packed row -> flat (sequence*T + t) row of the dense tensors
  OutT* dH_pair; OutT* dH_ref; int H;
  OutT* packed;                                  // non-null: keep the rows packed (vocab-parallel: the partial sums of the
  int packed_row0;                               // chunk, first packed row packed_row0, are all-reduced in this compact form)
};
template <int BN, class OutT = __nv_bfloat16>      // OutT = float: the fp32 path (split-bf16 operands)
struct EpiStoreRows {
  using Params = StoreRowsParams<OutT>;            // (independent of the tile width: the 256- and 512-column kernels share it)
  struct State { OutT* out; };
  __device__ static void unit_begin(const Params& p, State& s, int erow, int) {
    const int p0 = p.meta->p0;
    if (p.packed) { s.out = p.packed + (size_t)(erow - p.packed_row0) * p.H; return; }
    s.out = nullptr;
    if (erow < p0) { if (erow < p.meta->n_pair) s.out = p.dH_pair + (size_t)p.src_pair[erow] * p.H; }
    else { const int j = erow - p0; if (j < p.meta->n_ref) s.out = p.dH_ref + (size_t)p.src_ref[j] * p.H; }
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.H) break;
      float v[32];
      load_chunk(taddr + c * 32, v);                     // warp-collective: every lane loads
      if (s.out) store_out_chunk(s.out + cbase, v);      // H % 32 == 0 is checked on the host
    }
  }
  __device__ static void unit_end(const Params&, State&, int) {}
};

template <int BN> using EpiStoreRowsBf16 = EpiStoreRows<BN, __nv_bfloat16>;

// ---- dW += dZ^T * H: fp32 accumulate into the [V][ld] fp32 buffer (thread == vocabulary row)
template <int BN>
struct EpiAccumF32 {
  struct Params { float* out; int ld; int m_extent; int n_extent; int accumulate; };
  struct State { float* out; };
  __device__ static void unit_begin(const Params& p, State& s, int, int trow_global) {
    s.out = trow_global < p.m_extent ? p.out + (size_t)trow_global * p.ld : nullptr;
  }
  // (the chunked bf16 dW accumulation of the DPA path uses EpiAccumDW below; this one serves the fp32 path, whose fp32 dW
  //  is its own accumulator, and the debug GEMM)
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.n_extent) break;
      float v[32];
      load_chunk(taddr + c * 32, v);
      if (s.out) {
        float4* o = reinterpret_cast<float4*>(s.out + cbase);
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 t = o[i];
            t.x += v[4 * i]; t.y += v[4 * i + 1]; t.z += v[4 * i + 2]; t.w += v[4 * i + 3];
            o[i] = t;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  __device__ static void unit_end(const Params&, State&, int) {}
};

// ---- dW += dZ^T * H over the row chunks: the first chunk overwrites the fp32 accumulator (no memset), middle chunks add,
//      the last chunk that has rows adds and writes bf16 straight to dW (no conversion pass).  Which chunk is first / last
//      is decided on the device from DpaMeta (the host does not know the active row count).
struct AccumDWParams {
  const DpaMeta* meta; int chunk, chunk_rows;
  float* acc32; __nv_bfloat16* out; int ld; int m_extent; int n_extent;
};
template <int BN>
struct EpiAccumDW {
  using Params = AccumDWParams;
  struct State { float* acc; __nv_bfloat16* out; int first, last; };
  __device__ static void unit_begin(const Params& p, State& s, int, int trow_global) {
    const bool ok = trow_global < p.m_extent;
    s.acc = ok ? p.acc32 + (size_t)trow_global * p.ld : nullptr;
    s.out = ok ? p.out + (size_t)trow_global * p.ld : nullptr;
    s.first = p.chunk == 0;
    s.last = (long long)(p.chunk + 1) * p.chunk_rows >= p.meta->n_tot;
  }
  __device__ static void tile(const Params& p, State& s, uint32_t taddr, int col0) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int cbase = col0 + c * 32;
      if (cbase >= p.n_extent) break;
      float v[32];
      load_chunk(taddr + c * 32, v);
      if (!s.acc) continue;
      float4* a = reinterpret_cast<float4*>(s.acc + cbase);
      if (!s.first) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = a[i];
          v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
        }
      }
      if (s.last) {
        store_bf16_chunk(s.out + cbase, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
  __device__ static void unit_end(const Params&, State&, int) {}
};

}  // namespace dpa
