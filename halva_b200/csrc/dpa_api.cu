// C ABI of libhalva_dpa.so (include/halva_dpa.h): workspace layout, TMA descriptors, launch sequences.
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <cstring>
#include <cstdarg>
#include <mutex>
#include <atomic>
#include <unordered_map>
#include <algorithm>
#include <vector>

#include "dpa_kernels.cuh"
#include "dpa_peer.cuh"

using namespace dpa;

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(DPA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)
#define LAUNCH_CHECK(what) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(DPA_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e_)); } while (0)

// ---------------------------------------------------------------------------------- optional per-kernel timing
// bench.py asks for live CUDA-event timings of the individual kernels (they are launched inside this library,
// so the caller cannot bracket them). Off by default; when on, every launch group is bracketed by two events.
enum ProfKind { PK_META = 0, PK_GATHER, PK_FWD_PAIR, PK_FWD_REF, PK_REDUCE, PK_COEF, PK_DZ_PAIR, PK_DZ_REF, PK_DH, PK_DW, PK_MISC,
                PK_STASH_REF, PK_DZ_ELEM, PK_SCALE, PK_STASH_GIVEN, PK_COUNT };
struct ProfSpan { int kind; int launches; cudaEvent_t a, b; };
std::vector<ProfSpan> g_spans;
std::vector<cudaEvent_t> g_event_pool;      // timing events are recycled by dpa_profile_collect, never leaked
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
cudaEvent_t pooled_event() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  cudaStream_t st; bool on; ProfSpan sp;
  ProfScope(int kind, int launches, cudaStream_t s) : st(s), on(g_prof_on.load(std::memory_order_relaxed)) {
    if (!on) return;
    sp.kind = kind; sp.launches = launches;
    sp.a = pooled_event(); sp.b = pooled_event();
    cudaEventRecord(sp.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(sp.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_spans.push_back(sp);
  }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int round_up_i(int x, int a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------- device / driver entry
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct DeviceInfo { int sms = 0; int cc_major = 0; EncodeTiledFn encode = nullptr; bool ok = false; char why[128] = ""; };
DeviceInfo g_dev[16];
std::mutex g_mu;

int device_info(DeviceInfo** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(DPA_ERR_CUDA, "device index %d unsupported", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.ok) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (d.cc_major != 10) return fail(DPA_ERR_ARCH, "libhalva_dpa needs an sm_100 GPU (found compute capability %d.x)", d.cc_major);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(DPA_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    d.encode = reinterpret_cast<EncodeTiledFn>(fn);
    d.ok = true;
  }
  *out = &d;
  return DPA_OK;
}

// Encoded tensor maps are cached per (device, pointer, dims, stride, box): a trainer calls with the same buffers every
// step (the caching allocator hands the same blocks back), so steady-state calls encode nothing.
struct TmapKey {
  const void* ptr; uint64_t inner, outer, stride; uint32_t bi, bo; int dev;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && bi == o.bi && bo == o.bo && dev == o.dev;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner + 0x632BE59BD9B4E019ull) * 0xFF51AFD7ED558CCDull; h ^= h >> 29;
    h ^= (k.outer * 0xC4CEB9FE1A85EC53ull) ^ (k.stride << 7) ^ ((uint64_t)k.bi << 40) ^ ((uint64_t)k.bo << 20) ^ (uint64_t)k.dev;
    return (size_t)(h ^ (h >> 32));
  }
};
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
std::mutex g_tmap_mu;
std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};
constexpr size_t kTmapCacheMax = 4096;

int make_tmap_uncached(const DeviceInfo* d, CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer);
// 2-D bf16 tensor map, 128-byte swizzle. inner = contiguous dimension.
int make_tmap(const DeviceInfo* d, CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
              uint32_t box_inner, uint32_t box_outer) {
  int dev = 0; cudaGetDevice(&dev);
  const TmapKey key{ptr, inner, outer, row_stride_bytes, box_inner, box_outer, dev};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *tm = it->second; g_tmap_hits.fetch_add(1, std::memory_order_relaxed); return DPA_OK; }
  }
  int rc = make_tmap_uncached(d, tm, ptr, inner, outer, row_stride_bytes, box_inner, box_outer);
  if (rc) return rc;
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *tm);
  return DPA_OK;
}
int make_tmap_uncached(const DeviceInfo* d, CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = d->encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DPA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r, ptr,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner, box_outer);
  return DPA_OK;
}
// operand tile maps: K-major = [rows][K] row-major, box {64 k, rows}; MN-major = [K][mn] row-major, box {64 mn, 64 k}
int make_operand_tmap(const DeviceInfo* d, CUtensorMap* tm, const void* ptr, int major, uint64_t mn_extent, uint64_t k_extent,
                      uint64_t ld_elems, uint32_t tile_rows) {
  if (major == KMAJOR) return make_tmap(d, tm, ptr, k_extent, mn_extent, ld_elems * 2, BK, tile_rows);
  return make_tmap(d, tm, ptr, mn_extent, k_extent, ld_elems * 2, 64, BK);
}

template <class Cfg, class Epi> constexpr auto gemm_kernel_of() {
  if constexpr (Cfg::NSEG == 1) return dpa_gemm_kernel<Cfg, Epi>; else return dpa_gemm_split_kernel<Cfg, Epi>;
}
// sm: tensor maps of the mid / lo operand parts, only for Cfg::NSEG == 6 (split-bf16 operands of the fp32 path)
template <class Cfg, class Epi>
int launch_gemm(cudaStream_t st, int sms, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                const CUtensorMap& tmB2, const GemmWork& work, const typename Epi::Params& ep, const char* what,
                const SplitMaps* sm = nullptr) {
  auto kern = gemm_kernel_of<Cfg, Epi>();
  static std::once_flag once[16];
  int dev = 0; cudaGetDevice(&dev);
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[dev & 15], [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return fail(DPA_ERR_CUDA, "cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(attr_err));
  // persistent: one CTA per SM; cta_group::2 kernels run as clusters of two CTAs on an SM pair.  The units are
  // statically strided over the grid, so every cluster must be co-resident: size the grid by the occupancy query.
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(sms / Cfg::CG * Cfg::CG));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg::CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = Cfg::CG > 1 ? 1 : 0;
  if constexpr (Cfg::CG > 1) {
    static int max_clusters[16] = {0};
    if (max_clusters[dev & 15] == 0) {
      int n = 0;
      cudaError_t oe = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (oe != cudaSuccess || n < 1) return fail(DPA_ERR_CUDA, "cudaOccupancyMaxActiveClusters(%s): %s (n=%d)", what, cudaGetErrorString(oe), n);
      max_clusters[dev & 15] = n;
      if (getenv("DPA_VERBOSE")) fprintf(stderr, "dpa: %s: %d co-resident clusters of %d (SMs %d)\n", what, n, Cfg::CG, sms);
    }
    cfg.gridDim = dim3((unsigned)(std::min(sms / Cfg::CG, max_clusters[dev & 15]) * Cfg::CG));
  }
  cudaError_t e;
  if constexpr (Cfg::NSEG == 1) e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA2, tmB2, work, ep);
  else e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA2, tmB2, *sm, work, ep);
  if (e != cudaSuccess) return fail(DPA_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
  return DPA_OK;
}

// L2 super-block height (row-tile units that walk the column groups together) per GEMM class; 0 = the kernel's default.
// The statistics GEMMs (K = H, short) want every row unit of the launch in one super-block -- the packed rows of a chunk
// (<= 134 MB) then stay L2-resident while the lm_head streams through once; the gradient GEMMs (K = V or rows, long)
// want a square-ish wave.  Overridable for measurements: DPA_SUPER_STATS / DPA_SUPER_DH / DPA_SUPER_DW.
struct SuperM { int stats, dh, dw; };
int env_int(const char* n, int dflt) { const char* e = getenv(n); return e ? atoi(e) : dflt; }
const SuperM& super_m() {
  static const SuperM v = SuperM{env_int("DPA_SUPER_STATS", 0), env_int("DPA_SUPER_DH", 0), env_int("DPA_SUPER_DW", 0)};
  return v;
}
// Tile width of the two long-K gradient GEMMs (dH = dz W, dW = dz^T H): 512 (default) = one 256 x 512 tile per CTA pair in a
// single all-of-TMEM accumulator, 256 = the double-buffered 256 x 256 tile of the statistics GEMMs.  DPA_GRAD_BN overrides.
int grad_bn() {
  static const int v = env_int("DPA_GRAD_BN", 512) == 256 ? 256 : 512;
  return v;
}
// L2 eviction-priority hints of the operand loads per GEMM class, "ab" digits (0 none, 1 evict_first, 2 evict_last):
// DPA_HINT_STATS / DPA_HINT_DH / DPA_HINT_DW, e.g. DPA_HINT_STATS=21 keeps the packed hidden rows and streams the lm_head.
struct Hints { int a, b; };
Hints hints_of(const char* name, int dflt) {
  const int v = env_int(name, dflt);
  return Hints{(v / 10) % 10, v % 10};
}
const Hints& hint_stats() { static const Hints h = hints_of("DPA_HINT_STATS", 0); return h; }
const Hints& hint_dh() { static const Hints h = hints_of("DPA_HINT_DH", 0); return h; }
const Hints& hint_dw() { static const Hints h = hints_of("DPA_HINT_DW", 0); return h; }
// K-lockstep of the CTA groups (GemmWork::sync_every; > 0: every so many k-blocks, -1: once per unit, 0: off).  DPA_KSYNC:
// the long-K gradient GEMMs (32 k-blocks); DPA_KSYNC_STATS: the statistics GEMMs (K = H is only 64 k-blocks per tile: a
// barrier per tile costs more than it saves, one per five-tile unit pays -- profiles/r02e_klockstep.md).
int ksync() { static const int v = env_int("DPA_KSYNC", 32); return v; }
// (units of a single tile -- small vocabulary shards, see tiles_per_unit() -- would pay that barrier per tile: off there)
int ksync_stats(int tpu) { static const int v = env_int("DPA_KSYNC_STATS", -2); return v != -2 ? v : (tpu >= 4 ? -1 : 0); }
int ksync_split() { static const int v = env_int("DPA_KSYNC_SPLIT", 0); return v; }     // fp32 inputs: the split gradient GEMMs
// cta_group of the single-accumulator GEMMs: 2 = CTA pairs on 256 x 256 tiles (default), 1 = one CTA per 128 x 256 tile.
// Both are compiled; dpa_set_cta_group() / DPA_CTA_GROUP pick one (A/B measurements, bring-up).
std::atomic<int> g_cta_group{0};
int cta_group() {
  int v = g_cta_group.load(std::memory_order_relaxed);
  if (v == 0) {
    const char* e = getenv("DPA_CTA_GROUP");
    v = (e && e[0] == '1') ? 1 : 2;
    g_cta_group.store(v, std::memory_order_relaxed);
  }
  return v;
}
// tmB1 / tmB2: the B-operand maps whose box holds BN (cta_group 1) or BN / 2 (cta_group 2) rows; equal for MN-major B
template <class Epi, int AMAJ, int BMAJ>
int launch_single(cudaStream_t st, const DeviceInfo* d, const CUtensorMap& tmA, const CUtensorMap& tmB1, const CUtensorMap& tmB2,
                  const GemmWork& work, const typename Epi::Params& ep, const char* what) {
  if (cta_group() == 2)
    return launch_gemm<GemmCfg<256, false, AMAJ, BMAJ, 2>, Epi>(st, d->sms, tmA, tmB2, tmA, tmB2, work, ep, what);
  return launch_gemm<GemmCfg<256, false, AMAJ, BMAJ, 1>, Epi>(st, d->sms, tmA, tmB1, tmA, tmB1, work, ep, what);
}

// the long-K gradient GEMMs on 256 x 512 tiles (CTA pairs only; B is MN-major in both, so the 64-column TMA boxes serve)
template <template <int> class EpiT, int AMAJ, int BMAJ, class Params>
int launch_grad(cudaStream_t st, const DeviceInfo* d, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmWork& work,
                const Params& ep, const char* what, bool wide = true) {
  if (wide && cta_group() == 2 && grad_bn() == 512)
    return launch_gemm<GemmCfg<512, false, AMAJ, BMAJ, 2>, EpiT<512>>(st, d->sms, tmA, tmB, tmA, tmB, work, ep, what);
  return launch_single<EpiT<256>, AMAJ, BMAJ>(st, d, tmA, tmB, tmB, work, ep, what);
}

// dual-accumulator GEMMs (policy + frozen logits of the same rows): 128-wide tiles; tmB*_1 boxes hold 128 rows of W
// (cta_group 1), tmB*_2 boxes 64 rows (cta_group 2: a CTA pair shares a 256 x 128 x 2 tile)
template <class Epi>
int launch_dual(cudaStream_t st, const DeviceInfo* d, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB_1,
                const CUtensorMap& tmB2_1, const CUtensorMap& tmB_2, const CUtensorMap& tmB2_2, const GemmWork& work,
                const typename Epi::Params& ep, const char* what) {
  if (cta_group() == 2)
    return launch_gemm<GemmCfg<128, true, KMAJOR, KMAJOR, 2>, Epi>(st, d->sms, tmA, tmB_2, tmA2, tmB2_2, work, ep, what);
  return launch_gemm<GemmCfg<128, true, KMAJOR, KMAJOR, 1>, Epi>(st, d->sms, tmA, tmB_1, tmA2, tmB2_1, work, ep, what);
}

// ---------------------------------------------------------------------------------- workspace layout
// Column tiles per work unit of the statistics GEMMs.  A unit = (256-row tile, `tpu` consecutive 256-wide column tiles) and
// leaves one partial-statistics record per row; the units are dealt to the CTA pairs round-robin.  Full vocabulary
// (V = 32000: 125 tiles): 5 tiles per unit -> 25 column groups, small partial arrays, 24 even rounds.  A vocabulary SHARD
// has few tiles (V / 8 = 4000: 16): groups of 5 would be 5 + 5 + 5 + 1 tiles and the pair with the most work would run 20
// tiles against an average of 15.4 (cfg-3 x 8) -- single-tile units deal the same work out to within one tile.
// The 128-wide (dual) and the split-bf16 tilings use the same column groups (shared partials): 2x / 4x as many tiles.
constexpr int kBnPair = 256, kBnRef = 128;
int tiles_per_unit(int v_local) {
  const int forced = env_int("DPA_TPU", 0);             // measurements / tests; read per call so that a test can switch it
  if (forced > 0) return std::min(forced, 64);
  return (v_local + kBnPair - 1) / kBnPair <= 40 ? 1 : 5;
}

// Rows per chunk of the packed row space (kept logits / dLogits scratch = chunk_rows x V x 2 B, twice in the fused-gradient
// forward).  32768 holds every contributing row of a cfg-2 batch (24.8 K at 10 % phrase density) in ONE chunk: each GEMM class
// then runs once per step, dW needs no fp32 accumulation pass, and 4 of 6 launch groups of the 16384-row chunking disappear.
constexpr int kDefaultChunkRows = 32768;

constexpr int kSyncCtrs(int n_chunks) { return 8 * n_chunks + 8; }     // upper bound of GEMM launches per call

struct Layout {
  int n_pos_pair, n_pos_ref, cap_pair, cap_ref, nblk_pair, nblk_ref, ng_pair, ng_ref, ldz, chunk_rows, n_chunks;
  int tpu;          // 256-wide column tiles per unit of the statistics GEMMs (tiles_per_unit)
  int n_slots;      // fused-gradient forward: row chunks whose pair-row logits are KEPT (slot c = chunk c); later chunks reuse slot 0
  // state offsets
  size_t meta, bitmap, sync_ctr, block_counts, src_pair, lab_pair, col_pair, src_ref, act_ref, stats_pair, stats_ref, stats_u,
      lse_pair, logp_pair, lse_ref_p, lse_ref_r, kl_row, PN, sigma, coef_pair, coef_ref, rinv, dnw32, formed_for, state_total;
  size_t row_stats_bytes;
  // forward scratch offsets
  size_t f_Hp, f_Hf, f_part_pair, f_part_ref, f_part_u, f_gather, fwd_total;
  // backward scratch offsets
  size_t b_Hp, b_Hf, b_dz, b_dW32, bwd_total;
  // fp32 inputs (split-bf16 operands): Hp / Hf / dz hold three parts (hi, mid, lo) one after the other, and the lm_head
  // weights are split into scratch as well (three parts each)
  struct Split { size_t W3, Wr3; } f_split, b_split, u_split;
  size_t u_zp32;    // fp32 inputs, fused-gradient forward: kept fp32 policy logits of the chunk (u_zr then holds fp32 too)
  bool split;
  // fused-gradient forward scratch offsets (DPA_FLAG_FUSED_GRAD)
  size_t u_Hp, u_Hf, u_part_pair, u_part_ref, u_part_u, u_dz, u_zr, u_dW32, fused_total;
};

int g_layout_comm_world = 1;   // set by dpa_comm_init / dpa_comm_destroy (declared before the communicator itself)
int make_layout(const dpa_shape_t& s, Layout& L) {
  const int comm_world = g_layout_comm_world;
  if (s.B < 1 || s.T < 2 || s.B_ref < 1 || s.T_ref < 2) return fail(DPA_ERR_BAD_SHAPE, "need B>=1, T>=2, B_ref>=1, T_ref>=2");
  if (s.H < 64 || s.H % 64) return fail(DPA_ERR_BAD_SHAPE, "H=%d must be a positive multiple of 64", s.H);
  if (s.V_local < 1) return fail(DPA_ERR_BAD_SHAPE, "V_local=%d must be positive", s.V_local);
  if (s.v_offset < 0 || s.v_offset + s.V_local > s.V) return fail(DPA_ERR_BAD_SHAPE, "vocab shard [%d,%d) outside V=%d", s.v_offset, s.v_offset + s.V_local, s.V);
  if (s.dtype != DPA_BF16 && s.dtype != DPA_F32) return fail(DPA_ERR_DTYPE, "hidden states / weights must be bf16 or fp32");
  L.split = s.dtype == DPA_F32;
  if (L.split && (s.flags & (DPA_FLAG_ROUND_LOGITS | DPA_FLAG_FUSED_NORM)))
    return fail(DPA_ERR_ARG, "fp32 inputs: DPA_FLAG_ROUND_LOGITS / FUSED_NORM describe the bf16 execution and are not available");
  if (L.split && comm_world > 1 && s.V_local < s.V) return fail(DPA_ERR_ARG, "fp32 inputs are not available with the library-side vocab-parallel exchange");
  if (s.index_dtype != DPA_I64 && s.index_dtype != DPA_I32) return fail(DPA_ERR_DTYPE, "labels/signs must be int64 or int32");
  if ((long long)2 * s.B * s.T >= (1ll << 31) / 2 || 2 * s.B >= 32768) return fail(DPA_ERR_BAD_SHAPE, "batch too large for 32-bit row ids");
  L.n_pos_pair = 2 * s.B * (s.T - 1);
  L.n_pos_ref = s.B_ref * (s.T_ref - 1);
  L.cap_pair = round_up_i(L.n_pos_pair, kRowPad);
  L.cap_ref = round_up_i(L.n_pos_ref, kRowPad);
  L.nblk_pair = (L.n_pos_pair + kScanBlock - 1) / kScanBlock;
  L.nblk_ref = (L.n_pos_ref + kScanBlock - 1) / kScanBlock;
  L.tpu = tiles_per_unit(s.V_local);
  L.ng_pair = ((s.V_local + kBnPair - 1) / kBnPair + L.tpu - 1) / L.tpu;
  L.ng_ref = ((s.V_local + kBnRef - 1) / kBnRef + 2 * L.tpu - 1) / (2 * L.tpu);      // == ng_pair: the same column groups
  L.ldz = round_up_i(s.V_local, 64);
  // (fp32 inputs keep 14 B per logit of a chunk -- two fp32 logit buffers and a split-bf16 dLogits triple -- so their default chunk is smaller)
  L.chunk_rows = s.chunk_rows > 0 ? s.chunk_rows : (L.split ? kDefaultChunkRows / 4 : kDefaultChunkRows);
  // a batch of more than four default chunks of positions (cfg-4: 196 K) takes chunks twice as large: its contributing rows
  // (59 K at 10 % phrase density) then fit ONE chunk instead of two -- no fp32 dW accumulation pass (cfg-4: 58.4 -> 57.7 ms)
  if (s.chunk_rows <= 0 && !L.split && L.cap_pair + L.cap_ref > 4 * kDefaultChunkRows) L.chunk_rows *= 2;
  // a vocabulary shard keeps V_local / V of the logits per row: the same scratch holds proportionally more rows per chunk
  // (8 shards of cfg-3: every row in one chunk -- one statistics exchange per segment, no empty chunk launches)
  // (with the library's communicator the factor is the world size: every rank must cut the rows into the SAME chunks even
  //  when the shards are uneven -- the chunks are the units of the exchanges)
  if (s.chunk_rows <= 0 && s.V_local < s.V)
    L.chunk_rows = (int)std::min<long long>((long long)L.chunk_rows * (comm_world > 1 ? comm_world : s.V / s.V_local), 1 << 20);
  if (L.chunk_rows % kRowPad) return fail(DPA_ERR_BAD_SHAPE, "chunk_rows=%d must be a multiple of %d", L.chunk_rows, kRowPad);
  const int cap_tot = L.cap_pair + L.cap_ref;
  L.chunk_rows = std::min(L.chunk_rows, cap_tot);
  L.n_chunks = (cap_tot + L.chunk_rows - 1) / L.chunk_rows;

  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.meta = take(sizeof(DpaMeta));
  L.bitmap = take(16 * 4);
  L.sync_ctr = take((size_t)kSyncCtrs(L.n_chunks) * 4);   // K-lockstep counters, one per GEMM launch of a call (zeroed with DpaMeta)
  L.block_counts = take((size_t)(L.nblk_pair + L.nblk_ref) * 4);
  L.src_pair = take((size_t)L.cap_pair * 4); L.lab_pair = take((size_t)L.cap_pair * 4); L.col_pair = take((size_t)L.cap_pair * 4);
  L.src_ref = take((size_t)L.cap_ref * 4); L.act_ref = take((size_t)L.cap_ref * 4);
  // the three statistics arrays are contiguous: they are the vocab-parallel exchange record
  L.stats_pair = o; o += (size_t)L.cap_pair * 16;
  L.stats_ref = o; o += (size_t)L.cap_ref * 16;
  L.stats_u = o; o += (size_t)L.cap_ref * 4;
  L.row_stats_bytes = o - L.stats_pair;
  o = align_up(o, 256);
  L.lse_pair = take((size_t)L.cap_pair * 4); L.logp_pair = take((size_t)L.cap_pair * 4);
  L.lse_ref_p = take((size_t)L.cap_ref * 4); L.lse_ref_r = take((size_t)L.cap_ref * 4); L.kl_row = take((size_t)L.cap_ref * 4);
  L.PN = take((size_t)2 * s.B * (kMaxSign + 1) * 4);
  L.sigma = take((size_t)s.B * (kMaxSign + 1) * 4);
  L.coef_pair = take((size_t)L.cap_pair * 4); L.coef_ref = take((size_t)L.cap_ref * 4);
  L.rinv = take((size_t)cap_tot * 4);             // DPA_FLAG_FUSED_NORM: rsqrt(var + eps) of the packed policy rows
  L.dnw32 = take((size_t)s.H * 4);                //                      fp32 accumulator of d_norm_w
  L.formed_for = take(4);                         // DPA_FLAG_FUSED_GRAD: the upstream gradient the forward formed the gradients for
  L.state_total = o;

  const size_t parts = L.split ? 3 : 1;
  o = 0;
  L.f_Hp = take(parts * cap_tot * s.H * 2);
  L.f_Hf = take(parts * L.cap_ref * s.H * 2);
  L.f_part_pair = take((size_t)L.ng_pair * L.cap_pair * 16);
  L.f_part_ref = take((size_t)L.ng_ref * L.cap_ref * 16);
  L.f_part_u = take((size_t)L.ng_ref * L.cap_ref * 4);
  auto take_split = [&](Layout::Split& sp) {
    sp = Layout::Split{};
    if (!L.split) return;
    sp.W3 = take((size_t)3 * s.V_local * s.H * 2); sp.Wr3 = take((size_t)3 * s.V_local * s.H * 2);
  };
  take_split(L.f_split);
  L.f_gather = o;                       // records of all vocab shards when the library's own communicator exchanges them
  if (comm_world > 1 && s.V_local < s.V) take((size_t)comm_world * L.row_stats_bytes);
  L.fwd_total = o;

  o = 0;
  L.b_Hp = take(parts * cap_tot * s.H * 2);
  L.b_Hf = take(parts * L.cap_ref * s.H * 2);
  L.b_dz = take(parts * L.chunk_rows * L.ldz * 2);
  L.b_dW32 = ((s.flags & DPA_FLAG_NEED_DW) && !L.split) ? take((size_t)s.V_local * s.H * 4) : o;   // fp32: dW itself accumulates
  take_split(L.b_split);
  // (library-side vocab-parallel exchange: the packed dH partial sums live in the communicator's peer memory)
  L.bwd_total = o;

  o = 0;
  L.u_Hp = take(parts * cap_tot * s.H * 2);
  L.u_Hf = take(parts * L.cap_ref * s.H * 2);
  L.u_part_pair = take((size_t)L.ng_pair * L.cap_pair * 16);
  L.u_part_ref = take((size_t)L.ng_ref * L.cap_ref * 16);
  L.u_part_u = take((size_t)L.ng_ref * L.cap_ref * 4);
  // Kept bf16 logits: chunk c < n_slots owns slot c (its pair rows' logits are written by the statistics GEMM and turned into
  // dLogits in place, never recomputed); chunks past the pair segment reuse slot 0.  One slot unless every position is
  // processed (DPA_FLAG_DENSE: the pair segment then spans several chunks and all of them are kept -- 180 GB of HBM).
  L.n_slots = ((s.flags & DPA_FLAG_DENSE) && !L.split) ? std::max(1, (L.cap_pair + L.chunk_rows - 1) / L.chunk_rows) : 1;
  L.n_slots = std::min(L.n_slots, L.n_chunks);
  L.u_dz = take(parts * L.n_slots * L.chunk_rows * L.ldz * 2);
  L.u_zr = take((size_t)L.chunk_rows * L.ldz * (L.split ? 4 : 2));
  L.u_zp32 = L.split ? take((size_t)L.chunk_rows * L.ldz * 4) : o;
  L.u_dW32 = ((s.flags & DPA_FLAG_NEED_DW) && !L.split) ? take((size_t)s.V_local * s.H * 4) : o;
  take_split(L.u_split);
  // (library-side exchange: the shards' statistics records meet in peer memory, not in this scratch)
  L.fused_total = o;
  return DPA_OK;
}

template <class T> T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_common(const dpa_args_t* a, Layout& L, DeviceInfo** dev) {
  if (!a) return fail(DPA_ERR_ARG, "args is NULL");
  int rc = make_layout(a->shape, L);
  if (rc) return rc;
  rc = device_info(dev);
  if (rc) return rc;
  const void* ptrs[] = {a->h_pair, a->labels_pair, a->signs_pair, a->h_ref_pol, a->h_ref_frz, a->labels_ref, a->W, a->W_ref};
  const char* names[] = {"h_pair", "labels_pair", "signs_pair", "h_ref_pol", "h_ref_frz", "labels_ref", "W", "W_ref"};
  const bool given = a->ref_logits != nullptr;       // frozen logits given: h_ref_frz / W_ref are not needed
  for (int i = 0; i < 8; ++i) {
    if (given && (i == 4 || i == 7) && !ptrs[i]) continue;
    if (!ptrs[i]) return fail(DPA_ERR_ARG, "%s is NULL", names[i]);
    if (!aligned16(ptrs[i])) return fail(DPA_ERR_ALIGN, "%s must be 16-byte aligned", names[i]);
  }
  if (!a->state || a->state_bytes < L.state_total) return fail(DPA_ERR_WORKSPACE, "state workspace: need %zu bytes, got %zu", L.state_total, a->state_bytes);
  if (!aligned16(a->state) || (a->scratch && !aligned16(a->scratch))) return fail(DPA_ERR_ALIGN, "workspaces must be 16-byte aligned (256 recommended)");
  if (given) {
    if (!(a->shape.flags & DPA_FLAG_FUSED_GRAD))
      return fail(DPA_ERR_ARG, "ref_logits needs DPA_FLAG_FUSED_GRAD (+ DPA_FLAG_LOSS_ONLY for evaluation): the chunked forward keeps bf16 logits");
    if (a->ref_logits_dtype != DPA_BF16 && a->ref_logits_dtype != DPA_F32) return fail(DPA_ERR_DTYPE, "ref_logits must be bf16 or fp32");
    const long long esz = a->ref_logits_dtype == DPA_F32 ? 4 : 2;
    if (a->ref_logits_ld < a->shape.V_local) return fail(DPA_ERR_BAD_SHAPE, "ref_logits_ld=%lld < V_local=%d", (long long)a->ref_logits_ld, a->shape.V_local);
    if (!aligned16(a->ref_logits) || (a->ref_logits_ld * esz) % 16) return fail(DPA_ERR_ALIGN, "ref_logits and its row stride must be 16-byte aligned");
  }
  if ((a->shape.flags & DPA_FLAG_LOSS_ONLY) && !(a->shape.flags & DPA_FLAG_FUSED_GRAD))
    return fail(DPA_ERR_ARG, "DPA_FLAG_LOSS_ONLY modifies DPA_FLAG_FUSED_GRAD");
  if (a->shape.flags & DPA_FLAG_FUSED_NORM) {
    if (!a->norm_w || !a->norm_w_ref) return fail(DPA_ERR_ARG, "DPA_FLAG_FUSED_NORM set but norm_w / norm_w_ref is NULL");
    if (!aligned16(a->norm_w) || !aligned16(a->norm_w_ref) || (a->d_norm_w && !aligned16(a->d_norm_w)))
      return fail(DPA_ERR_ALIGN, "norm weights must be 16-byte aligned");
    if (!(a->norm_eps >= 0.f)) return fail(DPA_ERR_ARG, "norm_eps must be >= 0");
  }
  return DPA_OK;
}

MetaArgs meta_args(const dpa_args_t* a, const Layout& L) {
  MetaArgs m{};
  m.labels_pair = a->labels_pair; m.signs_pair = a->signs_pair; m.labels_ref = a->labels_ref;
  m.B = a->shape.B; m.T = a->shape.T; m.B_ref = a->shape.B_ref; m.T_ref = a->shape.T_ref; m.V = a->shape.V;
  m.dense = (a->shape.flags & DPA_FLAG_DENSE) ? 1 : 0;
  m.n_pos_pair = L.n_pos_pair; m.n_pos_ref = L.n_pos_ref; m.nblk_pair = L.nblk_pair; m.nblk_ref = L.nblk_ref;
  m.cap_pair = L.cap_pair; m.cap_ref = L.cap_ref;
  m.meta = at<DpaMeta>(a->state, L.meta); m.bitmap = at<unsigned>(a->state, L.bitmap);
  m.block_counts = at<int>(a->state, L.block_counts);
  m.src_pair = at<int>(a->state, L.src_pair); m.lab_pair = at<int>(a->state, L.lab_pair); m.col_pair = at<int>(a->state, L.col_pair);
  m.src_ref = at<int>(a->state, L.src_ref); m.act_ref = at<int>(a->state, L.act_ref);
  m.status = a->status;
  return m;
}

int run_gather(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, void* scratch, size_t oHp, size_t oHf, cudaStream_t st) {
  if (L.split) {            // fp32 inputs: packed rows as split-bf16 pairs
    GatherSplitArgs g{at<DpaMeta>(a->state, L.meta), at<int>(a->state, L.src_pair), at<int>(a->state, L.src_ref),
                      static_cast<const float*>(a->h_pair), static_cast<const float*>(a->h_ref_pol), static_cast<const float*>(a->h_ref_frz),
                      at<__nv_bfloat16>(scratch, oHp), (size_t)(L.cap_pair + L.cap_ref) * a->shape.H,
                      at<__nv_bfloat16>(scratch, oHf), (size_t)L.cap_ref * a->shape.H, a->shape.H};
    ProfScope ps(PK_GATHER, 1, st);
    gather_rows_split_kernel<<<d->sms * 8, 256, 0, st>>>(g);
    LAUNCH_CHECK("gather_rows_split");
    return DPA_OK;
  }
  GatherArgs g{};
  g.meta = at<DpaMeta>(a->state, L.meta);
  g.src_pair = at<int>(a->state, L.src_pair); g.src_ref = at<int>(a->state, L.src_ref);
  g.h_pair = static_cast<const __nv_bfloat16*>(a->h_pair);
  g.h_ref_pol = static_cast<const __nv_bfloat16*>(a->h_ref_pol);
  g.h_ref_frz = static_cast<const __nv_bfloat16*>(a->h_ref_frz);
  g.Hp = at<__nv_bfloat16>(scratch, oHp); g.Hf = at<__nv_bfloat16>(scratch, oHf); g.H = a->shape.H;
  g.norm_w = static_cast<const __nv_bfloat16*>(a->norm_w); g.norm_w_ref = static_cast<const __nv_bfloat16*>(a->norm_w_ref);
  g.norm_eps = a->norm_eps; g.rinv = at<float>(a->state, L.rinv);
  ProfScope ps(PK_GATHER, 1, st);
  if (a->shape.flags & DPA_FLAG_FUSED_NORM) gather_rows_kernel<true><<<d->sms * 8, 256, 0, st>>>(g);
  else gather_rows_kernel<false><<<d->sms * 8, 256, 0, st>>>(g);
  LAUNCH_CHECK("gather_rows");
  return DPA_OK;
}

// fp32 inputs: lm_head weights -> split-bf16 triples in scratch (W_ref shares W's when it aliases W)
int run_split_weights(const dpa_args_t* a, const Layout&, const DeviceInfo* d, const Layout::Split& sp, cudaStream_t st) {
  const size_t n = (size_t)a->shape.V_local * a->shape.H;
  const bool alias = a->W_ref == a->W;
  ProfScope ps(PK_GATHER, alias ? 1 : 2, st);
  split_f32_kernel<<<d->sms * 8, 256, 0, st>>>(static_cast<const float*>(a->W), at<__nv_bfloat16>(a->scratch, sp.W3), n);
  if (!alias) split_f32_kernel<<<d->sms * 8, 256, 0, st>>>(static_cast<const float*>(a->W_ref), at<__nv_bfloat16>(a->scratch, sp.Wr3), n);
  LAUNCH_CHECK("split_weights");
  return DPA_OK;
}

// DPA_FLAG_FUSED_NORM: turn the dH rows (gradients w.r.t. the normed hidden states) into gradients w.r.t. the pre-norm
// stream, in place, and finish d_norm_w.  Runs once every dH chunk is complete.
int run_norm_backward(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, cudaStream_t st) {
  if (!(a->shape.flags & DPA_FLAG_FUSED_NORM)) return DPA_OK;
  const dpa_shape_t& s = a->shape;
  float* dnw32 = a->d_norm_w ? at<float>(a->state, L.dnw32) : nullptr;
  ProfScope ps(PK_MISC, a->d_norm_w ? 2 : 1, st);
  if (dnw32) CUDA_TRY(cudaMemsetAsync(dnw32, 0, (size_t)s.H * 4, st));
  NormBwdArgs n{at<DpaMeta>(a->state, L.meta), at<int>(a->state, L.src_pair), at<int>(a->state, L.src_ref),
                static_cast<const __nv_bfloat16*>(a->h_pair), static_cast<const __nv_bfloat16*>(a->h_ref_pol),
                static_cast<__nv_bfloat16*>(a->dH_pair), static_cast<__nv_bfloat16*>(a->dH_ref),
                static_cast<const __nv_bfloat16*>(a->norm_w), at<float>(a->state, L.rinv), dnw32, s.H};
  const size_t smem = dnw32 ? (size_t)s.H * 4 : 0;
  if (smem > 48 * 1024) {
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(norm_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); });
  }
  norm_backward_kernel<<<d->sms * 2, 256, smem, st>>>(n);
  LAUNCH_CHECK("norm_backward");
  if (dnw32) {
    f32_to_bf16_kernel<<<(s.H / 4 + 255) / 256, 256, 0, st>>>(dnw32, static_cast<__nv_bfloat16*>(a->d_norm_w), (size_t)s.H);
    LAUNCH_CHECK("d_norm_w convert");
  }
  return DPA_OK;
}

enum : int { MERGE_ONLY = 0, MERGE_FINISH = 1, MERGE_FINISH_KL_LATER = 2 };
int run_merge(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, const float4* part_pair, const float4* part_ref,
              const float* part_u, int n_pair, int n_ref, size_t stride_pair, size_t stride_ref, size_t stride_u, cudaStream_t st,
              RowRange rr = RowRange{SEG_ALL, 0, 0}, int finish = MERGE_ONLY, const PeerView* push = nullptr, size_t push_off = 0);

}  // namespace

// merge kernel with explicit strides lives here so that local partials and gathered shard records share it
namespace dpa {
struct MergeStrided {
  const DpaMeta* meta;
  RowRange rr;
  const float4* part_pair; const float4* part_ref; const float* part_u;
  int n_parts_pair, n_parts_ref;
  size_t stride_pair, stride_ref, stride_u;     // element strides between partial records
  float4* stats_pair; float4* stats_ref; float* stats_u;
  // finish != 0: also the per-row results (what finish_rows_kernel computes), in the same pass over the merged statistics
  int finish, kl_from_stats;
  const int* act_ref;
  float* lse_pair; float* logp_pair; float* lse_ref_p; float* lse_ref_r; float* kl_row;
  // vocab-parallel, library-side exchange: push the merged rows of THIS shard into slot `rank` of the statistics gather
  // region of every peer (pv.world == 0: no push).  Record layout = [stats_pair | stats_ref | stats_u] of this state.
  PeerView pv; size_t push_off, rec_bytes, ref_off;
};
__device__ __forceinline__ void push_float4(const PeerView& pv, size_t off, const float4& v) {
  const uint4 u = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
#pragma unroll
  for (int r = 0; r < kPeerMax; ++r)
    if (r < pv.world) st_peer_v4(reinterpret_cast<uint4*>(pv.base[r] + off), u);
}
__global__ void merge_strided_kernel(MergeStrided a) {
  const int p0 = a.meta->p0;
  int lo, hi; row_range(a.meta, a.rr, lo, hi);
  for (int r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x) {
    if (r < p0) {
      float m = -INFINITY;
      for (int g = 0; g < a.n_parts_pair; ++g) m = fmaxf(m, a.part_pair[g * a.stride_pair + r].x);
      float s = 0.f, zy = 0.f;
      for (int g = 0; g < a.n_parts_pair; ++g) {
        const float4 p = a.part_pair[g * a.stride_pair + r];
        s += p.y * expf(p.x - m); zy += p.z;
      }
      a.stats_pair[r] = make_float4(m, s, zy, 0.f);
      if (a.pv.world) push_float4(a.pv, a.push_off + (size_t)a.pv.rank * a.rec_bytes + (size_t)r * 16, make_float4(m, s, zy, 0.f));
      if (a.finish) {
        const float lse = m + logf(s);
        a.lse_pair[r] = lse; a.logp_pair[r] = zy - lse;          // z[y] - LSE (halva_trainer.py:407)
      }
    } else {
      const int j = r - p0;
      float mp = -INFINITY, mr = -INFINITY;
      for (int g = 0; g < a.n_parts_ref; ++g) {
        const float4 p = a.part_ref[g * a.stride_ref + j];
        mp = fmaxf(mp, p.x); mr = fmaxf(mr, p.z);
      }
      float sp = 0.f, sr = 0.f, u = 0.f;
      for (int g = 0; g < a.n_parts_ref; ++g) {
        const float4 p = a.part_ref[g * a.stride_ref + j];
        const float wr = expf(p.z - mr);
        sp += p.y * expf(p.x - mp); sr += p.w * wr;
        if (a.part_u) u += a.part_u[g * a.stride_u + j] * wr;
      }
      a.stats_ref[j] = make_float4(mp, sp, mr, sr);
      if (a.part_u) a.stats_u[j] = u;
      if (a.pv.world) push_float4(a.pv, a.push_off + (size_t)a.pv.rank * a.rec_bytes + a.ref_off + (size_t)j * 16, make_float4(mp, sp, mr, sr));
      if (a.finish) {                                  // same maths as finish_rows_kernel (dpa_kernels.cuh)
        a.lse_ref_p[j] = mp + logf(sp); a.lse_ref_r[j] = mr + logf(sr);
        const float dlse = (mr - mp) + logf(sr / sp);
        if (a.kl_from_stats) a.kl_row[j] = a.act_ref[j] ? (u / sr - dlse) : 0.f;
        else a.kl_row[j] = dlse;                       // fused-gradient forward: dz_kl_elem_kernel finishes the row
      }
    }
  }
  if (a.pv.world) __threadfence_system();        // the pushed rows are ordered before the barrier that follows
}
// every shard's part of the divergence -> all peers (slot = rank); summed in fixed order by finish_div_kernel
__global__ void push_div_kernel(PeerView pv, const float* losses) {
  if (threadIdx.x < pv.world)
    *reinterpret_cast<volatile float*>(pv.base[threadIdx.x] + kPeerLossOff + 4 * pv.rank) = losses[2];
  __threadfence_system();
}
__global__ void finish_div_kernel(PeerView pv, float* losses, float alpha) {
  if (threadIdx.x == 0) {
    const float* part = reinterpret_cast<const float*>(pv.base[pv.rank] + kPeerLossOff);
    float div = 0.f;
    for (int r = 0; r < pv.world; ++r) div += __ldcg(part + r);
    losses[2] = div; losses[0] = losses[1] + alpha * div;
  }
}
}  // namespace dpa

namespace {

int run_merge(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, const float4* part_pair, const float4* part_ref,
              const float* part_u, int n_pair, int n_ref, size_t stride_pair, size_t stride_ref, size_t stride_u, cudaStream_t st,
              RowRange rr, int finish, const PeerView* push, size_t push_off) {
  MergeStrided m{at<DpaMeta>(a->state, L.meta), rr, part_pair, part_ref, part_u, n_pair, n_ref, stride_pair, stride_ref, stride_u,
                 at<float4>(a->state, L.stats_pair), at<float4>(a->state, L.stats_ref), at<float>(a->state, L.stats_u),
                 finish != MERGE_ONLY, finish == MERGE_FINISH, at<int>(a->state, L.act_ref), at<float>(a->state, L.lse_pair),
                 at<float>(a->state, L.logp_pair), at<float>(a->state, L.lse_ref_p), at<float>(a->state, L.lse_ref_r),
                 at<float>(a->state, L.kl_row), push ? *push : PeerView{}, push_off, L.row_stats_bytes, L.stats_ref - L.stats_pair};
  ProfScope ps(PK_REDUCE, 1, st);
  merge_strided_kernel<<<d->sms * 2, 256, 0, st>>>(m);
  LAUNCH_CHECK("merge_partials");
  return DPA_OK;
}

int run_finish_rows(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, cudaStream_t st, RowRange rr, bool kl_from_stats = true) {
  FinishArgs f{at<DpaMeta>(a->state, L.meta), rr, at<float4>(a->state, L.stats_pair), at<float4>(a->state, L.stats_ref),
               kl_from_stats ? at<float>(a->state, L.stats_u) : nullptr, at<int>(a->state, L.act_ref), at<float>(a->state, L.lse_pair),
               at<float>(a->state, L.logp_pair), at<float>(a->state, L.lse_ref_p), at<float>(a->state, L.lse_ref_r),
               at<float>(a->state, L.kl_row)};
  ProfScope ps(PK_REDUCE, 1, st);
  finish_rows_kernel<<<d->sms * 2, 256, 0, st>>>(f);
  LAUNCH_CHECK("finish_rows");
  return DPA_OK;
}

// P[b,k] / N[b,k] sums and the loss phases asked for, one launch (LOSS_ALIGN needs the pair rows' log-probs, LOSS_TOTAL every
// kl_row).  `ticket`: DpaMeta::pad slot that counts finished blocks -- a different one per launch of the same call.
int run_losses(const dpa_args_t* a, const Layout& L, cudaStream_t st, int phases, int ticket) {
  LossArgs l{at<DpaMeta>(a->state, L.meta), at<int>(a->state, L.src_pair), at<int>(a->state, L.col_pair), at<float>(a->state, L.logp_pair),
             at<float>(a->state, L.PN), at<float>(a->state, L.kl_row), at<float>(a->state, L.sigma), a->losses,
             a->shape.B, a->shape.B_ref, a->shape.T, a->alpha, phases, ticket};
  ProfScope ps(PK_REDUCE, 1, st);
  loss_kernel<<<(phases & LOSS_ALIGN) ? 2 * a->shape.B : 1, 256, 0, st>>>(l);
  LAUNCH_CHECK("loss");
  return DPA_OK;
}

int run_finish(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, cudaStream_t st) {
  int rc = run_finish_rows(a, L, d, st, RowRange{SEG_ALL, 0, 0});
  if (rc) return rc;
  return run_losses(a, L, st, LOSS_ALIGN | LOSS_TOTAL, 0);
}

int run_meta(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, cudaStream_t st) {
  CUDA_TRY(cudaMemsetAsync(at<char>(a->state, L.meta), 0, L.block_counts - L.meta, st));   // DpaMeta + bitmap
  MetaArgs m = meta_args(a, L);
  const bool i64 = a->shape.index_dtype == DPA_I64;
  const int nblk = L.nblk_pair + L.nblk_ref;
  const int bm_grid = std::min((L.n_pos_pair + 255) / 256, d->sms * 4);
  ProfScope ps(PK_META, 4, st);
  if (i64) {
    meta_bitmap_kernel<true><<<bm_grid, 256, 0, st>>>(m);
    meta_count_kernel<true><<<nblk, kScanBlock, 0, st>>>(m);
    meta_fill_kernel<true><<<nblk, kScanBlock, 0, st>>>(m);
  } else {
    meta_bitmap_kernel<false><<<bm_grid, 256, 0, st>>>(m);
    meta_count_kernel<false><<<nblk, kScanBlock, 0, st>>>(m);
    meta_fill_kernel<false><<<nblk, kScanBlock, 0, st>>>(m);
  }
  meta_pad_kernel<<<1, 256, 0, st>>>(m);
  LAUNCH_CHECK("meta");
  return DPA_OK;
}

int check_grad_outputs(const dpa_args_t* a) {
  if (!a->dH_pair || !a->dH_ref) return fail(DPA_ERR_ARG, "dH_pair/dH_ref is NULL");
  if ((a->shape.flags & DPA_FLAG_NEED_DW) && !a->dW) return fail(DPA_ERR_ARG, "DPA_FLAG_NEED_DW set but dW is NULL");
  if (!aligned16(a->dH_pair) || !aligned16(a->dH_ref) || (a->dW && !aligned16(a->dW))) return fail(DPA_ERR_ALIGN, "gradient buffers must be 16-byte aligned");
  return DPA_OK;
}

// ---------------------------------------------------------------------------------- optional NCCL communicator
// Vocab-parallel lm_head has one exchange step each way (all-gather of the per-row statistics record in the forward,
// all-reduce of the dH partial sums in the backward).  With dpa_comm_init() the library issues them itself on the
// caller's stream, so dpa_forward / dpa_backward stay ONE call each; without it the caller exchanges (dpa_forward_stats /
// dpa_forward_finish).  NCCL is bound at run time (the copy the process already loaded, else libnccl.so.2): the library
// has no link-time dependency on it and still loads on a box without NCCL.
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
// Symmetric peer memory (dpa_peer.cuh): one cudaMalloc per rank, mapped into every peer with CUDA IPC.  (Re)allocated
// collectively, on first use or when a larger shape arrives -- the only place the communicator synchronises the device.
struct PeerMem {
  char* base = nullptr; size_t bytes = 0;
  char* peer[kPeerMax] = {};           // peer[rank] == base
  bool failed = false; char why[160] = "";
};
struct Comm { NcclComm comm = nullptr; int rank = 0, world = 1, device = -1; cudaStream_t stream = nullptr; PeerMem pm; };   // stream: the
// exchange of the backward runs beside the caller's stream (fork / join with events) so that it overlaps the remaining GEMMs
Comm g_comm;
constexpr int kNcclUint8 = 1;

int nccl_bind() {
  if (g_nccl.lib) return DPA_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy torch (or the host program) already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(DPA_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define BIND(field, sym) do { *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, sym); \
    if (!g_nccl.field) return fail(DPA_ERR_NCCL, "libnccl: symbol %s missing", sym); } while (0)
  BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
  BIND(AllGather, "ncclAllGather"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  g_nccl.lib = h;
  return DPA_OK;
}
#define NCCL_TRY(x) do { int r_ = (x); if (r_ != 0) return fail(DPA_ERR_NCCL, "%s: %s", #x, g_nccl.GetErrorString(r_)); } while (0)
bool comm_active(const dpa_shape_t& s) { return g_comm.comm && g_comm.world > 1 && s.V_local < s.V; }

void peer_mem_release() {
  PeerMem& pm = g_comm.pm;
  for (int r = 0; r < kPeerMax; ++r)
    if (pm.peer[r] && pm.peer[r] != pm.base) cudaIpcCloseMemHandle(pm.peer[r]);
  if (pm.base) cudaFree(pm.base);
  pm = PeerMem{};
}
// Collective over the communicator: make sure every rank's symmetric allocation holds `data_bytes` behind the control block.
int ensure_peer_mem(size_t data_bytes) {
  PeerMem& pm = g_comm.pm;
  if (pm.failed) return fail(DPA_ERR_NCCL, "peer memory unavailable: %s", pm.why);
  const size_t need = kPeerCtrlBytes + align_up(data_bytes, 256);
  if (pm.base && pm.bytes >= need) return DPA_OK;
  const int G = g_comm.world, me = g_comm.rank;
  if (G > kPeerMax) { pm.failed = true; snprintf(pm.why, sizeof(pm.why), "world %d > %d", G, kPeerMax); return fail(DPA_ERR_NCCL, "peer memory unavailable: %s", pm.why); }
  auto bad = [&](const char* what, cudaError_t e) {
    pm.failed = true;
    snprintf(pm.why, sizeof(pm.why), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return fail(DPA_ERR_NCCL, "peer memory unavailable: %s", pm.why);
  };
  cudaError_t e = cudaDeviceSynchronize();           // nothing may still use the old allocation
  if (e != cudaSuccess) return bad("cudaDeviceSynchronize", e);
  peer_mem_release();
  const size_t bytes = align_up(need + need / 4, 2u << 20);       // some headroom: shapes grow a little from batch to batch
  if ((e = cudaMalloc(&pm.base, bytes)) != cudaSuccess) return bad("cudaMalloc(symmetric buffer)", e);
  pm.bytes = bytes;
  cudaStream_t st = g_comm.stream;
  if ((e = cudaMemsetAsync(pm.base, 0, kPeerCtrlBytes, st)) != cudaSuccess) return bad("cudaMemsetAsync", e);
  cudaIpcMemHandle_t mine;
  if ((e = cudaIpcGetMemHandle(&mine, pm.base)) != cudaSuccess) return bad("cudaIpcGetMemHandle", e);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  char* stage = nullptr;                              // [G + 1][64]: every rank's handle, then mine as the send buffer
  if ((e = cudaMalloc(&stage, (size_t)(G + 1) * 64)) != cudaSuccess) return bad("cudaMalloc(stage)", e);
  if ((e = cudaMemcpyAsync(stage + (size_t)G * 64, &mine, 64, cudaMemcpyHostToDevice, st)) != cudaSuccess) { cudaFree(stage); return bad("cudaMemcpyAsync", e); }
  // (stream order: the control block is zeroed before this rank's handle can reach any peer)
  int nr = g_nccl.AllGather(stage + (size_t)G * 64, stage, 64, kNcclUint8, g_comm.comm, st);
  if (nr != 0) { cudaFree(stage); pm.failed = true; snprintf(pm.why, sizeof(pm.why), "ncclAllGather(handles): %s", g_nccl.GetErrorString(nr)); return fail(DPA_ERR_NCCL, "%s", pm.why); }
  std::vector<cudaIpcMemHandle_t> all(G);
  e = cudaMemcpyAsync(all.data(), stage, (size_t)G * 64, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(stage);
  if (e != cudaSuccess) return bad("handle read-back", e);
  for (int r = 0; r < G; ++r) {
    if (r == me) { pm.peer[r] = pm.base; continue; }
    void* ptr = nullptr;
    if ((e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) return bad("cudaIpcOpenMemHandle", e);
    pm.peer[r] = static_cast<char*>(ptr);
  }
  if (getenv("DPA_VERBOSE")) fprintf(stderr, "dpa: rank %d/%d: symmetric peer buffer of %.1f MB mapped\n", me, G, bytes / 1e6);
  return DPA_OK;
}
PeerView peer_view() {
  PeerView v{};
  for (int r = 0; r < kPeerMax; ++r) v.base[r] = g_comm.pm.peer[r];
  v.rank = g_comm.rank; v.world = g_comm.world;
  return v;
}

// ---------------------------------------------------------------------------------- one call's GEMM plan
// Tensor maps over the packed operands of this call and one method per GEMM class.  `scratch` offsets differ between
// the forward / backward / fused-forward layouts, everything else is shared.
struct Plan {
  const dpa_args_t* a; const Layout& L; const DeviceInfo* d = nullptr; cudaStream_t st;
  const dpa_shape_t& s;
  const DpaMeta* meta;
  __nv_bfloat16* dz = nullptr; __nv_bfloat16* zr = nullptr; float* dW32 = nullptr;
  float* zp32 = nullptr; float* zr32 = nullptr;      // fp32 inputs, fused-gradient forward: the chunk's kept fp32 logits
  int n_slots = 1;                        // kept-logit slots behind `dz` (fused-gradient forward; 1 elsewhere)
  int sync_used = 0;                      // K-lockstep counters handed out to this call's GEMM launches
  unsigned* next_sync() {
    unsigned* base = at<unsigned>(a->state, L.sync_ctr);
    const int i = sync_used < kSyncCtrs(L.n_chunks) ? sync_used++ : -1;
    return i >= 0 ? base + i : nullptr;
  }
  void set_sync(GemmWork& w, int every) { w.sync_ctr = every != 0 ? next_sync() : nullptr; w.sync_every = w.sync_ctr ? every : 0; }
  int slot_of(int ch) const { return ch < n_slots ? ch : 0; }
  __nv_bfloat16* dz_of(int ch) const { return dz + (size_t)slot_of(ch) * L.chunk_rows * L.ldz; }
  __nv_bfloat16* dH_packed = nullptr;     // vocab-parallel with the library's communicator: dH rows stay packed until reduced
  float4* part_pair = nullptr; float4* part_ref = nullptr; float* part_u = nullptr;
  CUtensorMap tmHp, tmHf, tmW256, tmW128, tmW64, tmWr256, tmWr128, tmWr64, tmDzK, tmWmn, tmDzMN, tmHpMN;
  // fp32 inputs: the maps above cover the high-order bf16 parts (in scratch), these the mid ([0]) and low-order ([1]) parts
  // vocab-parallel with the library's communicator: after chunk ch's dH GEMM its packed rows are all-reduced on the
  // communicator's own stream while this stream goes on with dW and the next chunk (joined before the final scatter)
  const Layout::Split* split = nullptr;
  size_t dz_part = 0;                      // elements between the hi / mid / lo copies of dz
  CUtensorMap tmHpS[2], tmHfS[2], tmW128S[2], tmW64S[2], tmWr128S[2], tmWr64S[2], tmDzKS[2], tmWmnS[2], tmDzMNS[2], tmHpMNS[2];
  CUtensorMap tmW32, tmWr32, tmW32S[2], tmWr32S[2];       // 32-row boxes: each CTA's half of a 64-column dual tile (CTA pairs)

  Plan(const dpa_args_t* a_, const Layout& L_, const DeviceInfo* d_, cudaStream_t st_)
      : a(a_), L(L_), d(d_), st(st_), s(a_->shape), meta(at<DpaMeta>(a_->state, L_.meta)) {}

  int init(size_t oHp, size_t oHf, size_t oPartPair, size_t oPartRef, size_t oPartU, bool has_parts, __nv_bfloat16* dz_, __nv_bfloat16* zr_,
           float* dW32_, const Layout::Split* split_ = nullptr, int n_slots_ = 1) {
    const int cap_tot = L.cap_pair + L.cap_ref;
    dz = dz_; zr = zr_; dW32 = dW32_; split = split_; n_slots = n_slots_;
    const int dz_rows = n_slots * L.chunk_rows;
    if (has_parts) {
      part_pair = at<float4>(a->scratch, oPartPair); part_ref = at<float4>(a->scratch, oPartRef); part_u = at<float>(a->scratch, oPartU);
    }
    int rc;
    char* Hp = at<char>(a->scratch, oHp);
    if ((rc = make_operand_tmap(d, &tmHp, Hp, KMAJOR, cap_tot, s.H, s.H, BM))) return rc;
    const bool frozen = a->h_ref_frz && a->W_ref;      // false: the frozen model's logits are given (ref_logits)
    if (frozen && (rc = make_operand_tmap(d, &tmHf, at<char>(a->scratch, oHf), KMAJOR, L.cap_ref, s.H, s.H, BM))) return rc;
    // the bf16 lm_head operands: the caller's tensors, or (fp32 inputs) the high-order halves split into scratch
    const bool alias = a->W_ref == a->W;
    const size_t w_part = (size_t)s.V_local * s.H * 2;                     // bytes between the parts of a split weight
    const char* W3 = split ? at<char>(a->scratch, split->W3) : nullptr;
    const char* Wr3 = split ? at<char>(a->scratch, alias ? split->W3 : split->Wr3) : nullptr;
    const void* Wb = split ? W3 : a->W;
    const void* Wrb = split ? Wr3 : a->W_ref;
    // lm_head rows as the K-major B operand: boxes of 256 (single CTA), 128 (CTA pair / dual single CTA), 64 (dual pair)
    if ((rc = make_operand_tmap(d, &tmW256, Wb, KMAJOR, s.V_local, s.H, s.H, 256))) return rc;
    if ((rc = make_operand_tmap(d, &tmW128, Wb, KMAJOR, s.V_local, s.H, s.H, 128))) return rc;
    if ((rc = make_operand_tmap(d, &tmW64, Wb, KMAJOR, s.V_local, s.H, s.H, 64))) return rc;
    if (frozen) {
      if ((rc = make_operand_tmap(d, &tmWr256, Wrb, KMAJOR, s.V_local, s.H, s.H, 256))) return rc;
      if ((rc = make_operand_tmap(d, &tmWr128, Wrb, KMAJOR, s.V_local, s.H, s.H, 128))) return rc;
      if ((rc = make_operand_tmap(d, &tmWr64, Wrb, KMAJOR, s.V_local, s.H, s.H, 64))) return rc;
    }
    if (dz) {
      // dH = dZ[rows x V] * W[V x H]:      A = dz K-major (K = vocab),  B = W MN-major (N = hidden contiguous)
      // dW = dZ^T[V x rows] * H[rows x H]: A = dz MN-major (M = vocab), B = packed hidden rows MN-major
      if ((rc = make_operand_tmap(d, &tmDzK, dz, KMAJOR, dz_rows, s.V_local, L.ldz, BM))) return rc;
      if ((rc = make_operand_tmap(d, &tmWmn, Wb, MNMAJOR, s.H, s.V_local, s.H, 0))) return rc;
      if ((rc = make_operand_tmap(d, &tmDzMN, dz, MNMAJOR, s.V_local, dz_rows, L.ldz, 0))) return rc;
      if ((rc = make_operand_tmap(d, &tmHpMN, Hp, MNMAJOR, s.H, cap_tot, s.H, 0))) return rc;
    }
    if (split) {
      const size_t hp_part = (size_t)cap_tot * s.H * 2, hf_part = (size_t)L.cap_ref * s.H * 2;
      dz_part = (size_t)L.chunk_rows * L.ldz;
      if ((rc = make_operand_tmap(d, &tmW32, Wb, KMAJOR, s.V_local, s.H, s.H, 32))) return rc;
      if ((rc = make_operand_tmap(d, &tmWr32, Wrb, KMAJOR, s.V_local, s.H, s.H, 32))) return rc;
      for (int p = 0; p < 2; ++p) {
        char* HpP = Hp + (p + 1) * hp_part;
        const char* WP = W3 + (p + 1) * w_part;
        if ((rc = make_operand_tmap(d, &tmHpS[p], HpP, KMAJOR, cap_tot, s.H, s.H, BM))) return rc;
        if ((rc = make_operand_tmap(d, &tmHfS[p], at<char>(a->scratch, oHf) + (p + 1) * hf_part, KMAJOR, L.cap_ref, s.H, s.H, BM))) return rc;
        if ((rc = make_operand_tmap(d, &tmW128S[p], WP, KMAJOR, s.V_local, s.H, s.H, 128))) return rc;
        if ((rc = make_operand_tmap(d, &tmW64S[p], WP, KMAJOR, s.V_local, s.H, s.H, 64))) return rc;
        if ((rc = make_operand_tmap(d, &tmWr64S[p], Wr3 + (p + 1) * w_part, KMAJOR, s.V_local, s.H, s.H, 64))) return rc;
        if ((rc = make_operand_tmap(d, &tmWr128S[p], Wr3 + (p + 1) * w_part, KMAJOR, s.V_local, s.H, s.H, 128))) return rc;
        if ((rc = make_operand_tmap(d, &tmW32S[p], WP, KMAJOR, s.V_local, s.H, s.H, 32))) return rc;
        if ((rc = make_operand_tmap(d, &tmWr32S[p], Wr3 + (p + 1) * w_part, KMAJOR, s.V_local, s.H, s.H, 32))) return rc;
        if (dz) {
          __nv_bfloat16* dzP = dz + (p + 1) * dz_part;
          if ((rc = make_operand_tmap(d, &tmDzKS[p], dzP, KMAJOR, L.chunk_rows, s.V_local, L.ldz, BM))) return rc;
          if ((rc = make_operand_tmap(d, &tmWmnS[p], WP, MNMAJOR, s.H, s.V_local, s.H, 0))) return rc;
          if ((rc = make_operand_tmap(d, &tmDzMNS[p], dzP, MNMAJOR, s.V_local, L.chunk_rows, L.ldz, 0))) return rc;
          if ((rc = make_operand_tmap(d, &tmHpMNS[p], HpP, MNMAJOR, s.H, cap_tot, s.H, 0))) return rc;
        }
      }
    }
    return DPA_OK;
  }
  static SplitMaps split_maps(const CUtensorMap (&A)[2], const CUtensorMap (&B)[2], const CUtensorMap (&A2)[2], const CUtensorMap (&B2)[2]) {
    SplitMaps m;
    for (int p = 0; p < 2; ++p) { m.A[p] = A[p]; m.B[p] = B[p]; m.A2[p] = A2[p]; m.B2[p] = B2[p]; }
    return m;
  }
  // split-bf16 GEMM configurations (fp32 inputs): 128-column accumulators (the K-chunk promotion keeps them in registers);
  // the column groups of the statistics stay those of the bf16 tilings (256 * Layout::tpu columns)
  static constexpr int kBnSplit = 128, kBnSplitDual = 64;
  int tpu_pair() const { return L.tpu; }                                  // tiles per unit of each tiling: the same columns
  int tpu_ref() const { return L.tpu * (kBnPair / kBnRef); }
  int tpu_split() const { return L.tpu * (kBnPair / kBnSplit); }
  int tpu_split_dual() const { return L.tpu * (kBnPair / kBnSplitDual); }
  template <int AMAJ, int BMAJ, int CG> using SplitCfg = GemmCfg<kBnSplit, false, AMAJ, BMAJ, CG, 6>;
  template <int CG> using SplitDualCfg = GemmCfg<kBnSplitDual, true, KMAJOR, KMAJOR, CG, 6>;
  // cta_group of the split kernels: 2 (default) = CTA pairs on 256-row tiles, each CTA staging half of the B tile -- less
  // L2 -> smem traffic per FLOP, which is what bounds these small-tile kernels (ncu: tensor pipe 36-48 % active with one
  // CTA per tile; cfg-1 fp32 12.1 -> 9.0 ms with pairs); 1 = one CTA per 128-row tile.  DPA_SPLIT_CG_SINGLE /
  // DPA_SPLIT_CG_DUAL override (measurements).
  static int split_cg(bool dual) {
    static const int v[2] = {[] { const char* e = getenv("DPA_SPLIT_CG_SINGLE"); return e ? atoi(e) : 2; }(),
                             [] { const char* e = getenv("DPA_SPLIT_CG_DUAL"); return e ? atoi(e) : 2; }()};
    return v[dual ? 1 : 0] == 1 ? 1 : 2;
  }
  // K-major B operand: bx1 / bx2 = maps whose box holds BN (one CTA) or BN / 2 (CTA pair) rows, with their mid / lo parts
  template <int AMAJ, int BMAJ, class Epi>
  int launch_split(const CUtensorMap& tmA, const CUtensorMap (&tmAS)[2], const CUtensorMap& bx1, const CUtensorMap (&bx1S)[2],
                   const CUtensorMap& bx2, const CUtensorMap (&bx2S)[2], GemmWork w, const typename Epi::Params& ep, const char* what) {
    if (ksync_split() == 0) { w.sync_every = 0; w.sync_ctr = nullptr; }      // (K-lockstep of the split kernels: DPA_KSYNC_SPLIT)
    if (split_cg(false) == 2) {
      const SplitMaps sm = split_maps(tmAS, bx2S, tmAS, bx2S);
      return launch_gemm<SplitCfg<AMAJ, BMAJ, 2>, Epi>(st, d->sms, tmA, bx2, tmA, bx2, w, ep, what, &sm);
    }
    const SplitMaps sm = split_maps(tmAS, bx1S, tmAS, bx1S);
    return launch_gemm<SplitCfg<AMAJ, BMAJ, 1>, Epi>(st, d->sms, tmA, bx1, tmA, bx1, w, ep, what, &sm);
  }
  template <class Epi>
  int launch_split_dual(GemmWork w, const typename Epi::Params& ep, const char* what) {
    if (ksync_split() == 0) { w.sync_every = 0; w.sync_ctr = nullptr; }
    if (split_cg(true) == 2) {
      const SplitMaps sm = split_maps(tmHpS, tmW32S, tmHfS, tmWr32S);
      return launch_gemm<SplitDualCfg<2>, Epi>(st, d->sms, tmHp, tmW32, tmHf, tmWr32, w, ep, what, &sm);
    }
    const SplitMaps sm = split_maps(tmHpS, tmW64S, tmHfS, tmWr64S);
    return launch_gemm<SplitDualCfg<1>, Epi>(st, d->sms, tmHp, tmW64, tmHf, tmWr64, w, ep, what, &sm);
  }

  // statistics of every pair row; store_rows > 0 also keeps the bf16 logits of the first store_rows packed rows in dz
  template <bool ROUND> int fwd_pair(int store_rows) {
    using Epi = EpiStatsPair<kBnPair, ROUND>;
    GemmWork w{meta, KIND_FWD_PAIR, 0, 0, 0, s.V_local, s.H, tpu_pair(), super_m().stats, hint_stats().a, hint_stats().b};
    set_sync(w, ksync_stats(L.tpu));
    typename Epi::Params ep{at<int>(a->state, L.lab_pair), part_pair, L.cap_pair, s.V_local, s.v_offset,
                            store_rows > 0 ? dz : nullptr, L.ldz, store_rows};
    ProfScope ps(PK_FWD_PAIR, 1, st);
    if constexpr (!ROUND) {
      if (split) {
        using EpiS = EpiStatsPair<kBnSplit, false, float>;
        GemmWork ws = w; ws.tiles_per_unit = tpu_split();
        typename EpiS::Params eps{ep.labels, ep.part, ep.row_cap, ep.v_local, ep.v_offset, store_rows > 0 ? zp32 : nullptr, L.ldz, store_rows};
        return launch_split<KMAJOR, KMAJOR, EpiS>(tmHp, tmHpS, tmW128, tmW128S, tmW64, tmW64S, ws, eps, "fwd_pair[split]");
      }
    }
    return launch_single<Epi, KMAJOR, KMAJOR>(st, d, tmHp, tmW256, tmW128, w, ep, "fwd_pair");
  }
  // policy + frozen logits of every reference-sample row side by side: statistics incl. the KL numerator
  template <bool ROUND> int fwd_ref() {
    using Epi = EpiStatsRef<kBnRef, ROUND>;
    GemmWork w{meta, KIND_FWD_REF, 0, 0, 0, s.V_local, s.H, tpu_ref(), super_m().stats, hint_stats().a, hint_stats().b};
    set_sync(w, ksync_stats(L.tpu));
    typename Epi::Params ep{part_ref, part_u, L.cap_ref, s.V_local};
    ProfScope ps(PK_FWD_REF, 1, st);
    if constexpr (!ROUND) {
      if (split) {
        using EpiS = EpiStatsRef<kBnSplitDual, false>;
        GemmWork ws = w; ws.tiles_per_unit = tpu_split_dual();
        typename EpiS::Params eps{part_ref, part_u, L.cap_ref, s.V_local};
        return launch_split_dual<EpiS>(ws, eps, "fwd_ref[split]");
      }
    }
    return launch_dual<Epi>(st, d, tmHp, tmHf, tmW128, tmWr128, tmW64, tmWr64, w, ep, "fwd_ref");
  }
  // recompute the pair rows of chunk ch -> dz
  template <bool ROUND> int dz_pair(int ch) {
    using Epi = EpiDzPair<kBnPair, ROUND>;
    GemmWork w{meta, KIND_DZ_PAIR, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_pair(), super_m().stats, hint_stats().a, hint_stats().b};
    set_sync(w, ksync_stats(L.tpu));
    typename Epi::Params ep{at<int>(a->state, L.lab_pair), at<float>(a->state, L.lse_pair), at<float>(a->state, L.coef_pair),
                            dz_of(ch), L.ldz, ch * L.chunk_rows, s.V_local, s.v_offset, 0};
    ProfScope ps(PK_DZ_PAIR, 1, st);
    if constexpr (!ROUND) {
      if (split) {
        using EpiS = EpiDzPair<kBnSplit, false, true>;
        GemmWork ws = w; ws.tiles_per_unit = tpu_split();
        typename EpiS::Params eps{ep.labels, ep.lse, ep.coef, dz, L.ldz, ch * L.chunk_rows, s.V_local, s.v_offset, dz_part};
        return launch_split<KMAJOR, KMAJOR, EpiS>(tmHp, tmHpS, tmW128, tmW128S, tmW64, tmW64S, ws, eps, "dz_pair[split]");
      }
    }
    return launch_single<Epi, KMAJOR, KMAJOR>(st, d, tmHp, tmW256, tmW128, w, ep, "dz_pair");
  }
  // recompute the reference-sample rows of chunk ch (both models) -> dz
  template <bool ROUND> int dz_ref(int ch) {
    using Epi = EpiDzRef<kBnRef, ROUND>;
    GemmWork w{meta, KIND_DZ_REF, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_ref(), super_m().stats, hint_stats().a, hint_stats().b};
    set_sync(w, ksync_stats(L.tpu));
    typename Epi::Params ep{meta, at<float>(a->state, L.lse_ref_p), at<float>(a->state, L.lse_ref_r), at<float>(a->state, L.coef_ref),
                            dz, L.ldz, ch * L.chunk_rows, s.V_local, 0};
    ProfScope ps(PK_DZ_REF, 1, st);
    if constexpr (!ROUND) {
      if (split) {
        using EpiS = EpiDzRef<kBnSplitDual, false, true>;
        GemmWork ws = w; ws.tiles_per_unit = tpu_split_dual();
        typename EpiS::Params eps{meta, ep.lse_p, ep.lse_r, ep.coef, dz, L.ldz, ch * L.chunk_rows, s.V_local, dz_part};
        return launch_split_dual<EpiS>(ws, eps, "dz_ref[split]");
      }
    }
    return launch_dual<Epi>(st, d, tmHp, tmHf, tmW128, tmWr128, tmW64, tmWr64, w, ep, "dz_ref");
  }
  // fused-gradient forward: policy and frozen logits of chunk ch's reference-sample rows: statistics + the bf16 logits
  int stash_ref(int ch) {
    using Epi = EpiStatsStore<256>;
    float2* part2 = reinterpret_cast<float2*>(part_ref);
    const bool given = a->ref_logits != nullptr;
    if (split) {              // fp32 inputs: the same two GEMMs on split-bf16 operands, fp32 logits kept
      using EpiS = EpiStatsStore<kBnSplit, float>;
      ProfScope ps(PK_STASH_REF, 2, st);
      GemmWork wp{meta, KIND_STASH_POL, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_split(), super_m().stats};
      typename EpiS::Params epp{meta, part2, zp32, L.cap_ref, s.V_local, L.ldz, ch * L.chunk_rows, 0};
      int rc = launch_split<KMAJOR, KMAJOR, EpiS>(tmHp, tmHpS, tmW128, tmW128S, tmW64, tmW64S, wp, epp, "stash_pol[split]");
      if (rc) return rc;
      GemmWork wf{meta, KIND_STASH_FRZ, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_split(), super_m().stats};
      typename EpiS::Params epf{meta, part2, zr32, L.cap_ref, s.V_local, L.ldz, ch * L.chunk_rows, 1};
      return launch_split<KMAJOR, KMAJOR, EpiS>(tmHf, tmHfS, tmWr128, tmWr128S, tmWr64, tmWr64S, wf, epf, "stash_frz[split]");
    }
    {
      ProfScope ps(PK_STASH_REF, given ? 1 : 2, st);
      GemmWork wp{meta, KIND_STASH_POL, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_pair(), super_m().stats, hint_stats().a, hint_stats().b};
      set_sync(wp, ksync_stats(L.tpu));
      typename Epi::Params epp{meta, part2, dz_of(ch), L.cap_ref, s.V_local, L.ldz, ch * L.chunk_rows, 0};
      int rc = launch_single<Epi, KMAJOR, KMAJOR>(st, d, tmHp, tmW256, tmW128, wp, epp, "stash_pol");
      if (rc) return rc;
      if (!given) {
        GemmWork wf{meta, KIND_STASH_FRZ, ch, L.chunk_rows, 0, s.V_local, s.H, tpu_pair(), super_m().stats, hint_stats().a, hint_stats().b};
        set_sync(wf, ksync_stats(L.tpu));
        typename Epi::Params epf{meta, part2, zr, L.cap_ref, s.V_local, L.ldz, ch * L.chunk_rows, 1};
        return launch_single<Epi, KMAJOR, KMAJOR>(st, d, tmHf, tmWr256, tmWr128, wf, epf, "stash_frz");
      }
    }
    return stash_given(ch, part2);
  }
  // the frozen model's logits are given (dpa_args_t::ref_logits): copy the chunk's rows + their statistics, no GEMM
  int stash_given(int ch, float2* part2) {
    StashGivenArgs g{meta, ch, L.chunk_rows, a->ref_logits, (long long)a->ref_logits_ld, at<int>(a->state, L.src_ref),
                     at<int>(a->state, L.act_ref), zr, L.ldz, s.V_local, part2, L.cap_ref, L.ng_pair, kBnPair * L.tpu};
    ProfScope pg(PK_STASH_GIVEN, 1, st);
    if (a->ref_logits_dtype == DPA_F32) stash_given_kernel<true><<<d->sms * 8, 256, 0, st>>>(g);
    else stash_given_kernel<false><<<d->sms * 8, 256, 0, st>>>(g);
    LAUNCH_CHECK("stash_given");
    return DPA_OK;
  }
  // dH = dz W (rows scattered to their [s,t] positions) and, if lm_head is trainable, dW += dz^T H for chunk ch
  int grad_gemms(int ch) {
    if (split) {              // fp32 inputs: split-bf16 dz and operands, fp32 gradients; dW accumulates in place
      {
        using Epi = EpiStoreRows<kBnSplit, float>;
        GemmWork w{meta, KIND_DH, ch, L.chunk_rows, 0, s.H, s.V_local, 1, super_m().dh};
        set_sync(w, ksync_split());
        typename Epi::Params ep{meta, at<int>(a->state, L.src_pair), at<int>(a->state, L.src_ref),
                                static_cast<float*>(a->dH_pair), static_cast<float*>(a->dH_ref), s.H, nullptr};
        ProfScope ps(PK_DH, 1, st);
        int rc = launch_split<KMAJOR, MNMAJOR, Epi>(tmDzK, tmDzKS, tmWmn, tmWmnS, tmWmn, tmWmnS, w, ep, "dH[split]");
        if (rc) return rc;
      }
      if (s.flags & DPA_FLAG_NEED_DW) {
        using Epi = EpiAccumF32<kBnSplit>;
        GemmWork w{meta, KIND_DW, ch, L.chunk_rows, s.V_local, s.H, 0, 1, super_m().dw};
        set_sync(w, ksync_split());
        typename Epi::Params ep{static_cast<float*>(a->dW), s.H, s.V_local, s.H, ch > 0};
        ProfScope ps(PK_DW, 1, st);
        return launch_split<MNMAJOR, MNMAJOR, Epi>(tmDzMN, tmDzMNS, tmHpMN, tmHpMNS, tmHpMN, tmHpMNS, w, ep, "dW[split]");
      }
      return DPA_OK;
    }
    {
      using P = typename EpiStoreRows<256>::Params;
      // 512-wide tiles have one accumulator stage, so their epilogue is not hidden: they pay for long K only.  K = V_local of a
      // small vocabulary shard is short (V / 8 = 4000: 63 k-blocks) -- there the double-buffered 256-wide tiles are 6 % faster
      // (tools/r2_tpu.sh: dH 0.775 -> 0.73 ms at the cfg-3 x 8 shard shape).
      const int wide = (cta_group() == 2 && grad_bn() == 512 && s.V_local > 8192) ? 1 : 0;
      // 512-wide tiles: a super-block of 9 row units x all column tiles is one wave of the 74 CTA pairs
      GemmWork w{meta, KIND_DH, ch, L.chunk_rows, 0, s.H, s.V_local, 1, super_m().dh ? super_m().dh : (wide ? 9 : 0), hint_dh().a, hint_dh().b,
                 slot_of(ch) * L.chunk_rows};
      set_sync(w, ksync());
      if (dH_packed) { int rc = before_dH_chunk(ch); if (rc) return rc; }
      P ep{meta, at<int>(a->state, L.src_pair), at<int>(a->state, L.src_ref),
           static_cast<__nv_bfloat16*>(a->dH_pair), static_cast<__nv_bfloat16*>(a->dH_ref), s.H,
           dH_packed ? reinterpret_cast<__nv_bfloat16*>(g_comm.pm.base + peer_buf_off(ch)) : nullptr, ch * L.chunk_rows};
      {
        ProfScope ps(PK_DH, 1, st);
        int rc = launch_grad<EpiStoreRowsBf16, KMAJOR, MNMAJOR>(st, d, tmDzK, tmWmn, w, ep, "dH", wide != 0);
        if (rc) return rc;
      }
      if (dH_packed) { int rc = exchange_dH_chunk(ch); if (rc) return rc; }
    }
    if (dW32) {
      using P = typename EpiAccumDW<256>::Params;
      // 512-wide tiles unless their wave count over the CTA pairs is clearly worse (a vocabulary shard: few row units of dW)
      auto waves = [&](int bn) { const long long t = (long long)((s.V_local + 255) / 256) * ((s.H + bn - 1) / bn); const int p = std::max(1, d->sms / 2);
                                 return (double)((t + p - 1) / p) * (bn / 256); };
      const int wide = (cta_group() == 2 && grad_bn() == 512 && waves(512) <= 1.08 * waves(256)) ? 1 : 0;
      GemmWork w{meta, KIND_DW, ch, L.chunk_rows, s.V_local, s.H, 0, 1, super_m().dw ? super_m().dw : (wide ? 9 : 0), hint_dw().a, hint_dw().b,
                 slot_of(ch) * L.chunk_rows};
      set_sync(w, ksync());
      P ep{meta, ch, L.chunk_rows, dW32, static_cast<__nv_bfloat16*>(a->dW), s.H, s.V_local, s.H};
      ProfScope ps(PK_DW, 1, st);
      return launch_grad<EpiAccumDW, MNMAJOR, MNMAJOR>(st, d, tmDzMN, tmHpMN, w, ep, "dW", wide != 0);
    }
    return DPA_OK;
  }
  // ---- vocab-parallel with the library's communicator: this rank's dH partial sums of chunk ch (packed rows, written by the
  //      dH GEMM into peer-memory buffer ch & 1) are all-reduced in place over NVLink and scattered, on the communicator's
  //      stream, while this stream goes on with dW and the next chunk.  Row ranges come from DpaMeta: no host count.
  size_t peer_stats_off() const { return kPeerCtrlBytes; }
  size_t peer_buf_off(int ch) const {
    return kPeerCtrlBytes + align_up((size_t)g_comm.world * L.row_stats_bytes, 256) + (size_t)(ch & 1) * L.chunk_rows * s.H * 2;
  }
  // DPA_PEER_OVERLAP=0: run the dH reduction on the caller's stream (serialised) instead of beside the dW GEMM
  static bool peer_overlap() { static const bool v = env_int("DPA_PEER_OVERLAP", 1) != 0; return v; }
  cudaEvent_t buf_free[2] = {nullptr, nullptr};      // recorded on the communicator's stream after chunk's scatter
  int before_dH_chunk(int ch) {                      // buffer ch & 1 must have been scattered (chunk ch - 2) before it is rewritten
    cudaEvent_t& e = buf_free[ch & 1];
    if (!e) return DPA_OK;
    CUDA_TRY(cudaStreamWaitEvent(st, e, 0));
    CUDA_TRY(cudaEventDestroy(e));
    e = nullptr;
    return DPA_OK;
  }
  int exchange_dH_chunk(int ch) {
    const bool overlap = peer_overlap();
    cudaStream_t cs = overlap ? g_comm.stream : st;
    if (overlap) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CUDA_TRY(cudaEventRecord(e, st));
      CUDA_TRY(cudaStreamWaitEvent(g_comm.stream, e, 0));
      CUDA_TRY(cudaEventDestroy(e));              // released once it has completed
    }
    ProfScope ps(PK_MISC, 4, cs);
    const PeerView v = peer_view();
    peer_barrier_kernel<<<1, 32, 0, cs>>>(v, 2 * (ch & 3));
    PeerReduceArgs ra{v, meta, ch, L.chunk_rows, s.H, peer_buf_off(ch)};
    // Blocks of the reduce kernel: it runs beside the dW GEMM, whose persistent CTAs own every SM -- each reduce block slows
    // the GEMM CTA it shares an SM with, and the statically scheduled GEMM is as slow as its slowest CTA.  64 blocks saturate
    // NVLink (the exchange takes the same 1.2 ms with 64, 148 or 296 blocks on 8 GPUs) and cost the GEMM 0.3 ms less
    // (cfg-3 x 8: step 4.28 -> 4.02 ms; 24 blocks no longer saturate the links).  DPA_PEER_BLOCKS overrides.
    static const int reduce_blocks = env_int("DPA_PEER_BLOCKS", 64);
    CUDA_TRY(launch_peer_reduce(ra, reduce_blocks > 0 ? reduce_blocks : d->sms * 2, cs));
    peer_barrier_kernel<<<1, 32, 0, cs>>>(v, 2 * (ch & 3) + 1);
    ScatterChunkArgs sc{meta, ch, L.chunk_rows, at<int>(a->state, L.src_pair), at<int>(a->state, L.src_ref),
                        reinterpret_cast<const __nv_bfloat16*>(g_comm.pm.base + peer_buf_off(ch)),
                        static_cast<__nv_bfloat16*>(a->dH_pair), static_cast<__nv_bfloat16*>(a->dH_ref), s.H};
    scatter_chunk_rows_kernel<<<d->sms * 4, 256, 0, cs>>>(sc);
    LAUNCH_CHECK("peer exchange");
    if (!overlap) return DPA_OK;
    CUDA_TRY(cudaEventCreateWithFlags(&buf_free[ch & 1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(buf_free[ch & 1], cs));
    return DPA_OK;
  }
  // the forward's exchange: this shard's merged statistics rows of `rr` were pushed into every peer by merge(..., push);
  // wait for all shards, then merge the G records (+ per-row results)
  int exchange_stats(RowRange rr, int finish, int slot) {
    const PeerView v = peer_view();
    {
      ProfScope ps(PK_MISC, 1, st);
      peer_barrier_kernel<<<1, 32, 0, st>>>(v, 8 + (slot & 3));
      LAUNCH_CHECK("peer barrier");
    }
    const char* g = g_comm.pm.base + peer_stats_off();
    const size_t rec = L.row_stats_bytes;
    const int G = g_comm.world;
    return run_merge(a, L, d, reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(g + (L.stats_ref - L.stats_pair)), nullptr,
                     G, G, rec / 16, rec / 16, rec / 4, st, rr, finish);
  }
  int exchange_div() {          // the divergence is the sum of the shards' parts
    const PeerView v = peer_view();
    ProfScope ps(PK_REDUCE, 3, st);
    push_div_kernel<<<1, 32, 0, st>>>(v, a->losses);
    peer_barrier_kernel<<<1, 32, 0, st>>>(v, 12);
    finish_div_kernel<<<1, 32, 0, st>>>(v, a->losses, a->alpha);
    LAUNCH_CHECK("exchange_div");
    return DPA_OK;
  }
  int join_exchange() {                          // the caller's stream waits for everything queued on the communicator's stream
    if (!peer_overlap()) return DPA_OK;
    for (int i = 0; i < 2; ++i)
      if (buf_free[i]) { CUDA_TRY(cudaEventDestroy(buf_free[i])); buf_free[i] = nullptr; }
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(e, g_comm.stream));
    CUDA_TRY(cudaStreamWaitEvent(st, e, 0));
    CUDA_TRY(cudaEventDestroy(e));
    return DPA_OK;
  }
  // collective part of enabling the library-side exchange for this call (peer memory sized for two chunk buffers)
  int enable_peer_exchange(bool with_dH = true) {
    const int bufs = L.n_chunks > 1 ? 2 : 1;
    int rc = ensure_peer_mem(align_up((size_t)g_comm.world * L.row_stats_bytes, 256) + (with_dH ? (size_t)bufs * L.chunk_rows * s.H * 2 : 0));
    if (rc) return rc;
    dH_packed = reinterpret_cast<__nv_bfloat16*>(g_comm.pm.base);     // (per-chunk buffer: see grad_gemms)
    return DPA_OK;
  }
  // (Tried: the fill on a side stream, hidden behind the statistics GEMM -- the GEMM slows down by what the fill costs,
  //  the step is bound by power / HBM energy, not by idle tensor cores.  profiles/r02a.)
  int zero_grad_outputs() {
    const size_t es = split ? 4 : 2;
    const size_t n_pair = (size_t)2 * s.B * s.T * s.H * es, n_ref = (size_t)s.B_ref * s.T_ref * s.H * es;   // (H % 64 == 0: 16-byte multiples)
    static const bool use_kernel = env_int("DPA_ZERO_KERNEL", 1) != 0;        // 0: cudaMemsetAsync (A/B)
    ProfScope ps(PK_MISC, use_kernel ? 1 : 0, st);
    if (use_kernel) {
      ZeroSegs z{{a->dH_pair, a->dH_ref}, {n_pair, n_ref}};
      zero_fill_kernel<<<d->sms * 4, 512, 0, st>>>(z);
      LAUNCH_CHECK("zero_fill");
      return DPA_OK;
    }
    CUDA_TRY(cudaMemsetAsync(a->dH_pair, 0, n_pair, st));
    CUDA_TRY(cudaMemsetAsync(a->dH_ref, 0, n_ref, st));
    return DPA_OK;      // (the dW accumulator needs no memset: the first chunk's epilogue overwrites it)
  }
  int finish_dW() {
    if (!(s.flags & DPA_FLAG_NEED_DW) || !a->dW) return DPA_OK;
    ProfScope ps(PK_MISC, 1, st);      // the last chunk with rows wrote dW; only the no-rows-at-all case is left
    zero_if_no_rows_kernel<<<d->sms * 2, 256, 0, st>>>(meta, a->dW, (size_t)s.V_local * s.H * (split ? 4 : 2));
    LAUNCH_CHECK("dW zero-if-empty");
    return DPA_OK;
  }
  int coef(const float* d_loss, bool remember = false) {
    CoefArgs c{meta, at<int>(a->state, L.col_pair), at<int>(a->state, L.act_ref), at<float>(a->state, L.sigma), d_loss,
               at<float>(a->state, L.coef_pair), at<float>(a->state, L.coef_ref), s.B, s.B_ref, a->alpha,
               remember ? at<float>(a->state, L.formed_for) : nullptr};
    ProfScope ps(PK_COEF, 1, st);
    coef_kernel<<<d->sms * 2, 256, 0, st>>>(c);
    LAUNCH_CHECK("coef");
    return DPA_OK;
  }
  int merge(RowRange rr, bool with_u, int finish = MERGE_ONLY, bool push = false) {
    const PeerView v = push ? peer_view() : PeerView{};
    return run_merge(a, L, d, part_pair, part_ref, with_u ? part_u : nullptr, L.ng_pair, L.ng_ref, L.cap_pair, L.cap_ref, L.cap_ref, st, rr, finish,
                     push ? &v : nullptr, peer_stats_off());
  }
};

// ---------------------------------------------------------------------------------- statistics-only forward
int forward_stats_impl(const dpa_args_t* a, cudaStream_t st, Layout& L, DeviceInfo** dev, int finish = MERGE_ONLY) {
  int rc = check_common(a, L, dev);
  if (rc) return rc;
  DeviceInfo* d = *dev;
  if (a->ref_logits) return fail(DPA_ERR_ARG, "ref_logits is only available under DPA_FLAG_FUSED_GRAD");
  if (!a->scratch || a->scratch_bytes < L.fwd_total) return fail(DPA_ERR_WORKSPACE, "forward scratch: need %zu bytes, got %zu", L.fwd_total, a->scratch_bytes);
  if (!a->losses || !a->status) return fail(DPA_ERR_ARG, "losses/status is NULL");
  if ((rc = run_meta(a, L, d, st))) return rc;
  const Layout::Split* sp = L.split ? &L.f_split : nullptr;
  if ((rc = run_gather(a, L, d, a->scratch, L.f_Hp, L.f_Hf, st))) return rc;
  if (sp && (rc = run_split_weights(a, L, d, *sp, st))) return rc;
  Plan p(a, L, d, st);
  if ((rc = p.init(L.f_Hp, L.f_Hf, L.f_part_pair, L.f_part_ref, L.f_part_u, true, nullptr, nullptr, nullptr, sp))) return rc;
  const bool round = (a->shape.flags & DPA_FLAG_ROUND_LOGITS) != 0;
  if ((rc = round ? p.fwd_pair<true>(0) : p.fwd_pair<false>(0))) return rc;
  if ((rc = round ? p.fwd_ref<true>() : p.fwd_ref<false>())) return rc;
  return p.merge(RowRange{SEG_ALL, 0, 0}, true, finish);
}

// ---------------------------------------------------------------------------------- fused-gradient forward
// Loss AND gradients (for d_loss = 1) in one pass; dpa_backward then only scales.  Pair rows: the statistics GEMM keeps
// the bf16 logits of the pair rows of row chunk 0 and an elementwise pass turns them into dz once sigma is known (pair
// rows beyond chunk 0 are recomputed).  Reference-sample rows, chunk by chunk: two GEMMs produce the statistics and
// keep the chunk's bf16 logits of both models, an elementwise pass turns them into dz and the KL rows.  Every chunk
// ends with dH = dz W and dW += dz^T H.  Executed GEMM units per row = the algorithmic count (pair 3, reference 4).
int forward_fused_impl(const dpa_args_t* a, cudaStream_t st) {
  Layout L; DeviceInfo* d = nullptr;
  int rc = check_common(a, L, &d);
  if (rc) return rc;
  const dpa_shape_t& s = a->shape;
  // the kept logits are the ones the reference's softmax sees: bf16-rounded for a bf16 model (DPA_FLAG_ROUND_LOGITS), fp32 for fp32 inputs
  if (!L.split && !(s.flags & DPA_FLAG_ROUND_LOGITS)) return fail(DPA_ERR_ARG, "DPA_FLAG_FUSED_GRAD needs DPA_FLAG_ROUND_LOGITS (the kept logits are bf16)");
  if (L.split && a->ref_logits) return fail(DPA_ERR_ARG, "fp32 inputs: ref_logits belongs to the bf16 execution");
  if (!a->scratch || a->scratch_bytes < L.fused_total) return fail(DPA_ERR_WORKSPACE, "fused scratch: need %zu bytes, got %zu", L.fused_total, a->scratch_bytes);
  if (!a->losses || !a->status) return fail(DPA_ERR_ARG, "losses/status is NULL");
  const bool loss_only = (s.flags & DPA_FLAG_LOSS_ONLY) != 0;      // evaluation: same chunked pass, no gradient GEMMs
  if (!loss_only && (rc = check_grad_outputs(a))) return rc;
  // Vocab-parallel lm_head with the library's communicator: every rank runs this pass on the same rows and its own lm_head
  // rows; the per-row statistics of the shards are exchanged (all-gather of the 36 B/row record + max-rescale merge) once for
  // the pair segment and once per row chunk, the dH partial sums through peer memory (dpa_peer.cuh), dW stays sharded.
  const bool vp = comm_active(s);
  if (s.V_local < s.V && !vp) return fail(DPA_ERR_ARG, "DPA_FLAG_FUSED_GRAD with a vocab-sharded lm_head needs the library's communicator (dpa_comm_init)");
  if (vp && (s.flags & DPA_FLAG_FUSED_NORM)) return fail(DPA_ERR_ARG, "DPA_FLAG_FUSED_NORM is not available with a vocab-sharded lm_head");
  if (vp && a->ref_logits) return fail(DPA_ERR_ARG, "ref_logits is not available with a vocab-sharded lm_head");
  Plan p(a, L, d, st);
  if (vp && (rc = p.enable_peer_exchange(!loss_only))) return rc;
  int n_exchange = 0;
  const Layout::Split* sp = L.split ? &L.u_split : nullptr;
  if ((rc = p.init(L.u_Hp, L.u_Hf, L.u_part_pair, L.u_part_ref, L.u_part_u, true, at<__nv_bfloat16>(a->scratch, L.u_dz),
                   at<__nv_bfloat16>(a->scratch, L.u_zr),
                   ((s.flags & DPA_FLAG_NEED_DW) && !loss_only && !sp) ? at<float>(a->scratch, L.u_dW32) : nullptr, sp, L.n_slots))) return rc;
  if (sp) { p.zp32 = at<float>(a->scratch, L.u_zp32); p.zr32 = at<float>(a->scratch, L.u_zr); }

  if ((rc = run_meta(a, L, d, st))) return rc;
  if ((rc = run_gather(a, L, d, a->scratch, L.u_Hp, L.u_Hf, st))) return rc;
  if (sp && (rc = run_split_weights(a, L, d, *sp, st))) return rc;
  if (!loss_only && (rc = p.zero_grad_outputs())) return rc;

  // ---- alignment half: statistics of every pair row -> P, N -> sigma, alignment loss -> row coefficients
  // (the statistics GEMM keeps the logits of the pair rows of every slot; nothing else touches a slot until its chunk's turn)
  const int keep_rows = loss_only ? 0 : L.n_slots * L.chunk_rows;
  if ((rc = sp ? p.fwd_pair<false>(keep_rows) : p.fwd_pair<true>(keep_rows))) return rc;
  if ((rc = p.merge(RowRange{SEG_PAIR, 0, 0}, true, vp ? MERGE_ONLY : MERGE_FINISH, vp))) return rc;
  if (vp && (rc = p.exchange_stats(RowRange{SEG_PAIR, 0, 0}, MERGE_FINISH, n_exchange++))) return rc;
  if ((rc = run_losses(a, L, st, LOSS_ALIGN, 0))) return rc;
  // gradients are formed for d_loss = 1, or for the caller's hint of the upstream gradient (DPA_FLAG_GRAD_HINT: *d_loss now)
  const float* hint = (s.flags & DPA_FLAG_GRAD_HINT) ? a->d_loss : nullptr;
  if ((s.flags & DPA_FLAG_GRAD_HINT) && !hint) return fail(DPA_ERR_ARG, "DPA_FLAG_GRAD_HINT set but d_loss is NULL at dpa_forward");
  if ((rc = p.coef(hint, true))) return rc;

  // ---- row chunks of the packed space [pair rows | reference-sample rows]
  for (int ch = 0; ch < L.n_chunks; ++ch) {
    if (loss_only) {
      // no dz of the pair rows is needed
    } else if ((long long)ch * L.chunk_rows >= L.cap_pair) {
      // past the pair segment
    } else if (ch < L.n_slots && sp) {                            // kept fp32 logits -> split-bf16 dLogits
      DzPairElemF32Args e{p.meta, ch * L.chunk_rows, (ch + 1) * L.chunk_rows, p.zp32, p.dz, p.dz_part, L.ldz, s.V_local, s.v_offset,
                          at<int>(a->state, L.lab_pair), at<float>(a->state, L.lse_pair), at<float>(a->state, L.coef_pair)};
      ProfScope ps(PK_DZ_ELEM, 1, st);
      dz_pair_elem_f32_kernel<<<d->sms * 8, 256, 0, st>>>(e);
      LAUNCH_CHECK("dz_pair_elem_f32");
    } else if (ch < L.n_slots) {                                  // kept logits -> dLogits in place
      DzPairElemArgs e{p.meta, ch * L.chunk_rows, (ch + 1) * L.chunk_rows, p.dz, L.ldz, s.V_local, s.v_offset, at<int>(a->state, L.lab_pair),
                       at<float>(a->state, L.lse_pair), at<float>(a->state, L.coef_pair)};
      ProfScope ps(PK_DZ_ELEM, 1, st);
      dz_pair_elem_kernel<<<d->sms * 8, 256, 0, st>>>(e);
      LAUNCH_CHECK("dz_pair_elem");
    } else {                                                       // pair rows beyond the kept slots (if any): recompute
      if ((rc = sp ? p.dz_pair<false>(ch) : p.dz_pair<true>(ch))) return rc;
    }
    if ((rc = p.stash_ref(ch))) return rc;
    const RowRange rr{SEG_REF_CHUNK, ch, L.chunk_rows};
    if ((rc = p.merge(rr, false, vp ? MERGE_ONLY : MERGE_FINISH_KL_LATER, vp))) return rc;
    if (vp && (rc = p.exchange_stats(rr, MERGE_FINISH_KL_LATER, n_exchange++))) return rc;
    if (sp) {
      DzElemF32Args e{p.meta, ch, L.chunk_rows, p.zp32, p.zr32, p.dz, p.dz_part, L.ldz, s.V_local, at<float>(a->state, L.lse_ref_p),
                      at<float>(a->state, L.lse_ref_r), at<float>(a->state, L.coef_ref), at<int>(a->state, L.act_ref), at<float>(a->state, L.kl_row)};
      ProfScope ps(PK_DZ_ELEM, 1, st);
      dz_kl_elem_f32_kernel<<<d->sms * 8, 256, 0, st>>>(e);
      LAUNCH_CHECK("dz_kl_elem_f32");
    } else {
      DzElemArgs e{p.meta, ch, L.chunk_rows, p.dz_of(ch), p.zr, L.ldz, s.V_local, at<float>(a->state, L.lse_ref_p), at<float>(a->state, L.lse_ref_r),
                   at<float>(a->state, L.coef_ref), at<int>(a->state, L.act_ref), at<float>(a->state, L.kl_row),
                   (!vp || g_comm.rank == 0) ? 1 : 0};
      ProfScope ps(PK_DZ_ELEM, 1, st);
      dz_kl_elem_kernel<<<d->sms * 8, 256, 0, st>>>(e);
      LAUNCH_CHECK("dz_kl_elem");
    }
    if (!loss_only && (rc = p.grad_gemms(ch))) return rc;
  }
  if ((rc = run_losses(a, L, st, LOSS_TOTAL, 1))) return rc;
  if (vp && (rc = p.exchange_div())) return rc;
  if (loss_only) return DPA_OK;
  if (vp && (rc = p.join_exchange())) return rc;
  if ((rc = run_norm_backward(a, L, d, st))) return rc;
  return p.finish_dW();
}

// fused-gradient backward: scale what the forward left in dH_pair / dH_ref / dW by the upstream gradient
int backward_scale_impl(const dpa_args_t* a, cudaStream_t st) {
  Layout L; DeviceInfo* d = nullptr;
  int rc = check_common(a, L, &d);
  if (rc) return rc;
  if (a->shape.flags & DPA_FLAG_LOSS_ONLY) return fail(DPA_ERR_ARG, "DPA_FLAG_LOSS_ONLY: the forward formed no gradients");
  if ((rc = check_grad_outputs(a))) return rc;
  const dpa_shape_t& s = a->shape;
  if (!a->d_loss && !(s.flags & DPA_FLAG_GRAD_HINT)) return DPA_OK;       // NULL = 1, and they were formed for 1
  if (!a->d_loss) return fail(DPA_ERR_ARG, "DPA_FLAG_GRAD_HINT: dpa_backward needs the actual d_loss");
  ProfScope ps(PK_SCALE, 1, st);
  ScaleSegs sg{};
  auto add = [&](void* p, size_t n) { sg.p[sg.count] = p; sg.n[sg.count] = n; ++sg.count; };
  add(a->dH_pair, (size_t)2 * s.B * s.T * s.H);
  add(a->dH_ref, (size_t)s.B_ref * s.T_ref * s.H);
  if (s.flags & DPA_FLAG_NEED_DW) add(a->dW, (size_t)s.V_local * s.H);
  if ((s.flags & DPA_FLAG_FUSED_NORM) && a->d_norm_w) add(a->d_norm_w, (size_t)s.H);
  const float* formed_for = at<float>(a->state, L.formed_for);
  if (L.split) scale_grads_kernel<float><<<d->sms * 8, 256, 0, st>>>(sg, a->d_loss, formed_for);
  else scale_grads_kernel<__nv_bfloat16><<<d->sms * 8, 256, 0, st>>>(sg, a->d_loss, formed_for);
  LAUNCH_CHECK("scale_grads");
  return DPA_OK;
}

// ---------------------------------------------------------------------------------- recompute backward
// The north star's wording: the forward kept O(rows) statistics; the logits are recomputed chunk-wise into dz, then dH / dW.
template <bool ROUND>
int run_backward(const dpa_args_t* a, const Layout& L, const DeviceInfo* d, cudaStream_t st) {
  const dpa_shape_t& s = a->shape;
  int rc;
  Plan p(a, L, d, st);
  const Layout::Split* sp = L.split ? &L.b_split : nullptr;
  if ((rc = p.init(L.b_Hp, L.b_Hf, 0, 0, 0, false, at<__nv_bfloat16>(a->scratch, L.b_dz), nullptr,
                   ((s.flags & DPA_FLAG_NEED_DW) && !sp) ? at<float>(a->scratch, L.b_dW32) : nullptr, sp))) return rc;
  // Vocab-parallel with the library's own communicator: the dH partial sums of this rank's vocabulary shard stay packed
  // (active rows only -- a quarter of the dense tensors at 10 % phrase density), are all-reduced chunk by chunk through
  // peer memory (dpa_peer.cuh; row counts read from DpaMeta on the device, nothing is read back) and scattered.
  const bool vp = comm_active(s);
  if (vp && (rc = p.enable_peer_exchange())) return rc;
  CUDA_TRY(cudaMemsetAsync(at<char>(a->state, L.sync_ctr), 0, (size_t)kSyncCtrs(L.n_chunks) * 4, st));   // (a retained graph may run twice)
  if ((rc = p.coef(a->d_loss))) return rc;
  if ((rc = run_gather(a, L, d, a->scratch, L.b_Hp, L.b_Hf, st))) return rc;
  if (sp && (rc = run_split_weights(a, L, d, *sp, st))) return rc;
  if ((rc = p.zero_grad_outputs())) return rc;
  for (int ch = 0; ch < L.n_chunks; ++ch) {
    if ((long long)ch * L.chunk_rows < L.cap_pair && (rc = p.dz_pair<ROUND>(ch))) return rc;
    if ((rc = p.dz_ref<ROUND>(ch))) return rc;
    if ((rc = p.grad_gemms(ch))) return rc;       // (vp: the chunk's dH rows go to the exchange stream right after the dH GEMM)
  }
  if (vp && (rc = p.join_exchange())) return rc;
  if ((rc = run_norm_backward(a, L, d, st))) return rc;
  return p.finish_dW();
}

}  // namespace

extern "C" {

int dpa_version(void) { return DPA_ABI_VERSION; }
const char* dpa_last_error(void) { return g_err; }

int dpa_query_workspace(const dpa_shape_t* shape, dpa_workspace_t* out) {
  if (!shape || !out) return fail(DPA_ERR_ARG, "NULL argument");
  Layout L;
  int rc = make_layout(*shape, L);
  if (rc) return rc;
  out->state_bytes = L.state_total; out->fwd_bytes = L.fwd_total; out->bwd_bytes = L.bwd_total; out->fused_bytes = L.fused_total;
  return DPA_OK;
}

int dpa_struct_sizes(int* shape_bytes, int* workspace_bytes, int* args_bytes) {
  if (shape_bytes) *shape_bytes = (int)sizeof(dpa_shape_t);
  if (workspace_bytes) *workspace_bytes = (int)sizeof(dpa_workspace_t);
  if (args_bytes) *args_bytes = (int)sizeof(dpa_args_t);
  return DPA_OK;
}

int dpa_forward_stats(const dpa_args_t* a, void* stream) {
  Layout L; DeviceInfo* d = nullptr;
  return forward_stats_impl(a, static_cast<cudaStream_t>(stream), L, &d);
}

int dpa_row_stats_view(const dpa_args_t* a, void** ptr, size_t* bytes) {
  if (!a || !ptr || !bytes) return fail(DPA_ERR_ARG, "NULL argument");
  Layout L;
  int rc = make_layout(a->shape, L);
  if (rc) return rc;
  *ptr = at<char>(a->state, L.stats_pair); *bytes = L.row_stats_bytes;
  return DPA_OK;
}

int dpa_forward_finish(const dpa_args_t* a, const void* gathered, int n_shards, void* stream) {
  Layout L; DeviceInfo* d = nullptr;
  int rc = check_common(a, L, &d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gathered && n_shards > 1) {
    // records of all vocab shards, each laid out like this rank's [stats_pair | stats_ref | stats_u]
    const char* g = static_cast<const char*>(gathered);
    const size_t rec = L.row_stats_bytes;
    if (rec % 16) return fail(DPA_ERR_BAD_SHAPE, "row-stats record not 16-byte sized");
    rc = run_merge(a, L, d, reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(g + (L.stats_ref - L.stats_pair)),
                   reinterpret_cast<const float*>(g + (L.stats_u - L.stats_pair)), n_shards, n_shards, rec / 16, rec / 16, rec / 4, st,
                   RowRange{SEG_ALL, 0, 0}, MERGE_FINISH);
    if (rc) return rc;
    return run_losses(a, L, st, LOSS_ALIGN | LOSS_TOTAL, 0);
  }
  return run_finish(a, L, d, st);
}

int dpa_forward(const dpa_args_t* a, void* stream) {
  Layout L; DeviceInfo* d = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a && (a->shape.flags & DPA_FLAG_FUSED_GRAD)) {
    return forward_fused_impl(a, st);
  }
  const bool exchange = a && comm_active(a->shape);
  int rc = forward_stats_impl(a, st, L, &d, exchange ? MERGE_ONLY : MERGE_FINISH);
  if (rc) return rc;
  if (exchange) {
    // the one exchange step of the forward: every shard's per-row statistics record -> all ranks -> max-rescale merge
    if (a->shape.flags & DPA_FLAG_FUSED_NORM) return fail(DPA_ERR_ARG, "DPA_FLAG_FUSED_NORM is not available with a vocab-sharded lm_head");
    char* g = at<char>(a->scratch, L.f_gather);
    const size_t rec = L.row_stats_bytes;
    NCCL_TRY(g_nccl.AllGather(at<char>(a->state, L.stats_pair), g, rec, kNcclUint8, g_comm.comm, st));
    const int G = g_comm.world;
    rc = run_merge(a, L, d, reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(g + (L.stats_ref - L.stats_pair)),
                   reinterpret_cast<const float*>(g + (L.stats_u - L.stats_pair)), G, G, rec / 16, rec / 16, rec / 4, st,
                   RowRange{SEG_ALL, 0, 0}, MERGE_FINISH);
    if (rc) return rc;
  }
  return run_losses(a, L, st, LOSS_ALIGN | LOSS_TOTAL, 0);
}

int dpa_backward(const dpa_args_t* a, void* stream) {
  Layout L; DeviceInfo* d = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a && (a->shape.flags & DPA_FLAG_FUSED_GRAD)) return backward_scale_impl(a, st);
  int rc = check_common(a, L, &d);
  if (rc) return rc;
  if (!a->scratch || a->scratch_bytes < L.bwd_total) return fail(DPA_ERR_WORKSPACE, "backward scratch: need %zu bytes, got %zu", L.bwd_total, a->scratch_bytes);
  if ((rc = check_grad_outputs(a))) return rc;
  rc = (a->shape.flags & DPA_FLAG_ROUND_LOGITS) ? run_backward<true>(a, L, d, st) : run_backward<false>(a, L, d, st);
  if (rc) return rc;
  return DPA_OK;     // (vocab-parallel: the dH exchange already happened inside run_backward, on the packed rows)
}

int dpa_comm_unique_id(void* out128) {
  if (!out128) return fail(DPA_ERR_ARG, "NULL argument");
  int rc = nccl_bind();
  if (rc) return rc;
  NCCL_TRY(g_nccl.GetUniqueId(static_cast<NcclId*>(out128)));
  return DPA_OK;
}

int dpa_comm_init(int rank, int world, const void* unique_id128) {
  if (!unique_id128 || world < 1 || rank < 0 || rank >= world) return fail(DPA_ERR_ARG, "bad rank / world / unique id");
  if (g_comm.comm) return fail(DPA_ERR_ARG, "communicator already initialised (call dpa_comm_destroy first)");
  int rc = nccl_bind();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, unique_id128, sizeof(id));
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  NcclComm c = nullptr;
  NCCL_TRY(g_nccl.CommInitRank(&c, world, id, rank));
  g_comm.comm = c; g_comm.rank = rank; g_comm.world = world; g_comm.device = dev;
  CUDA_TRY(cudaStreamCreateWithFlags(&g_comm.stream, cudaStreamNonBlocking));
  g_layout_comm_world = world;
  return DPA_OK;
}

int dpa_comm_destroy(void) {
  if (g_comm.stream) { cudaStreamSynchronize(g_comm.stream); }
  if (g_comm.comm) { cudaDeviceSynchronize(); peer_mem_release(); }
  if (g_comm.stream) cudaStreamDestroy(g_comm.stream);
  if (g_comm.comm) NCCL_TRY(g_nccl.CommDestroy(g_comm.comm));
  g_comm = Comm{};
  g_layout_comm_world = 1;
  return DPA_OK;
}

int dpa_comm_world(void) { return g_comm.comm ? g_comm.world : 1; }

int dpa_set_cta_group(int cg) {
  if (cg != 1 && cg != 2) return fail(DPA_ERR_ARG, "cta_group must be 1 or 2");
  g_cta_group.store(cg);
  return DPA_OK;
}

// Probe (profiles/): how many clusters of `cluster_size` CTAs of the 256-wide GEMM kernel (225 KB of shared memory per CTA)
// can be co-resident on this device -- the cost side of a multicast design with clusters of 4.
int dpa_debug_max_clusters(int cluster_size) {
  DeviceInfo* d = nullptr;
  int rc = device_info(&d);
  if (rc) return rc;
  using Cfg = GemmCfg<256, false, KMAJOR, KMAJOR, 2>;
  auto kern = dpa_gemm_kernel<Cfg, EpiAccumF32<256>>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (cluster_size > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(d->sms / cluster_size * cluster_size));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  if (e != cudaSuccess) return fail(DPA_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e));
  return n;
}

int dpa_tmap_cache_stats(long long* hits, long long* misses) {
  if (hits) *hits = g_tmap_hits.load();
  if (misses) *misses = g_tmap_misses.load();
  return DPA_OK;
}

int dpa_profile_enable(int on) {
  g_prof_on.store(on != 0);
  return DPA_OK;
}

// Synchronises on the recorded events. ms[k] / launches[k] are summed per kernel class (see dpa_profile_kind_name).
int dpa_profile_collect(double* ms, int* launches, int n) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (auto& sp : g_spans) {
    float t = 0.f;
    cudaEventSynchronize(sp.b);
    cudaEventElapsedTime(&t, sp.a, sp.b);
    if (sp.kind < n) { ms[sp.kind] += t; launches[sp.kind] += sp.launches; }
    g_event_pool.push_back(sp.a); g_event_pool.push_back(sp.b);
  }
  g_spans.clear();
  return PK_COUNT;
}

const char* dpa_profile_kind_name(int k) {
  static const char* names[PK_COUNT] = {"meta", "gather", "fwd_pair", "fwd_ref", "reduce", "coef", "dz_pair", "dz_ref", "dH", "dW", "misc",
                                        "stash_ref", "dz_elem", "scale", "stash_given"};
  return (k >= 0 && k < PK_COUNT) ? names[k] : "";
}

int dpa_debug_gemm(const void* A, const void* B, float* D, int M, int N, int K, int a_major, int b_major, int bn, void* stream) {
  DeviceInfo* d = nullptr;
  int rc = device_info(&d);
  if (rc) return rc;
  if (N % 32 || K % 8 || M % 8) return fail(DPA_ERR_BAD_SHAPE, "debug gemm: N%%32, K%%8, M%%8 required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmA, tmB;
  if ((rc = make_operand_tmap(d, &tmA, A, a_major, M, K, a_major == KMAJOR ? K : M, BM))) return rc;
  if ((rc = make_operand_tmap(d, &tmB, B, b_major, N, K, b_major == KMAJOR ? K : N, bn > 256 ? 256 : bn))) return rc;
  GemmWork w{nullptr, KIND_STATIC, 0, 0, M, N, K, 1, 0};
  if (bn == 128 && a_major == KMAJOR && b_major == KMAJOR) {     // the tile shape of the dual-accumulator kernels
    using Epi = EpiAccumF32<128>;
    typename Epi::Params ep{D, N, M, N, 0};
    return launch_gemm<GemmCfg<128, false, KMAJOR, KMAJOR, 1>, Epi>(st, d->sms, tmA, tmB, tmA, tmB, w, ep, "debug_gemm");
  }
  if (bn == 512) {                                                 // the 256 x 512 tile of the gradient GEMMs (CTA pairs)
    CUtensorMap tmB5;                                              // K-major B: boxes of 128 rows (one per MMA and CTA)
    if ((rc = make_operand_tmap(d, &tmB5, B, b_major, N, K, b_major == KMAJOR ? K : N, 128))) return rc;
    using Epi = EpiAccumF32<512>;
    typename Epi::Params ep{D, N, M, N, 0};
#define DBG_CASE5(AM, BMJ) \
    if (a_major == AM && b_major == BMJ) return launch_gemm<GemmCfg<512, false, AM, BMJ, 2>, Epi>(st, d->sms, tmA, tmB5, tmA, tmB5, w, ep, "debug_gemm512");
    DBG_CASE5(KMAJOR, KMAJOR)
    DBG_CASE5(KMAJOR, MNMAJOR)
    DBG_CASE5(MNMAJOR, MNMAJOR)
    DBG_CASE5(MNMAJOR, KMAJOR)
#undef DBG_CASE5
  }
  CUtensorMap tmBh;                                                // B box of BN / 2 rows for cta_group 2
  if ((rc = make_operand_tmap(d, &tmBh, B, b_major, N, K, b_major == KMAJOR ? K : N, bn / 2))) return rc;
#define DBG_CASE(AM, BMJ)                                                                             \
  if (bn == 256 && a_major == AM && b_major == BMJ) {                                                 \
    using Epi = EpiAccumF32<256>;                                                                     \
    typename Epi::Params ep{D, N, M, N, 0};                                                           \
    return launch_single<Epi, AM, BMJ>(st, d, tmA, tmB, tmBh, w, ep, "debug_gemm");                   \
  }
  DBG_CASE(KMAJOR, KMAJOR)
  DBG_CASE(KMAJOR, MNMAJOR)
  DBG_CASE(MNMAJOR, MNMAJOR)
  DBG_CASE(MNMAJOR, KMAJOR)
#undef DBG_CASE
  return fail(DPA_ERR_ARG, "debug gemm: unsupported (bn=%d, a_major=%d, b_major=%d)", bn, a_major, b_major);
}

}  // extern "C"
