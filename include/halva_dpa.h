/* libhalva_dpa.so -- C ABI of the B200-native DPA-loss path.
 *
 * Replaces, behind one call pair, what HalvaTrainer.compute_loss does after the decoder has produced the
 * final hidden states (citations into the reference, pritamqu/HALVA):
 *   lm_head projection + fp32 upcast        vila/train/transformers_replace/models/llama/modeling_llama.py:1022-1029
 *   cal_batch_logp                          llava/train/halva_trainer.py:392-409
 *   accumulate_logps                        llava/train/halva_trainer.py:411-419
 *   concatenated_forward (shift, split)     llava/train/halva_trainer.py:475-489
 *   reference_forward (shift)               llava/train/halva_trainer.py:522-531
 *   compute_loss (alignment + divergence)   llava/train/halva_trainer.py:550-592
 *   and the autograd backward of all of the above (dHidden, optional dW).
 * The reference has no FFI layer; its seam is the Python override compute_loss(model, inputs)
 * (llava/train/halva_trainer.py:534).  INTEGRATION.md shows the ctypes binding and the trainer patch.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns all memory (inputs, outputs, workspaces); the library allocates nothing persistent
 *   - all calls are asynchronous on `stream` and never synchronise the stream or the device on the step path
 *   - return value 0 = OK, negative = dpa_status_t; message via dpa_last_error() (thread local)
 *   - data-dependent errors (bad phrase ids, K_pos != K_neg, K == 0) cannot be seen by the host without a
 *     sync: they are reported in status[0] (device) and turn the losses into NaN, which is also what the
 *     reference returns for K == 0 (it raises RuntimeError for K_pos != K_neg).
 */
#ifndef HALVA_DPA_H
#define HALVA_DPA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPA_ABI_VERSION 5

typedef enum {
  DPA_OK = 0,
  DPA_ERR_BAD_SHAPE = -1,   /* dims not supported (H % 64, alignment of V_local, ...) */
  DPA_ERR_DTYPE = -2,
  DPA_ERR_ALIGN = -3,       /* a pointer is not 16-byte aligned (TMA) */
  DPA_ERR_WORKSPACE = -4,   /* workspace too small / NULL */
  DPA_ERR_CUDA = -5,
  DPA_ERR_NCCL = -6,
  DPA_ERR_ARG = -7,
  DPA_ERR_ARCH = -8         /* not an sm_100 device */
} dpa_status_t;

/* bits of status[0] (device int), written by dpa_forward */
#define DPA_STATUS_BAD_SIGN   1   /* a phrase id outside [0, 255] (after -100 -> 0) */
#define DPA_STATUS_K_MISMATCH 2   /* K_pos != K_neg: the reference raises RuntimeError (or silently broadcasts) */
#define DPA_STATUS_K_ZERO     4   /* no phrase at all: the reference returns NaN */
#define DPA_STATUS_BAD_LABEL  8   /* an active label outside [0, V) */

typedef enum { DPA_BF16 = 0, DPA_F32 = 1 } dpa_dtype_t;   /* DPA_F32: ref_logits only */
typedef enum { DPA_I64 = 0, DPA_I32 = 1 } dpa_index_t;

#define DPA_FLAG_ROUND_LOGITS 1   /* round logits to bf16 before the fp32 softmax, as nn.Linear does in bf16 */
#define DPA_FLAG_NEED_DW      2   /* lm_head is trainable: backward also produces dW */
#define DPA_FLAG_DENSE        4   /* do not compact: run every fed position (what the reference executes) */
#define DPA_FLAG_FUSED_GRAD   8   /* dpa_forward also forms dH (+dW) for d_loss = 1 (reference-sample logits are kept in
                                     bf16 per row chunk instead of being recomputed); dpa_backward then only scales by
                                     *d_loss.  Needs DPA_FLAG_ROUND_LOGITS, the gradient pointers at forward time and
                                     `fused_bytes` of scratch. */

#define DPA_FLAG_FUSED_NORM   16  /* h_pair / h_ref_pol / h_ref_frz are the PRE-norm residual stream: the decoder's final
                                     RMSNorm (llava/model/language_model/modelling_llama.py:55-70: fp32 variance, cast back,
                                     weight multiply) is applied while the active rows are gathered, and its backward to
                                     the dH rows.  Needs norm_w / norm_w_ref / norm_eps. */
#define DPA_FLAG_LOSS_ONLY    32  /* with DPA_FLAG_FUSED_GRAD: run the chunked forward for the losses only (evaluation,
                                     torch.no_grad()): no gradient buffers needed, no gradient GEMMs */

#define DPA_FLAG_GRAD_HINT    64  /* with DPA_FLAG_FUSED_GRAD: `d_loss` is already valid at dpa_forward and holds the upstream
                                     gradient the caller EXPECTS (1 / gradient-accumulation steps under HF Trainer /
                                     DeepSpeed).  The gradients are formed for that value, and dpa_backward -- given the
                                     actual *d_loss -- rescales them only if the two differ: with a correct hint the
                                     backward touches no gradient byte.  Without the flag they are formed for 1. */

typedef struct {
  int32_t B;          /* sample pairs; h_pair holds 2B sequences, first B "pos"(correct), last B "neg"(hallucinated) */
  int32_t T;          /* pair sequence length (post image splice) */
  int32_t B_ref;      /* reference samples */
  int32_t T_ref;
  int32_t H;          /* hidden size, multiple of 64 */
  int32_t V;          /* global vocabulary */
  int32_t V_local;    /* rows of W held by this rank (== V without vocab parallelism), multiple of 8 */
  int32_t v_offset;   /* first global vocabulary id of this rank's shard */
  int32_t dtype;      /* dpa_dtype_t of h_*, W, W_ref, dH_*, dW */
  int32_t index_dtype;/* dpa_index_t of labels / signs */
  int32_t flags;      /* DPA_FLAG_* */
  int32_t chunk_rows; /* row chunk of the packed row space (multiple of 256); 0 = default (32768; 65536 for batches
                         of more than 131072 positions; fp32 inputs 8192; times V / V_local for a vocabulary shard) */
} dpa_shape_t;

typedef struct {
  size_t state_bytes;   /* kept alive by the caller from dpa_forward until dpa_backward returns */
  size_t fwd_bytes;     /* scratch, dpa_forward only */
  size_t bwd_bytes;     /* scratch, dpa_backward only */
  size_t fused_bytes;   /* scratch of dpa_forward under DPA_FLAG_FUSED_GRAD (dpa_backward then needs none) */
} dpa_workspace_t;

typedef struct {
  dpa_shape_t shape;
  float alpha;                 /* loss = align + alpha * div  (halva_trainer.py:590) */
  /* inputs (unshifted, post-splice; the library applies the [:, :-1] / [:, 1:] shift) */
  const void* h_pair;          /* [2B, T, H] */
  const void* labels_pair;     /* [2B, T]   -100 = ignore */
  const void* signs_pair;      /* [2B, T]   0 / -100 = no phrase, k>=1 = phrase id */
  const void* h_ref_pol;       /* [B_ref, T_ref, H]  policy hidden states on the reference sample */
  const void* h_ref_frz;       /* [B_ref, T_ref, H]  frozen reference model's hidden states */
  const void* labels_ref;      /* [B_ref, T_ref] */
  const void* W;               /* [V_local, H] policy lm_head */
  const void* W_ref;           /* [V_local, H] frozen lm_head (may alias W) */
  /* workspaces */
  void* state;  size_t state_bytes;
  void* scratch; size_t scratch_bytes;
  /* forward outputs */
  float* losses;               /* [4]: loss, align, div, reserved */
  int32_t* status;             /* [8]: status bits, K, n_pair_active, n_ref_active, ... */
  /* backward */
  const float* d_loss;         /* device scalar: upstream gradient of `loss` (read by dpa_backward; NULL = 1.  Under
                                  DPA_FLAG_GRAD_HINT also read by dpa_forward, as the expected value) */
  void* dH_pair;               /* [2B, T, H]  (fully written, zeros for inactive rows) */
  void* dH_ref;                /* [B_ref, T_ref, H] */
  void* dW;                    /* [V_local, H] or NULL */
  /* DPA_FLAG_FUSED_NORM only */
  const void* norm_w;          /* [H] final-norm weight of the policy model */
  const void* norm_w_ref;      /* [H] final-norm weight of the frozen reference model (may alias norm_w) */
  void* d_norm_w;              /* [H] gradient of norm_w, or NULL (frozen under the LoRA recipe) */
  float norm_eps;
  /* Optional: the frozen reference model's logits on the reference sample, given instead of h_ref_frz / W_ref (both may
   * then be NULL) -- e.g. logits cached from a previous pass of the frozen model (halva_trainer.py:576-578 runs it under
   * no_grad every step).  Row (s, t) of the [B_ref * T_ref] rows starts at ref_logits + (s * T_ref + t) * ref_logits_ld
   * elements; column c is vocabulary id v_offset + c, V_local columns are read.  bf16, or fp32 values that are
   * bf16-representable (what `.float()` of a bf16 model's logits gives, halva_trainer.py:520; other fp32 values are
   * rounded to bf16).  Only with DPA_FLAG_FUSED_GRAD (the chunked forward keeps bf16 logits anyway). */
  const void* ref_logits;
  int64_t ref_logits_ld;       /* row stride in elements; (ref_logits_ld * element size) % 16 == 0 */
  int32_t ref_logits_dtype;    /* DPA_BF16 or DPA_F32 */
} dpa_args_t;

int dpa_version(void);
const char* dpa_last_error(void);
int dpa_query_workspace(const dpa_shape_t* shape, dpa_workspace_t* out);
/* sizeof(dpa_shape_t), sizeof(dpa_workspace_t), sizeof(dpa_args_t) as this library was compiled: lets a binding written in
 * another language (ctypes, cgo, JNI) check its mirror of the structs before the first real call */
int dpa_struct_sizes(int* shape_bytes, int* workspace_bytes, int* args_bytes);

/* forward: metadata + compaction + fused lm_head/log-softmax/gather/KL row statistics + loss assembly */
int dpa_forward(const dpa_args_t* args, void* cuda_stream);
/* backward: recompute logits chunk-wise -> dLogits (bf16, row chunks) -> dH (+ dW) */
int dpa_backward(const dpa_args_t* args, void* cuda_stream);

/* Vocab-parallel lm_head: the forward is split around the one exchange step (all-gather of per-row partial
 * statistics).  stats -> [caller or dpa_comm all-gathers `row_stats`] -> finish. */
int dpa_forward_stats(const dpa_args_t* args, void* cuda_stream);
int dpa_forward_finish(const dpa_args_t* args, const void* gathered_stats, int n_shards, void* cuda_stream);
/* byte size and device address of this rank's merged per-row statistics inside `state` */
int dpa_row_stats_view(const dpa_args_t* args, void** ptr, size_t* bytes);

/* Optional communicator for vocab-parallel lm_head (V_local < V): one call each way, no caller-side exchange.
 *   dpa_forward under DPA_FLAG_FUSED_GRAD: every rank runs the chunked pass on the same rows and its own lm_head rows; the
 *     per-row statistics of the shards (36 B/row) are pushed into CUDA-IPC peer memory and merged once for the pair segment
 *     and once per row chunk, the dH partial sums of a chunk are all-reduced in place through peer memory by the library's
 *     own kernels (packed rows only, row counts read from device memory: no host read-back, no stream synchronisation) and
 *     scattered; dW stays sharded.  No NCCL call on the step path.
 *   dpa_forward without DPA_FLAG_FUSED_GRAD (statistics forward + recompute dpa_backward): the statistics record is
 *     all-gathered with NCCL on the caller's stream, the backward's dH exchange is the same peer-memory one.
 * NCCL (bound at run time: the copy already loaded in the process, else libnccl.so.2) bootstraps the communicator and the
 * exchange of the IPC handles.  The symmetric peer buffer (statistics region + one or two dH chunk buffers) is allocated
 * collectively at first use and when a larger shape arrives -- the only moment the communicator synchronises the device.
 * Not thread-safe: call once per process at start-up, before dpa_query_workspace.
 *   rank 0:    dpa_comm_unique_id(id)  -> broadcast the 128 bytes by any means ->  every rank: dpa_comm_init(rank, world, id) */
int dpa_comm_unique_id(void* out_128_bytes);
int dpa_comm_init(int rank, int world, const void* unique_id_128_bytes);
int dpa_comm_destroy(void);
int dpa_comm_world(void);          /* 1 when no communicator is initialised */

/* ---- Host-buffer entry (HOST pointers): the batch as the data collator leaves it in pinned host memory
 * (llava/train/train_halva.py:963-989 builds labels / phrase ids on the host).  A hidden row (s, t) is read by the kernels
 * only when it predicts a contributing label (halva_trainer.py:403-407,550-562), so only those rows -- a few contiguous
 * spans per sequence -- need to cross the host->device link. */
typedef struct { int32_t seq, t0, t1; } dpa_span_t;            /* rows [t0, t1) of sequence seq */
/* Host-only, no CUDA call: spans of rows (s, t) with labels[s,t+1] != -100 (and, when signs != NULL, signs[s,t+1] not in
 * {0, -100}); spans less than merge_gap+1 rows apart are merged.  Returns the number of spans written (<= cap), or a
 * negative dpa_status_t (DPA_ERR_WORKSPACE: cap too small). */
int dpa_plan_host_spans(const void* labels, const void* signs, int S, int T, int index_dtype, int merge_gap,
                        dpa_span_t* out, int cap);
typedef struct {
  int32_t B, T, B_ref, T_ref, H;
  int32_t dtype;            /* dpa_dtype_t of the hidden states */
  int32_t index_dtype;      /* dpa_index_t of labels / signs */
  int32_t dense;            /* 1: copy every hidden row (no planning) */
  int32_t merge_gap;        /* merge spans less than merge_gap+1 rows apart (fewer, larger copies) */
  /* pinned HOST memory, laid out like the device tensors of dpa_args_t */
  const void* h_pair; const void* labels_pair; const void* signs_pair;
  const void* h_ref_pol; const void* h_ref_frz /* may be NULL */; const void* labels_ref;
  /* DEVICE destinations of the same shapes */
  void* d_h_pair; void* d_labels_pair; void* d_signs_pair; void* d_h_ref_pol; void* d_h_ref_frz; void* d_labels_ref;
} dpa_host_batch_t;
typedef struct {
  int64_t bytes;                         /* host->device bytes issued */
  int32_t spans_pair, spans_ref, rows_pair, rows_ref;
  float plan_us, issue_us;               /* host time spent planning / issuing the copies */
} dpa_stage_info_t;
/* Plans (per call, from this batch's labels) and issues the asynchronous copies on `stream`: labels / phrase ids in full,
 * hidden rows span-wise.  Rows outside the spans are left untouched on the device (never read by dpa_forward). */
int dpa_stage_host_batch(const dpa_host_batch_t* batch, void* cuda_stream, dpa_stage_info_t* info);

/* Optional live timing of the kernels this library launches (bench.py): when enabled every launch group is
 * bracketed by CUDA events on the caller's stream; collect() synchronises on them and sums milliseconds and
 * kernel-launch counts per kernel class k (name via dpa_profile_kind_name(k)); returns the number of classes. */
int dpa_profile_enable(int on);
int dpa_profile_collect(double* ms, int* launches, int n);
const char* dpa_profile_kind_name(int k);

/* Probe: co-resident clusters of `cluster_size` CTAs of the 256-wide GEMM kernel on the current device (>= 0), or a negative
 * dpa_status_t. */
int dpa_debug_max_clusters(int cluster_size);

/* Encoded TMA descriptors are cached per (device, pointer, dims, stride, box); cumulative hit / miss counts (tests). */
int dpa_tmap_cache_stats(long long* hits, long long* misses);

/* cta_group of the single-accumulator GEMMs: 2 (default; CTA pairs on 256 x 256 tiles, tcgen05 cta_group::2) or 1
 * (one CTA per 128 x 256 tile).  Also settable with the environment variable DPA_CTA_GROUP before the first call. */
int dpa_set_cta_group(int cg);

/* Debug / unit-test hook: D[M,N] (fp32) = A * B^T through the same tcgen05 mainloop.
 * a_major/b_major: 0 = K-major ([M,K] / [N,K] row-major), 1 = MN-major ([K,M] / [K,N] row-major). */
int dpa_debug_gemm(const void* A, const void* B, float* D, int M, int N, int K, int a_major, int b_major, int bn,
                   void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* HALVA_DPA_H */
