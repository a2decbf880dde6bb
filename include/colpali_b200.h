/*
 * colpali_b200 -- C ABI of the B200-native late-interaction hot path.
 *
 * The reference (illuin-tech/colpali, `colpali_engine`) is pure Python and has no FFI of its own:
 * the seams this library sits behind are three Python attributes.  Each
 * entry point below names the reference code it replaces.  The Python host side
 * (the .py files of colpali_b200/) binds these symbols with ctypes and keeps the reference signatures.
 *
 * Conventions
 *   - every pointer named `d_*` is a DEVICE pointer owned by the caller (PyTorch's caching
 *     allocator in practice); the library allocates nothing and never synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns;
 *   - bf16 tensors are raw uint16 storage, row-major, rows of `dim` elements;
 *   - return value: 0 on success, negative CPB_E_* on error; cpb_last_error() returns a
 *     thread-local human-readable message for the most recent failure on the calling thread;
 *   - argument blocks are plain structs whose first member is `struct_size` = sizeof(the struct)
 *     as the caller compiled it: members added by later versions are read only when present, so a
 *     zero-initialised struct with the fields you need is always a valid call;
 *   - launches may be issued concurrently from several host threads (different streams).  The
 *     only process-wide state is (a) caches of device properties / occupancy, which are
 *     idempotent and guarded, and (b) the tuning knobs of cpb_set_option, which are atomics read
 *     once per launch -- change them only for experiments.  Device scratch that a caller shares
 *     between launches (d_split_ws, d_done_counter) must not be shared by launches that can run
 *     at the same time: give every stream its own;
 *   - forward progress: the MaxSim kernels are persistent.  Whenever the documents are split
 *     into more than one partition per query-tile group, the grid is sized from the occupancy
 *     query so that ALL of its CTAs are resident at once on an otherwise idle device, and some
 *     CTAs then wait for a neighbour inside the kernel (a document cut by a partition boundary
 *     -- option "balanced" --, the fused top-k's per-group rendezvous, the fused all-gather's
 *     write-after-read guard; the fused loss never waits: the last CTA to arrive does the work).
 *     Kernels of other streams or processes that occupy SMs only delay those waits, as long as
 *     they finish on their own.  What breaks the assumption is a context that cannot use every
 *     SM the occupancy query counted (MPS with an active-thread percentage, green contexts):
 *     there, set cpb_set_option("balanced", 0) and do not pass d_topk_* / d_peer_bases.
 *     A wait that outlives "wait_timeout_ms" traps (in-kernel) or sets the status word
 *     (cpb_wait_flags) instead of hanging the device.
 */
#ifndef COLPALI_B200_H_
#define COLPALI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB_ABI_VERSION 3

/* error codes */
#define CPB_OK 0
#define CPB_E_INVALID (-1)   /* bad argument (shape, alignment, null pointer)      */
#define CPB_E_UNSUPPORTED (-2) /* valid request this build cannot serve (e.g. dim) */
#define CPB_E_CUDA (-3)      /* a CUDA runtime / driver call failed                 */
#define CPB_E_DEVICE (-4)    /* device is not sm_100 (tcgen05 / TMEM required)      */

/* flags of cpb_maxsim_args.flags / cpb_maxsim_fwd */
#define CPB_FLAG_ROUND_BF16 1u /* emulate the reference's bf16 result rounding: each per-token
                                  maximum and the final score are rounded to bf16 (what
                                  torch.einsum(...).max().sum() yields for bf16 inputs,
                                  processing_utils.py:179) instead of staying fp32. */
#define CPB_FLAG_CONTIGUOUS 2u /* the caller guarantees d_doc_start[j+1] == d_doc_start[j] + d_doc_len[j]
                                  for every j (documents stored back to back): tiles then run across
                                  document boundaries and no short per-document tail tiles are issued. */
#define CPB_FLAG_GRAD_BF16 8u /* cpb_maxsim_bwd_launch (hard max, no peer scatter): d_dq / d_dd are bf16 buffers -- the
                                 kernels round every gradient row once (round-to-nearest-even, the same value a cast of
                                 the fp32 row gives) instead of leaving two cast kernels to the caller */
#define CPB_FLAG_INDEPENDENT 4u /* this launch reads nothing that the previous kernel on the stream wrote
                                   (e.g. the next query batch against a resident bank): it may start
                                   while that kernel is still draining (programmatic dependent launch
                                   without the dependency wait). */

int cpb_abi_version(void);
const char* cpb_last_error(void);

/* Number of SMs / compute capability of `device` (for grid sizing on the host side). */
int cpb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/*
 * Tuning knobs for experiments (process-wide atomics, read once per launch):
 *   "cluster"         0 = auto, 1 / 2 / 4 = CTAs per cluster sharing document tiles by TMA multicast
 *   "qtiles_per_cta"  0 = auto, 1 / 2     = resident 128-row query tiles per CTA
 *   "balanced"        1 (default) / 0: allow tile-balanced partitions
 *   "pdl"             1 (default) / 0: programmatic dependent launch of the MaxSim kernels
 *   "boundary_mode"   1 (default) / 0: shifted-chunk epilogue for tiles that hold a document boundary
 *   "pair"            0 (default) / 1: CTA-pair MMAs (tcgen05.mma.cta_group::2, maxsim_pair_sm100.cu) where the shape
 *                     allows (dim 128, contiguous bank, query-tile count a multiple of 2 x qtiles_per_cta); bit-identical
 *                     results, measured 9-27 % slower than the default kernel
 *   "early_spin"      cycles (default 0 = one probe) the issuer polls those barriers before it finishes the current
 *                     job first; a huge value restores round 1's blocking wait
 *   "head_cluster"    0 = auto, 1 / 2: CTAs sharing the projection weight block (wide head)
 *   "wait_timeout_ms" time-out of cpb_wait_flags (default 120000)
 *   "dense_raster"    cpb_dense_dot_launch walks the output tiles in blocks of this many column tiles x all row tiles
 *                     (default 4; 1 = row tiles fastest, a huge value = column tiles fastest)
 *   "debug_delay", "debug_flags"  profiling only
 */
int cpb_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------------------------------
 * In-batch loss on a [n_queries, n_docs] matrix of raw MaxSim sums, with its gradient.
 *   replaces: ColbertLoss.forward            colpali_engine/loss/late_interaction_losses.py:152,155-164
 *             ColbertPairwiseCELoss.forward  :296,299-313      ColbertSigmoidLoss.forward :446-465
 *             ColbertModule._apply_normalization :46-71, ._filter_high_negatives :93-107
 *             explicit negatives: ColbertNegativeCELoss.forward :215-252, ColbertPairwiseNegativeCELoss.forward :361-398
 *               loss = (1 - w) * mean_{b,l} softplus((neg[b,l] - pos[b]) / T) + w * in_batch_loss(mode)
 *               pos[b] = scores[b, b + offset];  neg[b, l] = neg_scores[b, b * n_neg + l]  (both length-normalised)
 *   lengths (:152) are counted from column 0 of the query rows.  The positive of query b is document b + offset (:33-38).
 * ------------------------------------------------------------------------------------------------------------------ */
#define CPB_LOSS_CE 0       /* ColbertLoss: cross entropy over in-batch documents          */
#define CPB_LOSS_PAIRWISE 1 /* ColbertPairwiseCELoss: softplus(hardest negative - positive) */
#define CPB_LOSS_SIGMOID 2  /* ColbertSigmoidLoss: n_docs == n_queries, offset 0; with d_q == NULL BiSigmoidLoss
                               (bi_encoder_losses.py:372-418): +1 at column b + offset, -1 elsewhere, n_docs a multiple
                               of n_queries */
#define CPB_LOSS_SYMMETRIC_CE 3 /* BiPairedEncoderLoss (bi_encoder_losses.py:140-168): (CE over rows + CE over columns) / 2,
                                   square matrix, d_q == NULL */

typedef struct cpb_loss_desc {
  uint32_t struct_size;
  int32_t mode;                         /* CPB_LOSS_* */
  int32_t normalize_scores;             /* divide every row by its query length (:155-156) */
  int32_t pos_aware_negative_filtering; /* :161-162 */
  int32_t offset;                       /* positive of query b = document b + offset */
  float temperature;
  float filter_threshold, filter_factor;
  const float* d_neg_scores;            /* fp32 [n_queries, n_queries * n_neg] or NULL: every query against every query's
                                           negatives (only the block diagonal is used) */
  int32_t n_neg;
  float in_batch_term_weight;           /* w above; ignored without negatives */
  float* d_loss;                        /* fp32 [1] out: mean over queries */
  float* d_grad_scores;                 /* fp32 [n_queries, n_docs] out or NULL: d loss / d raw score */
  float* d_grad_neg_scores;             /* fp32 [n_queries, n_queries * n_neg] out or NULL */
  float* d_bounds;                      /* fp32 [2] out or NULL: min / max of the normalised scores (:64-70) */
  int32_t neg_pos_offset_delta;         /* with negatives: their term reads its positive from column b + offset + this
                                           (BiPairwiseNegativeCELoss, bi_encoder_losses.py:350 vs :299: the in-batch term of
                                           that class ignores `offset`, the explicit term does not).  Normally 0 */
} cpb_loss_desc;

/* Stand-alone launch of the loss kernel on an existing score matrix.  d_q: bf16 [n_queries * nq_pad, dim], or NULL when
 * the scores are dot products of single vectors (the bi-encoder losses, colpali_engine/loss/bi_encoder_losses.py:64-418:
 * BiEncoderLoss = CPB_LOSS_CE, BiPairwiseCELoss = CPB_LOSS_PAIRWISE, BiSigmoidLoss, BiPairedEncoderLoss, and the two
 * explicit-negative classes through d_neg_scores); nq_pad / dim are then ignored and normalize_scores must be 0. */
int cpb_colbert_loss_launch(const cpb_loss_desc* loss, const float* d_scores, const void* d_q, int n_queries, int nq_pad,
                            int n_docs, int dim, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused MaxSim forward.
 *   replaces: torch.einsum("bnd,csd->bcns", q, d).max(dim=3)[0].sum(dim=2)
 *             colpali_engine/utils/processing_utils.py:179  (score_multi_vector inner loop)
 *             colpali_engine/loss/late_interaction_losses.py:153-154 (ColbertLoss scores),
 *             :297-298 (ColbertPairwiseCELoss), :444-445 (ColbertSigmoidLoss), and with smooth_tau > 0 the
 *             tau * logsumexp(raw / tau) aggregation of :40-44 / :88-90.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cpb_maxsim_args {
  uint32_t struct_size;
  uint32_t flags;               /* CPB_FLAG_* */
  void* stream;
  /* queries: bf16 [n_queries * nq_pad, dim]; every query occupies nq_pad rows (a multiple of 32), zero rows after its
     real tokens (a zero row adds 0 to a hard-max score, exactly like the reference's zero padding, processing_utils.py:172) */
  const void* d_q;
  int32_t n_queries, nq_pad;
  int32_t nq_real;              /* smooth max only: rows of each query that belong to the caller's tensor (<= nq_pad) */
  int32_t dim;                  /* 128 (smaller dims zero-padded by the caller), 192, 256 or 320 */
  /* document bank: bf16 [doc_rows, dim] flat tokens + per-document (start row, rows) */
  const void* d_docs;
  int64_t doc_rows;
  const int32_t* d_doc_start;   /* [n_docs] */
  const int32_t* d_doc_len;     /* [n_docs], 0 allowed */
  const float* d_doc_floor;     /* [n_docs] or NULL: initial value of every per-token maximum of that document; 0 = the
                                   document had zero-padding rows in the reference batch (processing_utils.py:176-178:
                                   pad rows score 0 and take part in the max), -inf = plain max.  NULL = -inf everywhere */
  int32_t n_docs;
  int32_t uniform_len;          /* > 0 if every document has exactly this many rows, else 0 */
  int32_t max_doc_len;          /* longest document, 0 = unknown (no tile-balanced partitions) */
  /* outputs */
  float* d_scores;              /* fp32 [n_queries, n_docs] (may be NULL with the fused all-gather) */
  int32_t* d_argmax;            /* int32 [n_docs, n_queries * nq_pad] or NULL: document-relative row of the first maximal
                                   token per (document, query row), -1 when the floor won.  For cpb_maxsim_bwd_launch */
  float* d_lse;                 /* smooth max: fp32 [n_docs, n_queries * nq_pad] or NULL: tau * logsumexp per (document,
                                   query row).  For cpb_maxsim_bwd_launch */
  float* d_workspace;           /* fp32 [(nq_pad/32) * n_queries * n_docs] scratch, needed when nq_pad > 32 */
  /* tile-balanced partitions of a CPB_FLAG_CONTIGUOUS bank: every persistent CTA gets the same number of 256-row tiles
     even if that cuts a document; the two partial per-token maxima of a cut document are exchanged through d_split_ws */
  void* d_split_ws;             /* cpb_maxsim_split_workspace_bytes() bytes, ZERO-INITIALISED ONCE by the caller, reused by
                                   stream-ordered launches; NULL disables balancing */
  int64_t split_ws_bytes;
  uint32_t epoch;               /* non-zero, different for every launch that shares d_split_ws */
  /* aggregation */
  float smooth_tau;             /* 0 = maximum; > 0 = tau * logsumexp(raw / tau) (ColbertModule.tau) */
  /* Corpus-sharded scoring with the all-gather of the score slabs FUSED into the epilogue (BASELINE configs[3]; the
     reference scores on one device).  Every rank owns n_docs documents and the same queries; each score goes straight
     into all ranks' copies of a symmetric buffer over NVLink: word slab_word_offset + q * n_docs + doc.  Every CTA then
     adds 1 (release, system scope) to word flag_word_offset of every rank; a consumer waits until that word has grown by
     the grid size of the launch (grid_out, cpb_wait_flags).  nq_pad must be 32. */
  const uint64_t* d_peer_bases; /* device array of n_peers base addresses of the peers' symmetric buffers, or NULL */
  uint64_t mc_base;             /* NVSwitch multicast address of the same buffer (one multimem.st / multimem.red reaches
                                   every rank), or 0 to store through the n_peers peer mappings */
  int32_t n_peers;
  int64_t slab_word_offset;
  int64_t flag_word_offset;
  const uint32_t* d_wait_flags; /* or NULL: before its first score store the kernel waits until these n_wait LOCAL words
                                   have reached wait_value (wrap-safe) -- the completion counters of the previous launch on
                                   a double-buffered slab: write-after-read safety without a host round trip */
  int32_t n_wait;
  uint32_t wait_value;
  /* the in-batch loss "emitted directly": the last CTA of the grid turns d_scores into the loss (+ gradient) */
  const cpb_loss_desc* loss;    /* NULL = scores only */
  uint32_t* d_done_counter;     /* needed with `loss`: CPB_LOSS_WORKSPACE_WORDS device words (completion counters of the
                                   query-tile groups + their partial sums), ZERO before the first launch; the kernel leaves
                                   the counters zero.  Not shared by launches that can overlap.  (ABI 2: one word) */
  /* written by the call */
  int32_t grid_out;             /* CTAs launched */
  /* top-k selection fused into the kernel's tail (the sharded scorer's local top-k; replaces torch.topk on the
     [n_queries, n_docs] slab): once a query-tile group's scores are out, each of its CTAs selects the topk_k best documents
     of the group's queries inside its own slice of the score rows (still in L2) and the last one merges the lists -- larger
     score first, smaller document index on ties; fewer than topk_k documents (and scores of -inf) leave (-inf, INT32_MAX)
     filler.  Needs dim 128, nq_pad == 32, d_scores, no CPB_FLAG_INDEPENDENT. */
  float* d_topk_scores;         /* fp32 [n_queries, topk_k] out, or NULL */
  int32_t* d_topk_idx;          /* int32 [n_queries, topk_k] out: document index in this bank */
  void* d_topk_ws;              /* cpb_maxsim_topk_workspace_bytes() bytes, ZERO-INITIALISED ONCE by the caller (the kernel
                                   leaves its counters at zero); not shared by launches that can overlap */
  int32_t topk_k;               /* 1 .. CPB_TOPK_MAX */
} cpb_maxsim_args;
#define CPB_TOPK_MAX 16
#define CPB_LOSS_WORKSPACE_WORDS 4096

int cpb_maxsim_launch(cpb_maxsim_args* args);

/* Positional form of the plain case (dim 128, hard max, whole-document partitions). */
int cpb_maxsim_fwd(const void* d_q, int n_queries, int nq_pad,
                   const void* d_docs, int64_t doc_rows,
                   const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                   float* d_scores, int32_t* d_argmax, float* d_workspace,
                   uint32_t flags, void* stream);

/* Bytes of d_workspace for this shape (0 when nq_pad == 32) / of d_split_ws. */
int64_t cpb_maxsim_workspace_bytes(int n_queries, int nq_pad, int n_docs);
int64_t cpb_maxsim_split_workspace_bytes(int n_queries, int nq_pad);
int64_t cpb_maxsim_topk_workspace_bytes(void);

/* Enqueue a wait on `stream` until d_flags[i] has reached `value` for every i < n (wrap-safe: counters only grow) --
 * the consumer side of the fused all-gather.  A peer that is late is waited for up to "wait_timeout_ms"; on time-out bit
 * i of *d_status (may be NULL) is set and the stream continues -- check it after synchronising. */
int cpb_wait_flags(const uint32_t* d_flags, int n, uint32_t value, uint32_t* d_status, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward of the MaxSim forward: given g = d loss / d scores,
 *   hard max (d_argmax):  dq[row] = sum_c g[b(row), c] * docs[start_c + argmax[c, row]]
 *                         dd[start_c + s] = sum_{rows with argmax[c, row] == s} g[b(row), c] * q[row]
 *   smooth max (d_lse):   P[c, row, s] = exp((<q_row, d_s> - lse[c, row]) / tau) recomputed tile by tile,
 *                         dq[row] = sum_c g[b, c] sum_s P * d_s,   dd[start_c + s] = sum_row g[b, c] P * q_row
 *   replaces: autograd through torch.einsum / amax | logsumexp / sum (late_interaction_losses.py:153-154, :40-44); the
 *   reference keeps the [B, C, N_q, N_d] similarity tensor alive for it, this path keeps [C, rows] int32 or fp32.
 * Both outputs are WRITTEN (no pre-zeroing, no atomics): every bank row of every document is produced exactly once.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cpb_maxsim_bwd_args {
  uint32_t struct_size;
  uint32_t flags;               /* CPB_FLAG_CONTIGUOUS: the documents cover the bank (else rows outside are zeroed first) */
  void* stream;
  const float* d_grad_scores;   /* fp32 [n_queries, n_docs] */
  const float* d_grad_out;      /* fp32 [1] upstream gradient of the loss, or NULL for 1 */
  const int32_t* d_argmax;      /* hard max: [n_docs, n_queries * nq_pad] from the forward, else NULL */
  const float* d_lse;           /* smooth max: [n_docs, n_queries * nq_pad] from the forward, else NULL */
  float smooth_tau;             /* > 0 with d_lse */
  const void* d_q;
  int32_t n_queries, nq_pad, nq_real, dim;
  const void* d_docs;
  int64_t doc_rows;
  const int32_t* d_doc_start;   /* [n_docs] */
  const int32_t* d_doc_len;     /* [n_docs] (rows of the bank outside every document are zero-filled in dd) */
  int32_t n_docs;
  int32_t max_doc_len;          /* longest document */
  void* d_dq;                   /* fp32 (bf16 with CPB_FLAG_GRAD_BF16) [n_queries * nq_pad, dim] out, or NULL to skip */
  void* d_dd;                   /* fp32 (bf16 with CPB_FLAG_GRAD_BF16) [doc_rows, dim] out, or NULL to skip */
  const uint64_t* d_dd_doc_base; /* or NULL.  Hard max only, multi-GPU exchange: device array [n_docs] of addresses of each
                                   document's [len, dim] fp32 gradient block in its OWNER rank's pre-zeroed accumulator
                                   (NVLink peer mapping); gradient rows are then ADDED there (red.global.add) instead of
                                   being written to d_dd -- the reduce-scatter of the reference's gather backward */
} cpb_maxsim_bwd_args;

int cpb_maxsim_bwd_launch(const cpb_maxsim_bwd_args* args);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU training exchange without a collective kernel (replaces accelerator.pad_across_processes + concat_all_gather,
 * colpali_engine/trainer/contrastive_trainer.py:143-150, and pad_to_max_len_right + gather_with_grad,
 * trainer/colmodel_torch_training.py:155-175): every rank writes its zero-padded [n_docs, slot_len, dim] document block
 * into ALL ranks' copies of a symmetric buffer (multimem.st through the NVSwitch multicast mapping, or st.global through
 * the NVLink peer mappings), then every CTA adds 1 (release, system scope) to word flag_word_offset on every rank.
 * Consumers wait until that word has grown by grid_out (cpb_wait_flags, or cpb_maxsim_args.d_wait_flags in-kernel).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct cpb_exchange_push_args {
  uint32_t struct_size;
  uint32_t pad_first;           /* 1: zero rows in front of the data (HF trainer), 0: behind it (torch-loop trainer) */
  void* stream;
  const void* d_src;            /* bf16 [n_docs, len, dim] contiguous */
  int32_t n_docs, len, slot_len, dim;   /* slot_len >= len: rows per document in the gathered bank; dim % 8 == 0 */
  const uint64_t* d_peer_bases; /* device array of n_peers symmetric-buffer base addresses */
  uint64_t mc_base;             /* multicast address of the same buffer, or 0 */
  int32_t n_peers;
  int64_t bank_word_offset;     /* 4-byte words from the buffer base to this rank's block (multiple of 4) */
  int64_t flag_word_offset;
  int32_t grid_out;             /* written: CTAs launched = what the flag word grows by */
} cpb_exchange_push_args;

int cpb_exchange_push(cpb_exchange_push_args* args);

/* Enqueue a one-thread kernel that adds 1 (release, system scope) to word flag_word_offset of every rank's buffer:
 * publishes everything the stream did before (e.g. the peer adds of cpb_maxsim_bwd_launch with d_dd_doc_base). */
int cpb_signal_peers(const uint64_t* d_peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word_offset, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense dot products in fp32 (row f-4):  out[i, j] (+)= alpha * sum_k A[i, k] * B[row(j), k]
 *   replaces: torch.einsum("bd,cd->bc", qs, ps)        colpali_engine/utils/processing_utils.py:126 (score_single_vector)
 *             the same einsum of the bi-encoder losses  colpali_engine/loss/bi_encoder_losses.py:105,158,290,396 and
 *             "bd,bnd->bn" :235, with their backward products through strided operands
 *             torch.einsum("nk,ijk->nij", q, grid)      colpali_engine/interpretability/similarity_map_utils.py:50-52
 *             (d_b_rows = the masked patch rows regrouped "(h w) c -> w h c", :41-47)
 * Operands keep their dtype (bf16 is widened exactly, fp32 stays fp32); strides are in ELEMENTS.
 * ------------------------------------------------------------------------------------------------------------------ */
#define CPB_DOT_A_F32 1u      /* A is fp32 (default: bf16) */
#define CPB_DOT_B_F32 4u      /* B is fp32 (default: bf16) */
#define CPB_DOT_ACCUMULATE 2u /* out += instead of out = */

typedef struct cpb_dense_dot_args {
  uint32_t struct_size;
  uint32_t flags;               /* CPB_DOT_* */
  void* stream;
  const void* d_a;              /* element (i, k) at d_a[i * a_row_stride + k * a_k_stride] */
  int64_t a_row_stride, a_k_stride;
  const void* d_b;              /* element (j, k) at d_b[row(j) * b_row_stride + k * b_k_stride] */
  int64_t b_row_stride, b_k_stride;
  const int32_t* d_b_rows;      /* [n] or NULL: row(j) = d_b_rows[j] (gather), else j */
  int32_t m, n, k;
  float* d_out;                 /* fp32, element (i, j) at d_out[i * out_row_stride + j] */
  int64_t out_row_stride;
  const float* d_alpha;         /* device scalar or NULL (= 1) */
} cpb_dense_dot_args;

int cpb_dense_dot_launch(const cpb_dense_dot_args* args);

/* flags for cpb_head_fwd */
#define CPB_HEAD_CLAMP_NORM 1u      /* norm = max(norm, 1e-12): ColModernVBert variant (modeling_colmodernvbert.py:59) */
#define CPB_HEAD_SINGLE_ROUNDING 2u /* keep fp32 until the final store instead of emulating the reference's
                                       three bf16 roundings (Linear output, norm, quotient) */

/*
 * Fused multi-vector projection head.
 *   replaces: proj = self.custom_text_proj(h); proj = proj / proj.norm(dim=-1, keepdim=True);
 *             proj = proj * attention_mask.unsqueeze(-1) [; proj = proj * image_mask]
 *             colpali_engine/models/qwen2/colqwen2/modeling_colqwen2.py:65-74 (ctor :34-35) and the identical
 *             tails of ColPali :67-77, ColQwen2.5 :67-76, ColQwen3 (dim 320, modeling_colqwen3.py:87-96), ColQwen3.5
 *             :67-76, ColQwen2.5-Omni :64-73, ColGemma3 :84-93, ColIdefics3 :38-46, ColModernVBert :57-65
 *             (CPB_HEAD_CLAMP_NORM).
 *
 *   d_hidden          bf16 [n_tokens, hidden]   last_hidden_state, flattened over (batch, sequence)
 *   d_weight          bf16 [dim, hidden]        custom_text_proj.weight (nn.Linear layout)
 *   d_bias            bf16 [dim] or NULL        custom_text_proj.bias
 *   d_attention_mask  int64 [n_tokens] or NULL  multiplied in as a value (0/1 in practice)
 *   d_extra_mask      uint8 [n_tokens] or NULL  non-zero keeps the row (input_ids == image_token_id)
 *   d_out             bf16 [n_tokens, dim]
 *   dim: 128, or a multiple of 32 in (128, 320]; hidden a multiple of 64 (CPB_E_UNSUPPORTED otherwise).
 */
int cpb_head_fwd(const void* d_hidden, int64_t n_tokens, int hidden,
                 const void* d_weight, const void* d_bias, int dim,
                 const int64_t* d_attention_mask, const uint8_t* d_extra_mask,
                 void* d_out, uint32_t flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLPALI_B200_H_ */
