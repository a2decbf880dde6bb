/*
 * colpali_b200 -- C ABI of the B200-native late-interaction hot path.
 *
 * The reference (illuin-tech/colpali, `colpali_engine`) is pure Python and has no FFI of its own:
 * the seams this library sits behind are three Python attributes.  Each
 * entry point below names the reference code it replaces.  The Python host side
 * (colpali_b200/*.py) binds these symbols with ctypes and keeps the reference signatures.
 *
 * Conventions
 *   - every pointer named `d_*` is a DEVICE pointer owned by the caller (PyTorch's caching
 *     allocator in practice); the library allocates nothing and never synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns;
 *   - bf16 tensors are raw uint16 storage, row-major, rows of `dim` elements;
 *   - return value: 0 on success, negative CPB_E_* on error; cpb_last_error() returns a
 *     thread-local human-readable message for the most recent failure on the calling thread;
 *   - functions are re-entrant and keep no global mutable state besides a cached driver
 *     entry point.
 */
#ifndef COLPALI_B200_H_
#define COLPALI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB_ABI_VERSION 1

/* error codes */
#define CPB_OK 0
#define CPB_E_INVALID (-1)   /* bad argument (shape, alignment, null pointer)      */
#define CPB_E_UNSUPPORTED (-2) /* valid request this build cannot serve (e.g. dim) */
#define CPB_E_CUDA (-3)      /* a CUDA runtime / driver call failed                 */
#define CPB_E_DEVICE (-4)    /* device is not sm_100 (tcgen05 / TMEM required)      */

/* flags for cpb_maxsim_fwd */
#define CPB_FLAG_ROUND_BF16 1u /* emulate the reference's bf16 result rounding: each per-token
                                  maximum and the final score are rounded to bf16 (what
                                  torch.einsum(...).max().sum() yields for bf16 inputs,
                                  processing_utils.py:179) instead of staying fp32. */

#define CPB_FLAG_CONTIGUOUS 2u /* the caller guarantees d_doc_start[j+1] == d_doc_start[j] + d_doc_len[j]
                                  for every j (documents stored back to back): tiles then run across
                                  document boundaries and no short per-document tail tiles are issued. */

int cpb_abi_version(void);
const char* cpb_last_error(void);

/* Number of SMs / compute capability of `device` (for grid sizing on the host side). */
int cpb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/*
 * Tuning knobs for experiments (process-wide, not thread-safe against concurrent launches):
 *   "cluster"         0 = auto, 1 / 2 / 4 = CTAs per cluster sharing document tiles by TMA multicast
 *   "qtiles_per_cta"  0 = auto, 1 / 2     = resident 128-row query tiles per CTA
 *   "mma_split"       5..8 (default 6): K-steps of a job issued before the next job's barrier waits
 *   "balanced"        1 (default) / 0: allow tile-balanced partitions in cpb_maxsim_fwd_balanced
 *   "debug_delay"     profiling only
 *   "debug_flags"     profiling-only bits (upper 16), 0 in production
 */
int cpb_set_option(const char* name, int value);

/*
 * Fused MaxSim forward.
 *   replaces: torch.einsum("bnd,csd->bcns", q, d).max(dim=3)[0].sum(dim=2)
 *             colpali_engine/utils/processing_utils.py:179  (score_multi_vector inner loop)
 *             colpali_engine/loss/late_interaction_losses.py:153-154 (ColbertLoss scores),
 *             :297-298 (ColbertPairwiseCELoss), :444-445 (ColbertSigmoidLoss)
 *
 *   d_q          bf16 [n_queries * nq_pad, 128]; every query occupies nq_pad rows (nq_pad a
 *                multiple of 32), zero rows after its real tokens (a zero row adds 0, exactly
 *                like the reference's zero padding of queries, processing_utils.py:172).
 *   d_docs       bf16 [doc_rows, 128]: flat token bank.
 *   d_doc_start  int32 [n_docs]: first bank row of each document.
 *   d_doc_len    int32 [n_docs]: number of rows of each document (0 allowed).
 *   d_doc_floor  fp32 [n_docs] or NULL: initial value of every per-token maximum of that
 *                document; -inf = plain max, 0 = the document had zero-padding rows in the
 *                reference batch (processing_utils.py:176-178: pad rows score exactly 0 and take
 *                part in the max).  NULL means -inf everywhere.
 *   d_scores     fp32 [n_queries, n_docs] out.
 *   d_argmax     int32 [n_docs, n_queries * nq_pad] out, or NULL.  Row index (relative to the
 *                document start) of the first maximal token for every (document, query row);
 *                -1 when the maximum is the floor.  Needed by cpb_maxsim_bwd.
 *   d_workspace  fp32 [(nq_pad/32) * n_queries * n_docs] scratch, only read when nq_pad > 32
 *                (may be NULL otherwise).
 */
int cpb_maxsim_fwd(const void* d_q, int n_queries, int nq_pad,
                   const void* d_docs, int64_t doc_rows,
                   const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                   float* d_scores, int32_t* d_argmax, float* d_workspace,
                   uint32_t flags, void* stream);

/*
 * Same as cpb_maxsim_fwd, with tile-balanced partitioning for CPB_FLAG_CONTIGUOUS banks: every persistent CTA gets
 * the same number of 256-row tiles even if that cuts a document in two; the two partial per-token maxima of a cut
 * document are exchanged through d_split_ws (right neighbour publishes, left neighbour combines and emits the score).
 *   uniform_len     > 0 if every document has exactly this many rows (skips a binary search), else 0
 *   max_doc_len     length of the longest document (partitions shorter than this fall back to whole documents)
 *   d_split_ws      device scratch of cpb_maxsim_split_workspace_bytes() bytes, ZERO-INITIALISED ONCE by the caller and
 *                   then reused across calls on the same stream order; NULL disables balancing
 *   epoch           non-zero, different for every call that shares d_split_ws (a slot is valid when it holds `epoch`)
 */
int cpb_maxsim_fwd_balanced(const void* d_q, int n_queries, int nq_pad,
                            const void* d_docs, int64_t doc_rows,
                            const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                            float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags,
                            int uniform_len, int max_doc_len, void* d_split_ws, int64_t split_ws_bytes, uint32_t epoch,
                            void* stream);
int64_t cpb_maxsim_split_workspace_bytes(int n_queries, int nq_pad);

/*
 * Corpus-sharded scoring with the all-gather of the score slabs FUSED into the kernel epilogue (BASELINE configs[3];
 * the reference scores on one device only).  Every rank owns n_docs documents and the same n_queries queries; each
 * score is stored straight into all ranks' gathered buffers through NVLink peer mappings:
 *     peer_slabs[r][my_rank][q][doc] = score      for r in 0..n_peers-1      (fp32, [n_peers, n_queries, n_docs] each)
 * No collective kernel follows.  Completion: if d_done_counter is given, the last CTA of the grid stores `signal_value`
 * into 32-bit word (flag_word_offset + my_rank) of every peer's buffer once all of this rank's scores are written
 * (__threadfence_system ordered); a consumer waits until its own words [flag_word_offset, +n_peers) hold the value
 * (cpb_wait_flags).  Without d_done_counter the caller needs a cross-rank barrier instead.  nq_pad must be 32.
 *   d_done_counter   local device uint32, zero before the first launch (reset by the kernel), or NULL
 *   d_peer_slabs   device array of n_peers uint64 base addresses of the peers' gathered buffers (this rank's included),
 *                  e.g. torch.distributed._symmetric_memory handle.buffer_ptrs_dev
 * Other arguments as cpb_maxsim_fwd_balanced (d_split_ws may be NULL).
 */
int cpb_maxsim_fwd_allgather(const void* d_q, int n_queries, int nq_pad,
                             const void* d_docs, int64_t doc_rows,
                             const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                             const uint64_t* d_peer_slabs, int n_peers, int my_rank, uint32_t flags,
                             int uniform_len, int max_doc_len, void* d_split_ws, int64_t split_ws_bytes, uint32_t epoch,
                             uint32_t* d_done_counter, int64_t flag_word_offset, uint32_t signal_value, void* stream);

/* Enqueue a wait on `stream` until d_flags[0..n) all equal `value` (consumer side of cpb_maxsim_fwd_allgather). */
int cpb_wait_flags(const uint32_t* d_flags, int n, uint32_t value, void* stream);

/*
 * DRAFT: cpb_maxsim_fwd for embedding dims 192 / 256 / 320 (ColQwen3: models/qwen3/colqwen3/modeling_colqwen3.py:48).
 * d_q and d_docs are [rows, dim] bf16; every other argument as cpb_maxsim_fwd.  Whole-document partitions only.
 */
int cpb_maxsim_fwd_dim(const void* d_q, int n_queries, int nq_pad,
                       const void* d_docs, int64_t doc_rows,
                       const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                       float* d_scores, int32_t* d_argmax, float* d_workspace,
                       uint32_t flags, int dim, void* stream);

/* Bytes of d_workspace cpb_maxsim_fwd needs for this shape (0 when nq_pad == 32). */
int64_t cpb_maxsim_workspace_bytes(int n_queries, int nq_pad, int n_docs);

/* loss modes of cpb_colbert_loss_fwd */
#define CPB_LOSS_CE 0       /* ColbertLoss: cross entropy over in-batch documents          */
#define CPB_LOSS_PAIRWISE 1 /* ColbertPairwiseCELoss: softplus(hardest negative - positive) */
#define CPB_LOSS_SIGMOID 2  /* ColbertSigmoidLoss (late_interaction_losses.py:431-465): n_docs == n_queries, offset 0 */

/*
 * In-batch-negative loss on a [n_queries, n_docs] matrix of raw MaxSim sums, with its gradient.
 *   replaces: ColbertLoss.forward            colpali_engine/loss/late_interaction_losses.py:152,155-164
 *             ColbertPairwiseCELoss.forward  colpali_engine/loss/late_interaction_losses.py:296,299-313
 *             ColbertModule._apply_normalization :46-71, ._filter_high_negatives :93-107
 *   lengths (late_interaction_losses.py:152) are counted here from column 0 of d_q (same padded
 *   layout as cpb_maxsim_fwd).  The positive of query b is document b + offset (:33-38).
 *
 *   d_loss         fp32 [1] out: mean over queries.
 *   d_grad_scores  fp32 [n_queries, n_docs] out or NULL: d loss / d raw score.
 *   d_bounds       fp32 [2] out or NULL: min / max of the length-normalised scores (the reference
 *                  prints a warning when they leave [-norm_tol, 1 + norm_tol], :64-70).
 */
int cpb_colbert_loss_fwd(const float* d_scores, const void* d_q, int n_queries, int nq_pad, int n_docs, int mode,
                         float temperature, int normalize_scores, int pos_aware_negative_filtering,
                         float filter_threshold, float filter_factor, int offset,
                         float* d_loss, float* d_grad_scores, float* d_bounds, void* stream);

/*
 * Losses with explicit negative documents.
 *   replaces: ColbertNegativeCELoss.forward          colpali_engine/loss/late_interaction_losses.py:215-252
 *             ColbertPairwiseNegativeCELoss.forward  colpali_engine/loss/late_interaction_losses.py:361-398
 *   loss = (1 - w) * mean_{b,l} softplus((neg[b,l] - pos[b]) / T)  +  w * in_batch_loss(inner_mode)
 *   pos[b] = scores[b, b + offset];  neg[b, l] = neg_scores[b, b * n_neg + l]  (both length-normalised).
 *
 *   d_scores           fp32 [n_queries, n_docs]           raw MaxSim sums against the (gathered) positives
 *   d_neg_scores       fp32 [n_queries, n_queries * n_neg] raw MaxSim sums of every query against every query's
 *                      negatives (only the block diagonal is used; one dense MaxSim launch produces it)
 *   inner_mode         CPB_LOSS_CE (ColbertNegativeCELoss) or CPB_LOSS_PAIRWISE (ColbertPairwiseNegativeCELoss)
 *   d_grad_scores / d_grad_neg_scores   gradients of the loss w.r.t. both score matrices (zeros off the block diagonal)
 */
int cpb_colbert_neg_loss_fwd(const float* d_scores, const float* d_neg_scores, const void* d_q, int n_queries,
                             int nq_pad, int n_docs, int n_neg, int inner_mode, float temperature,
                             int normalize_scores, int pos_aware_negative_filtering, float filter_threshold,
                             float filter_factor, float in_batch_term_weight, int offset,
                             float* d_loss, float* d_grad_scores, float* d_grad_neg_scores, void* stream);

/*
 * Backward of cpb_maxsim_fwd: given g = d loss / d scores and the argmax saved by the forward,
 *   dq[row]                       = sum_c g[query(row), c] * docs[doc_start[c] + argmax[c, row]]
 *   dd[doc_start[c] + argmax[..]] += g[query(row), c] * q[row]            (fp32 vector atomics)
 *   replaces: autograd through torch.einsum / amax / sum (late_interaction_losses.py:153-154); the
 *   reference keeps the [B, C, N_q, N_d] similarity tensor alive for it, this path keeps [C, rows] int32.
 *
 *   d_grad_out  fp32 [1] upstream gradient of the loss, or NULL for 1.
 *   d_dq        fp32 [n_queries * nq_pad, 128] out, or NULL to skip.
 *   d_dd        fp32 [doc_rows, 128] in/out, MUST be zero-initialised by the caller, or NULL to skip.
 *   Rows whose argmax is -1 (the floor won) receive no gradient.
 */
int cpb_maxsim_bwd(const float* d_grad_scores, const float* d_grad_out, const int32_t* d_argmax,
                   const void* d_q, int n_queries, int nq_pad,
                   const void* d_docs, int64_t doc_rows, const int32_t* d_doc_start, int n_docs,
                   float* d_dq, float* d_dd, void* stream);

/* DRAFT: the two loss entry points for d_q of shape [n_queries * nq_pad, dim], dim in {128, 192, 256, 320}. */
int cpb_colbert_loss_fwd_dim(const float* d_scores, const void* d_q, int n_queries, int nq_pad, int n_docs, int mode,
                             float temperature, int normalize_scores, int pos_aware_negative_filtering,
                             float filter_threshold, float filter_factor, int offset,
                             float* d_loss, float* d_grad_scores, float* d_bounds, int dim, void* stream);
int cpb_colbert_neg_loss_fwd_dim(const float* d_scores, const float* d_neg_scores, const void* d_q, int n_queries,
                                 int nq_pad, int n_docs, int n_neg, int inner_mode, float temperature,
                                 int normalize_scores, int pos_aware_negative_filtering, float filter_threshold,
                                 float filter_factor, float in_batch_term_weight, int offset,
                                 float* d_loss, float* d_grad_scores, float* d_grad_neg_scores, int dim, void* stream);

/* DRAFT: cpb_maxsim_bwd for [rows, dim] operands and gradients, dim in {128, 192, 256, 320} (see cpb_maxsim_fwd_dim). */
int cpb_maxsim_bwd_dim(const float* d_grad_scores, const float* d_grad_out, const int32_t* d_argmax,
                       const void* d_q, int n_queries, int nq_pad,
                       const void* d_docs, int64_t doc_rows, const int32_t* d_doc_start, int n_docs,
                       float* d_dq, float* d_dd, int dim, void* stream);

/* flags for cpb_head_fwd */
#define CPB_HEAD_CLAMP_NORM 1u      /* norm = max(norm, 1e-12): ColModernVBert variant (modeling_colmodernvbert.py:59) */
#define CPB_HEAD_SINGLE_ROUNDING 2u /* keep fp32 until the final store instead of emulating the reference's
                                       three bf16 roundings (Linear output, norm, quotient) */

/*
 * Fused multi-vector projection head.
 *   replaces: proj = self.custom_text_proj(h); proj = proj / proj.norm(dim=-1, keepdim=True);
 *             proj = proj * attention_mask.unsqueeze(-1) [; proj = proj * image_mask]
 *             colpali_engine/models/qwen2/colqwen2/modeling_colqwen2.py:65-74 (ctor :34-35) and the identical
 *             tails of ColPali :67-77, ColQwen2.5 :67-76, ColQwen3.5 :67-76, ColQwen2.5-Omni :64-73,
 *             ColGemma3 :84-93, ColIdefics3 :38-46, ColModernVBert :57-65 (CPB_HEAD_CLAMP_NORM).
 *
 *   d_hidden          bf16 [n_tokens, hidden]   last_hidden_state, flattened over (batch, sequence)
 *   d_weight          bf16 [dim, hidden]        custom_text_proj.weight (nn.Linear layout)
 *   d_bias            bf16 [dim] or NULL        custom_text_proj.bias
 *   d_attention_mask  int64 [n_tokens] or NULL  multiplied in as a value (0/1 in practice)
 *   d_extra_mask      uint8 [n_tokens] or NULL  non-zero keeps the row (input_ids == image_token_id)
 *   d_out             bf16 [n_tokens, dim]
 *   dim must be 128 and hidden a multiple of 64 in this build (CPB_E_UNSUPPORTED otherwise).
 */
int cpb_head_fwd(const void* d_hidden, int64_t n_tokens, int hidden,
                 const void* d_weight, const void* d_bias, int dim,
                 const int64_t* d_attention_mask, const uint8_t* d_extra_mask,
                 void* d_out, uint32_t flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLPALI_B200_H_ */
