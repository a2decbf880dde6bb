// Parameter blocks of the loss / backward kernels (loss_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cpb {

struct LossParams {
  const float* scores;        // [B, C] raw sums of per-token maxima
  const __nv_bfloat16* q;     // [B * nq_pad, q_dim] padded queries (lengths are counted from column 0); nullptr = the
                              // scores are plain dot products of single vectors (bi-encoder losses): no lengths
  int q_dim;                  // row stride of q in elements (128, or 192 / 256 / 320)
  float* loss;                // [1]
  float* grad;                // [B, C] dLoss/dScores (raw), or nullptr
  float* bounds;              // [2] min / max of the normalised scores, or nullptr
  int B, C, nq_pad, offset;
  int mode;                   // 0 = cross entropy (ColbertLoss), 1 = pairwise softplus (ColbertPairwiseCELoss),
                              // 2 = sigmoid (+1 at column b + offset, -1 elsewhere, mean over B * C),
                              // 3 = symmetric cross entropy (BiPairedEncoderLoss: rows and columns, C == B)
  int normalize, filter;
  float temperature, filter_threshold, filter_factor;
  // explicit negatives (ColbertNegativeCELoss / ColbertPairwiseNegativeCELoss); neg_scores == nullptr: none
  const float* neg_scores;    // [B, B * n_neg] raw sums of every query against every query's negatives
  float* grad_neg;            // [B, B * n_neg] or nullptr
  int n_neg;
  float in_batch_weight;      // weight of the in-batch term (mode 0 / 1) when negatives are present
  int neg_pos_delta;          // the explicit-negative term takes its positive from column b + offset + neg_pos_delta
                              // (BiPairwiseNegativeCELoss: its in-batch term ignores `offset`, its explicit term does not)
};

struct BwdParams {
  const float* g;             // [B, C]
  const float* grad_out;      // scalar upstream gradient, or nullptr (= 1)
  const int32_t* argmax;      // hard max: [C, q_rows] document-relative token index, -1 = floor (no gradient)
  const float* lse;           // smooth max: [C, q_rows] tau * logsumexp per (document, query row)
  float smooth_c;             // log2(e) / tau
  const __nv_bfloat16* q;     // [q_rows, dim]
  const __nv_bfloat16* docs;  // [doc_rows, dim]
  const int32_t* doc_start;   // [C]
  const int32_t* doc_len;     // [C]
  float* dq;                  // [q_rows, dim] written
  float* dd;                  // [doc_rows, dim] written (every row of every document)
  const uint64_t* dd_doc_base; // or nullptr.  Multi-GPU training exchange: address of document c's [len_c, dim] fp32
                              // gradient block in its OWNER rank's (pre-zeroed) accumulator, reached through the NVLink
                              // peer mapping; rows are then ADDED there (red.global.add) instead of written to dd
  int B, C, nq_pad, nq_real, q_rows;
  int dim;                    // padded embedding dim: 128, 192, 256 or 320
  int max_doc_len;            // longest document
  int contiguous;             // documents back to back and covering the bank: no row of dd is outside a document
  int out_bf16;               // dq / dd point to bf16 buffers (hard max without dd_doc_base)
  int64_t doc_rows;
};

cudaError_t colbert_loss_launch(const LossParams& p, cudaStream_t stream);
cudaError_t maxsim_bwd_launch(const BwdParams& p, cudaStream_t stream);
cudaError_t smooth_bwd_launch(const BwdParams& p, cudaStream_t stream);

}  // namespace cpb
