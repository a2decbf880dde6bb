// Body of the ColBERT loss kernels: [B, C] raw MaxSim sums -> scalar loss AND dLoss/dScores in one pass.
//   replaces late_interaction_losses.py:152 (lengths), :155-156 / :46-71 (normalise), :161-162 / :93-107 (pos-aware
//   negative filtering), :164 (cross entropy, ColbertLoss), :309-313 (top-2 / softplus, ColbertPairwiseCELoss),
//   :452-465 (ColbertSigmoidLoss) and the explicit-negative terms :234-250 / :380-396.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "loss_params.h"

namespace cpb {

__device__ __forceinline__ float warp_sum_f(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max_f(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_min_f(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// value/index pair ordered by (value desc, index asc): the first maximal index wins, like torch.max
struct Top {
  float v;
  int i;
};
__device__ __forceinline__ bool better(const Top& a, const Top& b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

// The row of scores a warp works on: column c = lane + 32 j.  Up to 32 * kLossCols columns are fetched from L2 ONCE
// into registers (kCached) -- every pass over the row was a dependent L2 round trip before, five per row, and the
// last CTA of the fused forward spent ~25 us in them at cfg3 (B = C = 64); wider rows are re-read pass by pass.
constexpr int kLossCols = 16;
template <bool kCached, typename F>
__device__ __forceinline__ void loss_for_cols(int C, int lane, const float (&sv)[kLossCols], const float* row, float inv,
                                              F&& fn) {
  if constexpr (kCached) {
#pragma unroll
    for (int j = 0; j < kLossCols; ++j) {
      const int c = lane + 32 * j;
      if (c < C) fn(c, sv[j]);
    }
  } else {
    for (int c = lane; c < C; c += 32) fn(c, __ldcg(row + c) * inv);
  }
}

// one query row b: loss contribution (returned, warp-uniform), gradients written, running min / max updated
template <bool kCached>
__device__ __forceinline__ float colbert_loss_row(const LossParams& p, int b, int lane, float& mn, float& mx) {
  const bool has_neg = p.neg_scores != nullptr;
  // late_interaction_losses.py:248-250 / :394-396; the row half of the symmetric loss weighs 1/2 (bi_encoder_losses.py:168)
  const float w_ib = (has_neg ? p.in_batch_weight : 1.f) * (p.mode == 3 ? 0.5f : 1.f);
  const float w_out = has_neg ? 1.f - p.in_batch_weight : 0.f;
  const float* row = p.scores + static_cast<int64_t>(b) * p.C;
  const int pidx = b + p.offset;                              // :33-38
  // all loads of the row first (scores, the positive, column 0 of the query's tokens), then the arithmetic
  float sv[kLossCols];
  if constexpr (kCached) {
#pragma unroll
    for (int j = 0; j < kLossCols; ++j) {
      const int c = lane + 32 * j;
      sv[j] = (c < p.C) ? __ldcg(row + c) : 0.f;
    }
  }
  const float pos_raw = __ldcg(row + pidx);
  // lengths = (q[:, :, 0] != 0).sum(1)                       late_interaction_losses.py:152
  float cnt = 0.f;
  if (p.q != nullptr) {
    for (int n = lane; n < p.nq_pad; n += 32)
      cnt += (__bfloat162float(p.q[(static_cast<int64_t>(b) * p.nq_pad + n) * p.q_dim]) != 0.f) ? 1.f : 0.f;
    cnt = warp_sum_f(cnt);
  }
  const float inv = (p.normalize && p.q != nullptr) ? 1.f / cnt : 1.f;   // :155-156 -> :59-62
  if constexpr (kCached) {
#pragma unroll
    for (int j = 0; j < kLossCols; ++j) sv[j] *= inv;
  }
  const float pos = pos_raw * inv;
  const float thr = p.filter_threshold * pos;                 // :101-104
  const float invT = 1.f / p.temperature;
  const float invB = w_ib / static_cast<float>(p.B);
  float loss = 0.f;

  // filtered score of column c (normalised score s) and the factor it was multiplied by      (:105-107)
  auto filtered = [&](int c, float s, float& f) {
    f = (p.filter && c != pidx && s > thr) ? p.filter_factor : 1.f;
    return s * f;
  };
  auto for_cols = [&](auto&& fn) { loss_for_cols<kCached>(p.C, lane, sv, row, inv, fn); };

  if (p.mode == 0 || p.mode == 3) {
    // cross entropy of scores / T against pidx                (:164; bi_encoder_losses.py:113, :165)
    float m = -INFINITY;
    for_cols([&](int c, float s) {
      float f;
      mn = fminf(mn, s);
      mx = fmaxf(mx, s);
      m = fmaxf(m, filtered(c, s, f) * invT);
    });
    m = warp_max_f(m);
    float se = 0.f;
    for_cols([&](int c, float s) {
      float f;
      se += __expf(filtered(c, s, f) * invT - m);
    });
    se = warp_sum_f(se);
    const float lse = m + __logf(se);
    loss += w_ib * (lse - pos * invT);  // the positive column is never filtered
    if (p.grad != nullptr) {
      float* g = p.grad + static_cast<int64_t>(b) * p.C;
      for_cols([&](int c, float s) {
        float f;
        const float sm = __expf(filtered(c, s, f) * invT - lse);
        g[c] = (sm - (c == pidx ? 1.f : 0.f)) * invT * f * inv * invB;
      });
    }
  } else if (p.mode == 1) {
    // pos = diagonal(offset); top-2 of the row; neg = top1 == pos ? top2 : top1      (:309-311)
    Top t1{-INFINITY, 0x7fffffff}, t2{-INFINITY, 0x7fffffff};
    for_cols([&](int c, float s) {
      float f;
      const Top x{filtered(c, s, f), c};
      mn = fminf(mn, s);
      mx = fmaxf(mx, s);
      if (better(x, t1)) {
        t2 = t1;
        t1 = x;
      } else if (better(x, t2)) {
        t2 = x;
      }
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Top a1{__shfl_xor_sync(0xffffffffu, t1.v, o), __shfl_xor_sync(0xffffffffu, t1.i, o)};
      Top a2{__shfl_xor_sync(0xffffffffu, t2.v, o), __shfl_xor_sync(0xffffffffu, t2.i, o)};
      // merge two sorted pairs
      if (better(a1, t1)) {
        t2 = better(t1, a2) ? t1 : a2;
        t1 = a1;
      } else {
        t2 = better(a1, t2) ? a1 : t2;
      }
    }
    const Top neg = (t1.v == pos) ? t2 : t1;
    const float x = (neg.v - pos) * invT;
    loss += w_ib * (fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))));   // softplus           (:313)
    if (p.grad != nullptr) {
      const float sig = 1.f / (1.f + __expf(-x));
      float* g = p.grad + static_cast<int64_t>(b) * p.C;
      const int ni = neg.i < p.C ? neg.i : pidx;
      float fneg;
      (void)filtered(ni, __ldcg(row + ni) * inv, fneg);
      for_cols([&](int c, float) {
        float v = 0.f;
        if (c == neg.i) v += sig * invT * fneg * inv * invB;
        if (c == pidx) v -= sig * invT * inv * invB;
        g[c] = v;
      });
    }
  } else {
    // sigmoid loss: softplus(-s/T * m), m = +1 at the positive, -1 elsewhere; mean over B*C   (:452-465; the block
    // loop of bi_encoder_losses.py:400-416 visits every column once when C is a multiple of B)
    float* g = p.grad ? p.grad + static_cast<int64_t>(b) * p.C : nullptr;
    const float invBB = invB / static_cast<float>(p.C);
    float part = 0.f;  // per-lane partial sum, reduced below (the returned loss must stay warp-uniform)
    for_cols([&](int c, float s0) {
      float f;
      const float s = filtered(c, s0, f);
      mn = fminf(mn, s0);
      mx = fmaxf(mx, s0);
      const float msk = (c == pidx) ? 1.f : -1.f;
      const float z = -s * invT * msk;
      part += (fmaxf(z, 0.f) + log1pf(__expf(-fabsf(z)))) / static_cast<float>(p.C);
      if (g) g[c] = -msk * invT * f * inv * invBB / (1.f + __expf(-z));
    });
    loss += warp_sum_f(part);
  }

  if (has_neg) {
    // softplus((neg - pos) / T) over this query's own negatives, mean over B * n_neg          (:235-246, :381-392)
    const float* nrow = p.neg_scores + static_cast<int64_t>(b) * p.B * p.n_neg;
    float* gn = p.grad_neg ? p.grad_neg + static_cast<int64_t>(b) * p.B * p.n_neg : nullptr;
    const float scale = w_out / (static_cast<float>(p.B) * static_cast<float>(p.n_neg));
    const int nidx = pidx + p.neg_pos_delta;
    const float npos = p.neg_pos_delta ? __ldcg(row + nidx) * inv : pos;
    float gpos = 0.f, part = 0.f;
    for (int c = lane; c < p.B * p.n_neg; c += 32) {
      float gv = 0.f;
      if (c / p.n_neg == b) {
        const float x = (__ldcg(nrow + c) * inv - npos) * invT;
        part += (fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)))) * w_out / static_cast<float>(p.n_neg);
        gv = scale * invT * inv / (1.f + __expf(-x));
        gpos -= gv;
      }
      if (gn) gn[c] = gv;
    }
    gpos = warp_sum_f(gpos);
    loss += warp_sum_f(part);
    if (p.grad != nullptr && lane == (nidx & 31)) p.grad[static_cast<int64_t>(b) * p.C + nidx] += gpos;
  }
  return loss;
}

// Column half of the symmetric loss (mode 3, bi_encoder_losses.py:166-168): cross entropy over the QUERIES b of
// scores[b, c] / T against b == c, weight 1/2, gradient ADDED to what the row pass wrote.  The in-place filter of :160-161
// was applied row-wise (threshold from row b's positive) before both halves, so it is re-derived per element here.
__device__ __forceinline__ float symmetric_loss_col(const LossParams& p, int c, int lane) {
  const float invT = 1.f / p.temperature;
  auto filtered = [&](int b, float& f) {
    const float* row = p.scores + static_cast<int64_t>(b) * p.C;
    const float s = __ldcg(row + c);
    const int pidx = b + p.offset;
    f = (p.filter && c != pidx && s > p.filter_threshold * __ldcg(row + pidx)) ? p.filter_factor : 1.f;
    return s * f;
  };
  float m = -INFINITY;
  for (int b = lane; b < p.B; b += 32) {
    float f;
    m = fmaxf(m, filtered(b, f) * invT);
  }
  m = warp_max_f(m);
  float se = 0.f;
  for (int b = lane; b < p.B; b += 32) {
    float f;
    se += __expf(filtered(b, f) * invT - m);
  }
  se = warp_sum_f(se);
  const float lse = m + __logf(se);
  float ftgt;
  const float tgt = filtered(c, ftgt) * invT;
  if (p.grad != nullptr) {
    const float w = 0.5f / static_cast<float>(p.C);
    for (int b = lane; b < p.B; b += 32) {
      float f;
      const float sm = __expf(filtered(b, f) * invT - lse);
      p.grad[static_cast<int64_t>(b) * p.C + c] += (sm - (b == c ? 1.f : 0.f)) * invT * f * w;
    }
  }
  return 0.5f * (lse - tgt);
}

// Whole-CTA device function (every thread of the block must call it: it ends with a __syncthreads reduction).
// Used by the stand-alone colbert_loss_kernel (loss_sm100.cu) and by the last CTA of the fused MaxSim kernels, which
// read the score matrix other CTAs have just written (hence the L2 loads).
__device__ __forceinline__ void colbert_loss_body(const LossParams& p) {
  __shared__ float s_loss[32];
  __shared__ float s_min[32];
  __shared__ float s_max[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float loss_acc = 0.f, mn = INFINITY, mx = -INFINITY;
  if (p.C <= 32 * kLossCols) {
    for (int b = warp; b < p.B; b += nwarps) loss_acc += colbert_loss_row<true>(p, b, lane, mn, mx);
  } else {
    for (int b = warp; b < p.B; b += nwarps) loss_acc += colbert_loss_row<false>(p, b, lane, mn, mx);
  }
  if (p.mode == 3) {  // C == B: the column terms join the same sum / B
    __syncthreads();  // the row pass's gradient stores (this CTA's own) before the column pass adds to them
    for (int c = warp; c < p.C; c += nwarps) loss_acc += symmetric_loss_col(p, c, lane);
  }

  // mean over the batch (CrossEntropyLoss default reduction / .mean())
  mn = warp_min_f(mn);
  mx = warp_max_f(mx);
  if (lane == 0) {
    s_loss[warp] = loss_acc;
    s_min[warp] = mn;
    s_max[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    float l = (lane < nwarps) ? s_loss[lane] : 0.f;
    float a = (lane < nwarps) ? s_min[lane] : INFINITY;
    float z = (lane < nwarps) ? s_max[lane] : -INFINITY;
    l = warp_sum_f(l);
    a = warp_min_f(a);
    z = warp_max_f(z);
    if (lane == 0) {
      p.loss[0] = l / static_cast<float>(p.B);
      if (p.bounds != nullptr) {
        p.bounds[0] = a;
        p.bounds[1] = z;
      }
    }
  }
}

// Rows [b0, b1) only: their loss sum and the min / max of their normalised scores -> partial3[0..2] (written by one
// thread).  Whole-CTA function like colbert_loss_body.  The fused MaxSim kernels call it from the last CTA of every
// query-tile GROUP (the CTAs that produced those rows' scores), so the row work of a batch runs on q_groups SMs at once
// instead of on the single last CTA of the grid; colbert_loss_combine then folds the partials.
__device__ __forceinline__ void colbert_loss_rows_partial(const LossParams& p, int b0, int b1, float* partial3) {
  __shared__ float s_loss[32];
  __shared__ float s_min[32];
  __shared__ float s_max[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float loss_acc = 0.f, mn = INFINITY, mx = -INFINITY;
  if (p.C <= 32 * kLossCols) {
    for (int b = b0 + warp; b < b1; b += nwarps) loss_acc += colbert_loss_row<true>(p, b, lane, mn, mx);
  } else {
    for (int b = b0 + warp; b < b1; b += nwarps) loss_acc += colbert_loss_row<false>(p, b, lane, mn, mx);
  }
  mn = warp_min_f(mn);
  mx = warp_max_f(mx);
  if (lane == 0) {
    s_loss[warp] = loss_acc;
    s_min[warp] = mn;
    s_max[warp] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    float l = (lane < nwarps) ? s_loss[lane] : 0.f;
    float a = (lane < nwarps) ? s_min[lane] : INFINITY;
    float z = (lane < nwarps) ? s_max[lane] : -INFINITY;
    l = warp_sum_f(l);
    a = warp_min_f(a);
    z = warp_max_f(z);
    if (lane == 0) {
      partial3[0] = l;
      partial3[1] = a;
      partial3[2] = z;
    }
  }
}

// One warp: loss = (sum of the n_groups partial sums, in group order within each lane, fixed butterfly across lanes) / B.
__device__ __forceinline__ void colbert_loss_combine(const LossParams& p, const float* partials, int n_groups, int lane) {
  float l = 0.f, a = INFINITY, z = -INFINITY;
  for (int g = lane; g < n_groups; g += 32) {
    l += __ldcg(partials + 3 * g);
    a = fminf(a, __ldcg(partials + 3 * g + 1));
    z = fmaxf(z, __ldcg(partials + 3 * g + 2));
  }
  l = warp_sum_f(l);
  a = warp_min_f(a);
  z = warp_max_f(z);
  if (lane == 0) {
    p.loss[0] = l / static_cast<float>(p.B);
    if (p.bounds != nullptr) {
      p.bounds[0] = a;
      p.bounds[1] = z;
    }
  }
}

}  // namespace cpb
