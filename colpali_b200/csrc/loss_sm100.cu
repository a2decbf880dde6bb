// ColBERT in-batch-negative losses on top of the fused MaxSim scores, and the MaxSim backward.
//
//   colbert_loss_kernel   : [B, C] raw MaxSim sums -> scalar loss AND dLoss/dScores in one pass
//                           replaces late_interaction_losses.py:152 (lengths), :155-156 / :46-71 (normalise),
//                           :161-162 / :93-107 (pos-aware negative filtering), :164 (cross entropy, ColbertLoss)
//                           and :309-313 (top-2 / softplus, ColbertPairwiseCELoss).
//   maxsim_bwd_dq_kernel  : dQ[row] = sum_c g[b(row), c] * D[start_c + argmax[c, row]]
//   maxsim_bwd_dd_kernel  : dD[start_c + argmax[c, row]] += g[b(row), c] * Q[row]
//                           the reference gets these from autograd through einsum/amax with a saved
//                           [B, C, N_q, N_d] tensor; here only the int32 argmax [C, rows] is saved.
//
// These kernels move a few MB and are latency bound; they are plain CUDA (no tensor cores needed).
#include <cfloat>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
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

__global__ void __launch_bounds__(1024, 1) colbert_loss_kernel(const LossParams p) {
  __shared__ float s_loss[32];
  __shared__ float s_min[32];
  __shared__ float s_max[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float loss_acc = 0.f, mn = INFINITY, mx = -INFINITY;
  const bool has_neg = p.neg_scores != nullptr;
  const float w_ib = has_neg ? p.in_batch_weight : 1.f;     // late_interaction_losses.py:248-250 / :394-396
  const float w_out = has_neg ? 1.f - p.in_batch_weight : 0.f;

  for (int b = warp; b < p.B; b += nwarps) {
    // lengths = (q[:, :, 0] != 0).sum(1)                       late_interaction_losses.py:152
    float cnt = 0.f;
    for (int n = lane; n < p.nq_pad; n += 32)
      cnt += (__bfloat162float(p.q[(static_cast<int64_t>(b) * p.nq_pad + n) * p.q_dim]) != 0.f) ? 1.f : 0.f;
    cnt = warp_sum_f(cnt);
    const float inv = p.normalize ? 1.f / cnt : 1.f;          // :155-156 -> :59-62
    const float* row = p.scores + static_cast<int64_t>(b) * p.C;
    const int pidx = b + p.offset;                              // :33-38
    const float pos = row[pidx] * inv;
    const float thr = p.filter_threshold * pos;                 // :101-104
    const float invT = 1.f / p.temperature;
    const float invB = w_ib / static_cast<float>(p.B);

    // filtered score of column c and the factor it was multiplied by      (:105-107)
    auto filtered = [&](int c, float& f) {
      float s = row[c] * inv;
      f = (p.filter && c != pidx && s > thr) ? p.filter_factor : 1.f;
      return s * f;
    };

    if (p.mode == 0) {
      // cross entropy of scores / T against pidx                (:164)
      float m = -INFINITY;
      for (int c = lane; c < p.C; c += 32) {
        float f;
        const float s = filtered(c, f);
        mn = fminf(mn, row[c] * inv);
        mx = fmaxf(mx, row[c] * inv);
        m = fmaxf(m, s * invT);
      }
      m = warp_max_f(m);
      float se = 0.f;
      for (int c = lane; c < p.C; c += 32) {
        float f;
        se += __expf(filtered(c, f) * invT - m);
      }
      se = warp_sum_f(se);
      const float lse = m + __logf(se);
      loss_acc += w_ib * (lse - pos * invT);  // the positive column is never filtered
      if (p.grad != nullptr) {
        float* g = p.grad + static_cast<int64_t>(b) * p.C;
        for (int c = lane; c < p.C; c += 32) {
          float f;
          const float s = filtered(c, f);
          const float sm = __expf(s * invT - lse);
          g[c] = (sm - (c == pidx ? 1.f : 0.f)) * invT * f * inv * invB;
        }
      }
    } else if (p.mode == 1) {
      // pos = diagonal(offset); top-2 of the row; neg = top1 == pos ? top2 : top1      (:309-311)
      Top t1{-INFINITY, 0x7fffffff}, t2{-INFINITY, 0x7fffffff};
      for (int c = lane; c < p.C; c += 32) {
        float f;
        const Top x{filtered(c, f), c};
        mn = fminf(mn, row[c] * inv);
        mx = fmaxf(mx, row[c] * inv);
        if (better(x, t1)) {
          t2 = t1;
          t1 = x;
        } else if (better(x, t2)) {
          t2 = x;
        }
      }
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
      loss_acc += w_ib * (fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))));   // softplus           (:313)
      if (p.grad != nullptr) {
        const float sig = 1.f / (1.f + __expf(-x));
        float* g = p.grad + static_cast<int64_t>(b) * p.C;
        float fneg;
        (void)filtered(neg.i < p.C ? neg.i : pidx, fneg);
        for (int c = lane; c < p.C; c += 32) {
          float v = 0.f;
          if (c == neg.i) v += sig * invT * fneg * inv * invB;
          if (c == pidx) v -= sig * invT * inv * invB;
          g[c] = v;
        }
      }
    } else {
      // sigmoid loss: softplus(-s/T * m), m = +1 on the diagonal, -1 elsewhere; mean over B*B   (:452-465)
      float* g = p.grad ? p.grad + static_cast<int64_t>(b) * p.C : nullptr;
      const float invBB = invB / static_cast<float>(p.C);
      float part = 0.f;  // per-lane partial sum, reduced below (loss_acc must stay warp-uniform)
      for (int c = lane; c < p.C; c += 32) {
        float f;
        const float s = filtered(c, f);
        mn = fminf(mn, row[c] * inv);
        mx = fmaxf(mx, row[c] * inv);
        const float msk = (c == b) ? 1.f : -1.f;
        const float z = -s * invT * msk;
        part += (fmaxf(z, 0.f) + log1pf(__expf(-fabsf(z)))) / static_cast<float>(p.C);
        if (g) g[c] = -msk * invT * f * inv * invBB / (1.f + __expf(-z));
      }
      loss_acc += warp_sum_f(part);
    }

    if (has_neg) {
      // softplus((neg - pos) / T) over this query's own negatives, mean over B * n_neg          (:235-246, :381-392)
      const float* nrow = p.neg_scores + static_cast<int64_t>(b) * p.B * p.n_neg;
      float* gn = p.grad_neg ? p.grad_neg + static_cast<int64_t>(b) * p.B * p.n_neg : nullptr;
      const float scale = w_out / (static_cast<float>(p.B) * static_cast<float>(p.n_neg));
      float gpos = 0.f, part = 0.f;
      for (int c = lane; c < p.B * p.n_neg; c += 32) {
        float gv = 0.f;
        if (c / p.n_neg == b) {
          const float x = (nrow[c] * inv - pos) * invT;
          part += (fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)))) * w_out / static_cast<float>(p.n_neg);
          gv = scale * invT * inv / (1.f + __expf(-x));
          gpos -= gv;
        }
        if (gn) gn[c] = gv;
      }
      gpos = warp_sum_f(gpos);
      loss_acc += warp_sum_f(part);
      if (p.grad != nullptr && lane == (pidx & 31)) p.grad[static_cast<int64_t>(b) * p.C + pidx] += gpos;
    }
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

cudaError_t colbert_loss_launch(const LossParams& p, cudaStream_t stream) {
  colbert_loss_kernel<<<1, 1024, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward of scores[b, c] = sum_n max_s <q[b,n], d[c,s]>  given g = dL/dscores
// ------------------------------------------------------------------------------------------------
// one warp per query row, lane owns dims [4*lane, 4*lane+4)
__global__ void __launch_bounds__(256) maxsim_bwd_dq_kernel(const BwdParams p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.q_rows) return;
  const int b = row / p.nq_pad;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* g = p.g + static_cast<int64_t>(b) * p.C;
#pragma unroll 4
  for (int c = 0; c < p.C; ++c) {
    const int idx = __ldg(p.argmax + static_cast<int64_t>(c) * p.q_rows + row);
    if (idx < 0) continue;
    const float w = __ldg(g + c) * scale;
    if (w == 0.f) continue;  // block-diagonal gradients of the explicit-negative losses are mostly zero
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.docs + (static_cast<int64_t>(__ldg(p.doc_start + c)) + idx) * 128) + lane);
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    a0 = fmaf(w, __low2float(lo), a0);
    a1 = fmaf(w, __high2float(lo), a1);
    a2 = fmaf(w, __low2float(hi), a2);
    a3 = fmaf(w, __high2float(hi), a3);
  }
  reinterpret_cast<float4*>(p.dq + static_cast<int64_t>(row) * 128)[lane] = make_float4(a0, a1, a2, a3);
}

// one warp per (document, query row): a 128-wide fp32 vector atomic add into the winning token's gradient
__global__ void __launch_bounds__(256) maxsim_bwd_dd_kernel(const BwdParams p) {
  const int64_t w = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<int64_t>(p.C) * p.q_rows) return;
  const int c = static_cast<int>(w / p.q_rows);
  const int row = static_cast<int>(w % p.q_rows);
  const int idx = __ldg(p.argmax + w);
  if (idx < 0) return;
  const int b = row / p.nq_pad;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  const float wgt = __ldg(p.g + static_cast<int64_t>(b) * p.C + c) * scale;
  if (wgt == 0.f) return;
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.q + static_cast<int64_t>(row) * 128) + lane);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  float4 v = make_float4(wgt * __low2float(lo), wgt * __high2float(lo), wgt * __low2float(hi), wgt * __high2float(hi));
  float4* dst = reinterpret_cast<float4*>(p.dd + (static_cast<int64_t>(__ldg(p.doc_start + c)) + idx) * 128) + lane;
  atomicAdd(dst, v);  // red.global.add.v4.f32 (sm_90+)
}

// ---- wide embeddings (dim = 64 * P, P in 3..5): lane owns the bf16 pairs {lane, lane + 32, ...} of a row --------
template <int P>
__global__ void __launch_bounds__(256) maxsim_bwd_dq_wide_kernel(const BwdParams p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.q_rows) return;
  constexpr int kDim = 64 * P;
  const int b = row / p.nq_pad;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  float2 acc[P];
#pragma unroll
  for (int j = 0; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
  const float* g = p.g + static_cast<int64_t>(b) * p.C;
#pragma unroll 2
  for (int c = 0; c < p.C; ++c) {
    const int idx = __ldg(p.argmax + static_cast<int64_t>(c) * p.q_rows + row);
    if (idx < 0) continue;
    const float w = __ldg(g + c) * scale;
    if (w == 0.f) continue;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.docs + (static_cast<int64_t>(__ldg(p.doc_start + c)) + idx) * kDim);
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const uint32_t raw = __ldg(src + j * 32 + lane);
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&raw);
      acc[j].x = fmaf(w, __low2float(v), acc[j].x);
      acc[j].y = fmaf(w, __high2float(v), acc[j].y);
    }
  }
  float2* dst = reinterpret_cast<float2*>(p.dq + static_cast<int64_t>(row) * kDim);
#pragma unroll
  for (int j = 0; j < P; ++j) dst[j * 32 + lane] = acc[j];
}

template <int P>
__global__ void __launch_bounds__(256) maxsim_bwd_dd_wide_kernel(const BwdParams p) {
  const int64_t w = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<int64_t>(p.C) * p.q_rows) return;
  constexpr int kDim = 64 * P;
  const int c = static_cast<int>(w / p.q_rows);
  const int row = static_cast<int>(w % p.q_rows);
  const int idx = __ldg(p.argmax + w);
  if (idx < 0) return;
  const int b = row / p.nq_pad;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  const float wgt = __ldg(p.g + static_cast<int64_t>(b) * p.C + c) * scale;
  if (wgt == 0.f) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(p.q + static_cast<int64_t>(row) * kDim);
  float2* dst = reinterpret_cast<float2*>(p.dd + (static_cast<int64_t>(__ldg(p.doc_start + c)) + idx) * kDim);
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const uint32_t raw = __ldg(src + j * 32 + lane);
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&raw);
    atomicAdd(dst + j * 32 + lane, make_float2(wgt * __low2float(v), wgt * __high2float(v)));  // red.global.add.v2.f32
  }
}

template <int P>
static cudaError_t maxsim_bwd_wide_launch(const BwdParams& p, cudaStream_t stream) {
  const int wpb = 8;
  if (p.dq != nullptr) {
    maxsim_bwd_dq_wide_kernel<P><<<(p.q_rows + wpb - 1) / wpb, wpb * 32, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (p.dd != nullptr) {
    const int64_t warps = static_cast<int64_t>(p.C) * p.q_rows;
    maxsim_bwd_dd_wide_kernel<P><<<static_cast<unsigned>((warps + wpb - 1) / wpb), wpb * 32, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t maxsim_bwd_launch(const BwdParams& p, cudaStream_t stream) {
  if (p.dim == 192) return maxsim_bwd_wide_launch<3>(p, stream);
  if (p.dim == 256) return maxsim_bwd_wide_launch<4>(p, stream);
  if (p.dim == 320) return maxsim_bwd_wide_launch<5>(p, stream);
  if (p.dim != 128) return cudaErrorInvalidValue;
  const int wpb = 8;  // warps per block
  if (p.dq != nullptr) {
    const int blocks = (p.q_rows + wpb - 1) / wpb;
    maxsim_bwd_dq_kernel<<<blocks, wpb * 32, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (p.dd != nullptr) {
    const int64_t warps = static_cast<int64_t>(p.C) * p.q_rows;
    const int64_t blocks = (warps + wpb - 1) / wpb;
    maxsim_bwd_dd_kernel<<<static_cast<unsigned>(blocks), wpb * 32, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace cpb
