// ColBERT in-batch-negative losses on top of the fused MaxSim scores, and the MaxSim backward.
//
//   colbert_loss_kernel   : [B, C] raw MaxSim sums -> scalar loss AND dLoss/dScores in one pass
//                           replaces late_interaction_losses.py:152 (lengths), :155-156 / :46-71 (normalise),
//                           :161-162 / :93-107 (pos-aware negative filtering), :164 (cross entropy, ColbertLoss)
//                           and :309-313 (top-2 / softplus, ColbertPairwiseCELoss).
//   maxsim_bwd_dq_kernel  : dQ[row] = sum_c g[b(row), c] * D[start_c + argmax[c, row]]            (gather)
//   maxsim_bwd_dd_kernel  : dD[start_c + s] = sum_{row: argmax[c, row] == s} g[b(row), c] * Q[row]  (per-document
//                           counting sort in shared memory, every gradient row written once, no atomics)
//                           the reference gets these from autograd through einsum/amax with a saved
//                           [B, C, N_q, N_d] tensor; here only the int32 argmax [C, rows] is saved.
//   (the smooth-max backward, which has to recompute the similarity tiles, is in smooth_bwd_sm100.cu)
//
// These kernels move a few MB and are latency bound; they are plain CUDA (no tensor cores needed).
#include <cfloat>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
#include "loss_body.cuh"
#include "loss_params.h"

namespace cpb {

__global__ void __launch_bounds__(1024, 1) colbert_loss_kernel(const LossParams p) { colbert_loss_body(p); }

cudaError_t colbert_loss_launch(const LossParams& p, cudaStream_t stream) {
  colbert_loss_kernel<<<1, 1024, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward of scores[b, c] = sum_n max_s <q[b,n], d[c,s]>  given g = dL/dscores (hard max: saved argmax)
// ------------------------------------------------------------------------------------------------
// Embedding dim = 64 * P (P = 2..5): a lane owns the bf16 pairs {lane, lane + 32, ...} of a row.

// one gradient row of 64 P elements, two per lane and step: fp32, or bf16 rounded to nearest even (what a cast of the
// fp32 row would give)
template <int P>
__device__ __forceinline__ void store_grad_row(float* base, int out_bf16, int64_t row, int lane, const float2 (&acc)[P]) {
  if (out_bf16) {
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(base) + row * (32 * P);
#pragma unroll
    for (int j = 0; j < P; ++j) dst[j * 32 + lane] = __float22bfloat162_rn(acc[j]);
  } else {
    float2* dst = reinterpret_cast<float2*>(base) + row * (32 * P);
#pragma unroll
    for (int j = 0; j < P; ++j) dst[j * 32 + lane] = acc[j];
  }
}

// dQ: one warp per query row, a gather over the C winning document tokens.  The (index, weight, start) triples of 32
// documents are fetched by the 32 lanes at once and broadcast by shuffles, and four row gathers are in flight before the
// first is consumed: the loop is a chain of dependent 256-byte loads otherwise (C = 64..512 iterations).
template <int P>
__device__ __forceinline__ void maxsim_bwd_dq_body(const BwdParams& p, int block) {
  const int row = block * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.q_rows) return;
  constexpr int kDim = 64 * P;
  const int b = row / p.nq_pad;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  float2 acc[P];
#pragma unroll
  for (int j = 0; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
  const float* g = p.g + static_cast<int64_t>(b) * p.C;
  auto src_of = [&](int64_t start, int idx) {
    return reinterpret_cast<const uint32_t*>(p.docs + (start + idx) * kDim);
  };
  auto fma_row = [&](const uint32_t (&raw)[P], float w) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&raw[j]);
      acc[j].x = fmaf(w, __low2float(v), acc[j].x);
      acc[j].y = fmaf(w, __high2float(v), acc[j].y);
    }
  };
  for (int c0 = 0; c0 < p.C; c0 += 32) {
    const int c = c0 + lane;
    int idx = -1;
    float w = 0.f;
    int start = 0;
    if (c < p.C) {
      idx = __ldg(p.argmax + static_cast<int64_t>(c) * p.q_rows + row);
      w = __ldg(g + c) * scale;  // block-diagonal gradients of the explicit-negative losses are mostly zero
      start = __ldg(p.doc_start + c);
    }
    unsigned act = __ballot_sync(0xffffffffu, idx >= 0 && w != 0.f);  // ascending document order: deterministic sums
    while (act != 0u) {
      int k[4];
      float wk[4];
      uint32_t raw[4][P];
      int n = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        k[u] = act ? __ffs(act) - 1 : -1;
        if (act) { act &= act - 1; ++n; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = k[u] < 0 ? 0 : k[u];
        const int idx_k = __shfl_sync(0xffffffffu, idx, kk);
        const int start_k = __shfl_sync(0xffffffffu, start, kk);
        wk[u] = (u < n) ? __shfl_sync(0xffffffffu, w, kk) : 0.f;
        if (u < n) {
          const uint32_t* src = src_of(start_k, idx_k);
#pragma unroll
          for (int j = 0; j < P; ++j) raw[u][j] = __ldg(src + j * 32 + lane);
        } else {
#pragma unroll
          for (int j = 0; j < P; ++j) raw[u][j] = 0u;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) fma_row(raw[u], wk[u]);
    }
  }
  store_grad_row<P>(p.dq, p.out_bf16, row, lane, acc);
}

// dD without atomics: a CTA owns tokens [t0, t0 + kDdTokens) of ONE document.  It buckets the (query row -> winning
// token) pairs of that document by token in shared memory (counting sort: histogram, scan, fill), then every warp sums
// the query rows of its tokens and WRITES the gradient row once -- rows nobody points at are written as zeros, so the
// caller does not pre-zero the [doc_rows, dim] fp32 buffer (34 MB at cfg3) and nothing is read-modify-written.
// Buckets of at most 32 rows are summed in query-row order (deterministic); larger ones in arrival order.
// The kernel is a chain of dependent small loads (index -> weight -> query row), so it is sized for latency: 64 tokens
// per CTA (8 per warp, ~1 100 CTAs at cfg3 = every SM full), index loads and row gathers issued in pairs.
constexpr int kDdTokens = 64;
constexpr int kDdThreads = 256;

template <int P>
__device__ __forceinline__ void maxsim_bwd_dd_body(const BwdParams& p, int c, int token_block) {
  extern __shared__ int s_list[];  // [q_rows] query rows grouped by token
  __shared__ int s_cnt[kDdTokens];
  __shared__ int s_off[kDdTokens + 1];
  constexpr int kDim = 64 * P;
  const int t0 = token_block * kDdTokens;
  const int len = __ldg(p.doc_len + c);
  if (t0 >= len) return;
  const int nt = min(kDdTokens, len - t0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  const int32_t* am = p.argmax + static_cast<int64_t>(c) * p.q_rows;
  if (tid < kDdTokens) s_cnt[tid] = 0;
  __syncthreads();
  // bucket of a query row: its winning token if that lies in my range and the row's gradient weight is non-zero
  auto bucket_of = [&](int row) {
    const int idx = __ldg(am + row) - t0;
    return (idx >= 0 && idx < nt && __ldg(p.g + static_cast<int64_t>(row / p.nq_pad) * p.C + c) != 0.f) ? idx : -1;
  };
  // histogram (two rows per thread and iteration: independent load chains)
  for (int row = tid; row < p.q_rows; row += 2 * kDdThreads) {
    const int b0 = bucket_of(row);
    const int b1 = (row + kDdThreads < p.q_rows) ? bucket_of(row + kDdThreads) : -1;
    if (b0 >= 0) atomicAdd(&s_cnt[b0], 1);
    if (b1 >= 0) atomicAdd(&s_cnt[b1], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 64 counters by one warp (two per lane)
    const int v0 = s_cnt[2 * lane], v1 = s_cnt[2 * lane + 1];
    int x = v0 + v1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    s_off[2 * lane] = x - v0 - v1;
    s_off[2 * lane + 1] = x - v1;
    if (lane == 31) s_off[kDdTokens] = x;
    s_cnt[2 * lane] = 0;  // reused as the fill cursors
    s_cnt[2 * lane + 1] = 0;
  }
  __syncthreads();
  for (int row = tid; row < p.q_rows; row += 2 * kDdThreads) {
    const int b0 = bucket_of(row);
    const int b1 = (row + kDdThreads < p.q_rows) ? bucket_of(row + kDdThreads) : -1;
    if (b0 >= 0) s_list[s_off[b0] + atomicAdd(&s_cnt[b0], 1)] = row;
    if (b1 >= 0) s_list[s_off[b1] + atomicAdd(&s_cnt[b1], 1)] = row + kDdThreads;
  }
  __syncthreads();
  const int64_t doc_row0 = static_cast<int64_t>(__ldg(p.doc_start + c)) + t0;
  for (int t = warp; t < nt; t += kDdThreads / 32) {
    const int lo = s_off[t], n = s_off[t + 1] - lo;
    float2 acc[P];
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] = make_float2(0.f, 0.f);
    auto load_row = [&](int row, float& w, uint32_t (&raw)[P]) {
      w = __ldg(p.g + static_cast<int64_t>(row / p.nq_pad) * p.C + c) * scale;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(p.q + static_cast<int64_t>(row) * kDim);
#pragma unroll
      for (int j = 0; j < P; ++j) raw[j] = __ldg(src + j * 32 + lane);
    };
    auto fma_row = [&](float w, const uint32_t (&raw)[P]) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&raw[j]);
        acc[j].x = fmaf(w, __low2float(v), acc[j].x);
        acc[j].y = fmaf(w, __high2float(v), acc[j].y);
      }
    };
    if (n <= 32) {
      const int mine = (lane < n) ? s_list[lo + lane] : 0x7fffffff;
      int rank = 0;
      for (int k = 0; k < n; ++k) rank += (__shfl_sync(0xffffffffu, mine, k) < mine) ? 1 : 0;
      // sorted[k] = the row of rank k, gathered into lane k
      int sorted = 0x7fffffff;
      for (int k = 0; k < n; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, lane < n && rank == k);
        const int r = __shfl_sync(0xffffffffu, mine, __ffs(m) - 1);
        if (lane == k) sorted = r;
      }
      for (int k = 0; k < n; k += 2) {  // two row gathers in flight
        float w0, w1 = 0.f;
        uint32_t r0[P], r1[P];
        load_row(__shfl_sync(0xffffffffu, sorted, k), w0, r0);
        const bool two = k + 1 < n;
        const int row1 = __shfl_sync(0xffffffffu, sorted, two ? k + 1 : k);
        if (two) load_row(row1, w1, r1);
        fma_row(w0, r0);
        if (two) fma_row(w1, r1);
      }
    } else {
      for (int k = 0; k < n; ++k) {
        float w;
        uint32_t r[P];
        load_row(s_list[lo + k], w, r);
        fma_row(w, r);
      }
    }
    if (p.dd_doc_base != nullptr) {
      // exchange mode: this document belongs to another rank's batch; several ranks add into the same rows
      if (n == 0) continue;
      float2* dst = reinterpret_cast<float2*>(__ldg(p.dd_doc_base + c)) + static_cast<int64_t>(t0 + t) * (kDim / 2);
#pragma unroll
      for (int j = 0; j < P; ++j) atomicAdd(dst + j * 32 + lane, acc[j]);  // red.global.add.v2.f32 over NVLink
    } else {
      store_grad_row<P>(p.dd, p.out_bf16, doc_row0 + t, lane, acc);
    }
  }
}

// ONE launch for both gradients: blocks [0, n_dd) are dD blocks (document = block % C, token block = block / C), the
// rest dQ blocks.  The two are independent chains of small dependent loads, so co-resident they overlap instead of
// running back to back (cfg3: 38 + 19 us as two launches); the longer-running dD blocks are scheduled first.
template <int P>
__global__ void __launch_bounds__(kDdThreads) maxsim_bwd_kernel(const BwdParams p, int n_dd) {
  const int block = static_cast<int>(blockIdx.x);
  if (block < n_dd) maxsim_bwd_dd_body<P>(p, block % p.C, block / p.C);
  else maxsim_bwd_dq_body<P>(p, block - n_dd);
}

template <int P>
static cudaError_t maxsim_bwd_launch_p(const BwdParams& p, cudaStream_t stream) {
  const int wpb = kDdThreads / 32;
  const bool want_dd = p.dd != nullptr || p.dd_doc_base != nullptr;
  const int n_dq = (p.dq != nullptr) ? (p.q_rows + wpb - 1) / wpb : 0;
  const int n_dd = want_dd ? p.C * ((p.max_doc_len + kDdTokens - 1) / kDdTokens) : 0;
  if (n_dq + n_dd == 0) return cudaSuccess;
  const size_t smem = want_dd ? static_cast<size_t>(p.q_rows) * sizeof(int) : 0;
  auto kern = maxsim_bwd_kernel<P>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  if (want_dd && !p.contiguous && p.dd_doc_base == nullptr) {  // rows between documents belong to nobody: zero them
    cudaError_t e = cudaMemsetAsync(p.dd, 0, static_cast<size_t>(p.doc_rows) * 64 * P * (p.out_bf16 ? 2 : 4), stream);
    if (e != cudaSuccess) return e;
  }
  kern<<<static_cast<unsigned>(n_dd + n_dq), kDdThreads, smem, stream>>>(p, n_dd);
  return cudaGetLastError();
}

cudaError_t maxsim_bwd_launch(const BwdParams& p, cudaStream_t stream) {
  switch (p.dim) {
    case 128: return maxsim_bwd_launch_p<2>(p, stream);
    case 192: return maxsim_bwd_launch_p<3>(p, stream);
    case 256: return maxsim_bwd_launch_p<4>(p, stream);
    case 320: return maxsim_bwd_launch_p<5>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cpb
