// Backward of the smooth-max (tau * logsumexp) MaxSim aggregation -- ColbertModule._smooth_max,
// colpali_engine/loss/late_interaction_losses.py:40-44, reached through use_smooth_max=True (:88-90).
//
//   score[b, c] = sum_n  tau * log sum_s exp(<q[b,n], d[c,s]> / tau)
//   P[c, row, s] = exp((<q_row, d_s> - lse[c, row]) / tau)                 (softmax over the document's tokens)
//   dq[row]       = sum_c g[b(row), c] * sum_s P[c, row, s] * d[c, s]
//   dd[c, s]      = sum_row g[b(row), c] * P[c, row, s] * q[row]
//
// The reference differentiates through a saved [B, C, N_q, N_d] tensor; here the forward saves lse [C, rows] and the
// similarity tiles are RECOMPUTED, flash-attention style: per 64 x 64 tile S = A B^T on the tensor cores, P in the
// accumulator registers, converted in place into the A operand of the second product (P D or P^T Q).
//
// This is a training-only option of the reference (off in every shipped config), so these two kernels use the
// warp-level mma.sync path (m16n8k16 bf16, fp32 accumulate) with operands staged in shared memory -- one 64-row
// tile per 4-warp CTA, no TMA / tcgen05 pipeline.  dim = 128 only.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "loss_params.h"

namespace cpb {
namespace {

constexpr int kTile = 64;          // rows per tile (query rows or document tokens)
constexpr int kDim = 128;
constexpr int kLd = kDim + 8;      // shared-memory row stride in bf16: 272 B -> conflict-free fragment loads / ldmatrix
constexpr int kThreads = 128;      // 4 warps x 16 rows

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four 8x8 b16 matrices, transposed on the way: thread (g = lane / 4, t = lane % 4) receives M_j[2t][g], M_j[2t+1][g]
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __nv_bfloat16* row_ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(row_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// rows [row0, row0 + 64) of a [n_rows, 128] bf16 matrix -> shared tile (rows past `row_end` are zero)
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t row0, int64_t row_end) {
  for (int i = threadIdx.x; i < kTile * (kDim / 8); i += kThreads) {
    const int r = i / (kDim / 8), ch = i % (kDim / 8);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < row_end) v = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * kDim) + ch);
    *reinterpret_cast<uint4*>(dst + r * kLd + ch * 8) = v;
  }
}

// A-operand fragments (16 rows x 128, row-major in global memory) of the two rows this thread owns
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[8][4], const __nv_bfloat16* m, int64_t r_lo, int64_t r_hi,
                                             int64_t n_rows, int t) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int col = kk * 16 + 2 * t;
    a[kk][0] = (r_lo < n_rows) ? __ldg(reinterpret_cast<const uint32_t*>(m + r_lo * kDim + col)) : 0u;
    a[kk][1] = (r_hi < n_rows) ? __ldg(reinterpret_cast<const uint32_t*>(m + r_hi * kDim + col)) : 0u;
    a[kk][2] = (r_lo < n_rows) ? __ldg(reinterpret_cast<const uint32_t*>(m + r_lo * kDim + col + 8)) : 0u;
    a[kk][3] = (r_hi < n_rows) ? __ldg(reinterpret_cast<const uint32_t*>(m + r_hi * kDim + col + 8)) : 0u;
  }
}

// S[16 x 64] = A[16 x 128] * tile[64 x 128]^T  (B operand = the tile's rows, "col-major" for mma: k contiguous per n)
__device__ __forceinline__ void gemm_s(float (&s)[8][4], const uint32_t (&a)[8][4], const __nv_bfloat16* tile, int g, int t) {
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const __nv_bfloat16* bp = tile + (nb * 8 + g) * kLd + kk * 16 + 2 * t;
      mma_bf16(s[nb], a[kk], *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
    }
  }
}

// acc[16 x 128] += P[16 x 64] * tile[64 x 128]   (P already packed as A fragments; B read transposed by ldmatrix)
__device__ __forceinline__ void gemm_pv(float (&acc)[16][4], const uint32_t (&pa)[4][4], const __nv_bfloat16* tile, int lane) {
  const int lr = (lane & 7) + ((lane >> 3) & 1) * 8;  // row of the 16-row k slice this lane addresses
  const int lc = (lane >> 4) * 8;                     // first or second 8-column block
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int nb = 0; nb < 16; nb += 2) {
      uint32_t r[4];
      ldmatrix_x4_trans(r, tile + (kk * 16 + lr) * kLd + nb * 8 + lc);
      mma_bf16(acc[nb], pa[kk], r[0], r[1]);
      mma_bf16(acc[nb + 1], pa[kk], r[2], r[3]);
    }
  }
}

// ---- dq: one CTA per (64 query rows, slice of the documents) --------------------------------------------------------
__global__ void __launch_bounds__(kThreads) smooth_bwd_dq_kernel(const BwdParams p, int doc_splits) {
  __shared__ __align__(16) __nv_bfloat16 tile[kTile * kLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t r_lo = static_cast<int64_t>(blockIdx.x) * kTile + warp * 16 + g, r_hi = r_lo + 8;
  const int c0 = static_cast<int>(static_cast<int64_t>(p.C) * blockIdx.y / doc_splits);
  const int c1 = static_cast<int>(static_cast<int64_t>(p.C) * (blockIdx.y + 1) / doc_splits);
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  const bool ok_lo = r_lo < p.q_rows && (r_lo % p.nq_pad) < p.nq_real;
  const bool ok_hi = r_hi < p.q_rows && (r_hi % p.nq_pad) < p.nq_real;
  const int b_lo = static_cast<int>(r_lo / p.nq_pad), b_hi = static_cast<int>(r_hi / p.nq_pad);

  uint32_t a[8][4];
  load_a_frags(a, p.q, r_lo, r_hi, p.q_rows, t);
  float acc[16][4];
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;

  for (int c = c0; c < c1; ++c) {
    const int64_t start = __ldg(p.doc_start + c);
    const int len = __ldg(p.doc_len + c);
    const float w_lo = ok_lo ? __ldg(p.g + static_cast<int64_t>(b_lo) * p.C + c) * scale : 0.f;
    const float w_hi = ok_hi ? __ldg(p.g + static_cast<int64_t>(b_hi) * p.C + c) * scale : 0.f;
    // exponent offsets in base-2 units: P = 2^(c * s - c * lse)
    const float l_lo = ok_lo ? __ldg(p.lse + static_cast<int64_t>(c) * p.q_rows + r_lo) * p.smooth_c : INFINITY;
    const float l_hi = ok_hi ? __ldg(p.lse + static_cast<int64_t>(c) * p.q_rows + r_hi) * p.smooth_c : INFINITY;
    for (int t0 = 0; t0 < len; t0 += kTile) {
      __syncthreads();  // the previous tile has been consumed by every warp
      load_tile(tile, p.docs, start + t0, start + len);
      __syncthreads();
      float s[8][4];
      gemm_s(s, a, tile, g, t);
      // rows of the tile past the document's end are zero in shared memory, so whatever P holds there multiplies 0
      uint32_t pa[4][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        // rows that do not count carry w = 0 and an exponent offset of +inf: 2^(-inf) = 0, no branch, no 0 * inf
        const float p0 = w_lo * ex2(fmaf(s[nb][0], p.smooth_c, -l_lo));
        const float p1 = w_lo * ex2(fmaf(s[nb][1], p.smooth_c, -l_lo));
        const float p2 = w_hi * ex2(fmaf(s[nb][2], p.smooth_c, -l_hi));
        const float p3 = w_hi * ex2(fmaf(s[nb][3], p.smooth_c, -l_hi));
        pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
        pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      gemm_pv(acc, pa, tile, lane);
    }
  }
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    const int col = nb * 8 + 2 * t;
    if (doc_splits == 1) {
      if (r_lo < p.q_rows) *reinterpret_cast<float2*>(p.dq + r_lo * kDim + col) = make_float2(acc[nb][0], acc[nb][1]);
      if (r_hi < p.q_rows) *reinterpret_cast<float2*>(p.dq + r_hi * kDim + col) = make_float2(acc[nb][2], acc[nb][3]);
    } else {  // several document slices add into the same rows (dq zeroed by the launcher)
      if (r_lo < p.q_rows) atomicAdd(reinterpret_cast<float2*>(p.dq + r_lo * kDim + col), make_float2(acc[nb][0], acc[nb][1]));
      if (r_hi < p.q_rows) atomicAdd(reinterpret_cast<float2*>(p.dq + r_hi * kDim + col), make_float2(acc[nb][2], acc[nb][3]));
    }
  }
}

// ---- dd: one CTA per 64 bank rows, looping over all query rows; every gradient row is written once ------------------
__device__ __forceinline__ int doc_of_row(const BwdParams& p, int64_t row) {
  if (row >= p.doc_rows) return -1;
  int lo = 0, hi = p.C - 1;  // last document starting at or before `row` (starts ascending)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(p.doc_start + mid) <= row) lo = mid; else hi = mid - 1;
  }
  const int64_t s = __ldg(p.doc_start + lo);
  return (row >= s && row < s + __ldg(p.doc_len + lo)) ? lo : -1;
}

__global__ void __launch_bounds__(kThreads) smooth_bwd_dd_kernel(const BwdParams p) {
  __shared__ __align__(16) __nv_bfloat16 tile[kTile * kLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t r_lo = static_cast<int64_t>(blockIdx.x) * kTile + warp * 16 + g, r_hi = r_lo + 8;
  const float scale = p.grad_out ? *p.grad_out : 1.f;
  const int c_lo = doc_of_row(p, r_lo), c_hi = doc_of_row(p, r_hi);

  uint32_t a[8][4];
  load_a_frags(a, p.docs, r_lo, r_hi, p.doc_rows, t);
  float acc[16][4];
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;

  for (int q0 = 0; q0 < p.q_rows; q0 += kTile) {
    __syncthreads();
    load_tile(tile, p.q, q0, p.q_rows);
    __syncthreads();
    float s[8][4];
    gemm_s(s, a, tile, g, t);  // S^T[token, query row]
    uint32_t pa[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int qrow = q0 + nb * 8 + 2 * t;  // and qrow + 1: same query (nq_pad is even)
      const bool okq0 = qrow < p.q_rows && (qrow % p.nq_pad) < p.nq_real;
      const bool okq1 = qrow + 1 < p.q_rows && ((qrow + 1) % p.nq_pad) < p.nq_real;
      const int b = min(qrow, p.q_rows - 1) / p.nq_pad;
      const int qr = min(qrow, p.q_rows - 2);  // clamped address of the (even-aligned) lse pair
      // select the operands, then compute unconditionally: invalid (row, column) pairs get weight 0 and exponent -inf
      const int cl = max(c_lo, 0), ch = max(c_hi, 0);
      const float w_l = (c_lo >= 0) ? __ldg(p.g + static_cast<int64_t>(b) * p.C + cl) * scale : 0.f;
      const float w_h = (c_hi >= 0) ? __ldg(p.g + static_cast<int64_t>(b) * p.C + ch) * scale : 0.f;
      const float2 ll = __ldg(reinterpret_cast<const float2*>(p.lse + static_cast<int64_t>(cl) * p.q_rows + qr));
      const float2 lh = __ldg(reinterpret_cast<const float2*>(p.lse + static_cast<int64_t>(ch) * p.q_rows + qr));
      const float e0 = (c_lo >= 0 && okq0) ? (s[nb][0] - ll.x) * p.smooth_c : -INFINITY;
      const float e1 = (c_lo >= 0 && okq1) ? (s[nb][1] - ll.y) * p.smooth_c : -INFINITY;
      const float e2 = (c_hi >= 0 && okq0) ? (s[nb][2] - lh.x) * p.smooth_c : -INFINITY;
      const float e3 = (c_hi >= 0 && okq1) ? (s[nb][3] - lh.y) * p.smooth_c : -INFINITY;
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(w_l * ex2(e0), w_l * ex2(e1));
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(w_h * ex2(e2), w_h * ex2(e3));
    }
    gemm_pv(acc, pa, tile, lane);  // rows of the tile past q_rows are zero
  }
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    const int col = nb * 8 + 2 * t;
    if (r_lo < p.doc_rows) *reinterpret_cast<float2*>(p.dd + r_lo * kDim + col) = make_float2(acc[nb][0], acc[nb][1]);
    if (r_hi < p.doc_rows) *reinterpret_cast<float2*>(p.dd + r_hi * kDim + col) = make_float2(acc[nb][2], acc[nb][3]);
  }
}

}  // namespace

cudaError_t smooth_bwd_launch(const BwdParams& p, cudaStream_t stream) {
  if (p.dim != kDim) return cudaErrorInvalidValue;
  if (p.dq != nullptr) {
    const int row_blocks = (p.q_rows + kTile - 1) / kTile;
    int splits = (4 * 148 + row_blocks - 1) / row_blocks;  // about four CTAs per SM
    if (splits > p.C) splits = p.C;
    if (splits < 1) splits = 1;
    if (splits > 1) {
      cudaError_t e = cudaMemsetAsync(p.dq, 0, static_cast<size_t>(p.q_rows) * kDim * sizeof(float), stream);
      if (e != cudaSuccess) return e;
    }
    smooth_bwd_dq_kernel<<<dim3(row_blocks, splits), kThreads, 0, stream>>>(p, splits);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (p.dd != nullptr) {
    const int64_t blocks = (p.doc_rows + kTile - 1) / kTile;
    smooth_bwd_dd_kernel<<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace cpb
