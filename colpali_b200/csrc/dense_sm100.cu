// Dense dot products out[i, j] = alpha * sum_k A[i, k] * B[row(j), k] in fp32 on the CUDA cores:
//   score_single_vector     torch.einsum("bd,cd->bc")   colpali_engine/utils/processing_utils.py:126
//   bi-encoder loss scores  torch.einsum("bd,cd->bc")   colpali_engine/loss/bi_encoder_losses.py:105, :158, :290, :396
//                           and their backward products  dQ = G D, dD = G^T Q   (strided operands: no transposes in HBM)
//   similarity maps         torch.einsum("nk,ijk->nij") colpali_engine/interpretability/similarity_map_utils.py:50-52
//                           (the mask selection :43 and the "(h w) c -> w h c" regrouping :42-47 are the row index b_rows)
// The Bi* models score ONE hidden-size vector per query / page (1536+ dims, fp32 or bf16): the contraction is a small
// fp32 GEMM, latency bound at the batch sizes of the losses (64 x 512 x 1536 = 0.1 GFLOP) -- no tensor cores, inputs are
// never demoted (fp32 stays fp32, bf16 is widened exactly), accumulation in fp32.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "dense_params.h"

namespace cpb {
namespace {

constexpr int kThreads = 256;  // 16 x 16 threads, each R x R outputs (rows ty + 16 i, columns tx + 16 j)
constexpr int kKC = 32;        // k values per shared-memory stage

__device__ __forceinline__ float ld_as_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_f32(const __nv_bfloat16* p) {
  return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

// One operand tile [TILE rows x kKC] -> smem[kk][row] as fp32.  Threads run along k when k is the contiguous axis of the
// operand (coalesced 128-byte rows), along the rows otherwise (transposed operands of the backward products).
template <int TILE, typename T>
__device__ __forceinline__ void load_tile(float (*sm)[TILE + 1], const T* base, const int32_t* rows, int64_t rs, int64_t ks,
                                          int row0, int n_rows, int k0, int k) {
  const bool k_fast = (ks == 1);
#pragma unroll
  for (int e = threadIdx.x; e < TILE * kKC; e += kThreads) {
    const int r = k_fast ? e / kKC : e % TILE;
    const int kk = k_fast ? e % kKC : e / TILE;
    float v = 0.f;
    if (row0 + r < n_rows && k0 + kk < k) {
      const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + row0 + r)) : static_cast<int64_t>(row0 + r);
      v = ld_as_f32(base + row * rs + static_cast<int64_t>(k0 + kk) * ks);
    }
    sm[kk][r] = v;
  }
}

template <int R, typename TA, typename TB>
__global__ void __launch_bounds__(kThreads) dense_dot_kernel(const DenseDotParams p) {
  constexpr int TILE = 16 * R;
  __shared__ float sa[kKC][TILE + 1];
  __shared__ float sb[kKC][TILE + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * TILE, j0 = blockIdx.x * TILE;
  const TA* a = static_cast<const TA*>(p.a);
  const TB* b = static_cast<const TB*>(p.b);
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.k; k0 += kKC) {
    load_tile<TILE, TA>(sa, a, nullptr, p.a_rs, p.a_ks, i0, p.m, k0, p.k);
    load_tile<TILE, TB>(sb, b, p.b_rows, p.b_rs, p.b_ks, j0, p.n, k0, p.k);
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kKC; ++kk) {
      float av[R], bv[R];
#pragma unroll
      for (int i = 0; i < R; ++i) av[i] = sa[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < R; ++j) bv[j] = sb[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int row = i0 + ty + 16 * i;
    if (row >= p.m) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int col = j0 + tx + 16 * j;
      if (col >= p.n) continue;
      float* o = p.out + static_cast<int64_t>(row) * p.out_rs + col;
      const float v = alpha * acc[i][j];
      *o = p.accumulate ? *o + v : v;
    }
  }
}

template <int R, typename TA, typename TB>
cudaError_t launch(const DenseDotParams& p, cudaStream_t stream) {
  constexpr int TILE = 16 * R;
  const dim3 grid((p.n + TILE - 1) / TILE, (p.m + TILE - 1) / TILE);
  dense_dot_kernel<R, TA, TB><<<grid, kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

template <int R>
cudaError_t launch_typed(const DenseDotParams& p, cudaStream_t stream) {
  using bf16 = __nv_bfloat16;
  if (p.a_f32) return p.b_f32 ? launch<R, float, float>(p, stream) : launch<R, float, bf16>(p, stream);
  return p.b_f32 ? launch<R, bf16, float>(p, stream) : launch<R, bf16, bf16>(p, stream);
}

}  // namespace

cudaError_t dense_dot_launch(const DenseDotParams& p, cudaStream_t stream) {
  // 64 x 64 tiles (4 x 4 per thread) once they fill the 148 SMs twice over, 32 x 32 tiles below (the loss shapes)
  const int64_t big_tiles = (static_cast<int64_t>(p.m + 63) / 64) * ((p.n + 63) / 64);
  const bool big = big_tiles >= 296;
  return big ? launch_typed<4>(p, stream) : launch_typed<2>(p, stream);
}

}  // namespace cpb
