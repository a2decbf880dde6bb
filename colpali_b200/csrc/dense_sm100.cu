// Dense dot products out[i, j] = alpha * sum_k A[i, k] * B[row(j), k] in fp32 on the CUDA cores:
//   score_single_vector     torch.einsum("bd,cd->bc")   colpali_engine/utils/processing_utils.py:126
//   bi-encoder loss scores  torch.einsum("bd,cd->bc")   colpali_engine/loss/bi_encoder_losses.py:105, :158, :290, :396
//                           and their backward products  dQ = G D, dD = G^T Q   (strided operands: no transposes in HBM)
//   similarity maps         torch.einsum("nk,ijk->nij") colpali_engine/interpretability/similarity_map_utils.py:50-52
//                           (the mask selection :43 and the "(h w) c -> w h c" regrouping :42-47 are the row index b_rows)
// The Bi* models score ONE hidden-size vector per query / page (1536+ dims, fp32 or bf16): operands are never demoted
// (fp32 stays fp32, bf16 is widened exactly), accumulation in fp32 -- so this is SIMT FFMA work, not tensor-core work.
//
// Two shapes matter and they want different kernels:
//   * the losses: 64 x 512 x 1536 and its backward products -- 0.1 GFLOP, pure latency.  dense_tile_kernel<64, ...>
//     with the K range SPLIT ACROSS A THREAD-BLOCK CLUSTER (up to 8 CTAs per output tile): every CTA runs K / 8 in two
//     or three 32-wide stages, the partial tiles meet in distributed shared memory and are summed in rank order
//     (deterministic, no atomics, no second kernel);
//   * retrieval: 1000 queries x 100 000 pages x 1536 -- dense_tile_kernel<128, ...>, the classic 128 x 128 tile with
//     8 x 8 outputs per thread, 16-byte global loads, register-staged double buffering and conflict-free 128-bit
//     shared-memory reads.
// Operands may be k-contiguous (vectors along k, transposed on the way into shared memory) or row-contiguous (transposed
// VIEWS of the backward products: vectors along the rows, stored as they are).  Anything else (odd strides, unaligned
// bases, a K that is not a multiple of the vector width) takes dense_generic_kernel.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "dense_params.h"

namespace cpb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ float ld_as_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_f32(const __nv_bfloat16* p) {
  return bf16_bits_to_f32(__ldg(reinterpret_cast<const unsigned short*>(p)));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <typename T>
struct Vec;  // one 16-byte global load
template <>
struct Vec<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x), f[1] = __uint_as_float(v.y), f[2] = __uint_as_float(v.z), f[3] = __uint_as_float(v.w);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// ---- operand tile [ROWS x BK] -> registers (global, 16-byte vectors) -> shared memory sm[kk][row] as fp32 --------------
// row_major = the operand's k axis is contiguous (stride 1): vectors run along k and are transposed by the store;
// otherwise its row axis is contiguous (a transposed view): vectors run along the rows and are stored as they are.
template <int ROWS, int BK, typename T>
struct TileLoader {
  static constexpr int V = Vec<T>::kElems;
  static constexpr int kVecs = ROWS * BK / V;                         // vectors per tile
  static constexpr int kPerThread = (kVecs + kThreads - 1) / kThreads;
  uint4 r[kPerThread];

  __device__ __forceinline__ void load(const T* base, const int32_t* rows, int64_t rs, int64_t ks, bool k_contig, int row0,
                                       int n_rows, int k0, int k_end) {
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = threadIdx.x + u * kThreads;
      r[u] = make_uint4(0u, 0u, 0u, 0u);
      if (kVecs % kThreads != 0 && e >= kVecs) continue;
      if (k_contig) {
        const int row = row0 + e / (BK / V), kk = k0 + (e % (BK / V)) * V;
        if (row < n_rows && kk < k_end) {
          const int64_t gr = rows ? static_cast<int64_t>(__ldg(rows + row)) : static_cast<int64_t>(row);
          r[u] = __ldg(reinterpret_cast<const uint4*>(base + gr * rs + kk));
        }
      } else {
        const int kk = k0 + e / (ROWS / V), row = row0 + (e % (ROWS / V)) * V;
        if (row < n_rows && kk < k_end) r[u] = __ldg(reinterpret_cast<const uint4*>(base + row + static_cast<int64_t>(kk) * ks));
      }
    }
  }
  template <int LD>
  __device__ __forceinline__ void store(float (*sm)[LD], bool k_contig) const {
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = threadIdx.x + u * kThreads;
      if (kVecs % kThreads != 0 && e >= kVecs) continue;
      float f[V];
      Vec<T>::unpack(r[u], f);
      if (k_contig) {
        const int row = e / (BK / V), kk = (e % (BK / V)) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) sm[kk + j][row] = f[j];
      } else {
        const int kk = e / (ROWS / V), row = (e % (ROWS / V)) * V;
#pragma unroll
        for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(&sm[kk][row + j]) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      }
    }
  }
};

template <int TILE>
__device__ __forceinline__ void tile_origin(const DenseDotParams& p, int& i0, int& j0) {
  const int m_tiles = (p.m + TILE - 1) / TILE, n_tiles = (p.n + TILE - 1) / TILE;
  const int64_t per_block = static_cast<int64_t>(m_tiles) * p.raster_group;
  const int nb = static_cast<int>(blockIdx.x / per_block), r = static_cast<int>(blockIdx.x % per_block);
  const int g = min(p.raster_group, n_tiles - nb * p.raster_group);  // the last block may be narrower
  i0 = (r / g) * TILE;
  j0 = (nb * p.raster_group + r % g) * TILE;
}

// BM = BN = TILE; TILE 128: 8 x 8 outputs per thread (rows ty*4 + {0..3} and 64 + ty*4 + {0..3}, same for columns: every
// 128-bit shared-memory read of a quarter warp covers 128 consecutive bytes); TILE 64: 4 x 4 outputs per thread.
template <int TILE, int BK, typename TA, typename TB>
__global__ void __launch_bounds__(kThreads) dense_tile_kernel(const DenseDotParams p) {
  constexpr int LD = TILE + 4;            // row stride of a stage in floats (16-byte aligned rows)
  constexpr int H = TILE / 64;            // 64-wide halves per thread: 1 or 2
  constexpr int R = 4 * H;                // outputs per thread and dimension
  __shared__ __align__(16) float sa[2][BK][LD];
  __shared__ __align__(16) float sb[2][BK][LD];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // Linear tile index -> (row tile, column tile): blocks of `raster_group` column tiles x ALL row tiles, column tiles
  // fastest inside a block.  The CTAs in flight then share a few B tiles (documents) and sweep all of A (queries), so
  // B -- the big operand of a retrieval call, larger than L2 -- leaves DRAM once (plain column-fastest order re-read it
  // once per row tile: 4.9 GB of DRAM reads for 0.62 GB of operands at 1000 x 100 000 x 1536), without every CTA of a
  // wave asking the same L2 lines at the same moment (plain row-fastest order: 1.0 x traffic but 10 % slower).
  int i0, j0;
  tile_origin<TILE>(p, i0, j0);
  const TA* a = static_cast<const TA*>(p.a);
  const TB* b = static_cast<const TB*>(p.b);
  const bool a_kc = (p.a_ks == 1), b_kc = (p.b_ks == 1);
  // this CTA's share of K: split_k CTAs of a cluster (gridDim.z) work on the same output tile
  const int k_per = ((p.k + static_cast<int>(gridDim.z) * BK - 1) / (static_cast<int>(gridDim.z) * BK)) * BK;
  const int k_begin = blockIdx.z * k_per, k_end = min(p.k, k_begin + k_per);

  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  TileLoader<TILE, BK, TA> la;
  TileLoader<TILE, BK, TB> lb;
  if (k_begin < k_end) {
    la.load(a, nullptr, p.a_rs, p.a_ks, a_kc, i0, p.m, k_begin, k_end);
    lb.load(b, p.b_rows, p.b_rs, p.b_ks, b_kc, j0, p.n, k_begin, k_end);
    la.template store<LD>(sa[0], a_kc);
    lb.template store<LD>(sb[0], b_kc);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {  // the next stage's global loads fly while this stage is multiplied
      la.load(a, nullptr, p.a_rs, p.a_ks, a_kc, i0, p.m, k0 + BK, k_end);
      lb.load(b, p.b_rows, p.b_rs, p.b_ks, b_kc, j0, p.n, k0 + BK, k_end);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[R], bv[R];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 x = *reinterpret_cast<const float4*>(&sa[buf][kk][64 * h + ty * 4]);
        const float4 y = *reinterpret_cast<const float4*>(&sb[buf][kk][64 * h + tx * 4]);
        av[4 * h] = x.x, av[4 * h + 1] = x.y, av[4 * h + 2] = x.z, av[4 * h + 3] = x.w;
        bv[4 * h] = y.x, bv[4 * h + 1] = y.y, bv[4 * h + 2] = y.z, bv[4 * h + 3] = y.w;
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      la.template store<LD>(sa[buf ^ 1], a_kc);
      lb.template store<LD>(sb[buf ^ 1], b_kc);
    }
    __syncthreads();
    buf ^= 1;
  }

  const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
  auto emit = [&](int row, int col, float v) {
    if (row < p.m && col < p.n) {
      float* o = p.out + static_cast<int64_t>(row) * p.out_rs + col;
      *o = p.accumulate ? *o + alpha * v : alpha * v;
    }
  };
  if (gridDim.z == 1) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) emit(i0 + 64 * (i / 4) + ty * 4 + (i % 4), j0 + 64 * (j / 4) + tx * 4 + (j % 4), acc[i][j]);
    return;
  }
  if constexpr (TILE == 64) {
    // Split-K over the cluster (1, 1, S): every CTA parks its partial 64 x 64 tile in its own shared memory; after the
    // cluster barrier CTA z sums rows [z * 64 / S, (z + 1) * 64 / S) over the S partials IN RANK ORDER, reading its peers
    // through distributed shared memory.  (The stage buffers are free: the loop above ended with a block barrier.)
    float(*red)[LD] = sa[0];  // 64 rows x 68 floats = 17 KiB of the 17.4 KiB of sa
    static_assert(sizeof(sa) >= 64 * LD * sizeof(float), "partial tile must fit in the A stages");
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&red[ty * 4 + i][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int S = gridDim.z, rows_per = 64 / S;  // S is 2, 4 or 8
    for (int e = threadIdx.x; e < rows_per * 16; e += kThreads) {
      const int r = blockIdx.z * rows_per + e / 16, c4 = (e % 16) * 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t local = smem_addr(&red[r][c4]);
      for (int z = 0; z < S; ++z) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(z));
        float4 v;
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
        s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      }
      emit(i0 + r, j0 + c4, s.x);
      emit(i0 + r, j0 + c4 + 1, s.y);
      emit(i0 + r, j0 + c4 + 2, s.z);
      emit(i0 + r, j0 + c4 + 3, s.w);
    }
    // no CTA may exit while a peer still reads its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// ---- fallback for operands the vector loads cannot take (odd strides / alignment / K) ---------------------------------
constexpr int kGenKC = 32;
template <typename T>
__device__ __forceinline__ void generic_load_tile(float (*sm)[33], const T* base, const int32_t* rows, int64_t rs, int64_t ks,
                                                  int row0, int n_rows, int k0, int k) {
  const bool k_fast = (ks == 1);
#pragma unroll
  for (int e = threadIdx.x; e < 32 * kGenKC; e += kThreads) {
    const int r = k_fast ? e / kGenKC : e % 32;
    const int kk = k_fast ? e % kGenKC : e / 32;
    float v = 0.f;
    if (row0 + r < n_rows && k0 + kk < k) {
      const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + row0 + r)) : static_cast<int64_t>(row0 + r);
      v = ld_as_f32(base + row * rs + static_cast<int64_t>(k0 + kk) * ks);
    }
    sm[kk][r] = v;
  }
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(kThreads) dense_generic_kernel(const DenseDotParams p) {
  __shared__ float sa[kGenKC][33];
  __shared__ float sb[kGenKC][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int i0, j0;
  tile_origin<32>(p, i0, j0);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < p.k; k0 += kGenKC) {
    generic_load_tile<TA>(sa, static_cast<const TA*>(p.a), nullptr, p.a_rs, p.a_ks, i0, p.m, k0, p.k);
    generic_load_tile<TB>(sb, static_cast<const TB*>(p.b), p.b_rows, p.b_rs, p.b_ks, j0, p.n, k0, p.k);
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kGenKC; ++kk) {
      const float a0 = sa[kk][ty], a1 = sa[kk][ty + 16], b0 = sb[kk][tx], b1 = sb[kk][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = i0 + ty + 16 * i, col = j0 + tx + 16 * j;
      if (row < p.m && col < p.n) {
        float* o = p.out + static_cast<int64_t>(row) * p.out_rs + col;
        *o = p.accumulate ? *o + alpha * acc[i][j] : alpha * acc[i][j];
      }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <typename T>
bool operand_vectorisable(const void* base, int64_t rs, int64_t ks, int n_rows, int k, bool gathered) {
  constexpr int V = Vec<T>::kElems;
  if (reinterpret_cast<uintptr_t>(base) & 15u) return false;
  if (ks == 1) return (rs % V) == 0 && (k % V) == 0;                 // k-contiguous rows, 16-byte aligned
  if (rs == 1) return !gathered && (ks % V) == 0 && (n_rows % V) == 0;  // row-contiguous (transposed view)
  return false;
}

template <int TILE, int BK, typename TA, typename TB>
cudaError_t launch_tile(const DenseDotParams& p, int split_k, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(((p.n + TILE - 1) / TILE) * static_cast<int64_t>((p.m + TILE - 1) / TILE)), 1, split_k);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = split_k;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, dense_tile_kernel<TILE, BK, TA, TB>, p);
}

template <typename TA, typename TB>
cudaError_t launch_typed(const DenseDotParams& p, int sm_count, cudaStream_t stream) {
  const bool fast = operand_vectorisable<TA>(p.a, p.a_rs, p.a_ks, p.m, p.k, false) &&
                    operand_vectorisable<TB>(p.b, p.b_rs, p.b_ks, p.n, p.k, p.b_rows != nullptr);
  if (!fast) {
    const dim3 grid(static_cast<unsigned>(((p.n + 31) / 32) * static_cast<int64_t>((p.m + 31) / 32)));
    dense_generic_kernel<TA, TB><<<grid, kThreads, 0, stream>>>(p);
    return cudaGetLastError();
  }
  const int64_t big_tiles = (static_cast<int64_t>(p.m + 127) / 128) * ((p.n + 127) / 128);
  if (big_tiles >= sm_count) return launch_tile<128, 16, TA, TB>(p, 1, stream);
  // small problems: 64 x 64 tiles, and the K range split over a cluster until the SMs are covered (>= 2 stages per CTA)
  const int64_t tiles = (static_cast<int64_t>(p.m + 63) / 64) * ((p.n + 63) / 64);
  int split = 1;
  while (split < 8 && tiles * split * 2 <= 2 * sm_count && p.k >= split * 2 * 2 * 32) split *= 2;
  return launch_tile<64, 32, TA, TB>(p, split, stream);
}

}  // namespace

cudaError_t dense_dot_launch(const DenseDotParams& p, cudaStream_t stream) {
  using bf16 = __nv_bfloat16;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (p.a_f32) return p.b_f32 ? launch_typed<float, float>(p, sms, stream) : launch_typed<float, bf16>(p, sms, stream);
  return p.b_f32 ? launch_typed<bf16, float>(p, sms, stream) : launch_typed<bf16, bf16>(p, sms, stream);
}

}  // namespace cpb
