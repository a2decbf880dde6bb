// Fused projection head for output dims above 128 (validated on B200 in round 2: tests/test_head_gpu.py)
// (ColQwen3: dim = 320, colpali_engine/models/qwen3/colqwen3/modeling_colqwen3.py:48-49,87-96; ColQwen3.5 takes the
// dim from its config, models/qwen3_5/colqwen3_5/modeling_colqwen3_5.py:35-36).
//
//   out[t, :] = round( (h[t, :] @ W^T + b) / || h[t, :] @ W^T + b ||_2 ) * attention_mask[t] [* extra_mask[t]]
//
// Differences from head_sm100.cu (dim = 128, validated):
//   * dim is a runtime value, a multiple of 32 in (128, 320].  The W block of one K step is dim x 64 bf16
//     (up to 40 KiB), loaded as two TMA boxes of dim/2 rows (the box limit is 256 rows).
//   * one 128-token tile per CTA and ring stage (16 KiB of h + dim * 128 B of W).  At dim = 320 the head is no longer
//     HBM-bound on h alone (AI = 2*dim*H / 2*(H+dim) ~ 284 FLOP/B at H = 2560): every SM would re-read all of W
//     (dim * H * 2 B) from L2 per 128 tokens, 2.4x the bytes of its h tile.  CTAs therefore run as clusters of
//     two: each loads one half of every W block and multicasts it to both (W leaves L2 once per 256 tokens).
//   * dim <= 256: one tcgen05.mma 128 x dim x 16 per K step, two accumulators (ping-pong across tiles);
//     dim  > 256: two MMAs of N = dim/2 per K step and ONE accumulator (2 x 320 columns do not fit in TMEM).
//   * a token row no longer fits in registers: the epilogue reads the accumulator twice (sum of squares, then
//     normalise + store); the accumulator is released after the second read.
#include <cfloat>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
#include "head_params.h"
#include "sm100_ptx.cuh"

namespace cpb {

constexpr int kWBM = 128;                      // tokens per tile
constexpr int kWBK = 64;                       // K block (one 128-byte swizzle row)
constexpr int kWHTileBytes = kWBM * kWBK * 2;  // 16 KiB of h per stage
constexpr int kWThreads = 192;
constexpr int kWMaxDim = 320;
constexpr int kWMaxStages = 4;

struct HeadWideSmem {
  // [ring: S x (h tile | W block)] [barriers] [tmem ptr] [bias]
  static constexpr int kNumBars = 2 * kWMaxStages + 4;
  static __host__ __device__ int stage_bytes(int dim) { return kWHTileBytes + dim * 128; }
  static __host__ __device__ int bar_off(int dim, int stages) { return stages * stage_bytes(dim); }
  static __host__ __device__ int tmem_ptr_off(int dim, int stages) { return bar_off(dim, stages) + kNumBars * 8; }
  static __host__ __device__ int bias_off(int dim, int stages) { return tmem_ptr_off(dim, stages) + 16; }
  static __host__ __device__ int bytes(int dim, int stages) { return bias_off(dim, stages) + kWMaxDim * 4 + 1024; }
};

__device__ __forceinline__ float rbf_w(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(kWThreads, 1)
head_wide_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w,
                 const HeadParams p) {
  using L = HeadWideSmem;
  const int dim = p.dim;
  const int S = p.stages;
  const int C = p.cluster;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off(dim, S));
  uint64_t* full = bars;
  uint64_t* empty = bars + kWMaxStages;
  uint64_t* tmem_full = bars + 2 * kWMaxStages;
  uint64_t* tmem_empty = bars + 2 * kWMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::tmem_ptr_off(dim, S));
  float* s_bias = reinterpret_cast<float*>(smem + L::bias_off(dim, S));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.hidden / kWBK;
  const int stage_bytes = L::stage_bytes(dim);
  const int half_rows = dim / 2;             // rows of W per TMA box
  const int half_bytes = half_rows * 128;    // multiple of 1024 (dim % 16 == 0): swizzle atoms stay aligned
  const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const int64_t num_tiles = (p.n_tokens + kWBM - 1) / kWBM;
  const int64_t cluster_id = blockIdx.x / C;
  const int64_t num_clusters = gridDim.x / C;
  const int64_t num_rounds = (num_tiles + C - 1) / C;   // both CTAs of a cluster run the same number of tiles;
  // a tile index past num_tiles is a phantom: its h rows are zero-filled by TMA and nothing is stored
  const int n_acc = (dim <= 256) ? 2 : 1;

  for (int j = threadIdx.x; j < dim; j += kWThreads) s_bias[j] = p.bias ? __bfloat162float(p.bias[j]) : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], static_cast<uint32_t>(C));  // every CTA of the cluster releases the slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  if (C > 1) cluster_sync_all(); else __syncthreads();  // barriers initialised cluster-wide before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t r = cluster_id; r < num_rounds; r += num_clusters) {
        const int64_t tile = r * C + crank;
        const int row0 = static_cast<int>(tile * kWBM);  // may lie past n_tokens (phantom): zero fill
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], static_cast<uint32_t>(stage_bytes));
          uint8_t* dst = smem + stage * stage_bytes;
          tma_load_2d(dst, &tmap_h, kb * kWBK, row0, &full[stage]);
          uint8_t* wdst = dst + kWHTileBytes;
          if (C > 1) {
            tma_load_2d_mc(wdst + crank * half_bytes, &tmap_w, kb * kWBK, static_cast<int>(crank) * half_rows,
                           &full[stage], cmask);
          } else {
            tma_load_2d(wdst, &tmap_w, kb * kWBK, 0, &full[stage]);
            tma_load_2d(wdst + half_bytes, &tmap_w, kb * kWBK, half_rows, &full[stage]);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int n_mma = (dim <= 256) ? 1 : 2;
      const int mma_n = dim / n_mma;
      const uint32_t idesc = make_idesc_bf16_f32(kWBM, mma_n);
      const uint32_t s_addr = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t r = cluster_id; r < num_rounds; r += num_clusters, ++it) {
        const uint32_t a = (n_acc == 2) ? (it & 1u) : 0u;
        const uint32_t use = (n_acc == 2) ? (it >> 1) : it;  // how many times accumulator a was used before
        mbar_wait(&tmem_empty[a], (use & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = s_addr + stage * stage_bytes;
#pragma unroll
          for (int k = 0; k < kWBK / 16; ++k) {
            const uint64_t adesc = make_sw128_kmajor_desc(base) + static_cast<uint64_t>(k * 2);
            for (int m = 0; m < n_mma; ++m) {
              const uint64_t bdesc = make_sw128_kmajor_desc(base + kWHTileBytes + m * half_bytes) + static_cast<uint64_t>(k * 2);
              umma_bf16(tmem_base + a * 256 + m * mma_n, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          // the slot is free (in every CTA of the cluster: both read the multicast W halves) once these MMAs retire
          if (C > 1) umma_commit_mc(&empty[stage], cmask); else umma_commit(&empty[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[a]);
      }
    }
  } else {
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const bool single = (p.flags & CPB_HEAD_SINGLE_ROUNDING) != 0;
    const bool clamp = (p.flags & CPB_HEAD_CLAMP_NORM) != 0;
    const int n_chunks = dim / 32;
    uint32_t it = 0;
    for (int64_t r = cluster_id; r < num_rounds; r += num_clusters, ++it) {
      const int64_t tile = r * C + crank;
      const uint32_t a = (n_acc == 2) ? (it & 1u) : 0u;
      const uint32_t use = (n_acc == 2) ? (it >> 1) : it;
      mbar_wait(&tmem_full[a], use & 1u);
      tc_fence_after();
      const int64_t row = tile * kWBM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + lane_base + a * 256;
      // Both passes read the accumulator in 32-column chunks through two register buffers: chunk c + 1 is in flight
      // while chunk c is consumed (the epilogue is ~7 k instructions per token row at dim 320 and, with a single
      // accumulator, serial with the MMAs of the next tile -- every cycle here is tensor idle time).
      // pass 1: squared norm of the (bf16-rounded) projection row
      float ss = 0.f;
      auto sq_chunk = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + s_bias[c * 32 + j];
          if (!single) x = rbf_w(x);      // nn.Linear output is bf16                     (modeling_colqwen3.py:87)
          ss = fmaf(x, x, ss);
        }
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld_x32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          tmem_ld_wait();
          reg_fence32(va);
          if (c + 1 < n_chunks) tmem_ld_x32(taddr + (c + 1) * 32, vb);
          sq_chunk(va, c);
          if (c + 1 < n_chunks) {
            tmem_ld_wait();
            reg_fence32(vb);
            if (c + 2 < n_chunks) tmem_ld_x32(taddr + (c + 2) * 32, va);
            sq_chunk(vb, c + 1);
          }
        }
      }
      float nrm = sqrtf(ss);
      if (!single) nrm = rbf_w(nrm);      // proj.norm(...) is a bf16 tensor               (:90)
      if (clamp) nrm = fmaxf(nrm, 1e-12f);
      // x / nrm as reciprocal + one Newton step on the residual: the correctly rounded quotient for all but a
      // vanishing fraction of operands at a third of the instructions of an IEEE division (320 of them per row)
      const float inv = 1.0f / nrm;
      auto div_nrm = [&](float x) {
        const float y = x * inv;
        return fmaf(fmaf(-y, nrm, x), inv, y);
      };
      float mk = 1.f;
      const bool live = row < p.n_tokens;
      if (live) {
        if (p.attention_mask) mk = static_cast<float>(p.attention_mask[row]);              // (:91)
        if (p.extra_mask) mk *= (p.extra_mask[row] != 0) ? 1.f : 0.f;                      // (:93-96)
      }
      // pass 2: normalise, mask, store
      auto out_chunk = [&](const uint32_t (&v)[32], int c) {
        if (!live) return;
        uint4* dst = reinterpret_cast<uint4*>(p.out + row * dim + c * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float x0 = __uint_as_float(v[8 * j + 2 * u]) + s_bias[c * 32 + 8 * j + 2 * u];
            float x1 = __uint_as_float(v[8 * j + 2 * u + 1]) + s_bias[c * 32 + 8 * j + 2 * u + 1];
            if (!single) {
              x0 = rbf_w(x0);
              x1 = rbf_w(x1);
            }
            float y0 = div_nrm(x0), y1 = div_nrm(x1);
            if (!single) {
              y0 = rbf_w(y0);           // the quotient is a bf16 tensor                 (:90)
              y1 = rbf_w(y1);
            }
            const __nv_bfloat162 pk = __floats2bfloat162_rn(y0 * mk, y1 * mk);
            w[u] = *reinterpret_cast<const uint32_t*>(&pk);
          }
          dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      };
      auto release = [&]() {  // the whole accumulator has been read twice: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[a]);
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld_x32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          tmem_ld_wait();
          reg_fence32(va);
          if (c + 1 < n_chunks) tmem_ld_x32(taddr + (c + 1) * 32, vb); else release();
          out_chunk(va, c);
          if (c + 1 < n_chunks) {
            tmem_ld_wait();
            reg_fence32(vb);
            if (c + 2 < n_chunks) tmem_ld_x32(taddr + (c + 2) * 32, va); else release();
            out_chunk(vb, c + 1);
          }
        }
      }
    }
  }

  tc_fence_before();
  if (C > 1) cluster_sync_all(); else __syncthreads();  // no CTA leaves while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int head_wide_stages(int dim) {
  const int budget = 227 * 1024;
  for (int s = kWMaxStages; s >= 2; --s)
    if (HeadWideSmem::bytes(dim, s) <= budget) return s;
  return 0;
}

cudaError_t head_wide_launch(const CUtensorMap& th, const CUtensorMap& tw, const HeadParams& p, int grid, cudaStream_t stream) {
  const int smem = HeadWideSmem::bytes(p.dim, p.stages);
  cudaError_t e = cudaFuncSetAttribute(head_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kWThreads, 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, head_wide_kernel, th, tw, p);
}

}  // namespace cpb
