// Fused multi-vector projection head for sm_100a:
//
//   out[t, :] = round( (h[t, :] @ W^T + b) / || h[t, :] @ W^T + b ||_2 ) * attention_mask[t] [* extra_mask[t]]
//
// replaces the tail of every Col* model forward, e.g. colpali_engine/models/qwen2/colqwen2/modeling_colqwen2.py:65-74
// (custom_text_proj -> proj / proj.norm(dim=-1, keepdim=True) -> * attention_mask -> optional image-token mask)
// and the clamp variant of models/modernvbert/colvbert/modeling_colmodernvbert.py:59.  The reference runs four
// kernels with three extra [B, L, 128] round trips; this is one pass that reads h once.
//
// HBM-bound (AI ~ 118 FLOP/B at H = 1536): the design goal is to stream h at HBM rate.
//   * tokens on M (TMEM lanes), the 128 output dims on N, K = hidden size in blocks of 64.
//   * a CTA works on PAIRS of 128-token tiles so that every 16 KiB block of W fetched from L2 is used twice.
//   * the token range is cut into equal shares of 64-token UNITS per CTA (one contiguous range each), not into whole
//     pairs: 65 920 tokens are 258 pairs = 1.74 rounds of 148 CTAs, i.e. 2 rounds (0.68 of the HBM peak measured), but
//     1030 units = 6.96 per CTA; a range that ends in half a tile loads 64 rows (its own TMA box) and multiplies a
//     128-row tile whose upper half is stale -- wasted tensor work is free in an HBM-bound kernel, wasted bytes are not.
//   * 4-stage TMA ring (2 x 16 KiB of h + 16 KiB of W per stage), tcgen05.mma 128x128x16 into two pairs of
//     128-column fp32 accumulators (ping-pong across token-tile pairs), 4 epilogue warps.
//   * epilogue: one thread owns one token row (128 fp32 in registers): bias, the reference's three bf16
//     roundings (Linear output, the norm, the quotient), masks, 16-byte stores.
#include <cfloat>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
#include "head_params.h"
#include "sm100_ptx.cuh"

namespace cpb {

constexpr int kHBM = 128;                     // tokens per tile
constexpr int kHBK = 64;                      // K block (one 128-byte swizzle row)
constexpr int kHDim = 128;                    // output dim
constexpr int kHTileBytes = kHBM * kHBK * 2;  // 16 KiB
constexpr int kHStageBytes = 3 * kHTileBytes; // h tile 0, h tile 1, W block
constexpr int kHStages = 4;
constexpr int kHThreads = 192;

struct HeadSmem {
  static constexpr int kBarOff = kHStages * kHStageBytes;
  static constexpr int kNumBars = 2 * kHStages + 4;  // full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kBiasOff = kTmemPtrOff + 16;
  static constexpr int kBytes = kBiasOff + kHDim * 4;
  static constexpr int kAlloc = kBytes + 1024;
};

__device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__device__ __forceinline__ int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(kHThreads, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_h64,
                const __grid_constant__ CUtensorMap tmap_w, const HeadParams p) {
  using L = HeadSmem;
  constexpr int S = kHStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tmem_full = bars + 2 * S;
  uint64_t* tmem_empty = bars + 2 * S + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.hidden / kHBK;
  // this CTA's tokens [t_begin, t_end): an equal share of the 64-token units; walked in steps of 256 rows (a pair)
  const int64_t units = (p.n_tokens + 63) / 64;
  const int64_t t_begin = 64 * ((units * blockIdx.x) / gridDim.x);
  const int64_t t_end = imin64(p.n_tokens, 64 * ((units * (static_cast<int64_t>(blockIdx.x) + 1)) / gridDim.x));

  if (threadIdx.x < kHDim)
    s_bias[threadIdx.x] = p.bias ? __bfloat162float(p.bias[threadIdx.x]) : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_h64);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t row = t_begin; row < t_end; row += 2 * kHBM) {
        const int row0 = static_cast<int>(row);
        // rows of the two tiles that belong to this CTA: 128 or 64 (0 for an absent second tile); a tile of at most 64
        // rows is fetched with the 64-row box so that no byte of the neighbour's range is read twice
        const int n0 = static_cast<int>(imin64(kHBM, t_end - row));
        const int n1 = static_cast<int>(imax64(0, imin64(kHBM, t_end - row - kHBM)));
        const uint32_t bytes = kHTileBytes + (n0 > 64 ? kHTileBytes : kHTileBytes / 2) +
                               (n1 > 64 ? kHTileBytes : (n1 > 0 ? kHTileBytes / 2 : 0));
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], bytes);
          uint8_t* dst = smem + stage * kHStageBytes;
          tma_load_2d(dst, n0 > 64 ? &tmap_h : &tmap_h64, kb * kHBK, row0, &full[stage]);
          if (n1 > 0) tma_load_2d(dst + kHTileBytes, n1 > 64 ? &tmap_h : &tmap_h64, kb * kHBK, row0 + kHBM, &full[stage]);
          tma_load_2d(dst + 2 * kHTileBytes, &tmap_w, kb * kHBK, 0, &full[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(kHBM, kHDim);
      const uint32_t s_addr = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t row = t_begin; row < t_end; row += 2 * kHBM, ++it) {
        const uint32_t a = it & 1u;
        const int n_tiles = (t_end - row > kHBM) ? 2 : 1;
        mbar_wait(&tmem_empty[a], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = s_addr + stage * kHStageBytes;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r >= n_tiles) break;  // (rows past a tile's share are stale shared memory: computed, never stored)
#pragma unroll
            for (int k = 0; k < kHBK / 16; ++k) {
              const uint64_t adesc = make_sw128_kmajor_desc(base + r * kHTileBytes) + static_cast<uint64_t>(k * 2);
              const uint64_t bdesc = make_sw128_kmajor_desc(base + 2 * kHTileBytes) + static_cast<uint64_t>(k * 2);
              umma_bf16(tmem_base + (a * 2 + r) * kHDim, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
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
    uint32_t it = 0;
    for (int64_t prow = t_begin; prow < t_end; prow += 2 * kHBM, ++it) {
      const uint32_t a = it & 1u;
      const int n_tiles = (t_end - prow > kHBM) ? 2 : 1;
      mbar_wait(&tmem_full[a], (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < n_tiles; ++r) {
        const int64_t row = prow + r * kHBM + quad * 32 + lane;
        const uint32_t taddr = tmem_base + lane_base + (a * 2 + r) * kHDim;
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_x32(taddr, v0);
        tmem_ld_x32(taddr + 32, v1);
        tmem_ld_x32(taddr + 64, v2);
        tmem_ld_x32(taddr + 96, v3);
        tmem_ld_wait();
        reg_fence32(v0);
        reg_fence32(v1);
        reg_fence32(v2);
        reg_fence32(v3);
        if (r == n_tiles - 1) {  // every tile of this accumulator pair is now in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[a]);
        }
        float x[128];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = __uint_as_float(v0[j]) + s_bias[j];
          x[32 + j] = __uint_as_float(v1[j]) + s_bias[32 + j];
          x[64 + j] = __uint_as_float(v2[j]) + s_bias[64 + j];
          x[96 + j] = __uint_as_float(v3[j]) + s_bias[96 + j];
        }
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          if (!single) x[j] = rbf(x[j]);  // nn.Linear output is bf16                (modeling_colqwen2.py:65)
          ss = fmaf(x[j], x[j], ss);
        }
        float nrm = sqrtf(ss);
        if (!single) nrm = rbf(nrm);      // proj.norm(...) is a bf16 tensor           (:68)
        if (clamp) nrm = fmaxf(nrm, 1e-12f);
        float mk = 1.f;
        if (row < t_end) {
          if (p.attention_mask) mk = static_cast<float>(p.attention_mask[row]);        // (:69)
          if (p.extra_mask) mk *= (p.extra_mask[row] != 0) ? 1.f : 0.f;                // (:71-74)
        }
        if (row < t_end) {  // rows past this CTA's share belong to the next CTA (or lie past the last token)
          uint4* dst = reinterpret_cast<uint4*>(p.out + row * kHDim);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float y0 = x[8 * j + 2 * u] / nrm, y1 = x[8 * j + 2 * u + 1] / nrm;
              if (!single) {
                y0 = rbf(y0);             // the quotient is a bf16 tensor              (:68)
                y1 = rbf(y1);
              }
              const __nv_bfloat162 pk = __floats2bfloat162_rn(y0 * mk, y1 * mk);
              w[u] = *reinterpret_cast<const uint32_t*>(&pk);
            }
            dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

cudaError_t head_launch(const CUtensorMap& th, const CUtensorMap& th64, const CUtensorMap& tw, const HeadParams& p, int grid,
                        cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadSmem::kAlloc);
  if (e != cudaSuccess) return e;
  head_fwd_kernel<<<grid, kHThreads, HeadSmem::kAlloc, stream>>>(th, th64, tw, p);
  return cudaGetLastError();
}

}  // namespace cpb
