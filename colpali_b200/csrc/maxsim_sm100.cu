// Fused late-interaction (MaxSim) scoring for sm_100a.
//
//   scores[i, j] = sum_n  max_s  < q[i, n, :], d[j, s, :] >
//
// replaces the einsum("bnd,csd->bcns").max(dim=3)[0].sum(dim=2) chain of
// colpali_engine/utils/processing_utils.py:179 and colpali_engine/loss/late_interaction_losses.py:153-154,
// without ever writing the [B_q, B_d, N_q, N_d] similarity tensor to HBM.
//
// Mapping onto the hardware
//   * GEMM view: rows (M) = query tokens, columns (N) = document tokens, K = embedding dim (128).
//     Query tokens sit on TMEM lanes, so max-over-document-tokens is a per-thread running
//     FMNMX3 over accumulator columns and the sum over a query's tokens is one warp reduction
//     (queries are padded to a multiple of 32 rows: one warp == one query segment).
//   * One persistent CTA per SM.  A CTA keeps R (1 or 2) 128-row query tiles resident in shared
//     memory and streams its share of the document bank through a ring of 256-token tiles that
//     TMA writes with the 128-byte swizzle.  Each document tile is multiplied against all R
//     resident query tiles (tcgen05.mma 128 x N x 16, 8 K-steps) into two 256-column fp32 TMEM
//     accumulators used as a ping-pong, so the epilogue of job j overlaps the MMAs of job j+1.
//   * Documents are addressed as (start row, length) in a flat [tokens, 128] bf16 bank, so ragged
//     banks, left/right padded batches and dense [B_d, N_d, 128] tensors are all the same kernel.
//     The last tile of a document is issued with a smaller MMA N (multiple of 16) and its
//     unused columns are masked in the epilogue.
//   * warp 0: TMA producer.  warp 1: TMEM allocator + MMA issuer.  warps 2-5: epilogue.
#include <cfloat>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "maxsim_params.h"
#include "sm100_ptx.cuh"

namespace cpb {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kDim = 128;                           // embedding dim handled per pass (2 swizzle panels of 64)
constexpr int kQTileBytes = kTileM * kDim * 2;      // 32 KiB
constexpr int kQPanelBytes = kTileM * 64 * 2;       // 16 KiB
constexpr int kDTileBytes = kTileN * kDim * 2;      // 64 KiB
constexpr int kDPanelBytes = kTileN * 64 * 2;       // 32 KiB
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;

template <int R>
struct SmemLayout {
  static constexpr int kStages = (R == 1) ? 3 : 2;
  static constexpr int kQOff = 0;
  static constexpr int kDOff = R * kQTileBytes;
  static constexpr int kBarOff = kDOff + kStages * kDTileBytes;
  // barriers: q_full, full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kNumBars = 1 + 2 * kStages + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemPtrOff + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for manual 1024-B alignment
};

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// max over 32 accumulator columns folded into m (16 FMNMX3, shallow dependency tree)
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    t[i] = fmax3(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]));
  float u0 = fmax3(t[0], t[1], __uint_as_float(v[3]));
  float u1 = fmax3(t[2], t[3], __uint_as_float(v[7]));
  float u2 = fmax3(t[4], t[5], __uint_as_float(v[11]));
  float u3 = fmax3(t[6], t[7], __uint_as_float(v[15]));
  float w0 = fmax3(u0, __uint_as_float(v[19]), __uint_as_float(v[23]));
  float w1 = fmax3(u1, __uint_as_float(v[27]), __uint_as_float(v[31]));
  float x0 = fmax3(w0, w1, u2);
  return fmax3(x0, u3, m);
}

__device__ __forceinline__ float max32_masked(const uint32_t (&v)[32], float m, int nvalid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = (i < nvalid) ? __uint_as_float(v[i]) : -INFINITY;
    m = fmaxf(m, x);
  }
  return m;
}

// running (value, first index) argmax over 32 columns; strict '>' keeps the earliest maximum,
// which is what torch.max(dim) returns on ties.
__device__ __forceinline__ void argmax32(const uint32_t (&v)[32], float& m, int& idx, int col0, int nvalid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = __uint_as_float(v[i]);
    bool take = (i < nvalid) && (x > m);
    m = take ? x : m;
    idx = take ? (col0 + i) : idx;
  }
}

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int R, bool kArgmax>
__global__ void __launch_bounds__(kThreads, 1)
maxsim_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                  const MaxSimParams p) {
  using L = SmemLayout<R>;
  constexpr int S = L::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem + L::kQOff;
  uint8_t* d_smem = smem + L::kDOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = bars + 1 + S;
  uint64_t* tmem_full = bars + 1 + 2 * S;
  uint64_t* tmem_empty = bars + 1 + 2 * S + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which slice of the problem is ours -------------------------------------------------
  const int g = blockIdx.x % p.q_groups;     // query-tile group
  const int part = blockIdx.x / p.q_groups;  // document partition
  const int r_cnt = min(R, p.num_qtiles - g * R);
  const int d0 = static_cast<int>((static_cast<int64_t>(p.n_docs) * part) / p.doc_parts);
  const int d1 = static_cast<int>((static_cast<int64_t>(p.n_docs) * (part + 1)) / p.doc_parts);

  // ---- one-time setup ---------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    mbar_init(q_full, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      mbar_expect_tx(q_full, static_cast<uint32_t>(r_cnt) * kQTileBytes);
      for (int r = 0; r < r_cnt; ++r) {
        const int row = (g * R + r) * kTileM;
        tma_load_2d(q_smem + r * kQTileBytes, &tmap_q, 0, row, q_full);
        tma_load_2d(q_smem + r * kQTileBytes + kQPanelBytes, &tmap_q, 64, row, q_full);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int d = d0; d < d1; ++d) {
        const int start = __ldg(p.doc_start + d);
        const int len = __ldg(p.doc_len + d);
        const int nch = max(1, (len + kTileN - 1) / kTileN);
        for (int c = 0; c < nch; ++c) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], kDTileBytes);
          uint8_t* dst = d_smem + stage * kDTileBytes;
          const int row = start + c * kTileN;
          tma_load_2d(dst, &tmap_d, 0, row, &full[stage]);
          tma_load_2d(dst + kDPanelBytes, &tmap_d, 64, row, &full[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t d_addr = smem_u32(d_smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t job = 0;
      for (int d = d0; d < d1; ++d) {
        const int len = __ldg(p.doc_len + d);
        const int nch = max(1, (len + kTileN - 1) / kTileN);
        for (int c = 0; c < nch; ++c) {
          const int n_valid = min(kTileN, len - c * kTileN);
          const uint32_t n_mma = static_cast<uint32_t>(max(16, (n_valid + 15) & ~15));
          const uint32_t idesc = make_idesc_bf16_f32(kTileM, n_mma);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          for (int r = 0; r < r_cnt; ++r) {
            const uint32_t a = job & 1u;
            const uint32_t aphase = (job >> 1) & 1u;
            mbar_wait(&tmem_empty[a], aphase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * kTileN;
#pragma unroll
            for (int k = 0; k < kDim / 16; ++k) {
              const int kp = k >> 2, kk = k & 3;
              const uint64_t adesc =
                  make_sw128_kmajor_desc(q_addr + r * kQTileBytes + kp * kQPanelBytes) + static_cast<uint64_t>(kk * 2);
              const uint64_t bdesc =
                  make_sw128_kmajor_desc(d_addr + stage * kDTileBytes + kp * kDPanelBytes) + static_cast<uint64_t>(kk * 2);
              umma_bf16(d_tmem, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
            }
            umma_commit(&tmem_full[a]);
            ++job;
          }
          umma_commit(&empty[stage]);  // smem slot is free once these MMAs have read it
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ================================ epilogue ==============================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float m[R];
    int am[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      m[r] = -INFINITY;
      am[r] = -1;
    }
    const bool round_ref = (p.flags & CPB_FLAG_ROUND_BF16) != 0;
    uint32_t job = 0;
    for (int d = d0; d < d1; ++d) {
      const int len = __ldg(p.doc_len + d);
      const int nch = max(1, (len + kTileN - 1) / kTileN);
      const float init = (p.doc_floor != nullptr) ? __ldg(p.doc_floor + d) : -INFINITY;
      for (int c = 0; c < nch; ++c) {
        const int n_valid = min(kTileN, len - c * kTileN);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < r_cnt) {
            const uint32_t a = job & 1u;
            const uint32_t aphase = (job >> 1) & 1u;
            mbar_wait(&tmem_full[a], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_base + a * kTileN;
            float mm = (c == 0) ? init : m[r];
            int ai = (c == 0) ? -1 : am[r];
            if constexpr (!kArgmax) {
              if (n_valid == kTileN) {
#pragma unroll
                for (int cc = 0; cc < kTileN / 64; ++cc) {
                  uint32_t v0[32], v1[32];
                  tmem_ld_x32(taddr + cc * 64, v0);
                  tmem_ld_x32(taddr + cc * 64 + 32, v1);
                  tmem_ld_wait();
                  mm = max32(v0, mm);
                  mm = max32(v1, mm);
                }
              } else {
                for (int col = 0; col < n_valid; col += 32) {
                  uint32_t v[32];
                  tmem_ld_x32(taddr + col, v);
                  tmem_ld_wait();
                  const int nv = n_valid - col;
                  mm = (nv >= 32) ? max32(v, mm) : max32_masked(v, mm, nv);
                }
              }
            } else {
              for (int col = 0; col < n_valid; col += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + col, v);
                tmem_ld_wait();
                argmax32(v, mm, ai, c * kTileN + col, n_valid - col);
              }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[a]);
            m[r] = mm;
            am[r] = ai;

            if (c == nch - 1) {
              // document finished: fold this query segment's 32 token maxima into one score
              const int row0 = (g * R + r) * kTileM + quad * 32;  // first padded query row of this warp
              const int q = row0 / p.nq_pad;
              const int seg = (row0 % p.nq_pad) >> 5;
              if (kArgmax && p.argmax != nullptr && row0 + lane < p.q_rows)
                p.argmax[static_cast<int64_t>(d) * p.q_rows + row0 + lane] = ai;
              float x = round_ref ? round_bf16(mm) : mm;
              x = warp_sum(x);
              if (round_ref && p.nq_pad == 32) x = round_bf16(x);
              if (lane == 0 && q < p.n_queries)
                p.scores[static_cast<int64_t>(seg) * p.plane_stride + static_cast<int64_t>(q) * p.n_docs + d] = x;
            }
            ++job;
          }
        }
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// scores[q, d] = sum_seg partial[seg, q, d]   (only used when a query spans more than 32 rows)
__global__ void maxsim_reduce_segments_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                              int64_t plane, int nseg, int round_ref) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  float s = 0.f;
  for (int k = 0; k < nseg; ++k) s += partial[k * plane + i];
  out[i] = round_ref ? round_bf16(s) : s;
}

template <int R, bool kArgmax>
static cudaError_t launch_variant(const CUtensorMap& tq, const CUtensorMap& td, const MaxSimParams& p, int grid,
                                  cudaStream_t stream) {
  using L = SmemLayout<R>;
  auto kern = maxsim_fwd_kernel<R, kArgmax>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, L::kAlloc, stream>>>(tq, td, p);
  return cudaGetLastError();
}

cudaError_t maxsim_launch(const CUtensorMap& tq, const CUtensorMap& td, const MaxSimParams& p, int r, bool argmax,
                          int grid, cudaStream_t stream) {
  if (r == 1) return argmax ? launch_variant<1, true>(tq, td, p, grid, stream) : launch_variant<1, false>(tq, td, p, grid, stream);
  return argmax ? launch_variant<2, true>(tq, td, p, grid, stream) : launch_variant<2, false>(tq, td, p, grid, stream);
}

cudaError_t maxsim_reduce_segments(const float* partial, float* out, int64_t plane, int nseg, int round_ref,
                                   cudaStream_t stream) {
  const int threads = 256;
  const int64_t blocks = (plane + threads - 1) / threads;
  maxsim_reduce_segments_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(partial, out, plane, nseg,
                                                                                       round_ref);
  return cudaGetLastError();
}

}  // namespace cpb
