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
//   * One persistent CTA per SM.  A CTA keeps R (1 or 2) 128-row query tiles RESIDENT IN TENSOR MEMORY
//     (bf16 pairs, 64 columns per tile, written once with tcgen05.st) and feeds them to the tensor core
//     as the TMEM A operand (tcgen05.mma ... [d], [a], b_desc): only the document operand crosses the
//     shared-memory read port.  (The first version read A from shared memory and ran at 86% of the MMA
//     rate with TMA and epilogue switched off -- profiles/r01_notes.md.)
//   * The CTA streams its share of the document bank through a 4-deep ring of 192-token tiles that TMA
//     writes with the 128-byte swizzle.  Each document tile is multiplied against all R resident query
//     tiles (tcgen05.mma 128 x N x 16, 8 K-steps) into two 192-column fp32 TMEM accumulators used as a
//     ping-pong, so the epilogue of job j overlaps the MMAs of job j+1.
//     TMEM columns: [0,128) query tiles, [128,320) accumulator 0, [320,512) accumulator 1.
//   * Documents are addressed as (start row, length) in a flat [tokens, 128] bf16 bank, so ragged
//     banks, left/right padded batches and dense [B_d, N_d, 128] tensors are all the same kernel.
//     The last tile of a document is fetched in 32-row boxes, issued with a smaller MMA N (multiple
//     of 16) and its unused columns are masked in the epilogue.
//   * CTAs are launched as clusters of C (1, 2 or 4): the C CTAs of a cluster hold different query tiles,
//     walk the same document partition in lock-step and each fetches 1/C of every document tile with a
//     TMA multicast, so a bank byte crosses the L2->SM fabric once per C*R query tiles.
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
constexpr int kTileN = 192;
constexpr int kDim = 128;                       // embedding dim handled per pass (2 swizzle panels of 64)
constexpr int kDTileBytes = kTileN * kDim * 2;  // 48 KiB
constexpr int kDPanelBytes = kTileN * 64 * 2;   // 24 KiB
constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemQCols = 64;     // one 128 x 128 bf16 query tile = 64 32-bit columns
constexpr uint32_t kTmemAccBase = 128;  // accumulators start after two query tiles

struct SmemLayout {
  static constexpr int kDOff = 0;
  static constexpr int kBarOff = kStages * kDTileBytes;
  // barriers: q_ready, full[S], empty[S], tmem_full[2], tmem_empty[2]
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
maxsim_fwd_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_tail,
                  const MaxSimParams p) {
  using L = SmemLayout;
  constexpr int S = kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* d_smem = smem + L::kDOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_ready = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = bars + 1 + S;
  uint64_t* tmem_full = bars + 1 + 2 * S;
  uint64_t* tmem_empty = bars + 1 + 2 * S + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which slice of the problem is ours -------------------------------------------------
  // A cluster of C CTAs shares one document partition: every CTA holds different query tiles and
  // loads 1/C of each document tile, multicast to all C (each bank byte leaves L2 once per cluster).
  const int C = p.cluster;
  const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const int cluster_id = blockIdx.x / C;
  const int g = (cluster_id % p.group_sets) * C + static_cast<int>(crank);  // query-tile group
  const int part = cluster_id / p.group_sets;                                // document partition
  const int r_cnt = max(0, min(R, p.num_qtiles - g * R));
  const int d0 = static_cast<int>((static_cast<int64_t>(p.n_docs) * part) / p.doc_parts);
  const int d1 = static_cast<int>((static_cast<int64_t>(p.n_docs) * (part + 1)) / p.doc_parts);

  // ---- one-time setup ---------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_tail);
    mbar_init(q_ready, 4);  // one arrive per epilogue warp once its query rows are in TMEM
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], static_cast<uint32_t>(C));  // every CTA of the cluster releases the slot
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
  if (C > 1) cluster_sync_all(); else __syncthreads();  // barriers initialised cluster-wide before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const long long dbg_c0 = clock64();
  const uint64_t dbg_t0 = global_timer_ns();

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      const int rows_per_cta = kTileN / C;
      int stage = 0;
      uint32_t phase = 0;
      for (int d = d0; d < d1; ++d) {
        const int start = __ldg(p.doc_start + d);
        const int len = __ldg(p.doc_len + d);
        const int nch = max(1, (len + kTileN - 1) / kTileN);
        for (int c = 0; c < nch; ++c) {
          const int n_valid = min(kTileN, len - c * kTileN);
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* dst = d_smem + stage * kDTileBytes;
          const int row = start + c * kTileN;
          if (p.flags & CPB_DBG_NO_TMA) {
            mbar_arrive(&full[stage]);
          } else if (n_valid == kTileN) {
            // full tile: this CTA fetches rows [crank*192/C, (crank+1)*192/C) for the whole cluster
            mbar_expect_tx(&full[stage], kDTileBytes);
            const int r0 = static_cast<int>(crank) * rows_per_cta;
            if (C > 1) {
              tma_load_2d_mc(dst + r0 * 128, &tmap_d, 0, row + r0, &full[stage], cmask);
              tma_load_2d_mc(dst + kDPanelBytes + r0 * 128, &tmap_d, 64, row + r0, &full[stage], cmask);
            } else {
              tma_load_2d(dst, &tmap_d, 0, row, &full[stage]);
              tma_load_2d(dst + kDPanelBytes, &tmap_d, 64, row, &full[stage]);
            }
          } else {
            // tail of a document: 32-row boxes, fetched by rank 0 only
            const int nbox = max(1, (n_valid + 31) >> 5);
            mbar_expect_tx(&full[stage], static_cast<uint32_t>(nbox) * 32u * 256u);
            if (crank == 0) {
              for (int b = 0; b < nbox; ++b) {
                if (C > 1) {
                  tma_load_2d_mc(dst + b * 4096, &tmap_tail, 0, row + b * 32, &full[stage], cmask);
                  tma_load_2d_mc(dst + kDPanelBytes + b * 4096, &tmap_tail, 64, row + b * 32, &full[stage], cmask);
                } else {
                  tma_load_2d(dst + b * 4096, &tmap_tail, 0, row + b * 32, &full[stage]);
                  tma_load_2d(dst + kDPanelBytes + b * 4096, &tmap_tail, 64, row + b * 32, &full[stage]);
                }
              }
            }
          }
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
      mbar_wait(q_ready, 0);  // query tiles are in TMEM
      tc_fence_after();
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
            const uint32_t d_tmem = tmem_base + kTmemAccBase + a * kTileN;
            const uint32_t a_tmem = tmem_base + r * kTmemQCols;
#pragma unroll
            for (int k = 0; k < kDim / 16; ++k) {
              const int kp = k >> 2, kk = k & 3;
              const uint64_t bdesc =
                  make_sw128_kmajor_desc(d_addr + stage * kDTileBytes + kp * kDPanelBytes) + static_cast<uint64_t>(kk * 2);
              umma_bf16_ts(d_tmem, a_tmem + k * 8, bdesc, idesc, k > 0 ? 1u : 0u);
            }
            umma_commit(&tmem_full[a]);
            ++job;
          }
          // smem slot is free (in every CTA of the cluster) once these MMAs have read it
          if (C > 1) umma_commit_mc(&empty[stage], cmask); else umma_commit(&empty[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ================================ epilogue ==============================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;

    // ---- stage the resident query tiles: global -> registers -> TMEM (row = lane, 2 bf16 per column)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r < r_cnt) {
        const int row = (g * R + r) * kTileM + quad * 32 + lane;
        const uint4* src = reinterpret_cast<const uint4*>(p.q) + static_cast<int64_t>(row) * (kDim * 2 / 16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 x = (row < p.q_rows) ? __ldg(src + h * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
            v[4 * i] = x.x;
            v[4 * i + 1] = x.y;
            v[4 * i + 2] = x.z;
            v[4 * i + 3] = x.w;
          }
          tmem_st_x32(tmem_base + lane_base + r * kTmemQCols + h * 32, v);
        }
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);

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
            const uint32_t taddr = tmem_base + lane_base + kTmemAccBase + a * kTileN;
            float mm = (c == 0) ? init : m[r];
            int ai = (c == 0) ? -1 : am[r];
            if (p.flags & CPB_DBG_SKIP_EPILOGUE) {
              // profiling aid: leave the accumulator unread (results are garbage)
            } else if constexpr (!kArgmax) {
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
  // no CTA may exit while a peer can still multicast into its shared memory or signal its barriers
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if ((p.flags & CPB_DBG_CLOCKS) && threadIdx.x == 0) {
    p.scores[2 * blockIdx.x] = static_cast<float>(clock64() - dbg_c0);
    p.scores[2 * blockIdx.x + 1] = static_cast<float>(global_timer_ns() - dbg_t0);
  }
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

static void fill_cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster,
                             cudaStream_t stream) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SmemLayout::kAlloc;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

template <int R, bool kArgmax>
static cudaError_t launch_variant(const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p, int grid,
                                  cudaStream_t stream) {
  auto kern = maxsim_fwd_kernel<R, kArgmax>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::kAlloc);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  fill_cluster_cfg(cfg, attr, grid, p.cluster, stream);
  return cudaLaunchKernelEx(&cfg, kern, td, tt, p);
}

cudaError_t maxsim_launch(const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p, int r, bool argmax,
                          int grid, cudaStream_t stream) {
  if (r == 1)
    return argmax ? launch_variant<1, true>(td, tt, p, grid, stream) : launch_variant<1, false>(td, tt, p, grid, stream);
  return argmax ? launch_variant<2, true>(td, tt, p, grid, stream) : launch_variant<2, false>(td, tt, p, grid, stream);
}

// How many clusters of `cluster` CTAs can be co-resident (persistent-grid sizing).
int maxsim_max_clusters(int cluster) {
  auto kern = maxsim_fwd_kernel<2, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::kAlloc) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  fill_cluster_cfg(cfg, attr, cluster, cluster, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
}

int maxsim_tile_n() { return kTileN; }

cudaError_t maxsim_reduce_segments(const float* partial, float* out, int64_t plane, int nseg, int round_ref,
                                   cudaStream_t stream) {
  const int threads = 256;
  const int64_t blocks = (plane + threads - 1) / threads;
  maxsim_reduce_segments_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(partial, out, plane, nseg,
                                                                                       round_ref);
  return cudaGetLastError();
}

}  // namespace cpb
