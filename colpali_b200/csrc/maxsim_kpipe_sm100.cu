// K-pipelined fused MaxSim for embedding dims up to 320 (ColQwen3: colpali_engine/models/qwen3/colqwen3/modeling_colqwen3.py:48).
//
// Same contract, epilogue (maxsim_epilogue.cuh, R = 1) and partitioning as maxsim_sm100.cu; what differs is the K loop:
//   * the embedding dim is KP panels of 64 (KP = 1..5); ONE 128-row query tile per CTA stays resident in shared memory
//     (KP x 16 KiB); a ring stage holds ONE K-panel of a 256-row document tile (32 KiB), 4 stages;
//   * per document tile the issuer runs KP x 4 tcgen05.mma 128x256x16 into one 256-column accumulator (two
//     accumulators ping-pong across tiles), releasing a ring stage after each panel;
//   * barrier waits are software-pipelined as in maxsim_sm100.cu: the wait for the next panel (and, on the last panel
//     of a tile, for the next accumulator) sits after the 2nd of the 4 MMAs of the current panel.
// Smem: 80 KiB (KP = 5) + 128 KiB.  L2->SM traffic per flop is 2x the R = 2 kernel's; clusters of 2 (TMA multicast)
// bring it back.
#include <cfloat>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "maxsim_epilogue.cuh"
#include "maxsim_params.h"
#include "sm100_ptx.cuh"

namespace cpb {
namespace kpipe {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kQPanelBytes = kTileM * 64 * 2;   // 16 KiB
constexpr int kDPanelBytes = kTileN * 64 * 2;   // 32 KiB = one ring stage
constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr int R = 1;

template <int KP>
struct SmemLayout {
  static constexpr int kQOff = 0;
  static constexpr int kDOff = KP * kQPanelBytes;
  static constexpr int kBcOff = kDOff + kStages * kDPanelBytes;  // argmax mode: per-lane best-chunk cache (one query tile)
  static constexpr int kBarOff = kBcOff + kBcBytesPerTile;
  static constexpr int kNumBars = 1 + 2 * kStages + 4;  // q_full, full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemPtrOff + 16;
  static constexpr int kAlloc = kBytes + 1024;
  static_assert(kAlloc <= 227 * 1024, "shared memory budget of one CTA");
};

template <int KP, int kMode>
__global__ void __launch_bounds__(kThreads, 1)
maxsim_kpipe_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                    const __grid_constant__ CUtensorMap tmap_tail, const MaxSimParams p, const LossParams lp) {
  using L = SmemLayout<KP>;
  constexpr int S = kStages;

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

  const int C = p.cluster;
  const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const int cluster_id = blockIdx.x / C;
  const int g = (cluster_id % p.group_sets) * C + static_cast<int>(crank);  // query tile (R = 1: group == tile)
  const int part = cluster_id / p.group_sets;
  const int r_cnt = max(0, min(R, p.num_qtiles - g * R));
  int d0 = static_cast<int>((static_cast<int64_t>(p.n_docs) * part) / p.doc_parts);
  int d1 = static_cast<int>((static_cast<int64_t>(p.n_docs) * (part + 1)) / p.doc_parts);
  int bal_r0 = 0, bal_r1 = 0;
  if (p.balanced) {
    const int64_t tiles = (p.bank_rows + kTileN - 1) / kTileN;
    bal_r0 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * part) / p.doc_parts)));
    bal_r1 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * (part + 1)) / p.doc_parts)));
    d0 = 0;
    d1 = p.n_docs;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_tail);
    mbar_init(q_full, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], static_cast<uint32_t>(C));
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  if (C > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  maxsim_pdl_entry(p);
  const long long dbg_c0 = clock64();
  const uint64_t dbg_t0 = global_timer_ns();

  if (warp == 0) {
    // ================================ TMA producer: one K-panel of one document tile per ring stage ==========
    if (lane == 0) {
      if (r_cnt > 0) {
        mbar_expect_tx(q_full, static_cast<uint32_t>(KP) * kQPanelBytes);
        for (int kp = 0; kp < KP; ++kp) tma_load_2d(q_smem + kp * kQPanelBytes, &tmap_q, kp * 64, g * kTileM, q_full);
      }
      const int rows_per_cta = kTileN / C;
      int stage = 0;
      uint32_t phase = 0;
      for (int d = d0; d < d1;) {
        const Run run = next_run(p, d, d1, bal_r0, bal_r1);
        d = run.e;
        for (int row = run.row0; row < run.row1; row += kTileN) {
          const int n_valid = min(kTileN, run.row1 - row);
          for (int kp = 0; kp < KP; ++kp) {
            mbar_wait(&empty[stage], phase ^ 1u);
            uint8_t* dst = d_smem + stage * kDPanelBytes;
            if (n_valid == kTileN) {
              mbar_expect_tx(&full[stage], kDPanelBytes);
              const int r0 = static_cast<int>(crank) * rows_per_cta;
              if (C > 1) tma_load_2d_mc(dst + r0 * 128, &tmap_d, kp * 64, row + r0, &full[stage], cmask);
              else tma_load_2d(dst, &tmap_d, kp * 64, row, &full[stage]);
            } else {
              const int nbox = (n_valid + 31) >> 5;
              mbar_expect_tx(&full[stage], static_cast<uint32_t>(nbox) * 32u * 128u);
              if (crank == 0) {
                for (int b = 0; b < nbox; ++b) {
                  if (C > 1) tma_load_2d_mc(dst + b * 4096, &tmap_tail, kp * 64, row + b * 32, &full[stage], cmask);
                  else tma_load_2d(dst + b * 4096, &tmap_tail, kp * 64, row + b * 32, &full[stage]);
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
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform, software-pipelined waits) ====================
    {
      if (r_cnt > 0) mbar_wait(q_full, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t d_addr = smem_u32(d_smem);
      // cursor over (tile, panel) units; a unit = 4 MMAs reading one ring stage
      int it_d = d0, it_row = 0, it_row1 = 0;
      bool it_have_tile = false;
      auto next_tile = [&]() {
        if (it_have_tile) it_row += kTileN;
        while (!it_have_tile || it_row >= it_row1) {
          if (it_d >= d1) return false;
          const Run run = next_run(p, it_d, d1, bal_r0, bal_r1);
          it_d = run.e;
          it_row = run.row0;
          it_row1 = run.row1;
          it_have_tile = true;
        }
        return true;
      };
      struct Unit {
        bool valid, first_of_tile, last_of_tile;
        int stage, kp;
        uint32_t phase, tile, idesc;
        uint64_t a_desc0, b_desc0;
      };
      int u_stage = 0, u_kp = 0;
      uint32_t u_phase = 0, u_tile = 0, u_idesc = 0;
      bool started = false;
      auto advance = [&](Unit& u) {
        if (!started || u_kp == 0) {
          if (started) ++u_tile;
          if (!next_tile()) { u.valid = false; return; }
          const int n_valid = min(kTileN, it_row1 - it_row);
          u_idesc = make_idesc_bf16_f32(kTileM, static_cast<uint32_t>((n_valid + 15) & ~15));
          if (!started) u_tile = 0;
          started = true;
        }
        u.valid = true;
        u.kp = u_kp;
        u.first_of_tile = (u_kp == 0);
        u.last_of_tile = (u_kp == KP - 1);
        u.stage = u_stage;
        u.phase = u_phase;
        u.tile = u_tile;
        u.idesc = u_idesc;
        u_kp = (u_kp + 1 == KP) ? 0 : u_kp + 1;
        if (++u_stage == S) {
          u_stage = 0;
          u_phase ^= 1u;
        }
      };
      auto prepare = [&](Unit& u) {
        mbar_wait(&full[u.stage], u.phase);
        if (u.first_of_tile && r_cnt > 0) mbar_wait(&tmem_empty[u.tile & 1u], ((u.tile >> 1) & 1u) ^ 1u);
        tc_fence_after();
        u.a_desc0 = make_sw128_kmajor_desc(q_addr + u.kp * kQPanelBytes);
        u.b_desc0 = make_sw128_kmajor_desc(d_addr + u.stage * kDPanelBytes);
      };
      auto issue = [&](const Unit& u, int k_lo, int k_hi) {
        const uint32_t d_tmem = tmem_base + (u.tile & 1u) * kTileN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= k_lo && k < k_hi)
            umma_bf16(d_tmem, u.a_desc0 + static_cast<uint64_t>(k * 2), u.b_desc0 + static_cast<uint64_t>(k * 2), u.idesc,
                      (u.kp > 0 || k > 0) ? 1u : 0u);
        }
      };
      Unit cur{};
      advance(cur);
      if (cur.valid) prepare(cur);
      while (cur.valid) {
        Unit nxt{};
        advance(nxt);
        if (r_cnt > 0 && elect_one()) issue(cur, 0, 2);
        __syncwarp();
        if (nxt.valid) prepare(nxt);
        if (elect_one()) {
          if (r_cnt > 0) issue(cur, 2, 4);
          // ring stage free (in every CTA of the cluster) once these MMAs have read it
          if (C > 1) umma_commit_mc(&empty[cur.stage], cmask); else umma_commit(&empty[cur.stage]);
          if (cur.last_of_tile && r_cnt > 0) umma_commit(&tmem_full[cur.tile & 1u]);
        }
        __syncwarp();
        cur = nxt;
      }
    }
  } else {
    // ================================ epilogue (maxsim_epilogue.cuh) ==========================
    const CtaSlice sl{g, part, r_cnt, d0, d1, bal_r0, bal_r1};
    maxsim_epilogue<R, kMode>(p, sl, tmem_base, tmem_full, tmem_empty, warp, lane, smem + L::kBcOff);
  }

  // ---- teardown ---------------------------------------------------------------------------
  maxsim_finish(p, lp, C, warp);
  if ((p.flags & CPB_DBG_CLOCKS) && threadIdx.x == 0) {
    p.scores[2 * blockIdx.x] = static_cast<float>(clock64() - dbg_c0);
    p.scores[2 * blockIdx.x + 1] = static_cast<float>(global_timer_ns() - dbg_t0);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


static void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster, int smem,
                     cudaStream_t stream, int pdl) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl != 0) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
}

template <int KP, int kMode>
static cudaError_t launch_variant(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt,
                                  const MaxSimParams& p, const LossParams& lp, int grid, cudaStream_t stream) {
  auto kern = maxsim_kpipe_kernel<KP, kMode>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<KP>::kAlloc);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, grid, p.cluster, SmemLayout<KP>::kAlloc, stream, p.pdl);
  return cudaLaunchKernelEx(&cfg, kern, tq, td, tt, p, lp);
}

template <int KP>
static int max_clusters_variant(int cluster) {
  auto kern = maxsim_kpipe_kernel<KP, kModeMax>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<KP>::kAlloc) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, cluster, cluster, SmemLayout<KP>::kAlloc, nullptr, 0);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
}

template <int KP>
static cudaError_t launch_mode(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                               const LossParams& lp, int mode, int grid, cudaStream_t stream) {
  if (mode == kModeArgmax) return launch_variant<KP, kModeArgmax>(tq, td, tt, p, lp, grid, stream);
  if (mode == kModeSmooth) return launch_variant<KP, kModeSmooth>(tq, td, tt, p, lp, grid, stream);
  return launch_variant<KP, kModeMax>(tq, td, tt, p, lp, grid, stream);
}

}  // namespace kpipe

int maxsim_kpipe_max_clusters(int dim_panels, int cluster) {
  switch (dim_panels) {
    case 3: return kpipe::max_clusters_variant<3>(cluster);
    case 4: return kpipe::max_clusters_variant<4>(cluster);
    case 5: return kpipe::max_clusters_variant<5>(cluster);
    default: return 0;
  }
}

// dim_panels = padded embedding dim / 64 (3, 4 or 5; dims <= 128 use maxsim_sm100.cu); mode as in maxsim_launch
cudaError_t maxsim_kpipe_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                                const LossParams& lp, int dim_panels, int mode, int grid, cudaStream_t stream) {
  using namespace kpipe;
  switch (dim_panels) {
    case 3: return launch_mode<3>(tq, td, tt, p, lp, mode, grid, stream);
    case 4: return launch_mode<4>(tq, td, tt, p, lp, mode, grid, stream);
    case 5: return launch_mode<5>(tq, td, tt, p, lp, mode, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cpb
