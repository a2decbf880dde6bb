// Fused MaxSim for sm_100a, CTA-pair variant (tcgen05.mma.cta_group::2), embedding dim 128.
//
// Same problem, same epilogue and same partitioning as maxsim_sm100.cu (see its header); what changes is how a
// document tile reaches the tensor cores.  There, the two CTAs of a cluster each hold the WHOLE 256-token tile (half
// fetched, half received by TMA multicast) and each runs its own 128 x 256 x 16 MMAs: per K-step an SM reads 4 KiB of
// query rows and 8 KiB of document rows from shared memory in 128 tensor cycles, and receives 64 KiB of tile per two
// jobs -- ~120 of the 128 B/clk the shared memory delivers.  Here the two CTAs form a PAIR: one
// tcgen05.mma.cta_group::2 of M = 256 takes the query tile of CTA 0 (rows 0-127 of A) and the one of CTA 1 (rows
// 128-255), and HALF of the document tile from each CTA's shared memory (128 tokens each).  Every CTA gets its own 128
// rows of the result in its own TMEM, so the epilogue does not change, but an SM now reads 8 KiB per K-step and
// receives 32 KiB per tile: the ring holds four tiles instead of two in less space.
//
//   * leader = cluster rank 0: issues every MMA, owns the barriers the issuer waits on (`q_full`, `full[s]`,
//     `tmem_empty[a]`).  Both CTAs' TMA loads complete on the LEADER's `full[s]` (cp.async.bulk.tensor ...
//     .cta_group::2 with the barrier's shared::cluster address); both CTAs' epilogue warps arrive on the leader's
//     `tmem_empty[a]` (8 arrivals).
//   * `empty[s]` and `tmem_full[a]` exist in both CTAs and are signalled by multicast tcgen05.commit.
//   * the host only picks this kernel when every CTA has all R query tiles (query-tile count a multiple of 2 R) and the
//     bank is contiguous; the last, short tile of a partition is fetched and multiplied as a full tile (rows past the
//     partition belong to other documents or are zero-filled by TMA; the epilogue never looks at their columns).
#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "maxsim_epilogue.cuh"
#include "maxsim_params.h"
#include "sm100_ptx.cuh"

#ifndef CPB_EARLY_SPIN_KSTEPS
#define CPB_EARLY_SPIN_KSTEPS 4  /* how long (in 128-cycle K-steps) the issuer probes for the next job's resources */
#endif

namespace cpb {
namespace pair {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kDim = 128;
constexpr int kQTileBytes = kTileM * kDim * 2;       // 32 KiB
constexpr int kQPanelBytes = kTileM * 64 * 2;        // 16 KiB
constexpr int kHalfRows = kTileN / 2;                // document tokens per CTA and tile
constexpr int kDHalfBytes = kHalfRows * kDim * 2;    // 32 KiB = one ring stage of one CTA
constexpr int kDHalfPanelBytes = kHalfRows * 64 * 2; // 16 KiB
// warp 0: TMA, warp 1: MMA issuer, then one or two groups of four epilogue warps (maxsim_epilogue.cuh: two groups, one per
// resident query tile, in the modes whose epilogue is the bottleneck)
template <int R, int kMode>
constexpr int kEpiGroups = (R == 2 && kMode != kModeMax) ? 2 : 1;
// With two groups the CTA is three warpgroups: {TMA, MMA, idle, idle} gives registers away (setmaxnreg) and each
// epilogue group is a warpgroup of its own with 224 registers per thread -- at a flat 168 (65536 / 384 threads) the
// epilogue spilled, and local memory has almost no L1 behind it in a CTA that takes 224 KiB of shared memory.
template <int R, int kMode>
constexpr int kEpiWarp0 = (kEpiGroups<R, kMode> == 2) ? 4 : 2;  // first epilogue warp
template <int R, int kMode>
constexpr int kThreads = 32 * (kEpiWarp0<R, kMode> + 4 * kEpiGroups<R, kMode>);
constexpr int kRegsLaunch = 168, kRegsProducer = 56, kRegsEpilogue = 224;  // 128 x 56 + 256 x 224 = 384 x 168
constexpr uint32_t kTmemCols = 512;
#ifndef CPB_PAIR_SPLIT
#define CPB_PAIR_SPLIT 6
#endif

template <int R, int kMode>
struct SmemLayout {
  static constexpr int kBcBytes = (kMode == kModeArgmax) ? R * kBcBytesPerTile : 0;
  static constexpr int kStages = 4;
  static constexpr int kQOff = 0;
  static constexpr int kDOff = R * kQTileBytes;
  static constexpr int kBcOff = kDOff + kStages * kDHalfBytes;
  static constexpr int kBarOff = kBcOff + kBcBytes;
  // barriers: q_full, full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kNumBars = 1 + 2 * kStages + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemPtrOff + 16;
  static constexpr int kAlloc = kBytes + 1024;  // slack for manual 1024-B alignment
  static_assert(kAlloc <= 227 * 1024, "shared memory budget of one CTA");
};

template <int R, int kMode>
__global__ void __launch_bounds__(kThreads<R, kMode>, 1)
maxsim_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                   const MaxSimParams p, const LossParams lp) {
  using L = SmemLayout<R, kMode>;
  constexpr int S = L::kStages;

  extern __shared__ uint8_t smem_raw[];
  // the same offset in both CTAs (dynamic shared memory starts at the same address in every CTA of a launch)
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

  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int g = (cluster_id % p.group_sets) * 2 + static_cast<int>(crank);  // query-tile group
  const int part = cluster_id / p.group_sets;                                // document partition
  const int r_cnt = R;                                                       // guaranteed by the host
  int d0 = static_cast<int>((static_cast<int64_t>(p.n_docs) * part) / p.doc_parts);
  int d1 = static_cast<int>((static_cast<int64_t>(p.n_docs) * (part + 1)) / p.doc_parts);
  int bal_r0 = 0, bal_r1 = 0;
  if (p.balanced) {  // equal shares of the bank's 256-row tiles (see maxsim_sm100.cu)
    const int64_t tiles = (p.bank_rows + kTileN - 1) / kTileN;
    bal_r0 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * part) / p.doc_parts)));
    bal_r1 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * (part + 1)) / p.doc_parts)));
    d0 = 0;
    d1 = p.n_docs;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    mbar_init(q_full, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);   // leader's: one arrive.expect_tx for the bytes of both CTAs
      mbar_init(&empty[s], 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);   // one multicast commit per job
      mbar_init(&tmem_empty[a], 8);  // leader's: four epilogue warps in each CTA
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  maxsim_pdl_entry(p);
  const long long dbg_c0 = clock64();
  const uint64_t dbg_t0 = global_timer_ns();

  // Register budgets follow the roles (see kEpiWarp0): set at the top of each branch so that the allocation is
  // unambiguous along every path, and restored to the launch value before the paths join again.
  constexpr bool kRealloc = kEpiGroups<R, kMode> == 2;
  if (warp < kEpiWarp0<R, kMode>) {
  if constexpr (kRealloc) setmaxnreg_dec<kRegsProducer>();
  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ==============================
    if (lane == 0) {
      const uint32_t q_full_leader = mapa_u32(smem_u32(q_full), 0u);
      const uint32_t full_leader = mapa_u32(smem_u32(full), 0u);
      if (leader) mbar_expect_tx(q_full, 2u * R * kQTileBytes);
      for (int r = 0; r < R; ++r) {
        const int row = (g * R + r) * kTileM;
        tma_load_2d_pair(q_smem + r * kQTileBytes, &tmap_q, 0, row, q_full_leader);
        tma_load_2d_pair(q_smem + r * kQTileBytes + kQPanelBytes, &tmap_q, 64, row, q_full_leader);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int d = d0; d < d1;) {
        const Run run = next_run(p, d, d1, bal_r0, bal_r1);
        d = run.e;
        for (int row = run.row0; row < run.row1; row += kTileN) {
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* dst = d_smem + stage * kDHalfBytes;
          if (leader) mbar_expect_tx(&full[stage], 2u * kDHalfBytes);
          // my half of the tile: tokens [row + 128 crank, + 128); always a full box (see the header)
          const int r0 = row + static_cast<int>(crank) * kHalfRows;
          tma_load_2d_pair(dst, &tmap_d, 0, r0, full_leader + stage * 8u);
          tma_load_2d_pair(dst + kDHalfPanelBytes, &tmap_d, 64, r0, full_leader + stage * 8u);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ==============================
    // Same software-pipelined issue stream as maxsim_sm100.cu: the waits and set-up of job j+1 happen after K-step
    // kSplit of job j, while its first K-steps are still queued in the tensor pipe.
    if (leader) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t d_addr = smem_u32(d_smem);
      const bool dbg = (p.flags & CPB_DBG_CLOCKS) != 0;
      long long w_full = 0, w_tmem = 0;
      constexpr uint32_t kIdesc = make_idesc_bf16_f32(2 * kTileM, kTileN);
      constexpr uint16_t kBoth = 0x3;

      struct Job {
        bool valid, first_of_tile, last_of_tile;
        int stage, r;
        uint32_t phase, job;
        uint64_t a_desc0, b_desc0;
      };
      int it_d = d0, it_row = 0, it_row1 = 0, it_stage = 0;
      uint32_t it_phase = 0, it_job = 0;
      bool it_have_tile = false;
      auto next_tile = [&]() {
        if (it_have_tile) {
          it_row += kTileN;
          if (++it_stage == S) {
            it_stage = 0;
            it_phase ^= 1u;
          }
        }
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
      Job cur{};
      int r_next = 0;
      auto advance = [&](Job& j) {
        if (r_next == 0) {
          if (!next_tile()) { j.valid = false; return; }
        }
        j.valid = true;
        j.r = r_next;
        j.first_of_tile = (r_next == 0);
        j.last_of_tile = (r_next == R - 1);
        j.stage = it_stage;
        j.phase = it_phase;
        j.job = it_job++;
        r_next = (r_next + 1 == R) ? 0 : r_next + 1;
      };
      auto prepare = [&](Job& j) {
        long long t0 = dbg ? clock64() : 0;
        if (j.first_of_tile) mbar_wait(&full[j.stage], j.phase);
        if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
        mbar_wait(&tmem_empty[j.job & 1u], ((j.job >> 1) & 1u) ^ 1u);
        if (dbg) w_tmem += clock64() - t0;
        tc_fence_after();
        j.a_desc0 = make_sw128_kmajor_desc(q_addr + j.r * kQTileBytes);
        j.b_desc0 = make_sw128_kmajor_desc(d_addr + j.stage * kDHalfBytes);
      };
      auto ready = [&](const Job& j) {  // would prepare(j) return without waiting?
        if (j.first_of_tile && !mbar_test_wait(&full[j.stage], j.phase)) return false;
        return mbar_test_wait(&tmem_empty[j.job & 1u], ((j.job >> 1) & 1u) ^ 1u);
      };
      auto issue = [&](const Job& j, int k_lo, int k_hi) {
        const uint32_t d_tmem = tmem_base + (j.job & 1u) * kTileN;
#pragma unroll
        for (int k = 0; k < kDim / 16; ++k) {
          if (k >= k_lo && k < k_hi) {
            const uint64_t adesc = j.a_desc0 + static_cast<uint64_t>((k >> 2) * (kQPanelBytes >> 4) + (k & 3) * 2);
            const uint64_t bdesc = j.b_desc0 + static_cast<uint64_t>((k >> 2) * (kDHalfPanelBytes >> 4) + (k & 3) * 2);
            umma_bf16_pair(d_tmem, adesc, bdesc, kIdesc, k > 0 ? 1u : 0u);
          }
        }
      };
      const long long kEarlySpin = p.early_spin;
      // K-steps issued before the next job's waits.  Compile-time: with run-time bounds every issue() became eight
      // predicated UTCHMMAs and the MMA stream ran 40 % slower (measured, profiles/r02_notes.md).
      constexpr int kSplit = 6;
      advance(cur);
      if (cur.valid) prepare(cur);
      while (cur.valid) {
        Job nxt{};
        advance(nxt);
        if (elect_one()) issue(cur, 0, kSplit);
        __syncwarp();
        // Prepare the next job between the K-steps only if that does not block: when the epilogue is the slower side
        // the wait for its accumulator would hold back the last K-steps (and the commit) of THIS job, and the epilogue
        // would in turn wait for them -- a serialisation of ~400 cycles per job in the argmax forward.
        // (the probe is repeated for at most ~kEarlySpin cycles: the K-steps already queued cover that, and in the
        // MMA-bound modes the accumulator usually frees within that time -- going the late way there costs ~50 cycles
        // per job because only two K-steps are then in flight across the next job's set-up)
        bool early = false;
        if (nxt.valid) {
          const long long t_probe = clock64();
          do {
            early = ready(nxt);
          } while (!early && clock64() - t_probe < kEarlySpin);
        }
        if (early) prepare(nxt);
        if (elect_one()) {
          issue(cur, kSplit, kDim / 16);
          umma_commit_pair(&tmem_full[cur.job & 1u], kBoth);
          if (cur.last_of_tile) umma_commit_pair(&empty[cur.stage], kBoth);  // the ring stage is free in both CTAs
        }
        __syncwarp();
        if (nxt.valid && !early) prepare(nxt);
        cur = nxt;
      }
      if (dbg && lane == 0) {
        p.scores[512 + 8 * blockIdx.x + 0] = static_cast<float>(w_full);
        p.scores[512 + 8 * blockIdx.x + 1] = static_cast<float>(w_tmem);
      }
    }
  }
  if constexpr (kRealloc) {
    // All four warps take their registers back TOGETHER, after the TMA and MMA loops: the two idle warps get here at
    // once, and if they grew back to 168 on their own (setmaxnreg is not a barrier) they would take half of what the
    // epilogue warpgroups are still waiting for in their setmaxnreg.inc -- a dead CTA about once in 300 launches.
    named_bar_sync(1, 128);
    setmaxnreg_inc<kRegsLaunch>();
  }
  } else {
    // ================================ epilogue (maxsim_epilogue.cuh) ==========================
    if constexpr (kRealloc) setmaxnreg_inc<kRegsEpilogue>();
    const CtaSlice sl{g, part, r_cnt, d0, d1, bal_r0, bal_r1};
    maxsim_epilogue<R, kMode, true, kEpiGroups<R, kMode>>(p, sl, tmem_base, tmem_full, tmem_empty, warp, lane,
                                                          smem + L::kBcOff, (warp - kEpiWarp0<R, kMode>) >> 2);
    if constexpr (kRealloc) setmaxnreg_dec<kRegsLaunch>();
  }

  // ---- teardown ---------------------------------------------------------------------------
  maxsim_finish(p, lp, 2, warp);  // cluster barrier inside: no CTA leaves while its peer can still signal it
  if ((p.flags & CPB_DBG_CLOCKS) && threadIdx.x == 0) {
    p.scores[2 * blockIdx.x] = static_cast<float>(clock64() - dbg_c0);
    p.scores[2 * blockIdx.x + 1] = static_cast<float>(global_timer_ns() - dbg_t0);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

static void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int threads, int smem,
                     cudaStream_t stream, int pdl) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
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

template <int R, int kMode>
static cudaError_t launch_variant(const CUtensorMap& tq, const CUtensorMap& td, const MaxSimParams& p,
                                  const LossParams& lp, int grid, cudaStream_t stream) {
  auto kern = maxsim_pair_kernel<R, kMode>;
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<R, kMode>::kAlloc);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, grid, kThreads<R, kMode>, SmemLayout<R, kMode>::kAlloc, stream, p.pdl);
  return cudaLaunchKernelEx(&cfg, kern, tq, td, p, lp);
}

template <int R>
static int max_clusters_variant() {
  auto kern = maxsim_pair_kernel<R, kModeArgmax>;  // the variant with the largest shared-memory footprint
  constexpr int kSmem = SmemLayout<R, kModeArgmax>::kAlloc;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, 2, kThreads<R, kModeArgmax>, kSmem, nullptr, 0);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
}

}  // namespace pair

// td: bank map with 128-row boxes.  The grid is group_sets * doc_parts clusters of 2.
cudaError_t maxsim_pair_launch(const CUtensorMap& tq, const CUtensorMap& td, const MaxSimParams& p, const LossParams& lp,
                               int r, int mode, int grid, cudaStream_t stream) {
  using namespace pair;
  if (r == 1) {
    if (mode == kModeArgmax) return launch_variant<1, kModeArgmax>(tq, td, p, lp, grid, stream);
    if (mode == kModeSmooth) return launch_variant<1, kModeSmooth>(tq, td, p, lp, grid, stream);
    return launch_variant<1, kModeMax>(tq, td, p, lp, grid, stream);
  }
  if (mode == kModeArgmax) return launch_variant<2, kModeArgmax>(tq, td, p, lp, grid, stream);
  if (mode == kModeSmooth) return launch_variant<2, kModeSmooth>(tq, td, p, lp, grid, stream);
  return launch_variant<2, kModeMax>(tq, td, p, lp, grid, stream);
}
int maxsim_pair_max_clusters(int r) { return r == 1 ? pair::max_clusters_variant<1>() : pair::max_clusters_variant<2>(); }

}  // namespace cpb
