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
//     resident query tiles (tcgen05.mma 128 x 256 x 16, 8 K-steps) into two 256-column fp32 TMEM
//     accumulators used as a ping-pong, so the epilogue of job j overlaps the MMAs of job j+1.
//   * Documents are (start row, length) intervals of a flat [tokens, 128] bf16 bank, so ragged banks,
//     left/right padded batches and dense [B_d, N_d, 128] tensors are all the same kernel.  Tiles cover
//     the bank in 256-row steps across document boundaries: back-to-back documents form a "run" that is
//     tiled without gaps, and the epilogue walks the document boundaries inside each accumulator tile
//     (a measured ~22-cycle fixed cost per tcgen05.mma makes short per-document tail MMAs expensive:
//     profiles/r01_notes.md).  Only the last tile of a run is short: it is fetched in 32-row boxes and
//     issued with a smaller MMA N (multiple of 16).
//   * CTAs are launched as clusters of C (1, 2 or 4): the C CTAs of a cluster hold different query tiles,
//     walk the same document partition in lock-step and each fetches 1/C of every document tile with a
//     TMA multicast, so a bank byte crosses the L2->SM fabric once per C*R query tiles.
//   * warp 0: TMA producer.  warp 1: TMEM allocator + MMA issuer.  warps 2-5: epilogue.
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

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kDim = 128;                       // embedding dim handled per pass (2 swizzle panels of 64)
constexpr int kQTileBytes = kTileM * kDim * 2;  // 32 KiB
constexpr int kQPanelBytes = kTileM * 64 * 2;   // 16 KiB
constexpr int kDTileBytes = kTileN * kDim * 2;  // 64 KiB
constexpr int kDPanelBytes = kTileN * 64 * 2;   // 32 KiB
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
#ifndef CPB_MMA_SPLIT
#define CPB_MMA_SPLIT 6
#endif
constexpr uint32_t kTmemCols = 512;

template <int R, int kMode>
struct SmemLayout {
  // argmax mode: 16 KiB per resident query tile for the per-lane best-chunk cache (maxsim_epilogue.cuh)
  static constexpr int kBcBytes = (kMode == kModeArgmax) ? R * kBcBytesPerTile : 0;
  static constexpr int kStages = (R == 1 && kMode != kModeArgmax) ? 3 : 2;
  static constexpr int kQOff = 0;
  static constexpr int kDOff = R * kQTileBytes;
  static constexpr int kBcOff = kDOff + kStages * kDTileBytes;
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
maxsim_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                  const __grid_constant__ CUtensorMap tmap_tail, const MaxSimParams p, const LossParams lp) {
  using L = SmemLayout<R, kMode>;
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
  // A cluster of C CTAs shares one document partition: every CTA holds different query tiles and
  // loads 1/C of each document tile, multicast to all C (each bank byte leaves L2 once per cluster).
  const int C = p.cluster;
  const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const int cluster_id = blockIdx.x / C;
  const int g = (cluster_id % p.group_sets) * C + static_cast<int>(crank);  // query-tile group
  const int part = cluster_id / p.group_sets;                                // document partition
  const int r_cnt = max(0, min(R, p.num_qtiles - g * R));
  int d0 = static_cast<int>((static_cast<int64_t>(p.n_docs) * part) / p.doc_parts);
  int d1 = static_cast<int>((static_cast<int64_t>(p.n_docs) * (part + 1)) / p.doc_parts);
  // Balanced mode: partitions are equal shares of the bank's 256-row TILES instead of whole documents, so every CTA
  // gets the same number of MMA jobs (at cfg2 one partition of 37 had 28 documents instead of 27: 3.5 % of the kernel).
  // A document cut by a partition boundary is folded by both neighbours; the right one publishes its partial per-token
  // maxima, the left one combines and emits the score (see the epilogue).
  int bal_r0 = 0, bal_r1 = 0;
  if (p.balanced) {
    const int64_t tiles = (p.bank_rows + kTileN - 1) / kTileN;
    bal_r0 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * part) / p.doc_parts)));
    bal_r1 = static_cast<int>(min(static_cast<int64_t>(p.bank_rows), kTileN * ((tiles * (part + 1)) / p.doc_parts)));
    d0 = 0;
    d1 = p.n_docs;
  }

  // ---- one-time setup ---------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_tail);
    mbar_init(q_full, 1);
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
  maxsim_pdl_entry(p);
  const long long dbg_c0 = clock64();
  const uint64_t dbg_t0 = global_timer_ns();

  // Register budgets follow the roles (see kEpiWarp0): set at the top of each branch so that the allocation is
  // unambiguous along every path, and restored to the launch value before the paths join again.
  constexpr bool kRealloc = kEpiGroups<R, kMode> == 2;
  if (warp < kEpiWarp0<R, kMode>) {
  if constexpr (kRealloc) setmaxnreg_dec<kRegsProducer>();
  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      if (r_cnt > 0) {
        mbar_expect_tx(q_full, static_cast<uint32_t>(r_cnt) * kQTileBytes);
        for (int r = 0; r < r_cnt; ++r) {
          const int row = (g * R + r) * kTileM;
          tma_load_2d(q_smem + r * kQTileBytes, &tmap_q, 0, row, q_full);
          tma_load_2d(q_smem + r * kQTileBytes + kQPanelBytes, &tmap_q, 64, row, q_full);
        }
      }
      const int rows_per_cta = kTileN / C;
      int stage = 0;
      uint32_t phase = 0;
      for (int d = d0; d < d1;) {
        const Run run = next_run(p, d, d1, bal_r0, bal_r1);
        d = run.e;
        for (int row = run.row0; row < run.row1; row += kTileN) {
          const int n_valid = min(kTileN, run.row1 - row);
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* dst = d_smem + stage * kDTileBytes;
          if (p.flags & CPB_DBG_NO_TMA) {
            mbar_arrive(&full[stage]);
          } else if (n_valid == kTileN) {
            // full tile: this CTA fetches rows [crank*256/C, (crank+1)*256/C) for the whole cluster
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
            // last tile of a run: 32-row boxes, fetched by rank 0 only
            const int nbox = (n_valid + 31) >> 5;
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
    // The WHOLE warp runs this loop (waits, descriptor arithmetic) so that every value is provably warp-uniform
    // and stays in uniform registers; only the tcgen05 instructions themselves are issued by one elected lane.
    //
    // The tensor pipe's instruction queue is shallow: whatever the issuer does between the last MMA of one job and
    // the first MMA of the next (two mbarrier try_waits at ~90 cycles each even when already complete, fences,
    // descriptor setup) shows up as tensor idle time (measured: ~165 cycles per 1024-cycle job, profiles/r01_notes.md).
    // So the issue stream is software-pipelined: the waits and set-up of job j+1 are performed after K-step kSplit
    // of job j, while the previous K-steps are still queued, and the remaining K-steps of job j follow immediately.
    {
      if (r_cnt > 0) mbar_wait(q_full, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t d_addr = smem_u32(d_smem);
      const bool dbg = (p.flags & CPB_DBG_CLOCKS) != 0;
      long long w_full = 0, w_tmem = 0;

      struct Job {
        bool valid, first_of_tile, last_of_tile;
        int stage, r;
        uint32_t phase, job, idesc;
        uint64_t a_desc0, b_desc0;
      };
      // tile cursor over the non-empty runs of this partition
      int it_d = d0, it_row = 0, it_row1 = 0, it_stage = 0;
      uint32_t it_phase = 0, it_job = 0;
      bool it_have_tile = false;
      auto next_tile = [&]() {  // advance to the next 256-row tile; returns false at the end of the partition
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
      auto advance = [&](Job& j) {  // fill j with the job after the current cursor position
        if (r_cnt == 0) { j.valid = false; return; }
        if (r_next == 0) {
          if (!next_tile()) { j.valid = false; return; }
        }
        const int n_valid = min(kTileN, it_row1 - it_row);
        j.valid = true;
        j.r = r_next;
        j.first_of_tile = (r_next == 0);
        j.last_of_tile = (r_next == r_cnt - 1);
        j.stage = it_stage;
        j.phase = it_phase;
        j.job = it_job++;
        j.idesc = make_idesc_bf16_f32(kTileM, static_cast<uint32_t>((n_valid + 15) & ~15));
        r_next = (r_next + 1 == r_cnt) ? 0 : r_next + 1;
      };
      auto prepare = [&](Job& j) {  // block until job j's operands and accumulator are available
        long long t0 = dbg ? clock64() : 0;
        if (j.first_of_tile) mbar_wait(&full[j.stage], j.phase);
        if (dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
        mbar_wait(&tmem_empty[j.job & 1u], ((j.job >> 1) & 1u) ^ 1u);
        if (dbg) w_tmem += clock64() - t0;
        tc_fence_after();
        j.a_desc0 = make_sw128_kmajor_desc(q_addr + j.r * kQTileBytes);
        j.b_desc0 = make_sw128_kmajor_desc(d_addr + j.stage * kDTileBytes);
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
            // K step k: panel k/4 (16 KiB / 32 KiB apart), 32 bytes (2 encoded units) per step inside it
            const uint64_t adesc = j.a_desc0 + static_cast<uint64_t>((k >> 2) * (kQPanelBytes >> 4) + (k & 3) * 2);
            const uint64_t bdesc = j.b_desc0 + static_cast<uint64_t>((k >> 2) * (kDPanelBytes >> 4) + (k & 3) * 2);
            umma_bf16(d_tmem, adesc, bdesc, j.idesc, k > 0 ? 1u : 0u);
          }
        }
      };
      if (r_cnt == 0) {
        // phantom CTA (odd number of query-tile groups in a cluster): no MMAs, but the cluster's TMA ring still
        // needs this CTA's release of every stage (the `empty` barriers count one arrival per CTA)
        while (next_tile()) {
          mbar_wait(&full[it_stage], it_phase);
          tc_fence_after();
          if (elect_one()) {
            if (C > 1) umma_commit_mc(&empty[it_stage], cmask); else umma_commit(&empty[it_stage]);
          }
          __syncwarp();
        }
      }
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
          umma_commit(&tmem_full[cur.job & 1u]);
          if (cur.last_of_tile) {  // smem slot is free (in every CTA of the cluster) once these MMAs have read it
            if (C > 1) umma_commit_mc(&empty[cur.stage], cmask); else umma_commit(&empty[cur.stage]);
          }
        }
        __syncwarp();
        if (nxt.valid && !early) prepare(nxt);
        cur = nxt;
      }
      if (dbg && lane == 0) {  // cycles the issuer spent blocked on TMA data / on the epilogue
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
    maxsim_epilogue<R, kMode, false, kEpiGroups<R, kMode>>(p, sl, tmem_base, tmem_full, tmem_empty, warp, lane,
                                                           smem + L::kBcOff, (warp - kEpiWarp0<R, kMode>) >> 2);
    if constexpr (kRealloc) setmaxnreg_dec<kRegsLaunch>();
  }

  // ---- teardown ---------------------------------------------------------------------------
  // fused top-k (nq_pad == 32 there): this group's queries are g * R * 4 ... (4 per 128-row query tile)
  const int tk_q0 = g * R * (kTileM / 32);
  maxsim_finish(p, lp, C, warp, g, part, tk_q0, max(0, min(r_cnt * (kTileM / 32), p.n_queries - tk_q0)));
  if ((p.flags & CPB_DBG_CLOCKS) && threadIdx.x == 0) {
    p.scores[2 * blockIdx.x] = static_cast<float>(clock64() - dbg_c0);
    p.scores[2 * blockIdx.x + 1] = static_cast<float>(global_timer_ns() - dbg_t0);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Wait (on the stream) until flags[0..n) have all reached `value` (launch counters only grow; wrap-safe compare):
// consumer side of the fused all-gather.
// values[i] is what flags[i] must have reached (nullptr: `value` for all).  A peer that is late is not an error (first
// call set-up, a straggler): the wait reports a time-out through *status instead of trapping, after timeout_ms.
__global__ void wait_flags_kernel(const uint32_t* flags, const uint32_t* values, int n, uint32_t value,
                                  uint32_t timeout_ms, uint32_t* status) {
  if (threadIdx.x >= n) return;
  const uint32_t* f = flags + threadIdx.x;
  const uint32_t want = values ? values[threadIdx.x] : value;
  const uint64_t t0 = global_timer_ns();
  const uint64_t limit = static_cast<uint64_t>(timeout_ms) * 1000000ull;
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (static_cast<int32_t>(v - want) >= 0) break;
    if (global_timer_ns() - t0 > limit) {
      if (status) atomicOr(status, 1u << (threadIdx.x & 31));
      break;
    }
    __nanosleep(64);
  }
}
cudaError_t wait_flags_launch(const uint32_t* flags, const uint32_t* values, int n, uint32_t value, uint32_t timeout_ms,
                              uint32_t* status, cudaStream_t stream) {
  wait_flags_kernel<<<1, 64, 0, stream>>>(flags, values, n, value, timeout_ms, status);
  return cudaGetLastError();
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

static void fill_cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int cluster, int threads,
                             int smem_bytes, cudaStream_t stream, int pdl = 0) {
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl != 0) {  // this launch may start while the previous kernel of the stream is still draining
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
}

template <int R, int kMode>
static cudaError_t launch_variant(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt,
                                  const MaxSimParams& p, const LossParams& lp, int grid, cudaStream_t stream) {
  auto kern = maxsim_fwd_kernel<R, kMode>;
  // per instantiation and per device: the attribute is sticky once set (atomic: launches may come from several threads)
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
  fill_cluster_cfg(cfg, attr, grid, p.cluster, kThreads<R, kMode>, SmemLayout<R, kMode>::kAlloc, stream, p.pdl);
  return cudaLaunchKernelEx(&cfg, kern, tq, td, tt, p, lp);
}

// mode: kModeMax / kModeArgmax / kModeSmooth (maxsim_epilogue.cuh)
cudaError_t maxsim_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                          const LossParams& lp, int r, int mode, int grid, cudaStream_t stream) {
  if (r == 1) {
    if (mode == kModeArgmax) return launch_variant<1, kModeArgmax>(tq, td, tt, p, lp, grid, stream);
    if (mode == kModeSmooth) return launch_variant<1, kModeSmooth>(tq, td, tt, p, lp, grid, stream);
    return launch_variant<1, kModeMax>(tq, td, tt, p, lp, grid, stream);
  }
  if (mode == kModeArgmax) return launch_variant<2, kModeArgmax>(tq, td, tt, p, lp, grid, stream);
  if (mode == kModeSmooth) return launch_variant<2, kModeSmooth>(tq, td, tt, p, lp, grid, stream);
  return launch_variant<2, kModeMax>(tq, td, tt, p, lp, grid, stream);
}

// How many clusters of `cluster` CTAs can be co-resident (persistent-grid sizing).
template <int R>
static int max_clusters_variant(int cluster) {
  auto kern = maxsim_fwd_kernel<R, kModeMax>;
  constexpr int kSmem = SmemLayout<R, kModeMax>::kAlloc;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  fill_cluster_cfg(cfg, attr, cluster, cluster, kThreads<R, kModeMax>, kSmem, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
}
int maxsim_max_clusters(int r, int cluster) {
  return r == 1 ? max_clusters_variant<1>(cluster) : max_clusters_variant<2>(cluster);
}

int64_t maxsim_topk_workspace_bytes() { return kTopkWorkspaceBytes; }
int maxsim_topk_slots() { return kTopkSlots; }
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
