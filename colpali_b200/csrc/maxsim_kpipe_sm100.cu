// K-pipelined fused MaxSim for embedding dims up to 320 (ColQwen3: colpali_engine/models/qwen3/colqwen3/modeling_colqwen3.py:48).
//
// DRAFT (branch r2-drafts): compiles for sm_100a, NOT yet run on a GPU.  Same contract, epilogue and partitioning as
// maxsim_sm100.cu (the epilogue / teardown text below is carved verbatim from it, R = 1); what differs is the K loop:
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
  static constexpr int kBarOff = kDOff + kStages * kDPanelBytes;
  static constexpr int kNumBars = 1 + 2 * kStages + 4;  // q_full, full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kBytes = kTmemPtrOff + 16;
  static constexpr int kAlloc = kBytes + 1024;
};

// A run = documents [d, e) stored back to back in the bank, rows [row0, row1): tiled without gaps.
// With CPB_FLAG_CONTIGUOUS the caller guarantees start[j+1] == start[j] + len[j] for the whole bank, so a
// CTA's partition is ONE run (two loads, no scan); otherwise every document is its own run.
struct Run {
  int e, row0, row1;
};
__device__ __forceinline__ Run next_run(const MaxSimParams& p, int d, int d1, int bal_r0, int bal_r1) {
  Run r;
  if (p.balanced) {  // the partition is the row range [bal_r0, bal_r1), whatever documents it cuts
    r.e = d1;
    r.row0 = bal_r0;
    r.row1 = bal_r1;
    return r;
  }
  r.row0 = __ldg(p.doc_start + d);
  if (p.flags & CPB_FLAG_CONTIGUOUS) {
    r.e = d1;
    r.row1 = __ldg(p.doc_start + d1 - 1) + __ldg(p.doc_len + d1 - 1);
  } else {
    r.e = d + 1;
    r.row1 = r.row0 + __ldg(p.doc_len + d);
  }
  return r;
}

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

// max of 32 accumulator columns (15 FMNMX3)
__device__ __forceinline__ float tree32(const uint32_t (&v)[32]) {
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
  return fmax3(fmax3(w0, w1, u2), u3, u3);
}

// max over columns lo <= i < hi of a 32-column chunk
__device__ __forceinline__ float max32_range(const uint32_t (&v)[32], float m, int lo, int hi) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = (i >= lo && i < hi) ? __uint_as_float(v[i]) : -INFINITY;
    m = fmaxf(m, x);
  }
  return m;
}

// running (value, first index) argmax over columns lo <= i < hi; strict '>' keeps the earliest maximum,
// which is what torch.max(dim) returns on ties.  idx0 = document-relative index of column 0 of the chunk.
__device__ __forceinline__ void argmax32_range(const uint32_t (&v)[32], float& m, int& idx, int idx0, int lo, int hi) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = __uint_as_float(v[i]);
    bool take = (i >= lo) && (i < hi) && (x > m);
    m = take ? x : m;
    idx = take ? (idx0 + i) : idx;
  }
}

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int KP, bool kArgmax>
__global__ void __launch_bounds__(kThreads, 1)
maxsim_kpipe_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                    const __grid_constant__ CUtensorMap tmap_tail, const MaxSimParams p) {
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
    // ================================ epilogue (verbatim from maxsim_sm100.cu, R = 1) =========================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const bool round_ref = (p.flags & CPB_FLAG_ROUND_BF16) != 0;
    const bool skip = (p.flags & CPB_DBG_SKIP_EPILOGUE) != 0;

    // document `doc` is complete for resident query tile r: fold this query segment's 32 token maxima
    auto finalize = [&](int doc, int r, float mm, int ai) {
      const int row0 = (g * R + r) * kTileM + quad * 32;  // first padded query row of this warp
      const int q = row0 / p.nq_pad;
      const int seg = (row0 % p.nq_pad) >> 5;
      if (kArgmax && p.argmax != nullptr && row0 + lane < p.q_rows)
        p.argmax[static_cast<int64_t>(doc) * p.q_rows + row0 + lane] = ai;
      float x = round_ref ? round_bf16(mm) : mm;
      x = warp_sum(x);
      if (round_ref && p.nq_pad == 32) x = round_bf16(x);
      if (lane == 0 && q < p.n_queries && !(p.flags & CPB_DBG_CLOCKS)) {
        if (p.peer_scores != nullptr) {
          // fused all-gather of the score slabs: one 4-byte store per peer GPU, straight into its copy of
          // gathered[my_rank] through the NVLink peer mapping (no collective kernel afterwards, only a barrier)
          const int64_t off = p.peer_slab_offset + static_cast<int64_t>(q) * p.n_docs + doc;
          for (int pr = 0; pr < p.n_peers; ++pr) reinterpret_cast<float*>(__ldg(p.peer_scores + pr))[off] = x;
        } else {
          p.scores[static_cast<int64_t>(seg) * p.plane_stride + static_cast<int64_t>(q) * p.n_docs + doc] = x;
        }
      }
    };
    auto doc_init = [&](int doc) { return (p.doc_floor != nullptr) ? __ldg(p.doc_floor + doc) : -INFINITY; };

    float m[R];
    int am[R];
    uint32_t job = 0;
    const bool dbg = (p.flags & CPB_DBG_CLOCKS) != 0;
    long long e_wait = 0, e_hold = 0, e_post = 0, e_hold2 = 0;
    int n_path2 = 0;

    // ---- balanced mode: which document contains the first row of my partition, and is it cut? ----------------
    int first_doc = d0;
    bool head_frag = false;
    if (p.balanced) {
      if (bal_r0 == 0) {
        first_doc = 0;
      } else if (p.uniform_len > 0) {
        first_doc = bal_r0 / p.uniform_len;
      } else {  // last document that starts at or before bal_r0 (starts are sorted in a contiguous bank)
        int lo = 0, hi = p.n_docs - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (__ldg(p.doc_start + mid) <= bal_r0) lo = mid; else hi = mid - 1;
        }
        first_doc = lo;
      }
      first_doc = min(first_doc, p.n_docs - 1);
      head_frag = __ldg(p.doc_start + first_doc) < bal_r0;
    }
    // slot of the boundary between partitions `bp` and `bp + 1` for resident query tile r
    auto split_slot = [&](int bp, int r) { return ((g * p.doc_parts + bp) * R + r); };
    auto publish = [&](int r, float mm, int ai) {  // I hold the RIGHT part of a cut document
      const int slot = split_slot(part - 1, r);
      p.split_max[slot * 128 + quad * 32 + lane] = mm;
      if (kArgmax) p.split_idx[slot * 128 + quad * 32 + lane] = ai;
      __threadfence();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(p.split_flag + slot * 4 + quad) = p.epoch;
    };
    auto consume = [&](int r, float& mm, int& ai) {  // I hold the LEFT part: wait for the neighbour's partial, combine
      const int slot = split_slot(part, r);
      if (lane == 0) {
        const volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(p.split_flag + slot * 4 + quad);
        const uint64_t t0 = global_timer_ns();
        while (*f != p.epoch) {
          if (global_timer_ns() - t0 > 4000000000ull) __trap();
        }
      }
      __syncwarp();
      __threadfence();
      const float om = __ldcg(p.split_max + slot * 128 + quad * 32 + lane);
      const int oi = kArgmax ? __ldcg(p.split_idx + slot * 128 + quad * 32 + lane) : -1;
      if (om > mm) {  // on a tie the earlier (left) token wins, like torch.max
        mm = om;
        ai = oi;
      }
    };

    for (int d = p.balanced ? first_doc : d0; d < d1;) {
      const Run run = next_run(p, d, d1, bal_r0, bal_r1);
      if (run.row1 == run.row0) {
        // nothing but empty documents: their score is the sum of the floors (balanced mode: an empty partition)
        if (!p.balanced)
          for (int e = d; e < run.e; ++e)
            for (int r = 0; r < r_cnt; ++r) finalize(e, r, doc_init(e), -1);
        d = run.e;
        continue;
      }
      // state at the start of each tile: current document, its first/last bank row, running maxima
      int cur = d;
      int cur_row0 = p.balanced ? __ldg(p.doc_start + cur) : run.row0;
      int cur_end = cur_row0 + __ldg(p.doc_len + cur);
      // length and floor of the FOLLOWING document are fetched when the cursor moves, long before they are needed:
      // a global load while the accumulator is held costs ~300 cycles of tensor idle time on boundary tiles
      int cur_nlen = (cur + 1 < run.e) ? __ldg(p.doc_len + cur + 1) : 0;
      float cur_ninit = (cur + 1 < run.e) ? doc_init(cur + 1) : -INFINITY;
      {
        const float init = doc_init(cur);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          m[r] = init;
          am[r] = -1;
        }
      }
      for (int row = run.row0; row < run.row1; row += kTileN) {
        const int n_valid = min(kTileN, run.row1 - row);
        const int tile_end = row + n_valid;
        int nxt = cur, nxt_row0 = cur_row0, nxt_end = cur_end, nxt_nlen = cur_nlen;
        float nxt_ninit = cur_ninit;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r < r_cnt) {
            const uint32_t a = job & 1u;
            const uint32_t aphase = (job >> 1) & 1u;
            const long long t0 = dbg ? clock64() : 0;
            mbar_wait(&tmem_full[a], aphase);
            const long long t1 = dbg ? clock64() : 0;
            long long t2 = 0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_base + a * kTileN;
            float mm = m[r];
            int ai = am[r];
            int doc = cur, doc_row0 = cur_row0, doc_end = cur_end, doc_nlen = cur_nlen;
            float doc_ninit = cur_ninit;

            // the current document is complete: emit it and step to the next one of the run
            auto finish_doc = [&]() {
              if (head_frag && doc == first_doc) publish(r, mm, ai); else finalize(doc, r, mm, ai);
              ++doc;
              if (doc >= run.e) {
                doc_end = 0x7fffffff;  // run exhausted
                return;
              }
              doc_row0 = doc_end;
              doc_end = doc_row0 + doc_nlen;
              mm = doc_ninit;
              ai = -1;
              doc_nlen = (doc + 1 < run.e) ? __ldg(p.doc_len + doc + 1) : 0;
              doc_ninit = (doc + 1 < run.e) ? doc_init(doc + 1) : -INFINITY;
            };
            auto release_acc = [&]() {  // accumulator drained: hand the TMEM stage back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[a]);
              if (dbg) t2 = clock64();
            };

            // Three warp-uniform cases.  (1) the whole 256-column tile belongs to one document: FMNMX3 trees.
            // (2) exactly one document boundary inside a full tile: the same trees, routed to the old or the
            // new document per 32-column chunk, plus one masked pass over the chunk that holds the boundary.
            // (3) anything else (short documents, last tile of a run, argmax): generic masked walk.
            int path = 3;
            if (!kArgmax && n_valid == kTileN) {
              if (doc_end >= tile_end) {
                path = 1;
              } else if (doc_end > row && doc + 1 < run.e) {
                if (doc_end + doc_nlen >= tile_end) path = 2;
              }
            }
            if (skip) {
              if (p.dbg_delay > 0) {  // profiling aid: hold the (unread) accumulator for a fixed number of cycles
                const long long t_start = clock64();
                while (clock64() - t_start < p.dbg_delay) {
                }
              }
              release_acc();
              while (doc_end <= tile_end) {  // advance the cursor without reading the accumulator
                ++doc;
                if (doc >= run.e) { doc_end = 0x7fffffff; break; }
                doc_row0 = doc_end;
                doc_end = doc_row0 + doc_nlen;
                doc_nlen = (doc + 1 < run.e) ? __ldg(p.doc_len + doc + 1) : 0;
              }
            } else if (path == 1) {
              // software pipeline: the loads of columns [64k+64, 64k+128) are in flight while [64k, 64k+64) fold
              // (TMEM reads are ~64 B/clk per lane quadrant: ~490 cycles for 128 x 256 fp32 whatever the warp count)
              uint32_t va[32], vb[32], vc[32], vd[32];
              tmem_ld_x32(taddr, va);
              tmem_ld_x32(taddr + 32, vb);
              tmem_ld_wait();
              reg_fence32(va);
              reg_fence32(vb);
              tmem_ld_x32(taddr + 64, vc);
              tmem_ld_x32(taddr + 96, vd);
              mm = max32(va, mm);
              mm = max32(vb, mm);
              tmem_ld_wait();
              reg_fence32(vc);
              reg_fence32(vd);
              tmem_ld_x32(taddr + 128, va);
              tmem_ld_x32(taddr + 160, vb);
              mm = max32(vc, mm);
              mm = max32(vd, mm);
              tmem_ld_wait();
              reg_fence32(va);
              reg_fence32(vb);
              tmem_ld_x32(taddr + 192, vc);
              tmem_ld_x32(taddr + 224, vd);
              mm = max32(va, mm);
              mm = max32(vb, mm);
              tmem_ld_wait();
              reg_fence32(vc);
              reg_fence32(vd);
              release_acc();  // every accumulator read has landed in registers
              mm = max32(vc, mm);
              mm = max32(vd, mm);
              while (doc_end <= tile_end) finish_doc();  // document (and empty followers) ending at the tile end
            } else if (path == 2) {
              // one boundary at column b: per 32-column chunk the FMNMX3 tree goes to the old document (chunk < kb)
              // or the new one (chunk > kb) through selects; the boundary chunk itself is re-read at the end and split
              // element-wise after the release (branching per chunk would be if-converted into doing everything).
              const int b = doc_end - row;
              const int kb = b >> 5;
              float mb = doc_ninit;
              auto route = [&](const uint32_t (&v)[32], int c) {
                const float t = tree32(v);
                mm = (c < kb) ? fmaxf(mm, t) : mm;
                mb = (c > kb) ? fmaxf(mb, t) : mb;
              };
              uint32_t va[32], vb[32], vc[32], vd[32];
              tmem_ld_x32(taddr, va);
              tmem_ld_x32(taddr + 32, vb);
              tmem_ld_wait();
              reg_fence32(va);
              reg_fence32(vb);
              tmem_ld_x32(taddr + 64, vc);
              tmem_ld_x32(taddr + 96, vd);
              route(va, 0);
              route(vb, 1);
              tmem_ld_wait();
              reg_fence32(vc);
              reg_fence32(vd);
              tmem_ld_x32(taddr + 128, va);
              tmem_ld_x32(taddr + 160, vb);
              route(vc, 2);
              route(vd, 3);
              tmem_ld_wait();
              reg_fence32(va);
              reg_fence32(vb);
              tmem_ld_x32(taddr + 192, vc);
              tmem_ld_x32(taddr + 224, vd);
              route(va, 4);
              route(vb, 5);
              tmem_ld_wait();
              reg_fence32(vc);
              reg_fence32(vd);
              tmem_ld_x32(taddr + kb * 32, va);  // the boundary chunk again
              route(vc, 6);
              route(vd, 7);
              tmem_ld_wait();
              reg_fence32(va);
              release_acc();
              {
                const int bl = b & 31;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float x = __uint_as_float(va[i]);
                  if (i < bl) mm = fmaxf(mm, x); else mb = fmaxf(mb, x);
                }
              }
              finish_doc();  // old document (running max mm); the cursor moves to the new one, whose max is mb
              mm = mb;
              while (doc_end <= tile_end) finish_doc();
            } else {
#pragma unroll 1
              for (int cb = 0; cb < n_valid; cb += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + cb, v);
                tmem_ld_wait();
                reg_fence32(v);
                const int abs0 = row + cb;
                const int abs1 = min(abs0 + 32, tile_end);
                int pos = abs0;
                while (true) {
                  const int seg_end = min(doc_end, abs1);
                  if (seg_end > pos) {
                    if constexpr (kArgmax) {
                      argmax32_range(v, mm, ai, abs0 - doc_row0, pos - abs0, seg_end - abs0);
                    } else {
                      mm = max32_range(v, mm, pos - abs0, seg_end - abs0);
                    }
                  }
                  pos = seg_end;
                  if (doc_end > abs1) break;  // the current document continues past this chunk
                  finish_doc();
                }
              }
              release_acc();
            }
            if (dbg) {
              const long long t3 = clock64();
              e_wait += t1 - t0;
              e_post += t3 - t2;
              if (path == 2) { e_hold2 += t2 - t1; ++n_path2; } else { e_hold += t2 - t1; }
            }
            m[r] = mm;
            am[r] = ai;
            nxt = doc;
            nxt_row0 = doc_row0;
            nxt_end = doc_end;
            nxt_nlen = doc_nlen;
            nxt_ninit = doc_ninit;
            ++job;
          }
        }
        cur = nxt;
        cur_row0 = nxt_row0;
        cur_end = nxt_end;
        cur_nlen = nxt_nlen;
        cur_ninit = nxt_ninit;
      }
      if (p.balanced && !skip && cur < p.n_docs && cur_end != 0x7fffffff && cur_row0 < bal_r1 && cur_end > bal_r1) {
        // my last document continues in the next partition: combine with the neighbour's partial and emit it
        for (int r = 0; r < r_cnt; ++r) {
          float mm = m[r];
          int ai = am[r];
          if (head_frag && cur == first_doc) {
            // (a document longer than a whole partition is excluded by the host: it would need a chain)
            __trap();
          }
          consume(r, mm, ai);
          finalize(cur, r, mm, ai);
        }
      }
      d = run.e;
    }
    if (dbg && warp == 2 && lane == 0) {  // epilogue: blocked on MMA / holding the accumulator / after release
      float* o = p.scores + 512 + 8 * blockIdx.x;
      o[2] = static_cast<float>(e_wait);
      o[3] = static_cast<float>(e_hold);
      o[4] = static_cast<float>(e_post);
      o[5] = static_cast<float>(e_hold2);
      o[6] = static_cast<float>(n_path2);
      o[7] = static_cast<float>(job);
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  if (p.peer_scores != nullptr && p.done_counter != nullptr && warp >= 2)
    __threadfence_system();  // my peer stores are ordered before the completion signal below
  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its shared memory or signal its barriers
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if ((p.flags & CPB_DBG_CLOCKS) && threadIdx.x == 0) {
    p.scores[2 * blockIdx.x] = static_cast<float>(clock64() - dbg_c0);
    p.scores[2 * blockIdx.x + 1] = static_cast<float>(global_timer_ns() - dbg_t0);
  }
  if (p.peer_scores != nullptr && p.done_counter != nullptr && threadIdx.x == 0) {
    // fused all-gather completion: the last CTA of the grid tells every peer that this rank's slab is complete
    __threadfence();
    const unsigned prev = atomicAdd(p.done_counter, 1u);
    if (prev + 1u == gridDim.x) {
      *p.done_counter = 0u;  // ready for the next launch (stream ordered)
      __threadfence_system();
      for (int pr = 0; pr < p.n_peers; ++pr) {
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(__ldg(p.peer_scores + pr)) + p.peer_flag_offset + p.my_rank;
        *f = p.signal_value;
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


template <int KP, bool kArgmax>
static cudaError_t launch_variant(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt,
                                  const MaxSimParams& p, int grid, cudaStream_t stream) {
  auto kern = maxsim_kpipe_kernel<KP, kArgmax>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<KP>::kAlloc);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SmemLayout<KP>::kAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tq, td, tt, p);
}

template <int KP>
static int max_clusters_variant(int cluster) {
  auto kern = maxsim_kpipe_kernel<KP, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<KP>::kAlloc) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(cluster));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SmemLayout<KP>::kAlloc;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
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

// dim_panels = padded embedding dim / 64 (3, 4 or 5; dims <= 128 use maxsim_sm100.cu)
cudaError_t maxsim_kpipe_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                                int dim_panels, bool argmax, int grid, cudaStream_t stream) {
  using namespace kpipe;
  switch (dim_panels) {
    case 3: return argmax ? launch_variant<3, true>(tq, td, tt, p, grid, stream) : launch_variant<3, false>(tq, td, tt, p, grid, stream);
    case 4: return argmax ? launch_variant<4, true>(tq, td, tt, p, grid, stream) : launch_variant<4, false>(tq, td, tt, p, grid, stream);
    case 5: return argmax ? launch_variant<5, true>(tq, td, tt, p, grid, stream) : launch_variant<5, false>(tq, td, tt, p, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cpb
