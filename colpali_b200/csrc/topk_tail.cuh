// Top-k selection fused into the tail of the MaxSim kernel ("fuse the top-k selection into the
// scoring epilogue"; replaces torch.topk on the materialised [n_queries, n_local] slab of the sharded scorer).
// The CTAs that share a query-tile group (one per document partition, all co-resident: the grid is at most one CTA per
// SM) first count themselves in on a per-group counter and wait until the whole group has emitted its scores; then EVERY
// CTA selects the k best documents of the group's queries inside ITS OWN slice of the score rows (1 / doc_parts of a
// row: cfg4, 1389 scores per query and CTA) -- the rows are still in L2, nothing is re-read from HBM, no second kernel --
// and the last CTA of the group to finish merges the doc_parts candidate lists.  (First version: only the last CTA
// scanned, whole rows: 250 us at cfg4 against 167 us for torch.topk -- 8/9 of the group's SMs sat idle.)
// Order: larger score first, smaller document index on ties (sharded.merge_topk's order, deterministic).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace cpb {

constexpr int kTopkMax = 16;      // per-lane candidate list length = largest k selected in the kernel
constexpr int kTopkLoads = 8;     // independent L2 loads in flight per lane

// one warp, one (slice of a) query row: row[0..n) -> out_s / out_i[0..k).  Reported index = idx_map[position] if given
// (the merge pass: positions in the candidate list -> document indices), else position + idx_base.
__device__ __forceinline__ void topk_row_warp(const float* __restrict__ row, int n, int k, float* out_s, int32_t* out_i,
                                              int lane, int idx_base, const int32_t* idx_map) {
  float s[kTopkMax];
  int id[kTopkMax];
#pragma unroll
  for (int j = 0; j < kTopkMax; ++j) {
    s[j] = -INFINITY;
    id[j] = 0x7fffffff;
  }
  // each lane scans positions lane, lane + 32, ... (ascending: among equal scores the earlier position stays ahead)
  for (int base = lane; base < n; base += 32 * kTopkLoads) {
    float v[kTopkLoads];
#pragma unroll
    for (int u = 0; u < kTopkLoads; ++u) {
      const int d = base + 32 * u;
      v[u] = (d < n) ? __ldcg(row + d) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < kTopkLoads; ++u) {
      // candidates must beat the lane's current last entry; NaN and -inf never enter (the comparison is false)
      if (v[u] > s[kTopkMax - 1]) {
        // insert behind equal scores and shift the rest down by one.  The list is sorted, so (cv > s[j]) is monotone in
        // j: every slot decides on its own from two comparisons -- 16 independent selects instead of a 16-step carry
        // chain of dependent compare / select pairs (a lone warp cannot hide that latency)
        const float cv = v[u];
        const int ci = base + 32 * u;
#pragma unroll
        for (int j = kTopkMax - 1; j > 0; --j) {  // high to low: s[j - 1] is still the old value
          const bool here = cv > s[j], above = cv > s[j - 1];
          s[j] = here ? (above ? s[j - 1] : cv) : s[j];
          id[j] = here ? (above ? id[j - 1] : ci) : id[j];
        }
        const bool first = cv > s[0];
        s[0] = first ? cv : s[0];
        id[0] = first ? ci : id[0];
      }
    }
  }
  // k rounds: the best head over the 32 lanes wins, its lane pops
  for (int r = 0; r < k; ++r) {
    float bs = s[0];
    int bi = id[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) {
        bs = os;
        bi = oi;
      }
    }
    if (id[0] == bi && bi != 0x7fffffff) {  // positions are unique across lanes: exactly one lane pops
#pragma unroll
      for (int j = 0; j + 1 < kTopkMax; ++j) {
        s[j] = s[j + 1];
        id[j] = id[j + 1];
      }
      s[kTopkMax - 1] = -INFINITY;
      id[kTopkMax - 1] = 0x7fffffff;
    }
    if (lane == 0) {
      out_s[r] = bs;  // fewer than k entries: (-inf, INT32_MAX) filler
      out_i[r] = (bi == 0x7fffffff) ? bi : (idx_map ? __ldcg(idx_map + bi) : bi + idx_base);
    }
  }
}

// Workspace (cpb_maxsim_topk_workspace_bytes, zero-initialised once by the caller, reset by the kernel):
//   uint32 arrived[kTopkSlots], uint32 scanned[kTopkSlots]  -- indexed by query-tile group
//   float cand_s[kTopkSlots * 128], int32 cand_i[kTopkSlots * 128] -- [group][query of the group][partition][k];
//   groups x partitions <= CTAs of the grid <= kTopkSlots, 8 queries per group, k <= 16.
constexpr int kTopkSlots = 256;
constexpr int kTopkQueriesPerGroup = 8;
constexpr int64_t kTopkWorkspaceBytes = 2 * kTopkSlots * 4 + 2ll * kTopkSlots * kTopkQueriesPerGroup * kTopkMax * 4;

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Whole-CTA tail: every thread of every CTA with queries calls it (it contains block barriers).  `group` = this CTA's
// query-tile group, `part` its document partition, [q_first, q_first + q_count) the group's queries, `parts` = CTAs per
// group.  All CTAs of the grid are resident (<= 1 per SM), so waiting for the group cannot deadlock.
__device__ __forceinline__ void topk_group_tail(const float* scores, int n_docs, int k, float* topk_scores,
                                                int32_t* topk_idx, void* workspace, int group, int part, int parts,
                                                int q_first, int q_count, uint64_t timeout_ns) {
  uint32_t* arrived = static_cast<uint32_t*>(workspace);
  uint32_t* scanned = arrived + kTopkSlots;
  float* cand_s = reinterpret_cast<float*>(scanned + kTopkSlots);
  int32_t* cand_i = reinterpret_cast<int32_t*>(cand_s + kTopkSlots * kTopkQueriesPerGroup * kTopkMax);
  __shared__ int s_last_of_group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // ---- 1. the whole group has emitted its scores -------------------------------------------------------------------
  if (threadIdx.x == 0) {
    __threadfence();  // (the score stores of all warps were fenced before the block barrier in maxsim_finish)
    atomicAdd(arrived + group, 1u);
    uint64_t t0 = 0;
    while (ld_acquire_gpu_u32(arrived + group) < static_cast<uint32_t>(parts)) {
      if (t0 == 0) t0 = global_timer_ns();
      if (global_timer_ns() - t0 > timeout_ns) __trap();  // a CTA of the group never arrived
      __nanosleep(64);
    }
  }
  __syncthreads();

  // ---- 2. every CTA: the k best of its slice of every query row of the group ---------------------------------------
  const int d0 = static_cast<int>((static_cast<int64_t>(n_docs) * part) / parts);
  const int d1 = static_cast<int>((static_cast<int64_t>(n_docs) * (part + 1)) / parts);
  const int64_t gbase = static_cast<int64_t>(group) * kTopkQueriesPerGroup * parts * k;  // <= slots * 8 * 16 entries
  for (int qi = warp; qi < q_count; qi += nwarps) {
    const int64_t c = gbase + (static_cast<int64_t>(qi) * parts + part) * k;
    topk_row_warp(scores + static_cast<int64_t>(q_first + qi) * n_docs + d0, d1 - d0, k, cand_s + c, cand_i + c, lane, d0,
                  nullptr);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(scanned + group, 1u);
    s_last_of_group = (prev + 1u == static_cast<unsigned>(parts)) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last_of_group) return;

  // ---- 3. the last CTA of the group: merge the `parts` candidate lists of every query ------------------------------
  // (candidate positions ascend with the partition, hence with the document index: position order IS the tie order)
  __threadfence();
  for (int qi = warp; qi < q_count; qi += nwarps) {
    const int64_t c = gbase + static_cast<int64_t>(qi) * parts * k;
    const int q = q_first + qi;
    topk_row_warp(cand_s + c, parts * k, k, topk_scores + static_cast<int64_t>(q) * k, topk_idx + static_cast<int64_t>(q) * k,
                  lane, 0, cand_i + c);
  }
  if (threadIdx.x == 0) {  // ready for the next (stream-ordered) launch: every CTA of the group is past both counters
    arrived[group] = 0u;
    scanned[group] = 0u;
  }
}

}  // namespace cpb
