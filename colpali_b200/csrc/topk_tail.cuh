// Top-k selection fused into the tail of the MaxSim kernel ("fuse the top-k selection into the
// scoring epilogue"; replaces torch.topk on the materialised [n_queries, n_local] slab of the sharded scorer).
// The CTAs that share a query-tile group (one per document partition) count themselves out on a per-group counter; the
// LAST of them selects the k best documents of the group's queries from the score rows the others have just written --
// the rows are still in L2 (cfg4: 12 500 scores = 50 KB per query), nothing is re-read from HBM, no second kernel.
// Order: larger score first, smaller document index on ties (sharded.merge_topk's order, deterministic).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace cpb {

constexpr int kTopkMax = 16;      // per-lane candidate list length = largest k selected in the kernel
constexpr int kTopkLoads = 8;     // independent L2 loads in flight per lane

// one warp, one query row: scores[0..n) -> out_s / out_i[0..k)
__device__ __forceinline__ void topk_row_warp(const float* __restrict__ row, int n, int k, float* out_s, int32_t* out_i,
                                              int lane) {
  float s[kTopkMax];
  int id[kTopkMax];
#pragma unroll
  for (int j = 0; j < kTopkMax; ++j) {
    s[j] = -INFINITY;
    id[j] = 0x7fffffff;
  }
  // each lane scans documents lane, lane + 32, ... (ascending: among equal scores the earlier document stays ahead)
  for (int base = lane; base < n; base += 32 * kTopkLoads) {
    float v[kTopkLoads];
#pragma unroll
    for (int u = 0; u < kTopkLoads; ++u) {
      const int d = base + 32 * u;
      v[u] = (d < n) ? __ldcg(row + d) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < kTopkLoads; ++u) {
      // candidates must beat the lane's current k-th best; NaN never enters (every comparison with it is false)
      if (v[u] > s[kTopkMax - 1]) {
        float cv = v[u];
        int ci = base + 32 * u;
        bool shifting = false;
#pragma unroll
        for (int j = 0; j < kTopkMax; ++j) {  // insert behind equal scores, then shift the rest down by one
          shifting = shifting || (cv > s[j]);
          const float ts = s[j];
          const int ti = id[j];
          s[j] = shifting ? cv : ts;
          id[j] = shifting ? ci : ti;
          cv = shifting ? ts : cv;
          ci = shifting ? ti : ci;
        }
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
    if (id[0] == bi && bi != 0x7fffffff) {  // document indices are unique across lanes: exactly one lane pops
#pragma unroll
      for (int j = 0; j + 1 < kTopkMax; ++j) {
        s[j] = s[j + 1];
        id[j] = id[j + 1];
      }
      s[kTopkMax - 1] = -INFINITY;
      id[kTopkMax - 1] = 0x7fffffff;
    }
    if (lane == 0) {
      out_s[r] = bs;                          // fewer than k documents: (-inf, INT32_MAX) filler
      out_i[r] = bi;
    }
  }
}

// Whole-CTA tail: every thread calls it (it contains block barriers).  `group` = this CTA's query-tile group,
// [q_first, q_first + q_count) its queries, `expected` = CTAs per group (document partitions).
__device__ __forceinline__ void topk_group_tail(const float* scores, int n_docs, int k, float* topk_scores,
                                                int32_t* topk_idx, uint32_t* counters, int group, int expected,
                                                int q_first, int q_count) {
  __shared__ int s_last_of_group;
  if (threadIdx.x == 0) {
    __threadfence();  // (the score stores of all warps were fenced before the block barrier in maxsim_finish)
    const unsigned prev = atomicAdd(counters + group, 1u);
    s_last_of_group = (prev + 1u == static_cast<unsigned>(expected)) ? 1 : 0;
    if (s_last_of_group) counters[group] = 0u;  // ready for the next (stream-ordered) launch
  }
  __syncthreads();
  if (!s_last_of_group) return;
  __threadfence();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int qi = warp; qi < q_count; qi += nwarps) {
    const int q = q_first + qi;
    topk_row_warp(scores + static_cast<int64_t>(q) * n_docs, n_docs, k, topk_scores + static_cast<int64_t>(q) * k,
                  topk_idx + static_cast<int64_t>(q) * k, lane);
  }
}

}  // namespace cpb
