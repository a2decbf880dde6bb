// Kernel-side parameter block of the fused MaxSim kernel (maxsim_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>

#include "../../include/colpali_b200.h"

// internal (profiling only) flag bits, never set through the public header
#define CPB_DBG_SKIP_EPILOGUE 0x10000u
#define CPB_DBG_NO_TMA 0x20000u   /* producer signals 'full' without loading (MMA re-reads stale smem) */
#define CPB_DBG_CLOCKS 0x40000u   /* CTA b writes (SM cycles, ns) of its main loop to scores[2b], scores[2b+1] */

namespace cpb {

struct MaxSimParams {
  const void* q;             // bf16 [q_rows, 128] padded queries
  const int32_t* doc_start;  // [n_docs] first bank row of each document
  const int32_t* doc_len;    // [n_docs] rows per document
  const float* doc_floor;    // [n_docs] or nullptr (-inf)
  float* scores;             // [nseg, n_queries, n_docs] (nseg == 1: the final scores)
  int32_t* argmax;           // [n_docs, q_rows] or nullptr
  int64_t plane_stride;      // n_queries * n_docs
  int n_queries;
  int nq_pad;      // rows per query, multiple of 32
  int q_rows;      // n_queries * nq_pad
  int n_docs;
  int num_qtiles;  // ceil(q_rows / 128)
  int q_groups;    // ceil(num_qtiles / R)
  int cluster;     // CTAs per cluster (1, 2 or 4); they share a document partition via TMA multicast
  int group_sets;  // ceil(q_groups / cluster)
  int doc_parts;   // document partitions; grid = group_sets * doc_parts * cluster
  uint32_t flags;
  // tile-balanced partitioning of a contiguous bank (a document may straddle two partitions; its two partial
  // per-token maxima are combined through `split_*`): see maxsim_sm100.cu
  int balanced;        // 0 = partitions are whole documents, 1 = partitions are whole 256-row tiles
  int bank_rows;       // rows of the bank covered by documents (balanced mode)
  int uniform_len;     // > 0: every document has this many rows (first document of a partition = row / len)
  float* split_max;    // [q_groups_padded, doc_parts, R, 128]
  int32_t* split_idx;  // same shape (argmax variant)
  uint32_t* split_flag;  // [q_groups_padded, doc_parts, R, 4]; a slot is valid when it holds `epoch`
  uint32_t epoch;
  // fused all-gather: scores are stored straight into every peer GPU's slab [n_peers][n_queries][n_docs] over NVLink
  const uint64_t* peer_scores;  // device array of n_peers base pointers (symmetric memory), or nullptr
  int n_peers;
  int64_t peer_slab_offset;     // my_rank * n_queries * n_docs (floats)
  // completion signal of the fused all-gather: the last CTA to finish stores `signal_value` into word
  // peer_flag_offset + my_rank of every peer's buffer (consumers wait on their own copy)
  uint32_t* done_counter;       // local device word, zero before the first launch (the last CTA resets it)
  int64_t peer_flag_offset;     // in 4-byte words from the slab base
  uint32_t signal_value;
  int my_rank;
  int dbg_delay;   // profiling only: cycles the epilogue holds an unread accumulator in CPB_DBG_SKIP_EPILOGUE mode
  int mma_split;   // K-steps of a job issued before the next job's barrier waits (5..8)
};

}  // namespace cpb
