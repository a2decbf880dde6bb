// Kernel-side parameter block of the fused MaxSim kernel (maxsim_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>

#include "../../include/colpali_b200.h"

// internal (profiling only) flag bits, never set through the public header
#define CPB_DBG_SKIP_EPILOGUE 0x10000u
#define CPB_DBG_NO_TMA 0x20000u   /* producer signals 'full' without loading (MMA re-reads stale smem) */
#define CPB_DBG_CLOCKS 0x40000u   /* CTA b writes (SM cycles, ns) of its main loop to scores[2b], scores[2b+1] */

namespace cpb {

// layout of the fused loss's counter workspace (cpb_maxsim_args.d_done_counter): see maxsim_finish
constexpr int kLossWsGroups = 1024;  // counters in words [0, 1024): word 0 + one per query-tile group (at most 1023)
constexpr int kLossWsWords = 4096;   // + 3 floats per group from word 1024 on

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
  // fused all-gather: scores are stored straight into every peer GPU's slab over NVLink -- through the NVSwitch
  // multicast mapping of the symmetric buffer when there is one (one multimem.st reaches every rank), else through
  // one peer mapping per rank.  All offsets are in 4-byte words from the base of the symmetric buffer.
  const uint64_t* peer_scores;  // device array of n_peers base pointers (symmetric memory), or nullptr
  uint64_t mc_base;             // multicast address of the same buffer, or 0
  int n_peers;
  int64_t peer_slab_offset;     // first word of gathered[parity][my_rank] ([n_queries, n_docs] floats)
  // completion: every CTA, once its scores are out, adds 1 (release, system scope) to word peer_flag_offset of every
  // peer; a consumer waits until its copy of that word has grown by the grid size (cpb_wait_flags)
  int64_t peer_flag_offset;
  // write-after-read guard of the double-buffered slab: before its first score store the kernel waits until the LOCAL
  // flag words wait_flags[0..n_wait) have reached wait_value, i.e. every peer has finished the previous launch (and
  // therefore, by stream order on the peer, is done with the slab this launch overwrites)
  const uint32_t* wait_flags;
  int n_wait;
  uint32_t wait_value;
  uint32_t wait_timeout_ms;
  // fused loss: the last CTA to finish (this counter, reset by it) turns the score matrix into the loss + gradient
  uint32_t* done_counter;       // local device word, zero before the first launch
  // fused top-k (topk_tail.cuh): the CTAs of every query-tile group select the k best documents of its queries
  float* topk_scores;           // [n_queries, topk_k] or nullptr
  int32_t* topk_idx;            // [n_queries, topk_k] document index inside this bank
  void* topk_ws;                // kTopkWorkspaceBytes: group counters (zero before the first launch) + candidate lists
  int topk_k;                   // 1 .. kTopkMax
  // smooth-max aggregation (late_interaction_losses.py:40-44): tau * logsumexp(raw / tau) instead of the max
  float smooth_c;               // log2(e) / tau (0 = hard max)
  float smooth_out;             // tau * ln 2
  int nq_real;                  // rows of each query that count (the rest of nq_pad is layout padding)
  float* lse;                   // [n_docs, q_rows] per-(document, query row) smooth maximum, or nullptr
  int pdl;                      // programmatic dependent launch: 0 off, 1 wait before the first global access, 2 never wait
  int boundary_mode;            // epilogue path for tiles holding one document boundary (0 re-read, 1 shifted chunks)
  int early_spin;  // cycles the issuer probes those barriers before finishing the current job first
  int dbg_delay;   // profiling only: cycles the epilogue holds an unread accumulator in CPB_DBG_SKIP_EPILOGUE mode
};

}  // namespace cpb
