// Kernel-side parameter block of the fused MaxSim kernel (maxsim_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>

#include "../../include/colpali_b200.h"

namespace cpb {

struct MaxSimParams {
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
  int doc_parts;   // CTAs per query group; grid = q_groups * doc_parts
  uint32_t flags;
};

}  // namespace cpb
