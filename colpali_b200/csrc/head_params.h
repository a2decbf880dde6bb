// Parameter block of the fused projection-head kernel (head_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cpb {

struct HeadParams {
  const __nv_bfloat16* bias;       // [dim] or nullptr
  const int64_t* attention_mask;   // [n_tokens] or nullptr
  const uint8_t* extra_mask;       // [n_tokens] or nullptr (image-token mask)
  __nv_bfloat16* out;              // [n_tokens, dim]
  int64_t n_tokens;
  int hidden;
  uint32_t flags;
  // head_wide_sm100.cu only: output dim in (128, 320], ring depth, CTAs per cluster (1 or 2)
  int dim, stages, cluster;
};

// th / th64: the hidden states with 128-row and 64-row boxes (a CTA's share may end in half a tile)
cudaError_t head_launch(const CUtensorMap& th, const CUtensorMap& th64, const CUtensorMap& tw, const HeadParams& p, int grid,
                        cudaStream_t stream);
int head_wide_stages(int dim);  // 0: does not fit
cudaError_t head_wide_launch(const CUtensorMap& th, const CUtensorMap& tw, const HeadParams& p, int grid, cudaStream_t stream);

}  // namespace cpb
