// Parameter block of the dense dot-product kernel (dense_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cpb {

struct DenseDotParams {
  const void* a;            // [m, k] elements at a[i * a_rs + kk * a_ks]
  const void* b;            // [n, k] elements at b[row(j) * b_rs + kk * b_ks], row(j) = b_rows ? b_rows[j] : j
  const int32_t* b_rows;    // or nullptr
  int64_t a_rs, a_ks, b_rs, b_ks;
  int m, n, k;
  float* out;               // fp32 out[i * out_rs + j]
  int64_t out_rs;
  const float* alpha;       // device scalar or nullptr (= 1)
  int accumulate;           // out += alpha * dot instead of out = alpha * dot
  int a_f32, b_f32;         // operand element type: float, else __nv_bfloat16
  int raster_group;         // output tiles are walked in blocks of this many column tiles x all row tiles (>= 1)
};

cudaError_t dense_dot_launch(const DenseDotParams& p, cudaStream_t stream);

}  // namespace cpb
