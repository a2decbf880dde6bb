// Developer micro-benchmark (not product code): what costs ~200 cycles per 8-MMA job in the tcgen05 pipe?
// One CTA per SM issues MMA patterns on whatever is in shared memory and times them with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_probe scripts/mma_bubble_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../colpali_b200/csrc/sm100_ptx.cuh"
using namespace cpb;

constexpr int kQ = 32768, kD = 65536;

__global__ void __launch_bounds__(64, 1) probe(long long* out, int pattern, int n_mma, int chains, int per_chain, int data_mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQ + 2 * kD);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * kQ + 2 * kD) / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (data_mode == 1) {  // pseudo-random bf16 pairs in about [-0.25, 0.25] (sign + 3 exponent steps + random mantissa)
      uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const uint32_t lo = (h & 0x807fu) | (0x3c00u + ((h >> 7) & 0x0100u)), hi = ((h >> 16) & 0x807fu) | (0x3c00u + ((h >> 23) & 0x0100u));
      v = lo | (hi << 16);
    } else if (data_mode == 2) v = 0u;
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (warp == 0 && lane == 0) {
    const uint32_t qa = smem_u32(smem), da = smem_u32(smem + 2 * kQ);
    const uint32_t idesc = make_idesc_bf16_f32(128, n_mma);
    uint32_t ph[8] = {0,0,0,0,0,0,0,0};
    auto mma = [&](int acc, int r, int st, int k, bool accum) {
      const int kp = (k & 7) >> 2, kk = k & 3;
      const uint64_t ad = make_sw128_kmajor_desc(qa + r * kQ + kp * 16384) + (uint64_t)(kk * 2);
      const uint64_t bd = make_sw128_kmajor_desc(da + st * kD + kp * 32768) + (uint64_t)(kk * 2);
      umma_bf16(tb + acc * 256, ad, bd, idesc, accum ? 1u : 0u);
    };
    // warm up
    for (int k = 0; k < 8; ++k) mma(0, 0, 0, k, k > 0);
    umma_commit(&bars[7]); mbar_wait(&bars[7], 0);
    const long long t0 = clock64();
    if (pattern == 0) {            // chains back to back, alternating accumulators, no commits in between
      for (int c = 0; c < chains; ++c) for (int k = 0; k < per_chain; ++k) mma(c & 1, c & 1, (c >> 1) & 1, k, k > 0);
    } else if (pattern == 1) {     // + one commit after every chain (nobody waits on it)
      for (int c = 0; c < chains; ++c) { for (int k = 0; k < per_chain; ++k) mma(c & 1, c & 1, (c >> 1) & 1, k, k > 0); umma_commit(&bars[c & 1]); }
    } else if (pattern == 2) {     // + a try_wait on an already-completed barrier and a fence before every chain
      for (int c = 0; c < chains; ++c) {
        mbar_wait(&bars[7], 0); tc_fence_after();
        for (int k = 0; k < per_chain; ++k) mma(c & 1, c & 1, (c >> 1) & 1, k, k > 0);
        umma_commit(&bars[c & 1]);
      }
    } else if (pattern == 3) {     // same accumulator for every chain (accumulate flag reset at chain start)
      for (int c = 0; c < chains; ++c) for (int k = 0; k < per_chain; ++k) mma(0, c & 1, (c >> 1) & 1, k, k > 0);
    } else if (pattern == 4) {     // two chains interleaved k-step by k-step
      for (int c = 0; c < chains; c += 2) for (int k = 0; k < per_chain; ++k) { mma(0, 0, (c >> 1) & 1, k, k > 0); mma(1, 1, (c >> 1) & 1, k, k > 0); }
    } else if (pattern == 5) {     // one long chain (all MMAs accumulate into one tile)
      for (int c = 0; c < chains; ++c) for (int k = 0; k < per_chain; ++k) mma(0, c & 1, (c >> 1) & 1, k, (c | k) > 0);
    } else if (pattern == 6) {     // alternating accumulators, same A and B operands every chain
      for (int c = 0; c < chains; ++c) for (int k = 0; k < per_chain; ++k) mma(c & 1, 0, 0, k, k > 0);
    }
    const long long t1 = clock64();
    umma_commit(&bars[6]); mbar_wait(&bars[6], 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  const int smem = 2 * kQ + 2 * kD + 1024 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  long long h[296];
  const char* names[] = {"alternate acc, no commit", "alternate acc + commit/chain", "+try_wait+fence/chain", "same acc, accumulate reset per chain",
                         "two chains interleaved", "one long chain", "alternate acc, same operands"};
  for (int data_mode : {0, 1, 2}) for (int n_mma : {256}) for (int per_chain : {8}) for (int pat : {0, 1, 2, 5}) {
    const int chains = 256;
    for (int rep = 0; rep < 2; ++rep) probe<<<148, 64, smem>>>(d, pat, n_mma, chains, per_chain, data_mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0; for (int i = 0; i < 148; ++i) { issue += h[2*i]; total += h[2*i+1]; }
    issue /= 148; total /= 148;
    const double ideal = (double)chains * per_chain * (128.0 * n_mma / 256.0);
    printf("data=%d N=%3d per_chain=%2d %-40s: issue %8.0f total %8.0f cyc  | per chain %7.1f (ideal %6.0f, +%5.1f) per MMA %6.1f\n", data_mode, n_mma, per_chain, names[pat],
           issue, total, total / chains, ideal / chains, (total - ideal) / chains, total / (chains * per_chain));
  }
  return 0;
}
