// Developer micro-benchmark (not product code): does the tmem_full / tmem_empty handshake with epilogue warps create
// the ~170-cycle per-job bubble?  192 threads: warp 0 idle, warp 1 = issuer (uniform), warps 2-5 = epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_hs.bin scripts/mma_handshake_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../colpali_b200/csrc/sm100_ptx.cuh"
using namespace cpb;
constexpr int kQ = 32768, kD = 65536;

// mode bits: 1 = epilogue reads the accumulator (8 x tcgen05.ld x32), 2 = issuer skips the tmem_empty wait (no handshake),
//            4 = NACC accumulators of 128 columns (N = 128) instead of 2 x 256
__global__ void __launch_bounds__(192, 1) probe(long long* out, int mode, int chains, int nacc, int n_mma) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQ + 2 * kD);
  uint64_t* tfull = bars;        // [4]
  uint64_t* tempty = bars + 4;   // [4]
  uint64_t* done = bars + 8;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * kQ + 2 * kD) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); } mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  const int cols = 512 / nacc;
  if (warp == 1) {
    const uint32_t qa = smem_u32(smem), da = smem_u32(smem + 2 * kQ);
    const uint32_t idesc = make_idesc_bf16_f32(128, n_mma);
    const long long t0 = clock64();
    for (int c = 0; c < chains; ++c) {
      const int a = c % nacc;
      const uint32_t ph = (c / nacc) & 1;
      if (!(mode & 2)) { mbar_wait(&tempty[a], ph ^ 1u); tc_fence_after(); }
      const uint64_t ad0 = make_sw128_kmajor_desc(qa + (c & 1) * kQ), bd0 = make_sw128_kmajor_desc(da + ((c >> 1) & 1) * kD);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tb + a * cols, ad0 + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), bd0 + (uint64_t)((k >> 2) * 2048 + (k & 3) * 2), idesc, k > 0);
        umma_commit(&tfull[a]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  } else if (warp >= 2) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float m = 0.f;
    for (int c = 0; c < chains; ++c) {
      const int a = c % nacc;
      const uint32_t ph = (c / nacc) & 1;
      mbar_wait(&tfull[a], ph);
      tc_fence_after();
      if (mode & 1) {
        for (int cc = 0; cc < n_mma; cc += 64) {
          uint32_t v0[32], v1[32];
          tmem_ld_x32(tb + lane_base + a * cols + cc, v0);
          tmem_ld_x32(tb + lane_base + a * cols + cc + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
    if (m == 123.f) out[200] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  const int smem = 2 * kQ + 2 * kD + 1024 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, 256 * sizeof(long long));
  long long h[148];
  struct Cfg { int mode, nacc, n; const char* name; } cfgs[] = {
    {2, 2, 256, "no handshake (issuer never waits), 2 x 256"},
    {0, 2, 256, "handshake, epilogue releases immediately, 2 x 256"},
    {1, 2, 256, "handshake, epilogue reads accumulator, 2 x 256"},
    {0, 4, 128, "handshake, release immediately, 4 x 128 (N=128)"},
    {1, 4, 128, "handshake, epilogue reads, 4 x 128 (N=128)"},
    {2, 4, 128, "no handshake, 4 x 128 (N=128)"},
  };
  for (auto& c : cfgs) {
    const int chains = 256;
    for (int rep = 0; rep < 2; ++rep) probe<<<148, 192, smem>>>(d, c.mode, chains, c.nacc, c.n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double t = 0; for (int i = 0; i < 148; ++i) t += h[i]; t /= 148;
    printf("%-52s: %8.0f cycles, per chain %7.1f (ideal %5.0f)\n", c.name, t, t / chains, 8 * 128.0 * c.n / 256);
  }
  return 0;
}
