// In-batch-negative exchange for multi-GPU training, without a collective kernel.
//
// What the reference's trainers do around the loss when world_size > 1 (contrastive_trainer.py:143-150,
// colmodel_torch_training.py:155-184): zero-pad every rank's [B, L_r, D] document block to the global maximum length,
// all-gather it to [world * B, L_max, D], call the loss with offset = rank * B; autograd's backward of the gather is a
// reduce-scatter of dD.  Here:
//   * exchange_push_kernel -- every rank writes its padded block straight into ALL ranks' copies of the gathered bank
//     (symmetric memory): one multimem.st per 16 bytes through the NVSwitch multicast mapping, or one st.global per
//     NVLink peer mapping.  The zero rows of the padding are written by the same pass (no separate pad / cat kernels,
//     no NCCL all-gather).  Every CTA then adds 1 (release, system scope) to its rank's counter on every rank; the
//     loss kernel (maxsim_sm100.cu) waits on those counters in its prologue.
//   * the reduce-scatter disappears: the dD kernel (loss_sm100.cu) adds each document's gradient rows directly into the
//     OWNER rank's accumulator through the peer mapping (red.global.add.v2.f32), and signal_peers_kernel publishes
//     completion.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "exchange_params.h"

namespace cpb {
namespace {

__device__ __forceinline__ void signal_all(const uint64_t* peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word) {
  if (mc_base != 0) {
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_base + 4ull * static_cast<uint64_t>(flag_word)), "r"(1u) : "memory");
  } else {
    for (int pr = 0; pr < n_peers; ++pr) {
      uint32_t* f = reinterpret_cast<uint32_t*>(__ldg(peer_bases + pr)) + flag_word;
      asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
    }
  }
}

__global__ void __launch_bounds__(256) exchange_push_kernel(const ExchangePushParams p) {
  const int cpr = p.dim / 8;  // 16-byte chunks per row
  const int64_t chunks = static_cast<int64_t>(p.n_docs) * p.slot_len * cpr;
  const int lead = p.pad_first ? p.slot_len - p.len : 0;  // zero rows in front of the data
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < chunks;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cpr);
    const int64_t row = i / cpr;
    const int r = static_cast<int>(row % p.slot_len) - lead;
    const int64_t doc = row / p.slot_len;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r >= 0 && r < p.len) v = __ldg(reinterpret_cast<const uint4*>(p.src + (doc * p.len + r) * p.dim) + ch);
    const uint64_t off = 4ull * static_cast<uint64_t>(p.bank_word_offset) + 16ull * static_cast<uint64_t>(i);
    if (p.mc_base != 0) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.mc_base + off),
                   "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                   : "memory");
    } else {
      for (int pr = 0; pr < p.n_peers; ++pr) *reinterpret_cast<uint4*>(__ldg(p.peer_bases + pr) + off) = v;
    }
  }
  __syncthreads();  // every store of this CTA happened before the (cumulative) release below
  if (threadIdx.x == 0) signal_all(p.peer_bases, p.mc_base, p.n_peers, p.flag_word_offset);
}

// one thread: everything this stream did before (kernel boundary) is published to every rank
__global__ void signal_peers_kernel(const uint64_t* peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word) {
  signal_all(peer_bases, mc_base, n_peers, flag_word);
}

}  // namespace

cudaError_t exchange_push_launch(const ExchangePushParams& p, int* grid_out, cudaStream_t stream) {
  const int64_t chunks = static_cast<int64_t>(p.n_docs) * p.slot_len * (p.dim / 8);
  int64_t blocks = (chunks + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  exchange_push_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  *grid_out = static_cast<int>(blocks);
  return cudaGetLastError();
}

cudaError_t signal_peers_launch(const uint64_t* peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word,
                                cudaStream_t stream) {
  signal_peers_kernel<<<1, 1, 0, stream>>>(peer_bases, mc_base, n_peers, flag_word);
  return cudaGetLastError();
}

}  // namespace cpb
