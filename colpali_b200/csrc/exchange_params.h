// Parameter block of the training-exchange kernels (exchange_sm100.cu), filled by cabi.cu.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cpb {

struct ExchangePushParams {
  const __nv_bfloat16* src;    // [n_docs, len, dim] this rank's document embeddings
  int n_docs, len, slot_len, dim;
  int pad_first;               // zero rows in front of the data (HF trainer) or behind it (torch-loop trainer)
  const uint64_t* peer_bases;  // device array of n_peers symmetric-buffer base addresses
  uint64_t mc_base;            // multicast address of the same buffer, or 0
  int n_peers;
  int64_t bank_word_offset;    // 4-byte words from the buffer base to this rank's [n_docs, slot_len, dim] block
  int64_t flag_word_offset;    // this rank's push counter (one word per rank on every rank)
};

cudaError_t exchange_push_launch(const ExchangePushParams& p, int* grid_out, cudaStream_t stream);
cudaError_t signal_peers_launch(const uint64_t* peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word,
                                cudaStream_t stream);

}  // namespace cpb
