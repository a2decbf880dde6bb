// extern "C" surface of libcolpali_b200.so (declared in include/colpali_b200.h).
// Host-side only: argument validation, TMA tensor-map encoding, grid sizing, launches.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
#include "head_params.h"
#include "loss_params.h"
#include "maxsim_params.h"

namespace cpb {
cudaError_t maxsim_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                          int r, bool argmax, int grid, cudaStream_t stream);
int maxsim_max_clusters(int r, int cluster);
int maxsim_tile_n();
cudaError_t maxsim_kpipe_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                                int dim_panels, bool argmax, int grid, cudaStream_t stream);
int maxsim_kpipe_max_clusters(int dim_panels, int cluster);
cudaError_t wait_flags_launch(const uint32_t* flags, int n, uint32_t value, cudaStream_t stream);
cudaError_t maxsim_reduce_segments(const float* partial, float* out, int64_t plane, int nseg, int round_ref,
                                   cudaStream_t stream);
}  // namespace cpb

namespace {

thread_local char g_err[512] = "";

// tuning knobs (cpb_set_option); 0 = choose automatically
int g_opt_cluster = 0;
static int g_head_cluster = 0;  // DRAFT: 0 = auto (2 when there are at least two token tiles), 1, 2
int g_opt_qtiles_per_cta = 0;
unsigned g_opt_debug_flags = 0;
int g_opt_mma_split = 6;
int g_opt_dbg_delay = 0;
int g_opt_balanced = 1;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CPB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return fail(CPB_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the driver entry point is fetched through the runtime, so the library
// loads (and its symbols can be checked) on a machine without a GPU driver.
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 [rows, 128] row-major -> tiles of [box_rows, 64] written with the 128-byte swizzle.
// Rows past `rows` are zero-filled by the TMA unit.
int make_bf16_rowmajor_map(CUtensorMap* map, const void* base, int64_t rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(CPB_E_CUDA, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(base) & 15u) return fail(CPB_E_INVALID, "tensor base %p is not 16-byte aligned", base);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CPB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return CPB_OK;
}

struct DevInfo {
  int sm_count = 0, major = 0, minor = 0;
};

int current_device_info(DevInfo* out) {
  int dev = 0;
  CPB_CUDA(cudaGetDevice(&dev));
  static DevInfo cache[64];
  if (dev >= 0 && dev < 64 && cache[dev].sm_count > 0) {
    *out = cache[dev];
    return CPB_OK;
  }
  CPB_CUDA(cudaDeviceGetAttribute(&out->sm_count, cudaDevAttrMultiProcessorCount, dev));
  CPB_CUDA(cudaDeviceGetAttribute(&out->major, cudaDevAttrComputeCapabilityMajor, dev));
  CPB_CUDA(cudaDeviceGetAttribute(&out->minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (dev >= 0 && dev < 64) cache[dev] = *out;
  return CPB_OK;
}

}  // namespace

extern "C" {

int cpb_abi_version(void) { return CPB_ABI_VERSION; }

const char* cpb_last_error(void) { return g_err; }

int cpb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int v = 0;
  if (sm_count) {
    CPB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (cc_major) {
    CPB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    CPB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  return CPB_OK;
}

int cpb_set_option(const char* name, int value) {
  if (!name) return fail(CPB_E_INVALID, "null option name");
  if (!strcmp(name, "cluster")) {
    if (value != 0 && value != 1 && value != 2 && value != 4) return fail(CPB_E_INVALID, "cluster must be 0, 1, 2 or 4");
    g_opt_cluster = value;
  } else if (!strcmp(name, "qtiles_per_cta")) {
    if (value < 0 || value > 2) return fail(CPB_E_INVALID, "qtiles_per_cta must be 0, 1 or 2");
    g_opt_qtiles_per_cta = value;
  } else if (!strcmp(name, "mma_split")) {
    if (value < 5 || value > 8) return fail(CPB_E_INVALID, "mma_split must be 5..8");
    g_opt_mma_split = value;
  } else if (!strcmp(name, "head_cluster")) {
    if (value < 0 || value > 2) return fail(CPB_E_INVALID, "head_cluster must be 0, 1 or 2");
    g_head_cluster = value;
  } else if (!strcmp(name, "balanced")) {
    g_opt_balanced = value != 0;
  } else if (!strcmp(name, "debug_delay")) {
    g_opt_dbg_delay = value;
  } else if (!strcmp(name, "debug_flags")) {
    g_opt_debug_flags = static_cast<unsigned>(value) & 0xffff0000u;
  } else {
    return fail(CPB_E_INVALID, "unknown option '%s'", name);
  }
  return CPB_OK;
}

int64_t cpb_maxsim_workspace_bytes(int n_queries, int nq_pad, int n_docs) {
  if (nq_pad <= 32) return 0;
  return static_cast<int64_t>(nq_pad / 32) * n_queries * n_docs * 4;
}

static int maxsim_fwd_impl(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                           const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                           float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags, int uniform_len,
                           int max_doc_len, void* d_split_ws, int64_t split_ws_bytes, uint32_t epoch,
                           const uint64_t* d_peer_ptrs, int n_peers, int my_rank, uint32_t* d_done_counter,
                           int64_t flag_word_offset, uint32_t signal_value, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
  if (reinterpret_cast<uintptr_t>(d_q) & 15u) return fail(CPB_E_INVALID, "d_q is not 16-byte aligned");
  if (!d_q || !d_docs || !d_doc_start || !d_doc_len || (!d_scores && !d_peer_ptrs)) return fail(CPB_E_INVALID, "null device pointer");
  if (doc_rows <= 0 || doc_rows > 0x7fffffffLL) return fail(CPB_E_INVALID, "doc_rows=%lld out of range (1..2^31-1)", static_cast<long long>(doc_rows));
  const int nseg = nq_pad / 32;
  if (nseg > 1 && !d_workspace) return fail(CPB_E_INVALID, "nq_pad=%d needs a workspace (cpb_maxsim_workspace_bytes)", nq_pad);
  const int64_t q_rows64 = static_cast<int64_t>(n_queries) * nq_pad;
  if (q_rows64 > 0x7fffffffLL) return fail(CPB_E_INVALID, "too many query rows");

  DevInfo di;
  int rc = current_device_info(&di);
  if (rc != CPB_OK) return rc;
  if (di.major != 10) return fail(CPB_E_DEVICE, "device is sm_%d%d; this library needs sm_100 (B200)", di.major, di.minor);

  cpb::MaxSimParams p{};
  p.q = d_q;
  p.doc_start = d_doc_start;
  p.doc_len = d_doc_len;
  p.doc_floor = d_doc_floor;
  p.argmax = d_argmax;
  p.plane_stride = static_cast<int64_t>(n_queries) * n_docs;
  p.n_queries = n_queries;
  p.nq_pad = nq_pad;
  p.q_rows = static_cast<int>(q_rows64);
  p.n_docs = n_docs;
  p.num_qtiles = (p.q_rows + 127) / 128;
  p.flags = flags;
  p.scores = (nseg == 1) ? d_scores : d_workspace;
  if (d_peer_ptrs) {
    if (nseg != 1) return fail(CPB_E_UNSUPPORTED, "the fused all-gather needs queries of at most 32 tokens (nq_pad == 32)");
    if (n_peers < 1 || n_peers > 64 || my_rank < 0 || my_rank >= n_peers) return fail(CPB_E_INVALID, "bad peer arguments (%d peers, rank %d)", n_peers, my_rank);
    p.peer_scores = d_peer_ptrs;
    p.n_peers = n_peers;
    p.peer_slab_offset = static_cast<int64_t>(my_rank) * n_queries * n_docs;
    p.done_counter = d_done_counter;
    p.peer_flag_offset = flag_word_offset;
    p.signal_value = signal_value;
    p.my_rank = my_rank;
  }

  // Two resident query tiles per CTA halve the L2->SMEM traffic per flop; a single tile only
  // when there is just one.
  int R = (p.num_qtiles >= 2) ? 2 : 1;
  if (g_opt_qtiles_per_cta == 1 || g_opt_qtiles_per_cta == 2) R = g_opt_qtiles_per_cta;
  p.q_groups = (p.num_qtiles + R - 1) / R;
  // CTAs of a cluster hold different query-tile groups and share every document tile through TMA
  // multicast.  Pairs tile the 148 SMs exactly; clusters of 4 strand SMs in GPCs whose SM count is
  // not a multiple of 4, so they are opt-in.
  int cluster = (p.q_groups >= 2) ? 2 : 1;
  if (g_opt_cluster == 1 || g_opt_cluster == 2 || g_opt_cluster == 4) cluster = g_opt_cluster;
  // co-resident cluster count is a property of (device, R, cluster): query the driver once
  static int occ_cache[3][5] = {};
  auto max_clusters_cached = [&](int r, int c) {
    if (occ_cache[r][c] == 0) occ_cache[r][c] = cpb::maxsim_max_clusters(r, c);
    return occ_cache[r][c];
  };
  int max_clusters = max_clusters_cached(R, cluster);
  if (max_clusters <= 0) {
    if (cluster == 1) return fail(CPB_E_CUDA, "kernel cannot be resident on this device (shared memory?)");
    cluster = 1;
    max_clusters = max_clusters_cached(R, 1);
    if (max_clusters <= 0) return fail(CPB_E_CUDA, "kernel cannot be resident on this device (shared memory?)");
  }
  p.cluster = cluster;
  p.group_sets = (p.q_groups + cluster - 1) / cluster;
  int parts = max_clusters / p.group_sets;
  if (parts < 1) parts = 1;
  if (parts > n_docs) parts = n_docs;
  // tile-balanced partitions (a document may be cut between two partitions): contiguous banks only, every
  // partition at least as long as the longest document, caller-provided exchange workspace
  const int64_t tiles = (doc_rows + 255) / 256;
  if (g_opt_balanced && (flags & CPB_FLAG_CONTIGUOUS) && d_split_ws && max_doc_len > 0 && doc_rows <= 0x7fffffffLL) {
    int bparts = parts;
    if (bparts > tiles) bparts = static_cast<int>(tiles);
    const int64_t min_rows = 256 * (tiles / (bparts > 0 ? bparts : 1));
    const int64_t need = static_cast<int64_t>(p.group_sets) * cluster * bparts * R * (128 * 8 + 16);
    if (bparts > 1 && min_rows >= max_doc_len && need <= split_ws_bytes) {
      parts = bparts;
      p.balanced = 1;
      p.bank_rows = static_cast<int>(doc_rows);
      p.uniform_len = uniform_len;
      const int64_t slots = static_cast<int64_t>(p.group_sets) * cluster * bparts * R;
      p.split_max = static_cast<float*>(d_split_ws);
      p.split_idx = reinterpret_cast<int32_t*>(p.split_max + slots * 128);
      p.split_flag = reinterpret_cast<uint32_t*>(p.split_idx + slots * 128);
      p.epoch = epoch;
    }
  }
  p.doc_parts = parts;
  p.flags = flags | g_opt_debug_flags;
  p.mma_split = g_opt_mma_split;
  p.dbg_delay = g_opt_dbg_delay;
  const int grid = p.group_sets * p.doc_parts * cluster;

  CUtensorMap tq, td, tt;
  rc = make_bf16_rowmajor_map(&tq, d_q, q_rows64, 128, 128);
  if (rc != CPB_OK) return rc;
  rc = make_bf16_rowmajor_map(&td, d_docs, doc_rows, 128, cpb::maxsim_tile_n() / cluster);
  if (rc != CPB_OK) return rc;
  rc = make_bf16_rowmajor_map(&tt, d_docs, doc_rows, 128, 32);
  if (rc != CPB_OK) return rc;

  CPB_CUDA(cpb::maxsim_launch(tq, td, tt, p, R, d_argmax != nullptr, grid, stream));
  if (nseg > 1)
    CPB_CUDA(cpb::maxsim_reduce_segments(d_workspace, d_scores, p.plane_stride, nseg,
                                         (flags & CPB_FLAG_ROUND_BF16) ? 1 : 0, stream));
  return CPB_OK;
}

static int loss_fwd_impl(const float* d_scores, const void* d_q, int n_queries, int nq_pad, int n_docs, int mode,
                         float temperature, int normalize_scores, int pos_aware_negative_filtering,
                         float filter_threshold, float filter_factor, int offset, const float* d_neg_scores, int n_neg,
                         float in_batch_weight, float* d_loss, float* d_grad_scores, float* d_grad_neg,
                         float* d_bounds, int q_dim, void* stream_) {
  if (q_dim != 128 && q_dim != 192 && q_dim != 256 && q_dim != 320)
    return fail(CPB_E_UNSUPPORTED, "embedding dim %d is not supported by this build (128, 192, 256, 320)", q_dim);
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
  if (mode != CPB_LOSS_CE && mode != CPB_LOSS_PAIRWISE && mode != CPB_LOSS_SIGMOID) return fail(CPB_E_INVALID, "unknown loss mode %d", mode);
  if (mode == CPB_LOSS_SIGMOID && (n_docs != n_queries || offset != 0))
    return fail(CPB_E_INVALID, "the sigmoid loss needs n_docs == n_queries and offset == 0 (got %d, %d, %d)", n_docs, n_queries, offset);
  if (d_neg_scores && (n_neg <= 0 || mode == CPB_LOSS_SIGMOID || in_batch_weight < 0.f || in_batch_weight > 1.f))
    return fail(CPB_E_INVALID, "bad explicit-negative arguments (n_neg=%d, mode=%d, weight=%g)", n_neg, mode, static_cast<double>(in_batch_weight));
  if (offset < 0 || offset + n_queries > n_docs)
    return fail(CPB_E_INVALID, "positive index out of range: offset=%d + n_queries=%d > n_docs=%d", offset, n_queries, n_docs);
  if (!(temperature > 0.f)) return fail(CPB_E_INVALID, "temperature must be positive");
  if (!d_scores || !d_q || !d_loss) return fail(CPB_E_INVALID, "null device pointer");
  cpb::LossParams p{};
  p.scores = d_scores;
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.q_dim = q_dim;
  p.loss = d_loss;
  p.grad = d_grad_scores;
  p.bounds = d_bounds;
  p.B = n_queries;
  p.C = n_docs;
  p.nq_pad = nq_pad;
  p.offset = offset;
  p.mode = mode;
  p.normalize = normalize_scores;
  p.filter = pos_aware_negative_filtering;
  p.temperature = temperature;
  p.filter_threshold = filter_threshold;
  p.filter_factor = filter_factor;
  p.neg_scores = d_neg_scores;
  p.grad_neg = d_grad_neg;
  p.n_neg = n_neg;
  p.in_batch_weight = in_batch_weight;
  CPB_CUDA(cpb::colbert_loss_launch(p, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int cpb_colbert_loss_fwd(const float* d_scores, const void* d_q, int n_queries, int nq_pad, int n_docs, int mode,
                         float temperature, int normalize_scores, int pos_aware_negative_filtering,
                         float filter_threshold, float filter_factor, int offset, float* d_loss, float* d_grad_scores,
                         float* d_bounds, void* stream_) {
  return loss_fwd_impl(d_scores, d_q, n_queries, nq_pad, n_docs, mode, temperature, normalize_scores,
                       pos_aware_negative_filtering, filter_threshold, filter_factor, offset, nullptr, 0, 1.f, d_loss,
                       d_grad_scores, nullptr, d_bounds, 128, stream_);
}

int cpb_colbert_loss_fwd_dim(const float* d_scores, const void* d_q, int n_queries, int nq_pad, int n_docs, int mode,
                             float temperature, int normalize_scores, int pos_aware_negative_filtering,
                             float filter_threshold, float filter_factor, int offset, float* d_loss,
                             float* d_grad_scores, float* d_bounds, int dim, void* stream_) {
  return loss_fwd_impl(d_scores, d_q, n_queries, nq_pad, n_docs, mode, temperature, normalize_scores,
                       pos_aware_negative_filtering, filter_threshold, filter_factor, offset, nullptr, 0, 1.f, d_loss,
                       d_grad_scores, nullptr, d_bounds, dim, stream_);
}

int cpb_maxsim_fwd(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                   const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                   float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags, void* stream_) {
  return maxsim_fwd_impl(d_q, n_queries, nq_pad, d_docs, doc_rows, d_doc_start, d_doc_len, d_doc_floor, n_docs, d_scores,
                         d_argmax, d_workspace, flags, 0, 0, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0, 0, stream_);
}

int cpb_maxsim_fwd_balanced(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                            const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                            float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags, int uniform_len,
                            int max_doc_len, void* d_split_ws, int64_t split_ws_bytes, uint32_t epoch, void* stream_) {
  if (epoch == 0) return fail(CPB_E_INVALID, "epoch must be non-zero (a zero-initialised workspace means 'nothing published')");
  return maxsim_fwd_impl(d_q, n_queries, nq_pad, d_docs, doc_rows, d_doc_start, d_doc_len, d_doc_floor, n_docs, d_scores,
                         d_argmax, d_workspace, flags, uniform_len, max_doc_len, d_split_ws, split_ws_bytes, epoch, nullptr,
                         0, 0, nullptr, 0, 0, stream_);
}

int cpb_maxsim_fwd_allgather(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                             const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                             const uint64_t* d_peer_slabs, int n_peers, int my_rank, uint32_t flags, int uniform_len,
                             int max_doc_len, void* d_split_ws, int64_t split_ws_bytes, uint32_t epoch,
                             uint32_t* d_done_counter, int64_t flag_word_offset, uint32_t signal_value, void* stream_) {
  if (!d_peer_slabs) return fail(CPB_E_INVALID, "null peer pointer array");
  return maxsim_fwd_impl(d_q, n_queries, nq_pad, d_docs, doc_rows, d_doc_start, d_doc_len, d_doc_floor, n_docs, nullptr,
                         nullptr, nullptr, flags, uniform_len, max_doc_len, d_split_ws, d_split_ws ? split_ws_bytes : 0,
                         epoch, d_peer_slabs, n_peers, my_rank, d_done_counter, flag_word_offset, signal_value, stream_);
}

int cpb_wait_flags(const uint32_t* d_flags, int n, uint32_t value, void* stream_) {
  if (!d_flags || n <= 0 || n > 64) return fail(CPB_E_INVALID, "bad flag array");
  CPB_CUDA(cpb::wait_flags_launch(d_flags, n, value, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int64_t cpb_maxsim_split_workspace_bytes(int n_queries, int nq_pad) {
  // worst case: every SM is a partition of some query-tile group
  const int64_t qtiles = (static_cast<int64_t>(n_queries) * nq_pad + 127) / 128;
  const int64_t groups = qtiles + 4;  // padded to the cluster size
  return groups * 160 * 2 * (128 * 8 + 16);
}

int cpb_colbert_neg_loss_fwd(const float* d_scores, const float* d_neg_scores, const void* d_q, int n_queries,
                             int nq_pad, int n_docs, int n_neg, int inner_mode, float temperature,
                             int normalize_scores, int pos_aware_negative_filtering, float filter_threshold,
                             float filter_factor, float in_batch_term_weight, int offset, float* d_loss,
                             float* d_grad_scores, float* d_grad_neg_scores, void* stream_) {
  if (!d_neg_scores) return fail(CPB_E_INVALID, "null device pointer");
  return loss_fwd_impl(d_scores, d_q, n_queries, nq_pad, n_docs, inner_mode, temperature, normalize_scores,
                       pos_aware_negative_filtering, filter_threshold, filter_factor, offset, d_neg_scores, n_neg,
                       in_batch_term_weight, d_loss, d_grad_scores, d_grad_neg_scores, nullptr, 128, stream_);
}

int cpb_colbert_neg_loss_fwd_dim(const float* d_scores, const float* d_neg_scores, const void* d_q, int n_queries,
                                 int nq_pad, int n_docs, int n_neg, int inner_mode, float temperature,
                                 int normalize_scores, int pos_aware_negative_filtering, float filter_threshold,
                                 float filter_factor, float in_batch_term_weight, int offset, float* d_loss,
                                 float* d_grad_scores, float* d_grad_neg_scores, int dim, void* stream_) {
  if (!d_neg_scores) return fail(CPB_E_INVALID, "null device pointer");
  return loss_fwd_impl(d_scores, d_q, n_queries, nq_pad, n_docs, inner_mode, temperature, normalize_scores,
                       pos_aware_negative_filtering, filter_threshold, filter_factor, offset, d_neg_scores, n_neg,
                       in_batch_term_weight, d_loss, d_grad_scores, d_grad_neg_scores, nullptr, dim, stream_);
}

int cpb_maxsim_bwd(const float* d_grad_scores, const float* d_grad_out, const int32_t* d_argmax, const void* d_q,
                   int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows, const int32_t* d_doc_start,
                   int n_docs, float* d_dq, float* d_dd, void* stream_) {
  return cpb_maxsim_bwd_dim(d_grad_scores, d_grad_out, d_argmax, d_q, n_queries, nq_pad, d_docs, doc_rows, d_doc_start,
                            n_docs, d_dq, d_dd, 128, stream_);
}

int cpb_maxsim_bwd_dim(const float* d_grad_scores, const float* d_grad_out, const int32_t* d_argmax, const void* d_q,
                       int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows, const int32_t* d_doc_start,
                       int n_docs, float* d_dq, float* d_dd, int dim, void* stream_) {
  if (dim != 128 && dim != 192 && dim != 256 && dim != 320)
    return fail(CPB_E_UNSUPPORTED, "embedding dim %d is not supported by this build (128, 192, 256, 320)", dim);
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
  if (!d_grad_scores || !d_argmax || !d_q || !d_docs || !d_doc_start) return fail(CPB_E_INVALID, "null device pointer");
  if (doc_rows <= 0) return fail(CPB_E_INVALID, "doc_rows must be positive");
  if ((reinterpret_cast<uintptr_t>(d_dq) | reinterpret_cast<uintptr_t>(d_dd) | reinterpret_cast<uintptr_t>(d_q) |
       reinterpret_cast<uintptr_t>(d_docs)) & 15u)
    return fail(CPB_E_INVALID, "tensor pointers must be 16-byte aligned");
  cpb::BwdParams p{};
  p.g = d_grad_scores;
  p.grad_out = d_grad_out;
  p.argmax = d_argmax;
  p.q = static_cast<const __nv_bfloat16*>(d_q);
  p.docs = static_cast<const __nv_bfloat16*>(d_docs);
  p.doc_start = d_doc_start;
  p.dq = d_dq;
  p.dd = d_dd;
  p.B = n_queries;
  p.C = n_docs;
  p.nq_pad = nq_pad;
  p.q_rows = n_queries * nq_pad;
  p.dim = dim;
  CPB_CUDA(cpb::maxsim_bwd_launch(p, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int cpb_head_fwd(const void* d_hidden, int64_t n_tokens, int hidden, const void* d_weight, const void* d_bias, int dim,
                 const int64_t* d_attention_mask, const uint8_t* d_extra_mask, void* d_out, uint32_t flags,
                 void* stream_) {
  if (n_tokens <= 0) return fail(CPB_E_INVALID, "n_tokens=%lld must be positive", static_cast<long long>(n_tokens));
  if (n_tokens > 0x7fffff00LL) return fail(CPB_E_INVALID, "n_tokens too large");
  const bool wide = dim > 128;  // DRAFT: head_wide_sm100.cu
  if (dim != 128 && !(wide && dim <= 320 && (dim % 32) == 0))
    return fail(CPB_E_UNSUPPORTED, "projection dim %d is not supported by this build (128, or a multiple of 32 up to 320)", dim);
  if (hidden <= 0 || (hidden % 64) != 0) return fail(CPB_E_UNSUPPORTED, "hidden size %d must be a positive multiple of 64", hidden);
  if (!d_hidden || !d_weight || !d_out) return fail(CPB_E_INVALID, "null device pointer");
  if (reinterpret_cast<uintptr_t>(d_out) & 15u) return fail(CPB_E_INVALID, "d_out is not 16-byte aligned");
  DevInfo di;
  int rc = current_device_info(&di);
  if (rc != CPB_OK) return rc;
  if (di.major != 10) return fail(CPB_E_DEVICE, "device is sm_%d%d; this library needs sm_100 (B200)", di.major, di.minor);
  CUtensorMap th, tw;
  rc = make_bf16_rowmajor_map(&th, d_hidden, n_tokens, hidden, 128);
  if (rc != CPB_OK) return rc;
  rc = make_bf16_rowmajor_map(&tw, d_weight, dim, hidden, wide ? dim / 2 : 128);
  if (rc != CPB_OK) return rc;
  cpb::HeadParams p{};
  p.bias = static_cast<const __nv_bfloat16*>(d_bias);
  p.attention_mask = d_attention_mask;
  p.extra_mask = d_extra_mask;
  p.out = static_cast<__nv_bfloat16*>(d_out);
  p.n_tokens = n_tokens;
  p.hidden = hidden;
  p.flags = flags;
  p.dim = dim;
  if (wide) {
    p.stages = cpb::head_wide_stages(dim);
    if (p.stages < 2) return fail(CPB_E_UNSUPPORTED, "projection dim %d does not fit the shared-memory ring", dim);
    const int64_t tiles = (n_tokens + 127) / 128;
    p.cluster = (g_head_cluster == 1 || tiles < 2) ? 1 : 2;
    const int64_t rounds = (tiles + p.cluster - 1) / p.cluster;
    const int64_t max_clusters = di.sm_count / p.cluster;
    const int grid = static_cast<int>((rounds < max_clusters ? rounds : max_clusters) * p.cluster);
    CPB_CUDA(cpb::head_wide_launch(th, tw, p, grid, static_cast<cudaStream_t>(stream_)));
    return CPB_OK;
  }
  const int64_t pairs = (n_tokens + 255) / 256;
  const int grid = static_cast<int>(pairs < di.sm_count ? pairs : di.sm_count);
  CPB_CUDA(cpb::head_launch(th, tw, p, grid, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

// DRAFT (r2-drafts): MaxSim forward for embedding dims 192 / 256 / 320 (K-pipelined kernel, one query tile per CTA).
// dim must be a multiple of 64 in (128, 320]; queries and documents are [rows, dim] bf16.  No balancing, no fused gather yet.
int cpb_maxsim_fwd_dim(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                       const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                       float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags, int dim, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dim <= 128 || dim > 320 || (dim % 64) != 0) return fail(CPB_E_UNSUPPORTED, "dim=%d: this entry point serves 192, 256 and 320", dim);
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
  if (!d_q || !d_docs || !d_doc_start || !d_doc_len || !d_scores) return fail(CPB_E_INVALID, "null device pointer");
  if (doc_rows <= 0 || doc_rows > 0x7fffffffLL) return fail(CPB_E_INVALID, "doc_rows out of range");
  const int nseg = nq_pad / 32;
  if (nseg > 1 && !d_workspace) return fail(CPB_E_INVALID, "nq_pad=%d needs a workspace", nq_pad);
  DevInfo di;
  int rc = current_device_info(&di);
  if (rc != CPB_OK) return rc;
  if (di.major != 10) return fail(CPB_E_DEVICE, "device is sm_%d%d; this library needs sm_100 (B200)", di.major, di.minor);
  const int panels = dim / 64;
  cpb::MaxSimParams p{};
  p.q = d_q;
  p.doc_start = d_doc_start;
  p.doc_len = d_doc_len;
  p.doc_floor = d_doc_floor;
  p.argmax = d_argmax;
  p.plane_stride = static_cast<int64_t>(n_queries) * n_docs;
  p.n_queries = n_queries;
  p.nq_pad = nq_pad;
  p.q_rows = n_queries * nq_pad;
  p.n_docs = n_docs;
  p.num_qtiles = (p.q_rows + 127) / 128;
  p.scores = (nseg == 1) ? d_scores : d_workspace;
  p.q_groups = p.num_qtiles;  // one query tile per CTA
  int cluster = (p.q_groups >= 2) ? 2 : 1;
  int max_clusters = cpb::maxsim_kpipe_max_clusters(panels, cluster);
  if (max_clusters <= 0) return fail(CPB_E_CUDA, "K-pipelined kernel cannot be resident on this device");
  p.cluster = cluster;
  p.group_sets = (p.q_groups + cluster - 1) / cluster;
  int parts = max_clusters / p.group_sets;
  if (parts < 1) parts = 1;
  if (parts > n_docs) parts = n_docs;
  p.doc_parts = parts;
  p.flags = flags | g_opt_debug_flags;
  const int grid = p.group_sets * p.doc_parts * cluster;
  CUtensorMap tq, td, tt;
  rc = make_bf16_rowmajor_map(&tq, d_q, p.q_rows, dim, 128);
  if (rc != CPB_OK) return rc;
  rc = make_bf16_rowmajor_map(&td, d_docs, doc_rows, dim, 256 / cluster);
  if (rc != CPB_OK) return rc;
  rc = make_bf16_rowmajor_map(&tt, d_docs, doc_rows, dim, 32);
  if (rc != CPB_OK) return rc;
  CPB_CUDA(cpb::maxsim_kpipe_launch(tq, td, tt, p, panels, d_argmax != nullptr, grid, stream));
  if (nseg > 1)
    CPB_CUDA(cpb::maxsim_reduce_segments(d_workspace, d_scores, p.plane_stride, nseg, (flags & CPB_FLAG_ROUND_BF16) ? 1 : 0, stream));
  return CPB_OK;
}

}  // extern "C"
