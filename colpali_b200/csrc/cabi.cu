// extern "C" surface of libcolpali_b200.so (declared in include/colpali_b200.h).
// Host-side only: argument validation, TMA tensor-map encoding, grid sizing, launches.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/colpali_b200.h"
#include "dense_params.h"
#include "exchange_params.h"
#include "head_params.h"
#include "loss_params.h"
#include "maxsim_params.h"

namespace cpb {
cudaError_t maxsim_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                          const LossParams& lp, int r, int mode, int grid, cudaStream_t stream);
int maxsim_max_clusters(int r, int cluster);
int maxsim_tile_n();
int64_t maxsim_topk_workspace_bytes();
int maxsim_topk_slots();
cudaError_t maxsim_kpipe_launch(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tt, const MaxSimParams& p,
                                const LossParams& lp, int dim_panels, int mode, int grid, cudaStream_t stream);
int maxsim_kpipe_max_clusters(int dim_panels, int cluster);
cudaError_t maxsim_pair_launch(const CUtensorMap& tq, const CUtensorMap& td, const MaxSimParams& p, const LossParams& lp,
                               int r, int mode, int grid, cudaStream_t stream);
int maxsim_pair_max_clusters(int r);
cudaError_t wait_flags_launch(const uint32_t* flags, const uint32_t* values, int n, uint32_t value, uint32_t timeout_ms,
                              uint32_t* status, cudaStream_t stream);
cudaError_t maxsim_reduce_segments(const float* partial, float* out, int64_t plane, int nseg, int round_ref,
                                   cudaStream_t stream);
}  // namespace cpb

namespace {

thread_local char g_err[512] = "";

// tuning knobs (cpb_set_option): atomics, read once per launch; 0 = choose automatically
std::atomic<int> g_opt_cluster{0};
std::atomic<int> g_head_cluster{0};
std::atomic<int> g_opt_qtiles_per_cta{0};
std::atomic<unsigned> g_opt_debug_flags{0};
std::atomic<int> g_opt_dbg_delay{0};
std::atomic<int> g_opt_balanced{1};
std::atomic<int> g_opt_pdl{1};
std::atomic<int> g_opt_boundary_mode{1};
std::atomic<int> g_opt_pair{0};  // CTA-pair MMAs (maxsim_pair_sm100.cu) where the shape allows
std::atomic<int> g_opt_early_spin{0};  // MMA issuer pacing (MaxSimParams::early_spin)
std::atomic<int> g_opt_wait_timeout_ms{120000};
std::atomic<int> g_opt_dense_raster{4};  // DenseDotParams::raster_group
std::mutex g_cache_mu;  // guards the device-property / occupancy caches below

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

// a struct member is present if the caller's struct (struct_size bytes) reaches past it
#define CPB_HAS(args, T, member) ((args)->struct_size >= offsetof(T, member) + sizeof((args)->member))

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
  std::lock_guard<std::mutex> lock(g_cache_mu);
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


// co-resident cluster count of a kernel variant: a property of (device, variant, cluster size), queried once
// variant: 1 / 2 = R of the dim-128 kernel, 3..5 = K panels, 6 / 7 = R + 5 of the CTA-pair kernel
int max_clusters_cached(int dev, int variant, int cluster) {
  static int cache[64][8][5];
  std::lock_guard<std::mutex> lock(g_cache_mu);
  int* slot = (dev >= 0 && dev < 64) ? &cache[dev][variant][cluster] : nullptr;
  if (slot && *slot != 0) return *slot;
  const int n = variant <= 2   ? cpb::maxsim_max_clusters(variant, cluster)
                : variant <= 5 ? cpb::maxsim_kpipe_max_clusters(variant, cluster)
                               : cpb::maxsim_pair_max_clusters(variant - 5);
  if (slot) *slot = n;
  return n;
}

int fill_loss_params(cpb::LossParams* p, const cpb_loss_desc* d, const float* d_scores, const void* d_q, int n_queries,
                     int nq_pad, int n_docs, int dim) {
  if (!d || d->struct_size < offsetof(cpb_loss_desc, d_bounds) + sizeof(float*))
    return fail(CPB_E_INVALID, "cpb_loss_desc is null or its struct_size is too small");
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  const bool bi = (d_q == nullptr);  // single-vector scores (bi-encoder losses): no query rows to count lengths from
  if (!bi) {
    if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
    if (dim != 128 && dim != 192 && dim != 256 && dim != 320)
      return fail(CPB_E_UNSUPPORTED, "embedding dim %d is not supported by this build (128, 192, 256, 320)", dim);
  } else if (d->normalize_scores) {
    return fail(CPB_E_INVALID, "normalize_scores needs the query rows (d_q)");
  }
  const int mode = d->mode;
  if (mode != CPB_LOSS_CE && mode != CPB_LOSS_PAIRWISE && mode != CPB_LOSS_SIGMOID && mode != CPB_LOSS_SYMMETRIC_CE)
    return fail(CPB_E_INVALID, "unknown loss mode %d", mode);
  // ColbertSigmoidLoss flattens a [B, B] matrix (late_interaction_losses.py:456-463); BiSigmoidLoss walks C / B blocks
  if (mode == CPB_LOSS_SIGMOID && !bi && (n_docs != n_queries || d->offset != 0))
    return fail(CPB_E_INVALID, "the sigmoid loss needs n_docs == n_queries and offset == 0 (got %d, %d, %d)", n_docs, n_queries, d->offset);
  if (mode == CPB_LOSS_SIGMOID && bi && (n_docs % n_queries) != 0)
    return fail(CPB_E_INVALID, "the bi-encoder sigmoid loss needs n_docs (%d) to be a multiple of n_queries (%d)", n_docs, n_queries);
  if (mode == CPB_LOSS_SYMMETRIC_CE && (n_docs != n_queries || d->offset != 0 || d->normalize_scores || d->d_neg_scores))
    return fail(CPB_E_INVALID, "the symmetric loss needs a square score matrix, offset 0, no normalisation, no negatives");
  if (d->d_neg_scores && (d->n_neg <= 0 || mode == CPB_LOSS_SIGMOID || d->in_batch_term_weight < 0.f || d->in_batch_term_weight > 1.f))
    return fail(CPB_E_INVALID, "bad explicit-negative arguments (n_neg=%d, mode=%d, weight=%g)", d->n_neg, mode, static_cast<double>(d->in_batch_term_weight));
  const int neg_delta = (d->d_neg_scores && CPB_HAS(d, cpb_loss_desc, neg_pos_offset_delta)) ? d->neg_pos_offset_delta : 0;
  if (d->offset < 0 || d->offset + n_queries > n_docs || d->offset + neg_delta < 0 || d->offset + neg_delta + n_queries > n_docs)
    return fail(CPB_E_INVALID, "positive index out of range: offset=%d (+%d) + n_queries=%d > n_docs=%d", d->offset, neg_delta, n_queries, n_docs);
  if (!(d->temperature > 0.f)) return fail(CPB_E_INVALID, "temperature must be positive");
  if (!d_scores || !d->d_loss) return fail(CPB_E_INVALID, "null device pointer");
  *p = cpb::LossParams{};
  p->scores = d_scores;
  p->q = static_cast<const __nv_bfloat16*>(d_q);
  p->q_dim = dim;
  p->loss = d->d_loss;
  p->grad = d->d_grad_scores;
  p->bounds = d->d_bounds;
  p->B = n_queries;
  p->C = n_docs;
  p->nq_pad = nq_pad;
  p->offset = d->offset;
  p->mode = mode;
  p->normalize = d->normalize_scores;
  p->filter = d->pos_aware_negative_filtering;
  p->temperature = d->temperature;
  p->filter_threshold = d->filter_threshold;
  p->filter_factor = d->filter_factor;
  p->neg_scores = d->d_neg_scores;
  p->grad_neg = d->d_grad_neg_scores;
  p->n_neg = d->n_neg;
  p->in_batch_weight = d->d_neg_scores ? d->in_batch_term_weight : 1.f;
  p->neg_pos_delta = neg_delta;
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
  } else if (!strcmp(name, "balanced")) {
    g_opt_balanced = value != 0;
  } else if (!strcmp(name, "pdl")) {
    if (value < 0 || value > 1) return fail(CPB_E_INVALID, "pdl must be 0 or 1");
    g_opt_pdl = value;
  } else if (!strcmp(name, "boundary_mode")) {
    if (value < 0 || value > 1) return fail(CPB_E_INVALID, "boundary_mode must be 0 or 1");
    g_opt_boundary_mode = value;
  } else if (!strcmp(name, "early_spin")) {
    if (value < 0) return fail(CPB_E_INVALID, "early_spin must be >= 0");
    g_opt_early_spin = value;
  } else if (!strcmp(name, "pair")) {
    if (value < 0 || value > 1) return fail(CPB_E_INVALID, "pair must be 0 or 1");
    g_opt_pair = value;
  } else if (!strcmp(name, "head_cluster")) {
    if (value < 0 || value > 2) return fail(CPB_E_INVALID, "head_cluster must be 0, 1 or 2");
    g_head_cluster = value;
  } else if (!strcmp(name, "dense_raster")) {
    if (value < 1) return fail(CPB_E_INVALID, "dense_raster must be >= 1");
    g_opt_dense_raster = value;
  } else if (!strcmp(name, "wait_timeout_ms")) {
    if (value <= 0) return fail(CPB_E_INVALID, "wait_timeout_ms must be positive");
    g_opt_wait_timeout_ms = value;
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

int64_t cpb_maxsim_split_workspace_bytes(int n_queries, int nq_pad) {
  // worst case: every SM is a partition of some query-tile group
  const int64_t qtiles = (static_cast<int64_t>(n_queries) * nq_pad + 127) / 128;
  const int64_t groups = qtiles + 4;  // padded to the cluster size
  return groups * 160 * 2 * (128 * 8 + 16);
}

int64_t cpb_maxsim_topk_workspace_bytes(void) { return cpb::maxsim_topk_workspace_bytes(); }

int cpb_maxsim_launch(cpb_maxsim_args* a) {
  if (!a || a->struct_size < offsetof(cpb_maxsim_args, d_workspace) + sizeof(float*))
    return fail(CPB_E_INVALID, "cpb_maxsim_args is null or its struct_size is too small");
  // members past d_workspace are optional (absent in a shorter struct = zero)
  void* d_split_ws = CPB_HAS(a, cpb_maxsim_args, d_split_ws) ? a->d_split_ws : nullptr;
  const int64_t split_ws_bytes = CPB_HAS(a, cpb_maxsim_args, split_ws_bytes) ? a->split_ws_bytes : 0;
  const uint32_t epoch = CPB_HAS(a, cpb_maxsim_args, epoch) ? a->epoch : 0;
  const float smooth_tau = CPB_HAS(a, cpb_maxsim_args, smooth_tau) ? a->smooth_tau : 0.f;
  const uint64_t* d_peer_bases = CPB_HAS(a, cpb_maxsim_args, d_peer_bases) ? a->d_peer_bases : nullptr;
  const uint64_t mc_base = CPB_HAS(a, cpb_maxsim_args, mc_base) ? a->mc_base : 0;
  const int n_peers = CPB_HAS(a, cpb_maxsim_args, n_peers) ? a->n_peers : 0;
  const int64_t slab_word_offset = CPB_HAS(a, cpb_maxsim_args, slab_word_offset) ? a->slab_word_offset : 0;
  const int64_t flag_word_offset = CPB_HAS(a, cpb_maxsim_args, flag_word_offset) ? a->flag_word_offset : 0;
  const uint32_t* d_wait_flags = CPB_HAS(a, cpb_maxsim_args, d_wait_flags) ? a->d_wait_flags : nullptr;
  const int n_wait = CPB_HAS(a, cpb_maxsim_args, n_wait) ? a->n_wait : 0;
  const uint32_t wait_value = CPB_HAS(a, cpb_maxsim_args, wait_value) ? a->wait_value : 0;
  const cpb_loss_desc* loss = CPB_HAS(a, cpb_maxsim_args, loss) ? a->loss : nullptr;
  uint32_t* d_done_counter = CPB_HAS(a, cpb_maxsim_args, d_done_counter) ? a->d_done_counter : nullptr;

  float* d_topk_scores = CPB_HAS(a, cpb_maxsim_args, d_topk_scores) ? a->d_topk_scores : nullptr;

  cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
  const int n_queries = a->n_queries, nq_pad = a->nq_pad, n_docs = a->n_docs, dim = a->dim;
  const uint32_t flags = a->flags;
  if (n_queries <= 0 || n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", n_queries, n_docs);
  if (nq_pad <= 0 || (nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", nq_pad);
  if (dim != 128 && dim != 192 && dim != 256 && dim != 320)
    return fail(CPB_E_UNSUPPORTED, "embedding dim %d is not supported by this build (128, 192, 256, 320)", dim);
  if (reinterpret_cast<uintptr_t>(a->d_q) & 15u) return fail(CPB_E_INVALID, "d_q is not 16-byte aligned");
  if (!a->d_q || !a->d_docs || !a->d_doc_start || !a->d_doc_len || (!a->d_scores && !d_peer_bases)) return fail(CPB_E_INVALID, "null device pointer");
  if (a->doc_rows <= 0 || a->doc_rows > 0x7fffffffLL) return fail(CPB_E_INVALID, "doc_rows=%lld out of range (1..2^31-1)", static_cast<long long>(a->doc_rows));
  const int nseg = nq_pad / 32;
  if (nseg > 1 && !a->d_workspace) return fail(CPB_E_INVALID, "nq_pad=%d needs a workspace (cpb_maxsim_workspace_bytes)", nq_pad);
  const int64_t q_rows64 = static_cast<int64_t>(n_queries) * nq_pad;
  if (q_rows64 > 0x7fffffffLL) return fail(CPB_E_INVALID, "too many query rows");
  const bool smooth = smooth_tau > 0.f;
  if (smooth_tau < 0.f || std::isnan(smooth_tau)) return fail(CPB_E_INVALID, "smooth_tau must be >= 0");
  if (smooth && (a->nq_real <= 0 || a->nq_real > nq_pad)) return fail(CPB_E_INVALID, "smooth max needs 0 < nq_real <= nq_pad (got %d, %d)", a->nq_real, nq_pad);
  if (smooth && (a->d_argmax || (flags & CPB_FLAG_ROUND_BF16) || d_peer_bases))
    return fail(CPB_E_UNSUPPORTED, "smooth max excludes argmax, bf16 rounding and the fused all-gather");
  if (a->d_lse && !smooth) return fail(CPB_E_INVALID, "d_lse is an output of the smooth-max mode (smooth_tau > 0)");

  DevInfo di;
  int rc = current_device_info(&di);
  if (rc != CPB_OK) return rc;
  if (di.major != 10) return fail(CPB_E_DEVICE, "device is sm_%d%d; this library needs sm_100 (B200)", di.major, di.minor);
  int dev = 0;
  CPB_CUDA(cudaGetDevice(&dev));

  cpb::MaxSimParams p{};
  p.q = a->d_q;
  p.doc_start = a->d_doc_start;
  p.doc_len = a->d_doc_len;
  p.doc_floor = smooth ? nullptr : a->d_doc_floor;
  p.argmax = a->d_argmax;
  p.lse = a->d_lse;
  p.plane_stride = static_cast<int64_t>(n_queries) * n_docs;
  p.n_queries = n_queries;
  p.nq_pad = nq_pad;
  p.nq_real = smooth ? a->nq_real : nq_pad;
  p.q_rows = static_cast<int>(q_rows64);
  p.n_docs = n_docs;
  p.num_qtiles = (p.q_rows + 127) / 128;
  p.scores = (nseg == 1) ? a->d_scores : a->d_workspace;
  if (smooth) {
    p.smooth_c = 1.4426950408889634f / smooth_tau;
    p.smooth_out = smooth_tau * 0.6931471805599453f;
  }
  if (d_peer_bases) {
    if (nseg != 1) return fail(CPB_E_UNSUPPORTED, "the fused all-gather needs queries of at most 32 tokens (nq_pad == 32)");
    if (n_peers < 1 || n_peers > 64) return fail(CPB_E_INVALID, "bad peer count %d", n_peers);
    if (slab_word_offset < 0 || flag_word_offset < 0) return fail(CPB_E_INVALID, "negative symmetric-buffer offset");
    p.peer_scores = d_peer_bases;
    p.mc_base = mc_base;
    p.n_peers = n_peers;
    p.peer_slab_offset = slab_word_offset;
    p.peer_flag_offset = flag_word_offset;
  }
  if (d_wait_flags) {  // fused all-gather: write-after-read guard; training exchange: the gathered bank is complete
    if (n_wait < 1 || n_wait > 64) return fail(CPB_E_INVALID, "bad n_wait %d", n_wait);
    p.wait_flags = d_wait_flags;
    p.n_wait = n_wait;
    p.wait_value = wait_value;
  }
  p.wait_timeout_ms = static_cast<uint32_t>(g_opt_wait_timeout_ms.load());
  cpb::LossParams lp{};
  if (loss) {
    if (nseg != 1 || d_peer_bases || !d_done_counter || !a->d_scores || (flags & CPB_FLAG_INDEPENDENT))
      return fail(CPB_E_UNSUPPORTED, "the fused loss needs nq_pad == 32, d_scores, d_done_counter, no fused all-gather and no CPB_FLAG_INDEPENDENT");
    rc = fill_loss_params(&lp, loss, a->d_scores, a->d_q, n_queries, nq_pad, n_docs, dim);
    if (rc != CPB_OK) return rc;
    p.done_counter = d_done_counter;
  }
  if (d_topk_scores) {
    if (!CPB_HAS(a, cpb_maxsim_args, topk_k) || !a->d_topk_idx || !a->d_topk_ws || a->topk_k < 1 || a->topk_k > CPB_TOPK_MAX)
      return fail(CPB_E_INVALID, "fused top-k needs d_topk_idx, d_topk_ws and 1 <= topk_k <= %d", CPB_TOPK_MAX);
    if (dim != 128 || nseg != 1 || !a->d_scores || d_peer_bases || (flags & CPB_FLAG_INDEPENDENT))
      return fail(CPB_E_UNSUPPORTED, "fused top-k needs dim 128, nq_pad == 32, d_scores, no fused all-gather and no CPB_FLAG_INDEPENDENT");
    p.topk_scores = d_topk_scores;
    p.topk_idx = a->d_topk_idx;
    p.topk_ws = a->d_topk_ws;
    p.topk_k = a->topk_k;
  }
  const int mode = smooth ? 2 : (a->d_argmax ? 1 : 0);
  p.pdl = g_opt_pdl.load() ? ((flags & CPB_FLAG_INDEPENDENT) ? 2 : 1) : 0;
  p.boundary_mode = g_opt_boundary_mode.load();
  p.flags = (flags & 0xffffu) | g_opt_debug_flags.load();
  p.dbg_delay = g_opt_dbg_delay.load();
  p.early_spin = g_opt_early_spin.load();

  const int opt_cluster = g_opt_cluster.load(), opt_r = g_opt_qtiles_per_cta.load();
  int grid = 0;
  CUtensorMap tq, td, tt;
  if (dim == 128) {
    // Two resident query tiles per CTA halve the L2->SMEM traffic per flop; a single tile only when there is just one.
    int R = (p.num_qtiles >= 2) ? 2 : 1;
    if (opt_r == 1 || opt_r == 2) R = opt_r;
    p.q_groups = (p.num_qtiles + R - 1) / R;
    // CTAs of a cluster hold different query-tile groups and share every document tile through TMA multicast.  Pairs
    // tile the 148 SMs exactly; clusters of 4 strand SMs in GPCs whose SM count is not a multiple of 4 (opt-in).
    int cluster = (p.q_groups >= 2) ? 2 : 1;
    if (opt_cluster == 1 || opt_cluster == 2 || opt_cluster == 4) cluster = opt_cluster;
    // CTA pairs: one M = 256 MMA per two query tiles (one per CTA), half of every document tile per CTA.  Needs every
    // CTA to own all R query tiles and a bank that can be read past the end of a partition (contiguous).
    bool use_pair = g_opt_pair.load() != 0 && cluster == 2 && !d_topk_scores && (flags & CPB_FLAG_CONTIGUOUS) != 0 &&
                    p.num_qtiles % (2 * R) == 0 && !(p.flags & CPB_DBG_NO_TMA);
    int max_clusters = use_pair ? max_clusters_cached(dev, 5 + R, 2) : 0;
    if (max_clusters <= 0) {
      use_pair = false;
      max_clusters = max_clusters_cached(dev, R, cluster);
    }
    if (max_clusters <= 0) {
      if (cluster == 1) return fail(CPB_E_CUDA, "kernel cannot be resident on this device (shared memory?)");
      cluster = 1;
      max_clusters = max_clusters_cached(dev, R, 1);
      if (max_clusters <= 0) return fail(CPB_E_CUDA, "kernel cannot be resident on this device (shared memory?)");
    }
    p.cluster = cluster;
    p.group_sets = (p.q_groups + cluster - 1) / cluster;
    int parts = max_clusters / p.group_sets;
    if (parts < 1) parts = 1;
    if (parts > n_docs) parts = n_docs;
    // tile-balanced partitions (a document may be cut between two partitions): contiguous banks only, every partition
    // at least as long as the longest document, caller-provided exchange workspace, hard max only
    const int64_t tiles = (a->doc_rows + 255) / 256;
    if (g_opt_balanced.load() && !smooth && (flags & CPB_FLAG_CONTIGUOUS) && d_split_ws && a->max_doc_len > 0) {
      if (epoch == 0) return fail(CPB_E_INVALID, "epoch must be non-zero (a zero-initialised workspace means 'nothing published')");
      int bparts = parts;
      if (bparts > tiles) bparts = static_cast<int>(tiles);
      const int64_t min_rows = 256 * (tiles / (bparts > 0 ? bparts : 1));
      const int64_t slots = static_cast<int64_t>(p.group_sets) * cluster * bparts * R;
      const int64_t need = slots * (128 * 8 + 16);
      // launches flagged independent may overlap their predecessors: four generations of slots, picked by the epoch
      const int gens = (p.pdl == 2) ? 4 : 1;
      if (bparts > 1 && min_rows >= a->max_doc_len && need * gens <= split_ws_bytes) {
        parts = bparts;
        p.balanced = 1;
        p.bank_rows = static_cast<int>(a->doc_rows);
        p.uniform_len = a->uniform_len;
        char* base = static_cast<char*>(d_split_ws) + (gens > 1 ? (epoch & 3u) * need : 0);
        p.split_max = reinterpret_cast<float*>(base);
        p.split_idx = reinterpret_cast<int32_t*>(p.split_max + slots * 128);
        p.split_flag = reinterpret_cast<uint32_t*>(p.split_idx + slots * 128);
        p.epoch = epoch;
      }
    }
    p.doc_parts = parts;
    grid = p.group_sets * p.doc_parts * cluster;
    if (loss && p.q_groups >= cpb::kLossWsGroups)
      return fail(CPB_E_UNSUPPORTED, "the fused loss handles at most %d query tiles groups (%d)", cpb::kLossWsGroups - 1, p.q_groups);
    if (d_topk_scores && (grid > cpb::maxsim_topk_slots() || R * 4 > 8))
      return fail(CPB_E_UNSUPPORTED, "fused top-k: grid of %d CTAs exceeds the candidate workspace", grid);
    rc = make_bf16_rowmajor_map(&tq, a->d_q, q_rows64, 128, 128);
    if (rc != CPB_OK) return rc;
    rc = make_bf16_rowmajor_map(&td, a->d_docs, a->doc_rows, 128, cpb::maxsim_tile_n() / cluster);
    if (rc != CPB_OK) return rc;
    rc = make_bf16_rowmajor_map(&tt, a->d_docs, a->doc_rows, 128, 32);
    if (rc != CPB_OK) return rc;
    if (use_pair) {
      CPB_CUDA(cpb::maxsim_pair_launch(tq, td, p, lp, R, mode, grid, stream));
    } else {
      CPB_CUDA(cpb::maxsim_launch(tq, td, tt, p, lp, R, mode, grid, stream));
    }
  } else {
    // K-pipelined kernel: one query tile per CTA, whole-document partitions
    const int panels = dim / 64;
    p.q_groups = p.num_qtiles;
    int cluster = (p.q_groups >= 2) ? 2 : 1;
    if (opt_cluster == 1 || opt_cluster == 2) cluster = opt_cluster;
    int max_clusters = max_clusters_cached(dev, panels, cluster);
    if (max_clusters <= 0) return fail(CPB_E_CUDA, "K-pipelined kernel cannot be resident on this device");
    p.cluster = cluster;
    p.group_sets = (p.q_groups + cluster - 1) / cluster;
    int parts = max_clusters / p.group_sets;
    if (parts < 1) parts = 1;
    if (parts > n_docs) parts = n_docs;
    p.doc_parts = parts;
    grid = p.group_sets * p.doc_parts * cluster;
    rc = make_bf16_rowmajor_map(&tq, a->d_q, q_rows64, dim, 128);
    if (rc != CPB_OK) return rc;
    rc = make_bf16_rowmajor_map(&td, a->d_docs, a->doc_rows, dim, 256 / cluster);
    if (rc != CPB_OK) return rc;
    rc = make_bf16_rowmajor_map(&tt, a->d_docs, a->doc_rows, dim, 32);
    if (rc != CPB_OK) return rc;
    CPB_CUDA(cpb::maxsim_kpipe_launch(tq, td, tt, p, lp, panels, mode, grid, stream));
  }
  if (nseg > 1)
    CPB_CUDA(cpb::maxsim_reduce_segments(a->d_workspace, a->d_scores, p.plane_stride, nseg,
                                         (flags & CPB_FLAG_ROUND_BF16) ? 1 : 0, stream));
  if (CPB_HAS(a, cpb_maxsim_args, grid_out)) a->grid_out = grid;
  return CPB_OK;
}

int cpb_maxsim_fwd(const void* d_q, int n_queries, int nq_pad, const void* d_docs, int64_t doc_rows,
                   const int32_t* d_doc_start, const int32_t* d_doc_len, const float* d_doc_floor, int n_docs,
                   float* d_scores, int32_t* d_argmax, float* d_workspace, uint32_t flags, void* stream_) {
  cpb_maxsim_args a{};
  a.struct_size = sizeof(a);
  a.flags = flags;
  a.stream = stream_;
  a.d_q = d_q;
  a.n_queries = n_queries;
  a.nq_pad = nq_pad;
  a.dim = 128;
  a.d_docs = d_docs;
  a.doc_rows = doc_rows;
  a.d_doc_start = d_doc_start;
  a.d_doc_len = d_doc_len;
  a.d_doc_floor = d_doc_floor;
  a.n_docs = n_docs;
  a.d_scores = d_scores;
  a.d_argmax = d_argmax;
  a.d_workspace = d_workspace;
  return cpb_maxsim_launch(&a);
}

int cpb_colbert_loss_launch(const cpb_loss_desc* loss, const float* d_scores, const void* d_q, int n_queries, int nq_pad,
                            int n_docs, int dim, void* stream_) {
  cpb::LossParams p{};
  int rc = fill_loss_params(&p, loss, d_scores, d_q, n_queries, nq_pad, n_docs, dim);
  if (rc != CPB_OK) return rc;
  CPB_CUDA(cpb::colbert_loss_launch(p, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int cpb_dense_dot_launch(const cpb_dense_dot_args* a) {
  if (!a || a->struct_size < offsetof(cpb_dense_dot_args, out_row_stride) + sizeof(int64_t))
    return fail(CPB_E_INVALID, "cpb_dense_dot_args is null or its struct_size is too small");
  if (a->m <= 0 || a->n <= 0 || a->k <= 0) return fail(CPB_E_INVALID, "m=%d, n=%d, k=%d must be positive", a->m, a->n, a->k);
  if (!a->d_a || !a->d_b || !a->d_out) return fail(CPB_E_INVALID, "null device pointer");
  if (a->out_row_stride < a->n) return fail(CPB_E_INVALID, "out_row_stride=%lld < n=%d", static_cast<long long>(a->out_row_stride), a->n);
  if (((static_cast<int64_t>(a->m) + 31) / 32) * ((static_cast<int64_t>(a->n) + 31) / 32) > 0x7fffffffLL)
    return fail(CPB_E_UNSUPPORTED, "m=%d x n=%d: more than 2^31 output tiles", a->m, a->n);
  cpb::DenseDotParams p{};
  p.a = a->d_a;
  p.b = a->d_b;
  p.b_rows = a->d_b_rows;
  p.a_rs = a->a_row_stride;
  p.a_ks = a->a_k_stride;
  p.b_rs = a->b_row_stride;
  p.b_ks = a->b_k_stride;
  p.m = a->m;
  p.n = a->n;
  p.k = a->k;
  p.out = a->d_out;
  p.out_rs = a->out_row_stride;
  p.alpha = CPB_HAS(a, cpb_dense_dot_args, d_alpha) ? a->d_alpha : nullptr;
  p.accumulate = (a->flags & CPB_DOT_ACCUMULATE) ? 1 : 0;
  p.raster_group = g_opt_dense_raster.load();
  p.a_f32 = (a->flags & CPB_DOT_A_F32) ? 1 : 0;
  p.b_f32 = (a->flags & CPB_DOT_B_F32) ? 1 : 0;
  CPB_CUDA(cpb::dense_dot_launch(p, static_cast<cudaStream_t>(a->stream)));
  return CPB_OK;
}

int cpb_wait_flags(const uint32_t* d_flags, int n, uint32_t value, uint32_t* d_status, void* stream_) {
  if (!d_flags || n <= 0 || n > 64) return fail(CPB_E_INVALID, "bad flag array");
  CPB_CUDA(cpb::wait_flags_launch(d_flags, nullptr, n, value, static_cast<uint32_t>(g_opt_wait_timeout_ms.load()), d_status,
                                  static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int cpb_exchange_push(cpb_exchange_push_args* a) {
  if (!a || a->struct_size < offsetof(cpb_exchange_push_args, flag_word_offset) + sizeof(int64_t))
    return fail(CPB_E_INVALID, "cpb_exchange_push_args is null or its struct_size is too small");
  if (a->n_docs <= 0 || a->len < 0 || a->slot_len < a->len || a->slot_len <= 0) return fail(CPB_E_INVALID, "bad document block shape (%d docs, len %d, slot %d)", a->n_docs, a->len, a->slot_len);
  if (a->dim <= 0 || (a->dim % 8) != 0) return fail(CPB_E_INVALID, "dim=%d must be a positive multiple of 8", a->dim);
  if (!a->d_peer_bases || (a->len > 0 && !a->d_src)) return fail(CPB_E_INVALID, "null device pointer");
  if (a->n_peers < 1 || a->n_peers > 64) return fail(CPB_E_INVALID, "bad peer count %d", a->n_peers);
  if (a->bank_word_offset < 0 || (a->bank_word_offset & 3) || a->flag_word_offset < 0) return fail(CPB_E_INVALID, "bad symmetric-buffer offsets");
  if (reinterpret_cast<uintptr_t>(a->d_src) & 15u) return fail(CPB_E_INVALID, "d_src is not 16-byte aligned");
  cpb::ExchangePushParams p{};
  p.src = static_cast<const __nv_bfloat16*>(a->d_src);
  p.n_docs = a->n_docs;
  p.len = a->len;
  p.slot_len = a->slot_len;
  p.dim = a->dim;
  p.pad_first = a->pad_first ? 1 : 0;
  p.peer_bases = a->d_peer_bases;
  p.mc_base = a->mc_base;
  p.n_peers = a->n_peers;
  p.bank_word_offset = a->bank_word_offset;
  p.flag_word_offset = a->flag_word_offset;
  int grid = 0;
  CPB_CUDA(cpb::exchange_push_launch(p, &grid, static_cast<cudaStream_t>(a->stream)));
  if (CPB_HAS(a, cpb_exchange_push_args, grid_out)) a->grid_out = grid;
  return CPB_OK;
}

int cpb_signal_peers(const uint64_t* d_peer_bases, uint64_t mc_base, int n_peers, int64_t flag_word_offset, void* stream_) {
  if (!d_peer_bases || n_peers < 1 || n_peers > 64 || flag_word_offset < 0) return fail(CPB_E_INVALID, "bad peer arguments");
  CPB_CUDA(cpb::signal_peers_launch(d_peer_bases, mc_base, n_peers, flag_word_offset, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

int cpb_maxsim_bwd_launch(const cpb_maxsim_bwd_args* a) {
  if (!a || a->struct_size < offsetof(cpb_maxsim_bwd_args, d_dd) + sizeof(float*)) return fail(CPB_E_INVALID, "cpb_maxsim_bwd_args is null or its struct_size is too small");
  const int dim = a->dim;
  if (dim != 128 && dim != 192 && dim != 256 && dim != 320)
    return fail(CPB_E_UNSUPPORTED, "embedding dim %d is not supported by this build (128, 192, 256, 320)", dim);
  if (a->n_queries <= 0 || a->n_docs <= 0) return fail(CPB_E_INVALID, "n_queries=%d and n_docs=%d must be positive", a->n_queries, a->n_docs);
  if (a->nq_pad <= 0 || (a->nq_pad % 32) != 0) return fail(CPB_E_INVALID, "nq_pad=%d must be a positive multiple of 32", a->nq_pad);
  const bool smooth = a->d_lse != nullptr;
  if (smooth == (a->d_argmax != nullptr)) return fail(CPB_E_INVALID, "exactly one of d_argmax (hard max) and d_lse (smooth max) must be given");
  if (!a->d_grad_scores || !a->d_q || !a->d_docs || !a->d_doc_start || !a->d_doc_len) return fail(CPB_E_INVALID, "null device pointer");
  if (a->doc_rows <= 0) return fail(CPB_E_INVALID, "doc_rows must be positive");
  if ((reinterpret_cast<uintptr_t>(a->d_dq) | reinterpret_cast<uintptr_t>(a->d_dd) | reinterpret_cast<uintptr_t>(a->d_q) |
       reinterpret_cast<uintptr_t>(a->d_docs)) & 15u)
    return fail(CPB_E_INVALID, "tensor pointers must be 16-byte aligned");
  if (smooth && !(a->smooth_tau > 0.f)) return fail(CPB_E_INVALID, "smooth_tau must be positive with d_lse");
  if (smooth && dim != 128) return fail(CPB_E_UNSUPPORTED, "the smooth-max backward serves dim 128 only (got %d)", dim);
  if (smooth && (a->nq_real <= 0 || a->nq_real > a->nq_pad)) return fail(CPB_E_INVALID, "smooth max needs 0 < nq_real <= nq_pad");
  if (a->max_doc_len <= 0) return fail(CPB_E_INVALID, "max_doc_len must be positive");
  const uint64_t* dd_doc_base = CPB_HAS(a, cpb_maxsim_bwd_args, d_dd_doc_base) ? a->d_dd_doc_base : nullptr;
  if (dd_doc_base && smooth) return fail(CPB_E_UNSUPPORTED, "the peer-scatter dD serves the hard max only");
  const bool grad_bf16 = (a->flags & CPB_FLAG_GRAD_BF16) != 0;
  if (grad_bf16 && (smooth || dd_doc_base))
    return fail(CPB_E_UNSUPPORTED, "CPB_FLAG_GRAD_BF16 serves the hard max without the peer scatter (those accumulate in fp32)");
  cpb::BwdParams p{};
  p.g = a->d_grad_scores;
  p.grad_out = a->d_grad_out;
  p.argmax = a->d_argmax;
  p.lse = a->d_lse;
  p.smooth_c = smooth ? 1.4426950408889634f / a->smooth_tau : 0.f;
  p.q = static_cast<const __nv_bfloat16*>(a->d_q);
  p.docs = static_cast<const __nv_bfloat16*>(a->d_docs);
  p.doc_start = a->d_doc_start;
  p.doc_len = a->d_doc_len;
  p.dq = static_cast<float*>(a->d_dq);
  p.dd = static_cast<float*>(a->d_dd);
  p.out_bf16 = grad_bf16 ? 1 : 0;
  p.dd_doc_base = dd_doc_base;
  p.B = a->n_queries;
  p.C = a->n_docs;
  p.nq_pad = a->nq_pad;
  p.nq_real = smooth ? a->nq_real : a->nq_pad;
  p.q_rows = a->n_queries * a->nq_pad;
  p.dim = dim;
  p.max_doc_len = a->max_doc_len;
  p.contiguous = (a->flags & CPB_FLAG_CONTIGUOUS) ? 1 : 0;
  p.doc_rows = a->doc_rows;
  if (smooth) CPB_CUDA(cpb::smooth_bwd_launch(p, static_cast<cudaStream_t>(a->stream)));
  else CPB_CUDA(cpb::maxsim_bwd_launch(p, static_cast<cudaStream_t>(a->stream)));
  return CPB_OK;
}

int cpb_head_fwd(const void* d_hidden, int64_t n_tokens, int hidden, const void* d_weight, const void* d_bias, int dim,
                 const int64_t* d_attention_mask, const uint8_t* d_extra_mask, void* d_out, uint32_t flags,
                 void* stream_) {
  if (n_tokens <= 0) return fail(CPB_E_INVALID, "n_tokens=%lld must be positive", static_cast<long long>(n_tokens));
  if (n_tokens > 0x7fffff00LL) return fail(CPB_E_INVALID, "n_tokens too large");
  const bool wide = dim > 128;  // head_wide_sm100.cu
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
    p.cluster = (g_head_cluster.load() == 1 || tiles < 2) ? 1 : 2;
    const int64_t rounds = (tiles + p.cluster - 1) / p.cluster;
    const int64_t max_clusters = di.sm_count / p.cluster;
    const int grid = static_cast<int>((rounds < max_clusters ? rounds : max_clusters) * p.cluster);
    CPB_CUDA(cpb::head_wide_launch(th, tw, p, grid, static_cast<cudaStream_t>(stream_)));
    return CPB_OK;
  }
  // one CTA per SM, each an equal share of the 64-token units (at least two units = one tile per CTA)
  CUtensorMap th64;
  rc = make_bf16_rowmajor_map(&th64, d_hidden, n_tokens, hidden, 64);
  if (rc != CPB_OK) return rc;
  const int64_t units = (n_tokens + 63) / 64;
  const int64_t want = (units + 1) / 2;
  const int grid = static_cast<int>(want < di.sm_count ? want : di.sm_count);
  CPB_CUDA(cpb::head_launch(th, th64, tw, p, grid, static_cast<cudaStream_t>(stream_)));
  return CPB_OK;
}

}  // extern "C"
