// Epilogue of the fused MaxSim kernels (shared by maxsim_sm100.cu and maxsim_kpipe_sm100.cu): the four epilogue warps
// read finished accumulator tiles from TMEM, fold the per-query-token maxima over document tokens while walking the
// document boundaries inside each 256-column tile, and emit one score per (query segment, document).
//
// The caller (the kernel) owns shared memory, the barriers and TMEM; this header only needs the two accumulator
// barriers, the TMEM base and the CTA's slice of the problem.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "loss_body.cuh"
#include "maxsim_params.h"
#include "sm100_ptx.cuh"
#include "topk_tail.cuh"

namespace cpb {

constexpr int kEpiTileM = 128;
constexpr int kEpiTileN = 256;

// A run = documents [d, e) stored back to back in the bank, rows [row0, row1): tiled without gaps.
// With CPB_FLAG_CONTIGUOUS the caller guarantees start[j+1] == start[j] + len[j] for the whole bank, so a
// CTA's partition is ONE run (two loads, no scan); otherwise every document is its own run.
struct Run {
  int e, row0, row1;
};
__device__ __forceinline__ Run next_run(const MaxSimParams& p, int d, int d1, int bal_r0, int bal_r1) {
  Run r;
  if (p.balanced) {  // the partition is the row range [bal_r0, bal_r1), whatever documents it cuts
    r.e = d1;
    r.row0 = bal_r0;
    r.row1 = bal_r1;
    return r;
  }
  r.row0 = __ldg(p.doc_start + d);
  if (p.flags & CPB_FLAG_CONTIGUOUS) {
    r.e = d1;
    r.row1 = __ldg(p.doc_start + d1 - 1) + __ldg(p.doc_len + d1 - 1);
  } else {
    r.e = d + 1;
    r.row1 = r.row0 + __ldg(p.doc_len + d);
  }
  return r;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// max over 32 accumulator columns folded into m (16 FMNMX3, shallow dependency tree)
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    t[i] = fmax3(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]));
  float u0 = fmax3(t[0], t[1], __uint_as_float(v[3]));
  float u1 = fmax3(t[2], t[3], __uint_as_float(v[7]));
  float u2 = fmax3(t[4], t[5], __uint_as_float(v[11]));
  float u3 = fmax3(t[6], t[7], __uint_as_float(v[15]));
  float w0 = fmax3(u0, __uint_as_float(v[19]), __uint_as_float(v[23]));
  float w1 = fmax3(u1, __uint_as_float(v[27]), __uint_as_float(v[31]));
  float x0 = fmax3(w0, w1, u2);
  return fmax3(x0, u3, m);
}

// max of 32 accumulator columns (15 FMNMX3)
__device__ __forceinline__ float tree32(const uint32_t (&v)[32]) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    t[i] = fmax3(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]));
  float u0 = fmax3(t[0], t[1], __uint_as_float(v[3]));
  float u1 = fmax3(t[2], t[3], __uint_as_float(v[7]));
  float u2 = fmax3(t[4], t[5], __uint_as_float(v[11]));
  float u3 = fmax3(t[6], t[7], __uint_as_float(v[15]));
  float w0 = fmax3(u0, __uint_as_float(v[19]), __uint_as_float(v[23]));
  float w1 = fmax3(u1, __uint_as_float(v[27]), __uint_as_float(v[31]));
  return fmax3(fmax3(w0, w1, u2), u3, u3);
}

// Column masks are written as ONE unsigned compare per column, (unsigned)(i - lo) < (unsigned)(hi - lo), feeding a
// select: the obvious `(i >= lo && i < hi) ? .. : ..` made nvcc emit a BSSY / BRA / BSYNC diamond per column (32 per
// chunk), which is what made the generic epilogue walk ~10x slower than the tree paths.
__device__ __forceinline__ void mask32(const uint32_t (&v)[32], uint32_t (&x)[32], int lo, int hi) {
  const unsigned span = static_cast<unsigned>(hi - lo);
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = (static_cast<unsigned>(i - lo) < span) ? v[i] : 0xff800000u;  // -inf outside
}

// max over columns lo <= i < hi of a 32-column chunk (mask, then the FMNMX3 tree: no 32-deep dependency chain)
__device__ __forceinline__ float max32_range(const uint32_t (&v)[32], float m, int lo, int hi) {
  if (lo == 0 && hi == 32) return max32(v, m);
  uint32_t x[32];
  mask32(v, x, lo, hi);
  return max32(x, m);
}

// Argmax (training forward).  Looking up WHERE a chunk's maximum sits costs ~95 instructions (compare + select per
// column, min tree) against 16 for the maximum itself, and one warp per scheduler issues them at 0.3-0.5 IPC.  So the
// position is resolved lazily: per lane the epilogue keeps the running maximum, the first document-relative index of the
// chunk it came from (`bidx0`) and that chunk's 32 values in a 128-byte slot of shared memory -- eight predicated
// 16-byte stores when a chunk takes over (keeping them in registers cost 32 selects per chunk, twice the maximum
// itself).  The first-occurrence lookup runs ONCE PER DOCUMENT in argmax_resolve.  Strict '>' between chunks and the
// first equal column inside the chunk give the earliest maximum, which is what torch.max(dim) returns on ties.
//
// Slot layout: lane l of epilogue warp w owns bytes [(32 w + l) * 128, +128) of the query tile's 16 KiB area; its k-th
// 16-byte piece sits at ((k ^ (l & 7)) << 4), so the eight lanes of a quarter-warp cover all 32 banks.
constexpr int kBcBytesPerTile = 4 * 32 * 128;  // per resident query tile

// the four stores of one half chunk, all under one predicate (a C++ `if` would become a BSSY / BRA / BSYNC diamond)
__device__ __forceinline__ void bc_store_half(uint32_t slot, uint32_t sw, int half, bool take, const uint32_t* v) {
  const uint32_t a0 = slot + (((4 * half + 0) << 4) ^ sw), a1 = slot + (((4 * half + 1) << 4) ^ sw);
  const uint32_t a2 = slot + (((4 * half + 2) << 4) ^ sw), a3 = slot + (((4 * half + 3) << 4) ^ sw);
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %20, 0;\n\t"
      "@p st.shared.v4.b32 [%0], {%4, %5, %6, %7};\n\t"
      "@p st.shared.v4.b32 [%1], {%8, %9, %10, %11};\n\t"
      "@p st.shared.v4.b32 [%2], {%12, %13, %14, %15};\n\t"
      "@p st.shared.v4.b32 [%3], {%16, %17, %18, %19};\n\t}"
      ::"r"(a0), "r"(a1), "r"(a2), "r"(a3),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(static_cast<uint32_t>(take))
      : "memory");
}
__device__ __forceinline__ void bc_store(uint32_t slot, uint32_t sw, bool take, const uint32_t (&v)[32]) {
  bc_store_half(slot, sw, 0, take, &v[0]);
  bc_store_half(slot, sw, 1, take, &v[16]);
}
// a whole 32-column chunk
__device__ __forceinline__ void argmax_fold_full(const uint32_t (&v)[32], float& m, int& bidx0, uint32_t slot, uint32_t sw,
                                                 int idx0) {
  const float t = tree32(v);
  const bool take = t > m;
  m = take ? t : m;
  bidx0 = take ? idx0 : bidx0;
  bc_store(slot, sw, take, v);
}
// columns lo <= i < hi of a chunk
__device__ __forceinline__ void argmax_fold(const uint32_t (&v)[32], float& m, int& bidx0, uint32_t slot, uint32_t sw,
                                            int idx0, int lo, int hi) {
  if (lo == 0 && hi == 32) {
    argmax_fold_full(v, m, bidx0, slot, sw, idx0);
  } else {
    uint32_t x[32];
    mask32(v, x, lo, hi);
    argmax_fold_full(x, m, bidx0, slot, sw, idx0);
  }
}
// `bidx0` of a lane whose running maximum is still the initial value.  (Not -1: the first chunk of a document that starts
// in the middle of a 32-column chunk has a NEGATIVE first index, down to -31.)
constexpr int kNoChunk = -0x40000000;
// document-relative index of the first maximal token, -1 if no token beat the initial value (the floor won)
__device__ __forceinline__ int argmax_resolve(uint32_t slot, uint32_t sw, float m, int bidx0) {
  int best = 64;
#pragma unroll
  for (int k = 0; k < 8; k += 2) {  // eight columns at a time: the slot is read back in pieces to keep few registers live
    uint32_t w[8];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(slot + ((k << 4) ^ sw)) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(slot + (((k + 1) << 4) ^ sw)) : "memory");
    int c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = (__uint_as_float(w[i]) == m) ? 4 * k + i : 64;
    best = min(best, min(min(min(c[0], c[1]), min(c[2], c[3])), min(min(c[4], c[5]), min(c[6], c[7]))));
  }
  return (bidx0 == kNoChunk) ? -1 : bidx0 + best;
}

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// The CTA's slice of the problem, computed once by the kernel and shared by its three roles.
struct CtaSlice {
  int g;        // query-tile group of this CTA
  int part;     // document partition
  int r_cnt;    // resident query tiles that exist (0 = phantom CTA of an odd group count)
  int d0, d1;   // documents [d0, d1) (whole-document partitions) / [0, n_docs) in balanced mode
  int bal_r0, bal_r1;  // balanced mode: bank rows [bal_r0, bal_r1)
};

// Aggregation over a document's tokens
constexpr int kModeMax = 0;     // running maximum (scorer, losses without gradient)
constexpr int kModeArgmax = 1;  // maximum + index of the first maximal token (training forward)
constexpr int kModeSmooth = 2;  // tau * logsumexp(raw / tau): online (max, sum) pair per query row

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// online log-sum-exp over columns lo <= i < hi of a 32-column chunk, in base-2 units scaled by c = log2(e) / tau:
// (m, l) <- (max(m, max_i y_i), l * 2^(m - m') + sum_i 2^(y_i - m')) with y_i = c * v_i
__device__ __forceinline__ void lse32_range(const uint32_t (&v)[32], float& m, float& l, float c, int lo, int hi) {
  float y[32];
  float cm = -INFINITY;
  const unsigned span = static_cast<unsigned>(hi - lo);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    y[i] = (static_cast<unsigned>(i - lo) < span) ? __uint_as_float(v[i]) * c : -INFINITY;
    cm = fmaxf(cm, y[i]);
  }
  const float mn = fmaxf(m, cm);
  const float ms = (mn == -INFINITY) ? 0.f : mn;  // nothing seen yet: every term below is 2^(-inf) = 0, never NaN
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += ex2_approx(y[i] - ms);
  l = l * ex2_approx(m - ms) + acc;
  m = mn;
}
// the same over a whole chunk (no masks): FMNMX3 tree for the maximum, one FFMA + MUFU.EX2 + FADD per element
__device__ __forceinline__ void lse32_full(const uint32_t (&v)[32], float& m, float& l, float c) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    t[i] = fmax3(fmax3(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2])),
                 __uint_as_float(v[4 * i + 3]), __uint_as_float(v[4 * i + 3]));
  const float cm = fmax3(fmax3(t[0], t[1], t[2]), fmax3(t[3], t[4], t[5]), fmaxf(t[6], t[7])) * c;  // c > 0
  const float mn = fmaxf(m, cm);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    a0 += ex2_approx(fmaf(__uint_as_float(v[i]), c, -mn));
    a1 += ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mn));
    a2 += ex2_approx(fmaf(__uint_as_float(v[i + 2]), c, -mn));
    a3 += ex2_approx(fmaf(__uint_as_float(v[i + 3]), c, -mn));
  }
  l = l * ex2_approx(m - mn) + ((a0 + a1) + (a2 + a3));  // m = -inf the first time: 2^(-inf) = 0
  m = mn;
}

// one fp32 to every rank's copy of the symmetric buffer through the NVSwitch multicast mapping
__device__ __forceinline__ void multimem_st_f32(uint64_t mc_addr, float x) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(x) : "memory");
}

// Runs on warps 2..5 (one TMEM lane quadrant each).  R = resident query tiles per CTA.
// kPair: the CTA is half of a cta_group::2 pair -- the accumulator is handed back on the LEADER's tmem_empty barrier
// (rank 0 of the cluster, where the MMA issuer waits for eight arrivals: four epilogue warps in each CTA).
// kGroups: 1 = warps 2..5 walk all R query tiles; 2 (R = 2 only) = warps 2..5 take query tile 0 and warps 6..9 query tile
// 1, i.e. one accumulator of the ping-pong each.  A warp alone on its scheduler issues at 0.3-0.5 IPC, so two warps per
// scheduler nearly double what the epilogue can fold per tensor cycle (it is the bottleneck in the argmax and smooth
// modes); the per-tile state never crosses warps because every query tile has its own running maxima.
template <int R, int kMode, bool kPair = false, int kGroups = 1>
__device__ __forceinline__ void maxsim_epilogue(const MaxSimParams& p, const CtaSlice& sl, uint32_t tmem_base,
                                                uint64_t* tmem_full, uint64_t* tmem_empty, int warp, int lane,
                                                uint8_t* bc_smem = nullptr, int grp = 0) {
  constexpr int kTileM = kEpiTileM;
  constexpr int kTileN = kEpiTileN;
  constexpr bool kArgmax = (kMode == kModeArgmax);
  constexpr bool kSmooth = (kMode == kModeSmooth);
  const int g = sl.g, part = sl.part, r_cnt = sl.r_cnt, d0 = sl.d0, d1 = sl.d1, bal_r0 = sl.bal_r0, bal_r1 = sl.bal_r1;
  static_assert(kGroups == 1 || (kGroups == 2 && R == 2), "two epilogue warp groups = one per resident query tile");
  const int quad = warp & 3;  // TMEM lane quadrant this warp may read
  const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
  const bool round_ref = (p.flags & CPB_FLAG_ROUND_BF16) != 0;
  const bool skip = (p.flags & CPB_DBG_SKIP_EPILOGUE) != 0;
  const uint32_t empty_leader = kPair ? mapa_u32(smem_u32(tmem_empty), 0u) : 0u;
  const bool shifted_boundary = p.boundary_mode == 1;  // read once: a constant-bank load inside the hold window costs ~60 cycles

  // document `doc` is complete for resident query tile r: fold this query segment's 32 token maxima
  // (smooth mode: mm / ll are the online log-sum-exp pair in base-2 units, late_interaction_losses.py:40-44)
  // first padded query row / query / 32-row segment of this warp for resident query tile 0 and R - 1 (computed once:
  // the two divisions sat in every inlined copy of the emission code)
  const int f_row0[2] = {(g * R) * kTileM + quad * 32, (g * R + R - 1) * kTileM + quad * 32};
  const int f_q[2] = {f_row0[0] / p.nq_pad, f_row0[1] / p.nq_pad};
  const int f_seg[2] = {(f_row0[0] % p.nq_pad) >> 5, (f_row0[1] % p.nq_pad) >> 5};
  auto finalize = [&](int doc, int r, float mm, int ai, float ll) {
    const int row0 = (r == 0) ? f_row0[0] : f_row0[1];
    const int q = (r == 0) ? f_q[0] : f_q[1];
    const int seg = (r == 0) ? f_seg[0] : f_seg[1];
    if (kArgmax && p.argmax != nullptr && row0 + lane < p.q_rows)
      p.argmax[static_cast<int64_t>(doc) * p.q_rows + row0 + lane] = ai;
    float x;
    if constexpr (kSmooth) {
      // tau * ln(sum_s exp(raw_s / tau)); EVERY row of the query tensor counts (an all-zero row adds tau * ln N_d,
      // as in the reference) except the rows QueryBlock appended to reach a multiple of 32
      const float v = (mm + lg2_approx(ll)) * p.smooth_out;
      const int rowi = row0 + lane;
      if (p.lse != nullptr && rowi < p.q_rows) p.lse[static_cast<int64_t>(doc) * p.q_rows + rowi] = v;
      x = warp_sum(((rowi % p.nq_pad) < p.nq_real && rowi < p.q_rows) ? v : 0.f);
    } else {
      x = round_ref ? round_bf16(mm) : mm;
      x = warp_sum(x);
      if (round_ref && p.nq_pad == 32) x = round_bf16(x);
    }
    if (lane == 0 && q < p.n_queries && !(p.flags & CPB_DBG_CLOCKS)) {
      if (p.peer_scores != nullptr) {
        // fused all-gather of the score slabs: the score goes straight into every rank's copy of
        // gathered[parity][my_rank] -- one multimem.st through the NVSwitch multicast mapping, or one 4-byte store
        // per peer mapping (no collective kernel afterwards; completion is signalled per CTA in the teardown)
        const int64_t off = p.peer_slab_offset + static_cast<int64_t>(q) * p.n_docs + doc;
        if (p.mc_base != 0) {
          multimem_st_f32(p.mc_base + 4ull * static_cast<uint64_t>(off), x);
        } else {
          for (int pr = 0; pr < p.n_peers; ++pr) reinterpret_cast<float*>(__ldg(p.peer_scores + pr))[off] = x;
        }
      } else {
        p.scores[static_cast<int64_t>(seg) * p.plane_stride + static_cast<int64_t>(q) * p.n_docs + doc] = x;
      }
    }
  };
  auto doc_init = [&](int doc) {
    return (!kSmooth && p.doc_floor != nullptr) ? __ldg(p.doc_floor + doc) : -INFINITY;
  };

  float m[R], ls[R];
  int am[R];  // argmax mode: first document-relative index of the chunk the running maximum came from (`bidx0`)
  constexpr int kNoIdx = kArgmax ? kNoChunk : -1;
  // argmax mode: this lane's best-chunk slot of resident query tile 0 (tile r: + r * kBcBytesPerTile) and its swizzle
  const uint32_t bc_slot0 = kArgmax ? smem_u32(bc_smem) + static_cast<uint32_t>((quad * 32 + lane) * 128) : 0u;
  const uint32_t bc_sw = static_cast<uint32_t>(lane & 7) << 4;
  uint32_t job = 0;
  const bool dbg = (p.flags & CPB_DBG_CLOCKS) != 0;
  // debug cycle counters (CPB_DBG_CLOCKS) live in shared memory: in registers they were spilled, and re-loaded on every
  // job whether the flag was set or not.  Only the warp whose counters are reported (group 0, quadrant 2) keeps them.
  __shared__ int s_dbg[6];  // wait for MMA, hold (fast paths), after release, hold (other paths), other jobs, own jobs
  const bool dbg_me = dbg && grp == 0 && quad == 2 && lane == 0;
  if (dbg_me) {
    for (int i = 0; i < 6; ++i) s_dbg[i] = 0;
  }

  // ---- balanced mode: which document contains the first row of my partition, and is it cut? ----------------
  int first_doc = d0;
  bool head_frag = false;
  if (p.balanced) {
    if (bal_r0 == 0) {
      first_doc = 0;
    } else if (p.uniform_len > 0) {
      first_doc = bal_r0 / p.uniform_len;
    } else {  // last document that starts at or before bal_r0 (starts are sorted in a contiguous bank)
      int lo = 0, hi = p.n_docs - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(p.doc_start + mid) <= bal_r0) lo = mid; else hi = mid - 1;
      }
      first_doc = lo;
    }
    first_doc = min(first_doc, p.n_docs - 1);
    head_frag = __ldg(p.doc_start + first_doc) < bal_r0;
  }
  // slot of the boundary between partitions `bp` and `bp + 1` for resident query tile r
  auto split_slot = [&](int bp, int r) { return ((g * p.doc_parts + bp) * R + r); };
  auto publish = [&](int r, float mm, int ai) {  // I hold the RIGHT part of a cut document
    const int slot = split_slot(part - 1, r);
    p.split_max[slot * 128 + quad * 32 + lane] = mm;
    if (kArgmax) p.split_idx[slot * 128 + quad * 32 + lane] = ai;
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(p.split_flag + slot * 4 + quad) = p.epoch;
  };
  auto consume = [&](int r, float& mm, int& ai) {  // I hold the LEFT part: wait for the neighbour's partial, combine
    const int slot = split_slot(part, r);
    if (lane == 0) {
      const volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(p.split_flag + slot * 4 + quad);
      // the neighbour CTA may not be resident yet (PDL overlap, an SM busy with another stream): wait for it as long
      // as for a remote rank before calling it a protocol failure
      const uint64_t t0 = global_timer_ns();
      const uint64_t limit = static_cast<uint64_t>(p.wait_timeout_ms) * 1000000ull;
      while (*f != p.epoch) {
        if (global_timer_ns() - t0 > limit) __trap();
        __nanosleep(32);
      }
    }
    __syncwarp();
    __threadfence();
    const float om = __ldcg(p.split_max + slot * 128 + quad * 32 + lane);
    const int oi = kArgmax ? __ldcg(p.split_idx + slot * 128 + quad * 32 + lane) : -1;
    // hand the slot back: a CUDA-graph replay repeats this launch with the SAME epoch, and must not mistake the flag
    // its previous replay left behind for the neighbour's new partial (all lanes have read the partial first)
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(p.split_flag + slot * 4 + quad) = 0u;
    if (om > mm) {  // on a tie the earlier (left) token wins, like torch.max
      mm = om;
      ai = oi;
    }
  };

  for (int d = p.balanced ? first_doc : d0; d < d1;) {
    const Run run = next_run(p, d, d1, bal_r0, bal_r1);
    if (run.row1 == run.row0) {
      // nothing but empty documents: their score is the sum of the floors (balanced mode: an empty partition)
      if (!p.balanced)
        for (int e = d; e < run.e; ++e)
          for (int r = 0; r < r_cnt; ++r)
            if (kGroups == 1 || r == grp) finalize(e, r, doc_init(e), -1, 0.f);
      d = run.e;
      continue;
    }
    // state at the start of each tile: current document, its first/last bank row, running maxima
    int cur = d;
    int cur_row0 = p.balanced ? __ldg(p.doc_start + cur) : run.row0;
    int cur_end = cur_row0 + __ldg(p.doc_len + cur);
    // length and floor of the FOLLOWING document are fetched when the cursor moves, long before they are needed:
    // a global load while the accumulator is held costs ~300 cycles of tensor idle time on boundary tiles
    int cur_nlen = (cur + 1 < run.e) ? __ldg(p.doc_len + cur + 1) : 0;
    float cur_ninit = (cur + 1 < run.e) ? doc_init(cur + 1) : -INFINITY;
    {
      const float init = doc_init(cur);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        m[r] = init;
        am[r] = kNoIdx;
        ls[r] = 0.f;
      }
    }
    for (int row = run.row0; row < run.row1; row += kTileN) {
      const int n_valid = min(kTileN, run.row1 - row);
      const int tile_end = row + n_valid;
      int nxt = cur, nxt_row0 = cur_row0, nxt_end = cur_end, nxt_nlen = cur_nlen;
      float nxt_ninit = cur_ninit;
      // One copy of the job body for the argmax / smooth modes: with the loop over the resident query tiles unrolled
      // the generic walk exists twice and the kernel outgrows the instruction cache (measured: the cfg3 argmax forward
      // ran at ~12 k cycles per job).  The max mode keeps the unrolled form its per-job budget was tuned with.
#pragma unroll(kMode == kModeMax ? R : 1)
      for (int r = 0; r < R; ++r) {
        if (r < r_cnt && (kGroups == 1 || r == grp)) {
          const uint32_t jb = job + static_cast<uint32_t>(r);  // jobs are numbered tile-major, query tile minor
          const uint32_t a = jb & 1u;
          const uint32_t aphase = (jb >> 1) & 1u;
          const long long t0 = dbg ? clock64() : 0;
          mbar_wait(&tmem_full[a], aphase);
          const long long t1 = dbg ? clock64() : 0;
          long long t2 = 0;
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_base + a * kTileN;
          // per-tile state by select, not by index: r is a run-time value when the loop is not unrolled
          float mm = (r == 0) ? m[0] : m[R - 1], ll = (r == 0) ? ls[0] : ls[R - 1];
          int ai = (r == 0) ? am[0] : am[R - 1];
          const uint32_t bc_slot = bc_slot0 + static_cast<uint32_t>(r * kBcBytesPerTile);
          int doc = cur, doc_row0 = cur_row0, doc_end = cur_end, doc_nlen = cur_nlen;
          float doc_ninit = cur_ninit;

          // the current document is complete: emit it and step to the next one of the run
          auto emit_doc = [&](int e_doc, float e_mm, int e_ai, float e_ll) {  // stores only: may run after the release
            if (head_frag && e_doc == first_doc) publish(r, e_mm, e_ai); else finalize(e_doc, r, e_mm, e_ai, e_ll);
          };
          auto advance_doc = [&]() {  // step the cursor to the next document of the run
            ++doc;
            if (doc >= run.e) {
              doc_end = 0x7fffffff;  // run exhausted
              return;
            }
            doc_row0 = doc_end;
            doc_end = doc_row0 + doc_nlen;
            mm = doc_ninit;
            ai = kNoIdx;
            ll = 0.f;
            doc_nlen = (doc + 1 < run.e) ? __ldg(p.doc_len + doc + 1) : 0;
            doc_ninit = (doc + 1 < run.e) ? doc_init(doc + 1) : -INFINITY;
          };
          auto finish_doc = [&]() {
            if constexpr (kArgmax) ai = argmax_resolve(bc_slot, bc_sw, mm, ai);
            emit_doc(doc, mm, ai, ll);
            advance_doc();
          };
          auto release_acc = [&]() {  // accumulator drained: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kPair) mbar_arrive_cluster(empty_leader + a * 8u); else mbar_arrive(&tmem_empty[a]);
            }
            if (dbg) t2 = clock64();
          };

          // Three warp-uniform cases.  (1) the whole 256-column tile belongs to one document: FMNMX3 trees.
          // (2) exactly one document boundary inside a full tile: the same trees, routed to the old or the
          // new document per 32-column chunk, plus one masked pass over the chunk that holds the boundary.
          // (3) anything else (short documents, last tile of a run, argmax): generic masked walk.
          int path = 3;
          if (!kSmooth && n_valid == kTileN) {
            if (doc_end >= tile_end) {
              path = 1;
            } else if (doc_end > row && doc + 1 < run.e) {
              if (doc_end + doc_nlen >= tile_end) path = 2;
            }
          }
          if (skip) {
            if (p.dbg_delay > 0) {  // profiling aid: hold the (unread) accumulator for a fixed number of cycles
              const long long t_start = clock64();
              while (clock64() - t_start < p.dbg_delay) {
              }
            }
            release_acc();
            while (doc_end <= tile_end) {  // advance the cursor without reading the accumulator
              ++doc;
              if (doc >= run.e) { doc_end = 0x7fffffff; break; }
              doc_row0 = doc_end;
              doc_end = doc_row0 + doc_nlen;
              doc_nlen = (doc + 1 < run.e) ? __ldg(p.doc_len + doc + 1) : 0;
            }
          } else if (path == 1) {
            // software pipeline: the loads of columns [64k+64, 64k+128) are in flight while [64k, 64k+64) fold
            // (TMEM reads are ~64 B/clk per lane quadrant: ~490 cycles for 128 x 256 fp32 whatever the warp count)
            const int idx_t = row - doc_row0;  // document-relative index of the tile's first column (argmax)
            auto fold = [&](const uint32_t (&v)[32], int k) {
              if constexpr (kArgmax) argmax_fold_full(v, mm, ai, bc_slot, bc_sw, idx_t + 32 * k);
              else mm = max32(v, mm);
            };
            uint32_t va[32], vb[32], vc[32], vd[32];
            tmem_ld_x32(taddr, va);
            tmem_ld_x32(taddr + 32, vb);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + 64, vc);
            tmem_ld_x32(taddr + 96, vd);
            fold(va, 0);
            fold(vb, 1);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            tmem_ld_x32(taddr + 128, va);
            tmem_ld_x32(taddr + 160, vb);
            fold(vc, 2);
            fold(vd, 3);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + 192, vc);
            tmem_ld_x32(taddr + 224, vd);
            fold(va, 4);
            fold(vb, 5);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            release_acc();  // every accumulator read has landed in registers
            fold(vc, 6);
            fold(vd, 7);
            while (doc_end <= tile_end) finish_doc();  // document (and empty followers) ending at the tile end
          } else if (kArgmax && path == 2 && doc_end - row >= 32 && doc_end - row <= kTileN - 32) {
            // one boundary, argmax: the same boundary-aligned chunks as the max mode below, but in order -- the old
            // document's chunks, its lookup (which reads the best-chunk slot back), then the new document's chunks into
            // the same slot.  Overlapping chunks are harmless: a repeated value is not GREATER than the running maximum.
            const int b = doc_end - row;
            const int n_old = (b + 31) >> 5;
            auto col = [&](int i) { return (i < n_old) ? min(32 * i, b - 32) : min(b + 32 * (i - n_old), kTileN - 32); };
            // the old document's result is looked up at the boundary (the slot is about to be reused) but STORED after
            // the accumulator has been handed back: the emission code is several hundred instructions
            int s_doc = 0, s_ai = 0;
            float s_mm = 0.f;
            auto step = [&](const uint32_t (&v)[32], int i) {
              if (i == n_old) {  // warp-uniform: the old document is complete, the cursor moves on
                s_doc = doc;
                s_mm = mm;
                s_ai = argmax_resolve(bc_slot, bc_sw, mm, ai);
                advance_doc();
              }
              argmax_fold_full(v, mm, ai, bc_slot, bc_sw, row + col(i) - doc_row0);
            };
            // two loads in flight while two chunks fold (tcgen05.wait::ld waits for ALL outstanding loads, so the
            // pipeline advances in pairs): with one in flight the walk ran at the TMEM load latency, ~150 cycles a chunk
            uint32_t va[32], vb[32], vc[32], vd[32];
            tmem_ld_x32(taddr + col(0), va);
            tmem_ld_x32(taddr + col(1), vb);
#pragma unroll 1
            for (int i = 0; i < 8; i += 4) {
              tmem_ld_wait();
              reg_fence32(va);
              reg_fence32(vb);
              tmem_ld_x32(taddr + col(i + 2), vc);
              tmem_ld_x32(taddr + col(i + 3), vd);
              step(va, i);
              step(vb, i + 1);
              tmem_ld_wait();
              reg_fence32(vc);
              reg_fence32(vd);
              tmem_ld_x32(taddr + col(i + 4), va);                  // chunk 4, then chunk 8
              if (i == 0) tmem_ld_x32(taddr + col(i + 5), vb);      // chunk 5
              step(vc, i + 2);
              step(vd, i + 3);
            }
            tmem_ld_wait();
            reg_fence32(va);
            release_acc();
            step(va, 8);
            emit_doc(s_doc, s_mm, s_ai, 0.f);
            while (doc_end <= tile_end) finish_doc();
          } else if (!kArgmax && path == 2 && shifted_boundary && doc_end - row >= 32 && doc_end - row <= kTileN - 32) {
            // one boundary at column b, at least 32 columns from either edge: read the tile as 32-column chunks
            // ALIGNED TO THE BOUNDARY -- the old document's columns [0, b) as chunks at min(32 i, b - 32), the new
            // one's [b, 256) at min(b + 32 j, 224).  Chunks of one document may overlap (a maximum is idempotent), so
            // every chunk is a plain FMNMX3 tree: no element-wise split, no re-read.  8 or 9 chunks (a 9th slot
            // that is not needed repeats the last chunk).
            const int b = doc_end - row;
            const int n_old = (b + 31) >> 5;
            float mb = doc_ninit;
            auto col = [&](int i) { return (i < n_old) ? min(32 * i, b - 32) : min(b + 32 * (i - n_old), kTileN - 32); };
            auto route = [&](const uint32_t (&v)[32], int i) {
              const float t = tree32(v);
              mm = (i < n_old) ? fmaxf(mm, t) : mm;
              mb = (i >= n_old) ? fmaxf(mb, t) : mb;
            };
            uint32_t va[32], vb[32], vc[32], vd[32];
            tmem_ld_x32(taddr + col(0), va);
            tmem_ld_x32(taddr + col(1), vb);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + col(2), vc);
            tmem_ld_x32(taddr + col(3), vd);
            route(va, 0);
            route(vb, 1);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            tmem_ld_x32(taddr + col(4), va);
            tmem_ld_x32(taddr + col(5), vb);
            route(vc, 2);
            route(vd, 3);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + col(6), vc);
            tmem_ld_x32(taddr + col(7), vd);
            route(va, 4);
            route(vb, 5);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            tmem_ld_x32(taddr + col(8), va);
            route(vc, 6);
            route(vd, 7);
            tmem_ld_wait();
            reg_fence32(va);
            release_acc();
            route(va, 8);
            finish_doc();  // old document (running max mm); the cursor moves to the new one, whose max is mb
            mm = mb;
            while (doc_end <= tile_end) finish_doc();
          } else if (!kArgmax && path == 2) {
            // one boundary at column b: per 32-column chunk the FMNMX3 tree goes to the old document (chunk < kb)
            // or the new one (chunk > kb) through selects; the boundary chunk itself is re-read at the end and split
            // element-wise after the release (branching per chunk would be if-converted into doing everything).
            const int b = doc_end - row;
            const int kb = b >> 5;
            float mb = doc_ninit;
            auto route = [&](const uint32_t (&v)[32], int c) {
              const float t = tree32(v);
              mm = (c < kb) ? fmaxf(mm, t) : mm;
              mb = (c > kb) ? fmaxf(mb, t) : mb;
            };
            uint32_t va[32], vb[32], vc[32], vd[32];
            tmem_ld_x32(taddr, va);
            tmem_ld_x32(taddr + 32, vb);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + 64, vc);
            tmem_ld_x32(taddr + 96, vd);
            route(va, 0);
            route(vb, 1);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            tmem_ld_x32(taddr + 128, va);
            tmem_ld_x32(taddr + 160, vb);
            route(vc, 2);
            route(vd, 3);
            tmem_ld_wait();
            reg_fence32(va);
            reg_fence32(vb);
            tmem_ld_x32(taddr + 192, vc);
            tmem_ld_x32(taddr + 224, vd);
            route(va, 4);
            route(vb, 5);
            tmem_ld_wait();
            reg_fence32(vc);
            reg_fence32(vd);
            tmem_ld_x32(taddr + kb * 32, va);  // the boundary chunk again
            route(vc, 6);
            route(vd, 7);
            tmem_ld_wait();
            reg_fence32(va);
            release_acc();
            {  // split the boundary chunk: columns < bl belong to the old document (two masked trees, branch-free)
              const int bl = b & 31;
              mm = max32_range(va, mm, 0, bl);
              mb = max32_range(va, mb, bl, 32);
            }
            finish_doc();  // old document (running max mm); the cursor moves to the new one, whose max is mb
            mm = mb;
            while (doc_end <= tile_end) finish_doc();
          } else {
            // generic walk, software-pipelined over two register buffers: chunk c + 1 is in flight while chunk c folds
            auto fold_chunk = [&](const uint32_t (&v)[32], int cb) {
              const int abs0 = row + cb;
              const int abs1 = min(abs0 + 32, tile_end);
              int pos = abs0;
              while (true) {
                const int seg_end = min(doc_end, abs1);
                if (seg_end > pos) {
                  if constexpr (kSmooth) {
                    if (pos == abs0 && seg_end == abs0 + 32) lse32_full(v, mm, ll, p.smooth_c);
                    else lse32_range(v, mm, ll, p.smooth_c, pos - abs0, seg_end - abs0);
                  } else if constexpr (kArgmax) {
                    argmax_fold(v, mm, ai, bc_slot, bc_sw, abs0 - doc_row0, pos - abs0, seg_end - abs0);
                  } else {
                    mm = max32_range(v, mm, pos - abs0, seg_end - abs0);
                  }
                }
                pos = seg_end;
                if (doc_end > abs1) break;  // the current document continues past this chunk
                finish_doc();
              }
            };
            uint32_t va[32], vb[32];
            tmem_ld_x32(taddr, va);
#pragma unroll 1
            for (int cb = 0; cb < n_valid; cb += 64) {
              tmem_ld_wait();
              reg_fence32(va);
              const bool has_b = cb + 32 < n_valid;
              if (has_b) tmem_ld_x32(taddr + cb + 32, vb);
              fold_chunk(va, cb);
              if (has_b) {
                tmem_ld_wait();
                reg_fence32(vb);
                if (cb + 64 < n_valid) tmem_ld_x32(taddr + cb + 64, va);
                fold_chunk(vb, cb + 32);
              }
            }
            release_acc();
          }
          if (dbg) {
            const long long t3 = clock64();
            if (dbg_me) {
            s_dbg[0] += static_cast<int>(t1 - t0);
            s_dbg[2] += static_cast<int>(t3 - t2);
            // max mode: boundary tiles apart; argmax / smooth: everything that is not a whole-tile fast path apart
            const bool other = (kMode == kModeMax) ? (path == 2) : (path != 1);
            if (other) { s_dbg[3] += static_cast<int>(t2 - t1); ++s_dbg[4]; } else { s_dbg[1] += static_cast<int>(t2 - t1); }
            ++s_dbg[5];
            }
          }
          if (r == 0) {
            m[0] = mm;
            am[0] = ai;
            ls[0] = ll;
          } else {
            m[R - 1] = mm;
            am[R - 1] = ai;
            ls[R - 1] = ll;
          }
          nxt = doc;
          nxt_row0 = doc_row0;
          nxt_end = doc_end;
          nxt_nlen = doc_nlen;
          nxt_ninit = doc_ninit;
        }
      }
      job += static_cast<uint32_t>(r_cnt);
      cur = nxt;
      cur_row0 = nxt_row0;
      cur_end = nxt_end;
      cur_nlen = nxt_nlen;
      cur_ninit = nxt_ninit;
    }
    if (p.balanced && !skip && cur < p.n_docs && cur_end != 0x7fffffff && cur_row0 < bal_r1 && cur_end > bal_r1) {
      // my last document continues in the next partition: combine with the neighbour's partial and emit it
      for (int r = 0; r < r_cnt; ++r) {
        if (kGroups == 2 && r != grp) continue;
        float mm = (r == 0) ? m[0] : m[R - 1];
        int ai = (r == 0) ? am[0] : am[R - 1];
        if constexpr (kArgmax)
          ai = argmax_resolve(bc_slot0 + static_cast<uint32_t>(r * kBcBytesPerTile), bc_sw, mm, ai);
        if (head_frag && cur == first_doc) {
          // (a document longer than a whole partition is excluded by the host: it would need a chain)
          __trap();
        }
        consume(r, mm, ai);
        finalize(cur, r, mm, ai, 0.f);
      }
    }
    d = run.e;
  }
  if (dbg && grp == 0 && quad == 2 && lane == 0) {  // epilogue: blocked on MMA / holding the accumulator / after release
    float* o = p.scores + 512 + 8 * blockIdx.x;
    o[2] = static_cast<float>(s_dbg[0]);
    o[3] = static_cast<float>(s_dbg[1]);
    o[4] = static_cast<float>(s_dbg[2]);
    o[5] = static_cast<float>(s_dbg[3]);
    o[6] = static_cast<float>(s_dbg[4]);
    o[7] = static_cast<float>(s_dbg[5]);  // jobs THIS warp folded (half of the CTA's with two epilogue groups)
  }
}

// Programmatic dependent launch: let the next kernel of the stream be scheduled as SMs free up, and (mode 1) wait for
// the previous kernel's completion + memory flush before this one touches global memory.  Mode 2 never waits: only for
// launches that do not consume anything the previous kernel wrote.
__device__ __forceinline__ void maxsim_pdl_entry(const MaxSimParams& p) {
  if (p.pdl != 0) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.pdl == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (p.wait_flags != nullptr) {  // fused all-gather: the peers must be done with the slab this launch overwrites
    if (static_cast<int>(threadIdx.x) < p.n_wait) {
      const uint32_t* f = p.wait_flags + threadIdx.x;
      const uint64_t t0 = global_timer_ns();
      const uint64_t limit = static_cast<uint64_t>(p.wait_timeout_ms) * 1000000ull;
      while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (static_cast<int32_t>(v - p.wait_value) >= 0) break;
        if (global_timer_ns() - t0 > limit) __trap();  // a peer is gone: proceeding would corrupt its unread results
        __nanosleep(32);
      }
    }
    __syncthreads();
  }
}

// End of the kernel, called by ALL threads of the CTA after their role code: makes the CTA's results visible, signals
// the fused all-gather's consumers, and lets the last CTA of the grid turn the score matrix into the loss.
// `group` / `part` / `q_first` / `q_count`: this CTA's query-tile group, document partition and queries.  With them
// (q_count >= 0: the dim-128 kernel) the fused loss runs per group -- the last CTA of every group turns the group's rows
// of the score matrix into their loss terms and gradients, the last group to finish folds the partial sums -- and the
// fused top-k can run; without them (K-pipelined and pair kernels) the last CTA of the grid does all rows.
// Counter workspace (d_done_counter, kLossWsWords words): word 0 = groups (or CTAs) done, words [1, 1 + G) = per-group
// CTA counters (G = q_groups < kLossWsGroups), and from word kLossWsGroups on 3 floats of partial (sum, min, max) per
// group.  The counters are zero between launches; the partial region is FIXED, not packed behind the G counters -- a
// launch with more groups would otherwise find an earlier launch's partial sums where it expects zeroed counters.
__device__ __forceinline__ void maxsim_finish(const MaxSimParams& p, const LossParams& lp, int cluster, int warp,
                                              int group = 0, int part = 0, int q_first = 0, int q_count = -1) {
  static_assert(kLossWsWords >= 4 * kLossWsGroups, "partials of every group fit behind the counters");
  __shared__ int s_last;
  const bool fused_loss = lp.loss != nullptr && p.done_counter != nullptr;
  const bool fused_topk = p.topk_scores != nullptr;
  if ((fused_loss || fused_topk) && warp >= 2) __threadfence();  // this warp's score stores are visible device-wide before the count
  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its shared memory or signal its barriers
  if (cluster > 1) cluster_sync_all(); else __syncthreads();
  if (p.peer_scores != nullptr && threadIdx.x == 0) {
    // fused all-gather completion.  Every score store of this CTA happened before the barrier above, so ONE release at
    // system scope (cumulative) publishes them; the flag word of every rank then grows by one per CTA.
    if (p.mc_base != 0) {
      asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(p.mc_base + 4ull * static_cast<uint64_t>(p.peer_flag_offset)), "r"(1u) : "memory");
    } else {
      for (int pr = 0; pr < p.n_peers; ++pr) {
        uint32_t* f = reinterpret_cast<uint32_t*>(__ldg(p.peer_scores + pr)) + p.peer_flag_offset;
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
      }
    }
  }
  if (fused_loss && q_count < 0) {
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(p.done_counter, 1u);
      s_last = (prev + 1u == gridDim.x) ? 1 : 0;
      if (s_last) *p.done_counter = 0u;  // ready for the next launch (stream ordered)
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      colbert_loss_body(lp);  // reads the [B, C] scores through L2 (__ldcg)
    }
  } else if (fused_loss && group < p.q_groups) {  // uniform over the CTA (groups past q_groups are cluster padding)
    uint32_t* group_ctr = p.done_counter + 1 + group;
    float* partials = reinterpret_cast<float*>(p.done_counter + kLossWsGroups);
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(group_ctr, 1u);
      s_last = (prev + 1u == static_cast<unsigned>(p.doc_parts)) ? 1 : 0;
      if (s_last) *group_ctr = 0u;
    }
    __syncthreads();
    if (s_last) {  // this group's rows of the score matrix are complete
      __threadfence();
      colbert_loss_rows_partial(lp, q_first, q_first + q_count, partials + 3 * group);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();  // partial + gradient rows before the count
        const unsigned prev = atomicAdd(p.done_counter, 1u);
        s_last = (prev + 1u == static_cast<unsigned>(p.q_groups)) ? 1 : 0;
        if (s_last) *p.done_counter = 0u;
      }
      __syncthreads();
      if (s_last && warp == 0) {  // every group has published its partial
        __threadfence();
        colbert_loss_combine(lp, partials, p.q_groups, threadIdx.x & 31);
      }
    }
  }
  if (fused_topk && group < p.q_groups)  // uniform over the CTA (groups past q_groups are cluster padding: no queries)
    topk_group_tail(p.scores, p.n_docs, p.topk_k, p.topk_scores, p.topk_idx, p.topk_ws, group, part, p.doc_parts, q_first,
                    q_count, static_cast<uint64_t>(p.wait_timeout_ms) * 1000000ull);
}

}  // namespace cpb
