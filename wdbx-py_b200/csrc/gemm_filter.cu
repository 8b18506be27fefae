// gemm_filter.cu -- K2b: single-pass bf16 tensor-core FILTER + exact fp32 REFINE (sm_100a only).
//
// The default search path (option c): every batch >= 16 and, on fp32 stores of
// 1 GB or more, every batch size.  Replaces FaissIndex.search (wdbx/core/indexing.py:1002-1024) + the shard
// merge of VectorStore.search (wdbx/core/vector_store.py:323-330); results are BIT-IDENTICAL to the
// streaming kernel K1 because the final scores are recomputed with K1's own fp32 arithmetic.
//
//   1. prep:    q -> bf16 (RNE), 1/|q|, |q|, |q|^2, padded to whole 128-query blocks; zeroes the per-search
//               state                                                              (prep_queries_kernel)
//   2. filter:  S~ = Qb . Xb^T on tcgen05 (kind::f16, bf16 operands, fp32 TMEM accumulators) over a
//               bf16 SHADOW of the stored matrix (half the HBM bytes of the fp32 rows).
//               |s~ - s| <= eps(row, query) is a RIGOROUS bound derived from the data (see "error bound"
//               below): eps = |r||q_b| + |x||t| + accumulation terms, r = x - bf16(x) and t = q - bf16(q) the
//               actual rounding residuals (|r| is stored per row next to the shadow).  Per query
//               ONE list of the k largest lower bounds (s~ - eps) seen by any thread of any CTA lives in
//               global memory (lock-free, see lower_insert); its minimum L never exceeds the exact k-th
//               best score, so every row with s~ + eps >= L is appended to a candidate region and
//               everything else provably cannot be in the exact top-k.
//               gemm_filter_small_kernel: B <= 16, X rows on the M side of the MMA, 16 queries on N;
//               gemm_filter_kernel<METRIC, NCTA>: 128-query tiles, NCTA = 2 = CTA pairs (cta_group::2).
//   3. refine:  candidates (a few hundred to a few thousand per query) are re-scored from the fp32 rows
//               with K1's lane mapping, summation order and score formula; register bitonic lists; several
//               CTAs per query for small batches, the last one folds the partial lists (refine_topk_kernel).
//   4. queries whose candidate region overflowed (adversarial data) are flagged and re-run by K1 itself.
//
// Filter CTA = 12 warps: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 idle | 4-11 epilogue;
// mbarrier ring of TMA stages, two accumulator tiles in TMEM so that the epilogue of tile t overlaps the
// MMAs of tile t+1.  DESIGN.md section 4 has the measurements behind every choice.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "scan_topk_kernel.cuh"
#include "tc05.cuh"

namespace wdbx {

namespace {

using namespace tc05;

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;        // bf16 elements per K block = 128 bytes
constexpr int STAGES_SINGLE = 4; // single-CTA ring (48 KB / stage)
constexpr int STAGES_PAIR = 6;   // CTA-pair ring (32 KB / stage per CTA)
constexpr int kThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr int SQ_I8 = 16;         // queries of the int8 small-batch operand (two digit rows each)
constexpr int kMaxKFilter = 128;  // shared lower-bound list = up to four 128-byte lines per query; lists = 1 or 4 keys per lane
constexpr uint32_t Q_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t TMEM_COLS = 2 * BN;

struct FilterParams {
  const float* inv_norm;   // [n_rows] 1/|x|  (exact, of the stored rows)
  const float* sqnorm;     // [n_rows] |x|^2
  const uint32_t* tomb;    // bitmap or NULL
  const float* rres;       // [n_rows] upper bound of |x - bf16(x)| (NULL for bf16 stores: the rows ARE the filter operand)
  const float* q_inv;      // [B]
  const float* q_nrm;      // [B]
  const float* q_sq;       // [B]
  const float* q_bn;       // [B] upper bound of |bf16(q)|
  const float* q_tn;       // [B] upper bound of |q - bf16(q)|
  long long n_rows;
  int B, k;
  int n_kblocks, n_tiles, n_slices;
  // (cap / slice_base / s_total below are shared by both filter kernels; the region index differs)
  int seg;                 // segment index stored with each candidate
  float acc_rel;           // accumulation terms of the bound, relative to |x||q| (see "error bound")
  float c_l2;              // l2 only: rounding of |x|^2, |q|^2 and of the expanded form, relative to |x|^2 + |q|^2
  unsigned long long* cand;   // [B][s_total][cap]  (seg << 32 | row): one private region per (query, row slice)
  unsigned int* cand_count;   // [B][s_total]
  unsigned int* lower_glob;   // [B] monotone-mapped float: best known lower bound of the exact k-th score
  unsigned int* lower_list;   // [B][kMaxKFilter] monotone-mapped lower bounds of the k best rows seen by ANY CTA
  int cap;                    // entries per (query, slice) region
  int slice_base, s_total;    // this launch fills slices [slice_base, slice_base + n_slices) of s_total
  unsigned int* tile_ctr;     // small-batch kernel: the launch's dynamic tile counter (zero at launch)
  unsigned int* part_max;     // small-batch kernel: [B][kMaxKFilter] best lower bound per warp PARTITION (see bound server)
  const float* rowscale;      // int8 shadow (small-batch kernel): x ~ rowscale[row] * xi  (NULL: bf16 operand)
  const float* q_s1;          // int8 queries: q ~ q_s1[j] * q1 + q_s2[j] * q2  (two int8 digits per element)
  const float* q_s2;
  const unsigned int* prep_count;   // overlap mode: CTAs of prep kernels finished on this workspace (monotone) ...
  unsigned int prep_target;         // ... and the count at which THIS search's prep is complete; NULL = griddepcontrol.wait
  int trace;                  // WDBX_B200_FILTER_TRACE=1: every CTA prints its phase timestamps (small-batch kernel)
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- error bound
// The filter must never drop a row that the exact path (K1's fp32 arithmetic) would rank in the top-k, so
// every filter score s~ carries eps(row, query) with |s~ - s_K1| <= eps.  With x_b = bf16(x) (RNE), q_b =
// bf16(q) and the ACTUAL residuals r = x - x_b, t = q - q_b (both exact in fp32):
//     x.q = x_b.q_b + r.q_b + x.t      =>      |x.q - x_b.q_b| <= |r||q_b| + |x||t|          (Cauchy-Schwarz)
// which holds for ANY data -- coherent rounding (every element rounding the same way, few-level /
// quantised embeddings) simply shows up as a larger |r| -- and is ~3.5x tighter on ordinary data than the
// worst-case 2 * 2^-8 |x||q| (bf16 has 8 significand bits: unit round-off 2^-8 per rounded operand).
// |r| is computed once per row when the shadow is built (shadow_rows_kernel, `rres`), |q_b| and |t| per
// query by prep_queries_kernel; all three are inflated by 1e-4, and A / C below by another 1e-3, which
// covers the fp32 rounding of the norms themselves.  On top of that:
//   * the tensor core accumulates the (exact) bf16 x bf16 products in fp32, possibly truncating: at most
//     dim * 2^-23 relative to sum|x_b,i q_b,i| <= 1.01 |x||q|;
//   * K1's own fp32 dot product (what the final ranking uses) is within (dpad/32 + 40) * 2^-24 |x||q| of the
//     exact one (fma chain per lane + butterfly), its score formula adds two roundings;
// both folded into acc_rel (host: filter_acc_rel).  Per metric, relative quantities for cosine:
//     cosine  eps = rho_x * A_q + C_q,            rho_x = |r|/|x|,  A_q = 1.001 |q_b|/|q|,  C_q = 1.001 |t|/|q| + acc_rel
//     ip      eps = |r| A'_q + |x| C'_q,          A'_q = 1.001 |q_b|,  C'_q = 1.001 (|t| + acc_rel |q|)
//     l2      eps = 2 (|r| A'_q + |x| C'_q) + c_l2 (|x|^2 + |q|^2)      (expanded form -(|x|^2 - 2 x.q + |q|^2):
//             c_l2 covers the fp32 rounding of the stored |x|^2, of |q|^2, of the three additions, and of K1's
//             direct sum of squared differences)
// bf16 stores: the stored rows are the filter operand, r = 0.  tests/test_gpu_filter.py::test_bound_is_rigorous
// holds the adversarial cases (all elements just above a bf16 midpoint, etc.).
struct QueryBound {
  float A, C;           // see above (cosine: relative; ip / l2: absolute)
  float qinv, qnrm, qsq;
};
template <int METRIC>
__device__ __forceinline__ QueryBound make_query_bound(float qinv, float qnrm, float qsq, float qbn, float qtn, float acc_rel) {
  QueryBound b;
  b.qinv = qinv; b.qnrm = qnrm; b.qsq = qsq;
  if (METRIC == kCosine) {
    b.A = 1.001f * qbn * qinv;
    b.C = fmaf(1.001f * qtn, qinv, acc_rel);
  } else {
    b.A = 1.001f * qbn;
    b.C = 1.001f * fmaf(acc_rel, qnrm, qtn);
  }
  return b;
}
// exact-form score estimate sv and its error bound eps for one (row, query): d = bf16 dot from TMEM,
// inx = 1/|x| (cosine), sq = |x|^2 (ip / l2), rr = |r| bound of the row
template <int METRIC>
__device__ __forceinline__ void bound_eval(float d, float inx, float sq, float rr, const QueryBound& b, float c_l2,
                                           float& sv, float& eps) {
  if (METRIC == kCosine) {
    sv = d * inx * b.qinv;
    eps = fmaf(rr * inx, b.A, b.C);
  } else {
    const float xn = sqrtf(sq) * 1.00001f;
    const float e = fmaf(rr, b.A, xn * b.C);
    if (METRIC == kL2) {
      sv = -((sq - 2.0f * d) + b.qsq);
      eps = fmaf(2.0f, e, c_l2 * (sq + b.qsq));
    } else {
      sv = d;
      eps = e;
    }
  }
}

// ---------------------------------------------------------------- prep: queries -> bf16 + norms
// (rows B .. Bpad-1 are zero padding up to a whole 128-query block, so that the TMA box of the query
// tile never leaves the tensor: a mostly out-of-bounds box measurably slows the load pipeline)
// Bounded wait (~3 s) until search number `sn` of this workspace has completed (see engine.cu, "overlapping
// consecutive searches"); sn == 0: nothing to wait for.
__device__ __forceinline__ void wait_search_done(const unsigned int* done_ctr, unsigned int sn) {
  if (done_ctr == nullptr || sn == 0u) return;
  const long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done_ctr) : "memory");
    if (v >= sn) break;
    if (clock64() - t0 > 6000000000ll) break;
    __nanosleep(200);
  } while (true);
}

__global__ void prep_queries_kernel(const float* __restrict__ q, int B, int Bpad, int dim, int ld,
                                    __nv_bfloat16* __restrict__ qb, float* __restrict__ q_inv, float* __restrict__ q_nrm,
                                    float* __restrict__ q_sq, float* __restrict__ q_bn, float* __restrict__ q_tn,
                                    unsigned int* __restrict__ zero, size_t n_zero, const unsigned int* done_ctr,
                                    unsigned int wait_sn, int pdl_wait, unsigned int* prep_count) {
  // programmatic dependent launch: the filter kernel behind us may start its set-up now.  The buffers rewritten
  // below are the ones search `wait_sn` used (per-search state is double-buffered on the fused small-batch path): wait for
  // that search by NUMBER rather than for the whole launch before us, so that this kernel -- and the filter behind it
  // -- can run while the previous search still finishes its tail.  Other paths (pdl_wait) wait for the stream.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) wait_search_done(done_ctr, wait_sn);
  __syncthreads();
  // per-search state of the filter (candidate counts, flags, tickets, lower-bound lists) starts at zero
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_zero;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    zero[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b < Bpad) {
    float ss = 0.0f, sb = 0.0f, st = 0.0f;
    for (int c = lane; c < ld; c += 32) {
      const float v = (c < dim && b < B) ? q[static_cast<size_t>(b) * dim + c] : 0.0f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      qb[static_cast<size_t>(b) * ld + c] = h;
      const float hb = __bfloat162float(h);
      const float t = v - hb;              // exact
      ss = fmaf(v, v, ss);
      sb = fmaf(hb, hb, sb);
      st = fmaf(t, t, st);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(FULL_MASK, ss, o);
      sb += __shfl_xor_sync(FULL_MASK, sb, o);
      st += __shfl_xor_sync(FULL_MASK, st, o);
    }
    if (lane == 0) {
      q_sq[b] = ss;
      q_nrm[b] = sqrtf(ss);
      q_inv[b] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
      q_bn[b] = sqrtf(sb) * 1.0001f;
      q_tn[b] = sqrtf(st) * 1.0001f;
    }
  }
  // overlap mode: the filter kernel behind us does not wait for this LAUNCH (griddepcontrol.wait would also wait for
  // everything before it on the stream) but for this count: every CTA reports once its writes are visible
  if (prep_count != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(prep_count, 1u);
    }
  }
}

// fp32 stored rows -> bf16 shadow rows (RNE) + the norm of the rounding residual |x - bf16(x)| (inflated by
// 1e-4: an upper bound whatever the fp32 summation does), one warp per row
__global__ void shadow_rows_kernel(const float* __restrict__ rows, long long n, int dpad, int ld16,
                                   __nv_bfloat16* __restrict__ dst, float* __restrict__ rres) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* s = rows + r * dpad;
  __nv_bfloat16* d = dst + r * ld16;
  float sr = 0.0f;
  for (int c = lane; c < ld16; c += 32) {
    const float v = c < dpad ? s[c] : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    d[c] = h;
    const float t = v - __bfloat162float(h);   // exact (NaN / inf rows give NaN: such a row is always a candidate)
    sr = fmaf(t, t, sr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sr += __shfl_xor_sync(FULL_MASK, sr, o);
  if (lane == 0) rres[r] = sqrtf(sr) * 1.0001f;
}

// ---------------------------------------------------------------- int8 operands (small-batch kernel)
// A 1-byte shadow halves the bytes a small-batch search streams once more (10M x 768: 7.7 GB instead of 15.4).
// Row: x ~ sx * xi, xi = rint(x / sx) in [-127, 127], sx = max|x| / 127 -- symmetric per-row scaling, whose
// residual on Gaussian-like rows is ~0.8 % of |x| (e4m3 would be 3.6 %).  Query: TWO int8 digits per element,
// q ~ s1 * q1 + s2 * q2 with s2 = s1 / 254 (residual ~3e-5 |q|); the digits are two rows of the N = 32 operand, so
// one kind::i8 MMA (exact int32 accumulation, SASS UTCIMMA) produces both partial dot products.  The bound is the
// same data-derived one as for bf16 operands: x.q = x~.q~ + r.q~ + x.t with the ACTUAL residual norms |r| (stored
// per row, rres8) and |t| -- heavy-tailed rows (one huge element eats the scale) simply carry a larger |r|.
template <bool SRC_BF16>
__global__ void shadow8_rows_kernel(const void* __restrict__ rows_raw, long long n, int dpad, int ld8,
                                    signed char* __restrict__ dst, float* __restrict__ sx_out, float* __restrict__ rres) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  // the STORED row (fp32, or bf16 for bf16 stores: the residual is measured against what the exact path scores)
  struct Row {
    const void* p;
    __device__ float operator[](int c) const {
      if (SRC_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[c]);
      return static_cast<const float*>(p)[c];
    }
  };
  const Row s{static_cast<const unsigned char*>(rows_raw) + static_cast<size_t>(r) * dpad * (SRC_BF16 ? 2 : 4)};
  signed char* d = dst + r * ld8;
  float mx = 0.0f;
  bool bad = false;
  for (int c = lane; c < dpad; c += 32) {
    const float v = s[c];
    bad = bad || !(fabsf(v) <= 3.0e38f);     // NaN / inf
    mx = fmaxf(mx, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
  bad = __any_sync(FULL_MASK, bad);
  const float sx = mx > 0.0f ? mx / 127.0f : 0.0f;
  const float rinv = mx > 0.0f ? 127.0f / mx : 0.0f;
  float sr = 0.0f;
  for (int c = lane; c < ld8; c += 32) {
    const float v = c < dpad ? s[c] : 0.0f;
    float xi = rintf(v * rinv);
    xi = fminf(fmaxf(xi, -127.0f), 127.0f);
    if (bad) xi = 0.0f;
    d[c] = static_cast<signed char>(static_cast<int>(xi));
    const float t = fmaf(-sx, xi, v);   // ONE rounding of the exact residual (an unfused sx * xi could round it to 0)
    sr = fmaf(t, t, sr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sr += __shfl_xor_sync(FULL_MASK, sr, o);
  if (lane == 0) {
    sx_out[r] = bad ? 0.0f : sx;
    // 1.001: fp32 rounding of sx * xi and of the sum; NaN / inf rows get NaN: such a row is always a candidate
    rres[r] = bad ? __int_as_float(0x7fc00000) : sqrtf(sr) * 1.001f;
  }
}

// queries -> two int8 digits per element (rows j and 16 + j of a [32][ld8] operand) + norms and scales
__global__ void prep_queries_i8_kernel(const float* __restrict__ q, int B, int dim, int ld8, signed char* __restrict__ qi,
                                       float* __restrict__ q_inv, float* __restrict__ q_nrm, float* __restrict__ q_sq,
                                       float* __restrict__ q_bn, float* __restrict__ q_tn, float* __restrict__ q_s1,
                                       float* __restrict__ q_s2, unsigned int* __restrict__ zero, size_t n_zero,
                                       const unsigned int* done_ctr, unsigned int wait_sn, int pdl_wait,
                                       unsigned int* prep_count) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) wait_search_done(done_ctr, wait_sn);
  __syncthreads();
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_zero;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    zero[i] = 0u;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b < SQ_I8) {
    const bool live = b < B;
    float mx = 0.0f;
    for (int c = lane; c < dim; c += 32) mx = fmaxf(mx, live ? fabsf(q[static_cast<size_t>(b) * dim + c]) : 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
    const bool ok = mx > 0.0f && mx <= 3.0e38f;    // zero / NaN / inf queries: all digits 0, t = q (an honest, huge bound)
    const float s1 = ok ? mx / 127.0f : 0.0f, r1 = ok ? 127.0f / mx : 0.0f;
    const float s2 = s1 / 254.0f, r2 = ok ? 254.0f * r1 : 0.0f;
    float ss = 0.0f, sb = 0.0f, st = 0.0f;
    for (int c = lane; c < ld8; c += 32) {
      const float v = (c < dim && live) ? q[static_cast<size_t>(b) * dim + c] : 0.0f;
      float d1 = fminf(fmaxf(rintf(v * r1), -127.0f), 127.0f);
      const float e1 = v - s1 * d1;
      float d2 = fminf(fmaxf(rintf(e1 * r2), -127.0f), 127.0f);
      if (!ok) { d1 = 0.0f; d2 = 0.0f; }
      qi[static_cast<size_t>(b) * ld8 + c] = static_cast<signed char>(static_cast<int>(d1));
      qi[static_cast<size_t>(SQ_I8 + b) * ld8 + c] = static_cast<signed char>(static_cast<int>(d2));
      const float qt = fmaf(s2, d2, s1 * d1);      // q~ (fp32 rounding of it is covered by the 1.01 below)
      const float t = v - qt;
      ss = fmaf(v, v, ss);
      sb = fmaf(qt, qt, sb);
      st = fmaf(t, t, st);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(FULL_MASK, ss, o);
      sb += __shfl_xor_sync(FULL_MASK, sb, o);
      st += __shfl_xor_sync(FULL_MASK, st, o);
    }
    if (lane == 0) {
      q_sq[b] = ss;
      q_nrm[b] = sqrtf(ss);
      q_inv[b] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
      q_bn[b] = sqrtf(sb) * 1.0001f;
      q_tn[b] = sqrtf(st) * 1.01f + 2e-7f * sqrtf(ss);
      q_s1[b] = s1;
      q_s2[b] = s2;
    }
  }
  if (prep_count != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(prep_count, 1u);
    }
  }
}

// ---------------------------------------------------------------- filter kernel
// Per-query list of the k largest LOWER bounds (s~ - eps) seen so far by any thread of any CTA, kept in
// global memory and updated lock-free: a new bound replaces the current minimum by compare-and-swap.
// Every slot always holds the lower bound of a distinct row (or 0 = empty) and slots only grow, so the
// minimum over the k slots -- even over a stale snapshot -- never exceeds the exact k-th best score.
// That minimum is published through lower_glob (atomicMax) for the per-tile refresh of every thread.
// Sharing ONE list (instead of one list per thread, whose k-th best only reflects 1/36 of the rows)
// cuts the candidates per query roughly by the number of threads that share the query.
__device__ __forceinline__ float lower_insert(unsigned int* slots, unsigned int* glob, int k, float lo, float L) {
  const unsigned int m = mono_u32(lo);
  unsigned int result = 0u;
  for (int attempt = 0; attempt < 8; ++attempt) {
    unsigned int vmin = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
    int imin = 0;
    // 32 slots (eight 16-byte loads) in flight per round: one L2 round trip for k <= 32, four for k = 128
    for (int i0 = 0; i0 < k; i0 += 32) {
      uint4 v4[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v4[j] = (i0 + 4 * j < k) ? __ldcg(reinterpret_cast<const uint4*>(slots + i0 + 4 * j))
                                 : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned int vv[4] = {v4[j].x, v4[j].y, v4[j].z, v4[j].w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = i0 + 4 * j + jj;
          if (i < k) {
            const unsigned int v = vv[jj];
            if (v < vmin) { second = vmin; vmin = v; imin = i; }
            else if (v < second) second = v;
          }
        }
      }
    }
    if (m <= vmin) { result = vmin; break; }           // not among the k best bounds (any more)
    if (atomicCAS(slots + imin, vmin, m) == vmin) {
      result = second < m ? second : m;                // minimum of my snapshot after the replacement
      if (k == 1) result = m;
      if (result > 0x007FFFFFu) atomicMax(glob, result);
      break;
    }
  }
  const float g = unmono_f32(max(result, 0x007FFFFFu));
  return g > L ? g : L;
}

// Cooperative form for the small-batch kernel's BOUND SERVER (warp 3): up to 8 offers (lower bounds of 8 distinct
// rows, one per lane 0..7 of a lane GROUP, 0 = none) enter one query's shared list in one go.  A group is GL
// lanes: 8 for k <= 32 (4 slots per lane; four queries are served side by side by one warp), 32 for k <= 128.
// One L2 round trip snapshots all slots, the i-th largest offer is paired with the i-th smallest slot and
// compare-and-swapped in parallel; offers that lost a race retry against a fresh snapshot.  Ties between equal
// slots (the empty list at the first tile!) are broken by a per-CTA rotation `rot`, so that the 148 servers do not
// all aim at the same slots.  Returns the minimum of the (stale) snapshot with this group's successful writes
// applied -- a valid lower bound of the exact k-th best score once every slot is filled (0 otherwise) -- and
// publishes it through `glob`.  `slots` / `glob` / `k_valid` are per group (k_valid = 0: the group idles); offers
// still unplaced after `max_rounds` stay in `o` (the bound server retries them on its next sweep).
template <int GL>
__device__ __forceinline__ unsigned int list_multi_insert(unsigned int* slots, unsigned int* glob, int k_valid, unsigned int& o,
                                                          int lane, unsigned int rot, int max_rounds) {
  const int gl = lane & (GL - 1);                       // lane inside the group
  const unsigned gmask = GL == 32 ? FULL_MASK : (0xFFu << (lane & ~7));
  const int gbase = lane & ~(GL - 1);
  const int k = k_valid;
  unsigned int result = 0u;
  for (int round = 0; round < max_rounds; ++round) {
    const unsigned offering = __ballot_sync(FULL_MASK, gl < 8 && o != 0u && k > 0);
    if (offering == 0u) break;                           // warp-uniform: every group is done
    const int n_off = __popc(offering & gmask);
    uint4 v4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    if (4 * gl < k && n_off > 0) v4 = __ldcg(reinterpret_cast<const uint4*>(slots + 4 * gl));
    unsigned int v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (4 * gl + jj >= k) v[jj] = 0xFFFFFFFFu;
    // rank of this lane's offer among the group's offers, largest first (ties by lane)
    int rank = 0;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const unsigned int ol = __shfl_sync(FULL_MASK, o, gbase + l);
      if (ol > o || (ol == o && l < gl)) ++rank;
    }
    // the smallest slots, smallest first; the lane whose offer has rank r takes the r-th
    unsigned int w[4] = {v[0], v[1], v[2], v[3]};
    unsigned int tv = 0xFFFFFFFFu;
    int ti = -1;
    for (int r = 0; r < 8; ++r) {                        // (groups with fewer offers just compute a few unused targets)
      const unsigned int lm = min(min(w[0], w[1]), min(w[2], w[3]));
      const unsigned int gm = __reduce_min_sync(gmask, lm);
      // owner among the lanes that hold the minimum: first one at or after the rotation point
      const unsigned eq = (__ballot_sync(FULL_MASK, lm == gm) & gmask) >> gbase;          // GL-bit field
      const unsigned r0 = rot & (GL - 1);
      const unsigned hi = eq >> r0;
      const int owner = hi ? static_cast<int>(r0) + __ffs(hi) - 1 : __ffs(eq) - 1;
      int idx = -1;
      if (gl == owner) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (idx < 0 && w[jj] == gm) { idx = 4 * gl + jj; w[jj] = 0xFFFFFFFFu; }
      }
      idx = __shfl_sync(FULL_MASK, idx, gbase + owner);
      if (gl < 8 && o != 0u && rank == r) { tv = gm; ti = idx; }
    }
    bool placed = false;
    if (gl < 8 && o != 0u && k > 0) {
      if (ti < 0 || tv == 0xFFFFFFFFu || o <= tv) o = 0u;            // not among the k best bounds of list + offers
      else placed = atomicCAS(slots + ti, tv, o) == tv;               // lost a race: retry on the next snapshot
    }
    // apply this group's successful writes to the snapshot
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const bool pl = __shfl_sync(FULL_MASK, placed ? 1 : 0, gbase + l) != 0;
      const int pi = __shfl_sync(FULL_MASK, ti, gbase + l);
      const unsigned int pv = __shfl_sync(FULL_MASK, o, gbase + l);
      if (pl && (pi >> 2) == gl) v[pi & 3] = pv;
    }
    if (placed) o = 0u;
    const unsigned int lm = min(min(v[0], v[1]), min(v[2], v[3]));
    const unsigned int gm = __reduce_min_sync(gmask, lm);
    if (k > 0 && gm > 0x007FFFFFu && gm != 0xFFFFFFFFu && gm > result) result = gm;
  }
  if (result > 0x007FFFFFu && gl == 0) atomicMax(glob, result);
  return result;
}

// NCTA = 1: one CTA per 128-query block.  NCTA = 2: a CTA PAIR (cluster of two SMs of one TPC, cta_group::2)
// owns two adjacent query blocks and one row slice: each CTA stages its own 128 queries and HALF of the
// 256-row X tile, the leader issues one M=256 MMA that reads both halves, each CTA's TMEM receives the
// accumulator of its own queries -- the X operand crosses L2 -> shared memory once per pair instead of
// once per CTA (2/3 of the single-CTA staging traffic per MAC).
template <int METRIC, int NCTA>
__global__ void __launch_bounds__(kThreads, 1)
gemm_filter_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
                   const FilterParams p) {
  constexpr int STAGES = (NCTA == 2) ? STAGES_PAIR : STAGES_SINGLE;
  constexpr int X_ROWS = BN / NCTA;                      // X rows staged by this CTA per tile
  constexpr uint32_t X_TILE_BYTES = X_ROWS * BK * 2;
  constexpr uint32_t STAGE_BYTES = X_TILE_BYTES + Q_TILE_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // keep the shared-memory address space visible to the compiler (LDS/STS instead of generic accesses)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stage_base = smem;
  float* colscale = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);    // [2][BN] score scale per column
  float* coleps = colscale + 2 * BN;                                          // [2][BN] |x| term of eps per column
  float* colres = coleps + 2 * BN;                                            // [2][BN] residual term of eps per column
  uint64_t* bars = reinterpret_cast<uint64_t*>(colres + 2 * BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, slice = blockIdx.y;
  const uint32_t cta_rank = (NCTA == 2) ? cluster_ctarank() : 0u;   // cluster = (2,1,1): rank == qb & 1
  const int k = p.k;
  const float NEG_INF = __int_as_float(0xff800000);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_q);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tmem_full + a), 1);
      mbar_init(smem_u32(tmem_empty + a), 8 * NCTA);  // epilogue warps of the whole pair
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (NCTA == 2) tmem_alloc_pair(smem_u32(tmem_ptr), TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_ptr), TMEM_COLS);
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();   // the peer's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int my_tiles = (p.n_tiles - slice + p.n_slices - 1) / p.n_slices;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the ring (uniform control flow), one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int row0 = (slice + t * p.n_slices) * BN;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
        const uint32_t sb = smem_u32(stage_base) + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t bar = smem_u32(full_bar) + static_cast<uint32_t>(stage) * 8u;
        if (elect_one()) {
          if (NCTA == 2) {
            // both CTAs' bytes are credited to the leader's barrier (the only one the MMA issuer waits on)
            if (cta_rank == 0) mbar_expect_tx(bar, 2 * STAGE_BYTES);
            tma_load_2d_pair(sb, &tm_x, kb * BK, row0 + static_cast<int>(cta_rank) * X_ROWS, bar);
            tma_load_2d_pair(sb + X_TILE_BYTES, &tm_q, kb * BK, qb * BM, bar);
          } else {
            mbar_expect_tx(bar, STAGE_BYTES);
            tma_load_2d(sb, &tm_x, kb * BK, row0, bar);
            tma_load_2d(sb + X_TILE_BYTES, &tm_q, kb * BK, qb * BM, bar);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only in pair mode): uniform loop, one elected lane issues
    if (cta_rank == 0) {
      const uint32_t idesc = make_idesc(BM * NCTA, BN, 1u);  // BF16 x BF16 -> F32
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int a = t & 1;
        mbar_wait(smem_u32(tmem_empty + a), ((static_cast<uint32_t>(t) >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * BN);
        for (int kb = 0; kb < p.n_kblocks; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(stage_base) + static_cast<uint32_t>(stage) * STAGE_BYTES;
          const uint64_t d_x = make_desc_kmajor(sb, 128, 2);
          const uint64_t d_q = make_desc_kmajor(sb + X_TILE_BYTES, 128, 2);
          const uint32_t ebar = smem_u32(empty_bar) + static_cast<uint32_t>(stage) * 8u;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              const uint64_t adv = static_cast<uint64_t>((kk * 16 * 2) >> 4);  // 32 bytes per K=16 step
              if (NCTA == 2) umma_f16_pair(d_tmem, d_q + adv, d_x + adv, idesc, (kb | kk) ? 1u : 0u);
              else umma_f16(d_tmem, d_q + adv, d_x + adv, idesc, (kb | kk) ? 1u : 0u);
            }
            if (NCTA == 2) umma_commit_pair(ebar);   // frees the stage in both CTAs
            else umma_commit(ebar);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) {
          if (NCTA == 2) umma_commit_pair(smem_u32(tmem_full + a));
          else umma_commit(smem_u32(tmem_full + a));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps.  TMEM lane (= query) quarter = warp & 3; warps 4-7 take columns [0,128)
    // of every tile, warps 8-11 columns [128,256): two threads per query, each with its own list of
    // lower bounds (both are valid bounds) and its own candidate region.
    const int et = (warp & 3) * 32 + lane;      // TMEM lane == query inside the block
    const int half = (warp - 4) >> 2;           // column half of the tile
    const int q = qb * BM + et;
    const bool q_valid = q < p.B;
    const QueryBound qb_ = make_query_bound<METRIC>(q_valid ? p.q_inv[q] : 0.0f, q_valid ? p.q_nrm[q] : 0.0f,
                                                    q_valid ? p.q_sq[q] : 0.0f, q_valid ? p.q_bn[q] : 0.0f,
                                                    q_valid ? p.q_tn[q] : 0.0f, p.acc_rel);
    const float qnrm = qb_.qnrm, qsq = qb_.qsq;
    const float c_l2 = p.c_l2;
    float L = NEG_INF;  // best known lower bound of this query's exact k-th best score (see lower_insert)
    unsigned int* my_slots = p.lower_list + static_cast<size_t>(q_valid ? q : 0) * kMaxKFilter;
    // FAST TEST.  "upper bound of the exact score >= L" (bound_eval: sv + eps >= L) is rewritten so that it costs
    // two or three fused multiply-adds and one compare per accumulator element, against a per-thread
    // threshold T that only changes when L does:
    //   cosine  d/|x| + rho_x (A_q |q|) >= (L - C_q)|q|        (the inequality multiplied by |q|: no division)
    //   ip      d + |r| A'_q + |x| C'_q >= L
    //   l2      2d + 2|r| A'_q + 2|x| C'_q - (1 - c_l2)|x|^2 >= L + (1 - c_l2)|q|^2
    // T carries a 1e-6 relative slack (and the staged |x|^2 is shrunk by 1e-6) so that the fast test can
    // only ADMIT more elements than the reference form; admitted groups are re-tested below with the
    // reference form (bound_eval), which alone decides what becomes a candidate.
    const float fA = (METRIC == kCosine) ? qb_.A * qnrm : qb_.A;
    const float fC = qb_.C;
    auto fast_thr = [&](float l) -> float {
      if (METRIC == kCosine) {
        const float t = (l - qb_.C) * qnrm;
        return t - 1e-6f * (fabsf(t) + qnrm);
      }
      if (METRIC == kL2) return (l + qsq * (1.0f - c_l2)) - 1e-6f * (fabsf(l) + 2.0f * qsq);
      return l - 1e-6f * fabsf(l);
    };
    float T = fast_thr(L);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    // private candidate region of this (query, slice, half): plain stores, no atomics on the hot path
    const size_t region = static_cast<size_t>(q_valid ? q : 0) * p.s_total + 2 * (p.slice_base + slice) + half;
    unsigned long long* my_cand = p.cand + region * p.cap;
    unsigned int n_cand = 0;
    // per-column terms of column half*128 + et of tile t, staged for the fast test:
    //   cosine  cs = 1/|x|                       cr = rho_x = |r|/|x|
    //   ip      ce = |x|                         cr = |r|
    //   l2      cs = (1 - c_l2)(1 - 1e-6)|x|^2   ce = 2|x|    cr = 2|r|
    // columns past the end get ce = -inf (cosine: all zero) and are rejected by the reference path's range
    // check anyway.  The loads for tile t+1 (and the shared bound) are issued BEFORE tile t is processed, so
    // their latency never sits on the epilogue's critical path.
    auto load_scale = [&](int t, float& sc, float& ep, float& rs) {
      const long long row = static_cast<long long>(slice + t * p.n_slices) * BN + half * 128 + et;
      sc = 0.0f;
      ep = (METRIC == kCosine) ? 0.0f : NEG_INF;
      rs = 0.0f;
      if (t < my_tiles && row < p.n_rows) {
        const float rr = p.rres != nullptr ? __ldg(p.rres + row) : 0.0f;
        if (METRIC == kCosine) {
          const float inx = __ldg(p.inv_norm + row);
          sc = inx;
          rs = rr * inx;
        } else {
          const float sq = __ldg(p.sqnorm + row);
          const float xn = sqrtf(sq) * 1.00001f;
          sc = (METRIC == kL2) ? sq * ((1.0f - c_l2) * (1.0f - 1e-6f)) : 0.0f;
          ep = (METRIC == kL2) ? 2.0f * xn : xn;
          rs = (METRIC == kL2) ? 2.0f * rr : rr;
        }
      }
    };
    float sc_next, ep_next, rs_next;
    load_scale(0, sc_next, ep_next, rs_next);
    unsigned int glob_next = q_valid ? __ldcg(p.lower_glob + q) : 0u;
    for (int t = 0; t < my_tiles; ++t) {
      const int a = t & 1;
      const long long row0 = static_cast<long long>(slice + t * p.n_slices) * BN;
      const long long rem = p.n_rows - row0;
      const int valid = rem < BN ? static_cast<int>(rem) : BN;
      float* cs = colscale + a * BN;
      float* ce = coleps + a * BN;
      float* cr = colres + a * BN;
      cs[half * 128 + et] = sc_next;   // buffer a was last read for tile t-2: every thread has passed the
      ce[half * 128 + et] = ep_next;   // barrier of tile t-1 since
      cr[half * 128 + et] = rs_next;
      // share the bound: any thread's k-th best lower bound is a valid global lower bound
      if (q_valid) {
        const float g = unmono_f32(max(glob_next, 0x007FFFFFu));
        if (g > L) { L = g; T = fast_thr(L); }
        glob_next = __ldcg(p.lower_glob + q);
      }
      load_scale(t + 1, sc_next, ep_next, rs_next);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(smem_u32(tmem_full + a), (static_cast<uint32_t>(t) >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
        if (c0 >= valid) break;                       // tile-uniform
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(lane_base + static_cast<uint32_t>(a * BN + c0), r);
        tmem_ld_wait();
        unsigned gmask = 0;                           // bit g: some column of group [4g, 4g+4) passed the fast test
        if (q_valid) {
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + c0 + 4 * g4);
            const float4 r4 = *reinterpret_cast<const float4*>(cr + c0 + 4 * g4);
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
            float u[4];
            if (METRIC == kCosine) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) u[jj] = fmaf(rr[jj], fA, __uint_as_float(r[4 * g4 + jj]) * cc[jj]);
            } else {
              const float4 e4 = *reinterpret_cast<const float4*>(ce + c0 + 4 * g4);
              const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float d = __uint_as_float(r[4 * g4 + jj]);
                const float e = fmaf(rr[jj], fA, ee[jj] * fC);
                if (METRIC == kL2) u[jj] = fmaf(2.0f, d, e - cc[jj]);
                else u[jj] = d + e;
              }
            }
            // !(x < T) also admits NaN (the refine ranks it like K1)
            if (!(u[0] < T) || !(u[1] < T) || !(u[2] < T) || !(u[3] < T)) gmask |= 1u << g4;
          }
        }
        // rare path, warp-uniform: groups admitted by any lane are re-read from TMEM (4 columns) and
        // decided with the reference form of the bound
        unsigned um = __reduce_or_sync(FULL_MASK, gmask);
        while (um) {
          const int g4 = __ffs(um) - 1;
          um &= um - 1;
          uint32_t v[4];
          tmem_ld4(lane_base + static_cast<uint32_t>(a * BN + c0 + 4 * g4), v);
          tmem_ld_wait();
          if ((gmask >> g4) & 1u) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int col = c0 + 4 * g4 + jj;
              if (col >= valid) continue;
              const long long row = row0 + col;
              const float d = __uint_as_float(v[jj]);
              const float rrow = p.rres != nullptr ? __ldg(p.rres + row) : 0.0f;
              float sv, ej;
              bound_eval<METRIC>(d, METRIC == kCosine ? cs[col] : 0.0f, METRIC == kCosine ? 0.0f : __ldg(p.sqnorm + row),
                                 rrow, qb_, c_l2, sv, ej);
              const float up = sv + ej;  // upper bound of the exact score
              if (!(up < L)) {
                bool dead = false;
                if (p.tomb != nullptr) dead = (__ldg(p.tomb + (row >> 5)) >> (row & 31)) & 1u;
                if (!dead) {
                  WDBX_ASSERT(row >= 0 && row < p.n_rows && region < static_cast<size_t>(p.B) * p.s_total);
                  WDBX_ASSERT(static_cast<uint32_t>(a * BN + c0 + 32) <= TMEM_COLS);
                  if (n_cand < static_cast<unsigned int>(p.cap))
                    my_cand[n_cand] = (static_cast<unsigned long long>(p.seg) << 32) | static_cast<unsigned long long>(row);
                  ++n_cand;
                  const float lo = sv - ej;  // NaN never enters the list of lower bounds
                  if (lo > L) {
                    const float nl = lower_insert(my_slots, p.lower_glob + q, k, lo, L);
                    if (nl > L) { L = nl; T = fast_thr(L); }
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(smem_u32(tmem_empty + a), 0u);   // the leader's MMA issuer waits
        else mbar_arrive(smem_u32(tmem_empty + a));
      }
    }
    if (q_valid) p.cand_count[region] = n_cand;
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();   // the peer may still be reading TMEM / signalling our barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- filter kernel, small batches (B <= 16)
// Same filter with the operands swapped: the X tile is the M side (two 128-row MMAs per 256-row tile), the
// queries are the N = 16 side.  For a handful of queries the 128-query tile above spends 8x the tensor work
// (and, at ~7 TB/s of HBM streaming, enough power to trip the board's power cap) on zero padding; here a
// tile costs 2 x 4 MMAs of 128x16x16, the query tile is 2 KB per stage (6 stages of 34 KB), an accumulator
// tile is 32 TMEM columns, and the epilogue is one X row per thread: one multiply + compare per query.
//
// ONE launch per search (per segment): after its last tile the CTA re-scores ITS OWN candidates from the stored
// rows with K1's arithmetic (small_tail), publishes one k-key list per query, and the last CTA of the search
// (atomic ticket over all segment launches) merges the lists, runs the cross-GPU key exchange and emits --
// what used to be refine_topk_kernel + exchange_merge_kernel (3 launches and ~40 us per search, which capped
// 8-GPU scaling at 0.71).  The CTA's first tile is PARKED: its accumulators stay in registers, its rows only
// OFFER their lower bounds to the shared list, and the tile is decided after the last one -- no thread ever
// waits for a first bound (the earlier spin stalled the TMA ring for ~13 us at every launch).
constexpr int SQ = 16;                                 // query columns per MMA
constexpr int S_STAGES = 6;
constexpr uint32_t SX_BYTES = BN * BK * 2;             // 32 KB: 256 rows x 64 dims
constexpr uint32_t SQ_BYTES = SQ * BK * 2;             // 2 KB
constexpr uint32_t S_STAGE_BYTES = SX_BYTES + SQ_BYTES;
constexpr uint32_t S_TMEM_COLS = 64;                   // 2 buffers x 2 sub-tiles x 16 columns
// int8 operands: a stage is 256 rows x 128 one-byte dims (the same 32 KB) + the 32-row query operand (4 KB)
constexpr uint32_t SQ8_BYTES = 2 * SQ * 128;           // 4 KB: 16 queries x 2 digits x 128 dims
constexpr uint32_t S8_STAGE_BYTES = SX_BYTES + SQ8_BYTES;
constexpr uint32_t S8_TMEM_COLS = 128;                 // 2 buffers x 2 sub-tiles x 32 columns
constexpr int S8_STAGES = 5;                           // 5 x 36 KB: leaves room for the co-resident gate / prep CTAs
constexpr int kSmallWarpRegions = 1;                   // candidate regions per (query, slice): one, shared by the CTA
constexpr int kTailWarps = kThreads / 32;

struct SmallTail {
  const float* q;              // [B][dim] fp32 queries
  const unsigned char* rows;   // stored rows of this segment
  const uint32_t* gids;
  const uint32_t* allow;       // per-search "allowed rows" bitmap or NULL
  int dim, dpad, row_bytes, cpr, lpr_log2, nch, bf16;
  float min_score;
  int* overflow;
  uint64_t* fin_keys;          // [B][kFinalCap] exact keys that reached the query's bound
  unsigned int* fin_count;     // [B] zero at launch
  unsigned int* ticket;
  XchgCtx xchg;
  const unsigned int* done_ctr;   // searches completed on this workspace
  unsigned int done_sn;           // this search's number
  uint64_t* keys_out;
  float* scores_out;
  long long* gids_out;
  int* counts_out;
};

constexpr int kFinalCap = 2048;   // exact keys per query that may reach the last CTA (normally a few dozen)

// Re-score CU x G candidates of one query exactly as K1 does: lanes-per-row lpr, chunk c = lig + j*lpr
// accumulated in j order with the x,y,z,w fmaf chain, xor-butterfly over the lpr lanes, K1's score formula.
// A key whose exact score reaches the query's bound L (a lower bound of the exact k-th best score, so no
// top-k row is below it) is appended to the query's FINAL list in global memory; everything else -- almost
// every candidate -- ends here.  All chunks of a row are requested before the first is consumed (NB = 8
// chunks per lane cover rows up to 4096 bytes x lanes-per-row / 32 in one round trip).
template <bool BF16, bool L2>
__device__ __forceinline__ void rescore_unit(const SmallTail& tp, const float* q_s, float qinv, bool cosine,
                                             const unsigned long long* cand, int cnt, int base, const float* inv_norm,
                                             long long n_rows, float L, unsigned int* fin_count, uint64_t* fin_keys, int lane) {
  (void)n_rows;
  constexpr int CU = 2, NB = 8;
  const int lpr_log2 = tp.lpr_log2, lpr = 1 << lpr_log2, G = 32 >> lpr_log2;
  const int g = lane >> lpr_log2, lig = lane & (lpr - 1);
  bool have[CU];
  long long row[CU];
  const unsigned char* rp[CU];
#pragma unroll
  for (int u = 0; u < CU; ++u) {
    const int ci = base + u * G + g;
    have[u] = ci < cnt;
    row[u] = have[u] ? static_cast<long long>(__ldcg(cand + ci) & 0xFFFFFFFFull) : 0ll;
    WDBX_ASSERT(row[u] >= 0 && row[u] < n_rows);
    rp[u] = tp.rows + static_cast<size_t>(row[u]) * tp.row_bytes;
  }
  float inx[CU];
  uint32_t gid[CU];
#pragma unroll
  for (int u = 0; u < CU; ++u) {
    inx[u] = cosine ? __ldg(inv_norm + row[u]) : 1.0f;
    gid[u] = __ldg(tp.gids + row[u]);
  }
  float dot[CU];
#pragma unroll
  for (int u = 0; u < CU; ++u) dot[u] = 0.0f;
  for (int j0 = 0; j0 < tp.nch; j0 += NB) {
    uint4 raw[CU][NB];
#pragma unroll
    for (int u = 0; u < CU; ++u) {
#pragma unroll
      for (int v = 0; v < NB; ++v) {
        const int c = lig + ((j0 + v) << lpr_log2);
        raw[u][v] = (j0 + v < tp.nch && c < tp.cpr) ? __ldg(reinterpret_cast<const uint4*>(rp[u] + c * 16))
                                                    : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      float acc1 = dot[u];
#pragma unroll
      for (int v = 0; v < NB; ++v) {
        const int c = lig + ((j0 + v) << lpr_log2);
        if (j0 + v < tp.nch && c < tp.cpr) {
          if (!BF16) {
            const float4 x = make_float4(__uint_as_float(raw[u][v].x), __uint_as_float(raw[u][v].y),
                                         __uint_as_float(raw[u][v].z), __uint_as_float(raw[u][v].w));
            const float4 qv = lds128(q_s + c * 4);
            if (L2) {
              const float d0 = x.x - qv.x, d1 = x.y - qv.y, d2 = x.z - qv.z, d3 = x.w - qv.w;
              acc1 = fmaf(d0, d0, acc1); acc1 = fmaf(d1, d1, acc1); acc1 = fmaf(d2, d2, acc1); acc1 = fmaf(d3, d3, acc1);
            } else {
              acc1 = fmaf(x.x, qv.x, acc1); acc1 = fmaf(x.y, qv.y, acc1); acc1 = fmaf(x.z, qv.z, acc1); acc1 = fmaf(x.w, qv.w, acc1);
            }
          } else {
            const uint32_t w[4] = {raw[u][v].x, raw[u][v].y, raw[u][v].z, raw[u][v].w};
            const float4 qa = lds128(q_s + c * 8);
            const float4 qb = lds128(q_s + c * 8 + 4);
            const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xFFFF0000u);
              if (L2) {
                const float d0 = lo - qq[2 * i], d1 = hi - qq[2 * i + 1];
                acc1 = fmaf(d0, d0, acc1); acc1 = fmaf(d1, d1, acc1);
              } else {
                acc1 = fmaf(lo, qq[2 * i], acc1); acc1 = fmaf(hi, qq[2 * i + 1], acc1);
              }
            }
          }
        }
      }
      dot[u] = acc1;
    }
  }
#pragma unroll
  for (int u = 0; u < CU; ++u) {
    float d = dot[u];
    for (int o = lpr >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(FULL_MASK, d, o);
    float sc = d;
    if (L2) sc = -sc;
    else if (cosine) sc = sc * inx[u] * qinv;
    sc = (sc != sc) ? __int_as_float(0xff800000) : sc;
    // L already includes the caller's score floor, applied exactly as K1 does (a row passes when s >= floor)
    if (have[u] && lig == 0 && sc >= L) {
      const unsigned int pos = atomicAdd(fin_count, 1u);
      if (pos < static_cast<unsigned int>(kFinalCap)) fin_keys[pos] = pack_key(sc, gid[u]);
    }
  }
}

// Sorted top-k list of n keys in global memory (one query's FINAL list): 32 keys per warp step, four steps in
// flight, each chunk sorted in registers and merged into the warp's list; a tree over the warps only when more
// than one warp had work.  All threads of the CTA call it; the result lands in final_list (shared memory).
template <int S>
__device__ __forceinline__ void final_topk(const uint64_t* keys, int n_keys, uint64_t* scratch, uint64_t* final_list,
                                           int k, int warp, int lane, int nwarps) {
  uint64_t acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0ull;
  const int nchunks = (n_keys + 31) >> 5;
  const bool solo = nchunks <= 4;   // CTA-uniform: warp 0 alone, no barrier
  const int wstride = solo ? 1 : nwarps;
  if (!solo || warp == 0) {
    for (int c0 = solo ? 0 : warp; c0 < nchunks; c0 += 4 * wstride) {
      uint64_t v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = (c0 + j * wstride) * 32 + lane;
        v[j] = idx < n_keys ? __ldcg(reinterpret_cast<const unsigned long long*>(keys + idx)) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c0 + j * wstride < nchunks) {   // warp-uniform
          uint64_t b4[S];
#pragma unroll
          for (int s = 0; s < S; ++s) b4[s] = 0ull;
          b4[0] = scan::warp_sort32_desc(v[j], lane);
          scan::warp_merge<S>(acc, b4, lane);
        }
      }
    }
  }
  if (!solo) scan::block_tree_merge<S>(acc, scratch, warp, lane, nwarps);
  if (warp == 0) scan::store_list<S>(final_list, acc, k, lane);
}

// Everything after the CTA's last tile (all kThreads threads call it; `ring` = the idle TMA stage ring).
template <int METRIC, int S>
__device__ __forceinline__ void small_tail(const FilterParams& p, const SmallTail& tp, unsigned char* ring, size_t ring_bytes,
                                           const unsigned int* wcount, const unsigned int* Lq, int slice,
                                           unsigned long long t_entry, unsigned long long t_loop) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long ts[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) ts[i] = 0ull;
  auto stamp = [&](int i) { if (p.trace) ts[i] = gtime_ns() - t_entry; };
  stamp(0);
  const int k = p.k, B = p.B, dpad = tp.dpad;
  uint64_t* scratch = reinterpret_cast<uint64_t*>(ring);                 // [kTailWarps][32 * S] merge scratch
  uint64_t* final_list = scratch + kTailWarps * 32 * S;                  // [32 * S]
  float* qinv_s = reinterpret_cast<float*>(final_list + 32 * S);         // [SQ] 1/|q| as K1 computes it
  float* L_s = qinv_s + SQ;                                              // [SQ] lower bound of the exact k-th best score
  int* cnt_s = reinterpret_cast<int*>(L_s + SQ);                         // [SQ] candidates of this CTA per query
  int* flag_s = cnt_s + SQ;                                              // [2 SQ + 2] last CTA | any overflow | overflow, count per query
  const size_t q_off = scan::align128(static_cast<size_t>(kTailWarps + 1) * 32 * S * 8 + (5 * SQ + 2) * 4);
  float* q_s = reinterpret_cast<float*>(ring + q_off);                   // [QG][dpad] fp32 queries, zero padded
  int QG = static_cast<int>((ring_bytes - q_off) / (static_cast<size_t>(dpad) * 4));
  QG = QG < B ? QG : B;                                                  // >= 1: the host routes larger rows to K1
  const size_t my_slice = static_cast<size_t>(p.slice_base + slice);
  if (tid < SQ) {
    int c = tid < B ? static_cast<int>(wcount[tid]) : 0;
    if (c > p.cap) {   // adversarial data: thousands of near-identical rows -> the flag-gated K1 launch re-runs the query
      tp.overflow[tid] = 1;
      c = 0;
    }
    cnt_s[tid] = c;
    if (tid < B) {
      p.cand_count[static_cast<size_t>(tid) * p.s_total + my_slice] = wcount[tid];   // statistics only
      // the settled shared bound (the k-th largest lower bound any CTA has seen) or the CTA's own / the floor
      const unsigned int g = __ldcg(p.lower_glob + tid);
      L_s[tid] = unmono_f32(max(max(g, Lq[tid]), 0x007FFFFFu));
    }
  }
  __syncthreads();
  const bool cosine = METRIC == kCosine;
  const int G = 32 >> tp.lpr_log2;
  const int upc = 2 * G;   // candidates per work unit (rescore_unit: CU = 2)
  for (int j0 = 0; j0 < B; j0 += QG) {
    const int nqg = (B - j0) < QG ? (B - j0) : QG;
    for (int i = tid; i < nqg * dpad; i += kThreads) {
      const int b = i / dpad, c = i - b * dpad;
      q_s[i] = c < tp.dim ? __ldg(tp.q + static_cast<size_t>(j0 + b) * tp.dim + c) : 0.0f;
    }
    __syncthreads();
    for (int b = warp; b < nqg; b += kTailWarps) {
      float ss = 0.0f;
      for (int i = lane; i < dpad; i += 32) ss = fmaf(q_s[b * dpad + i], q_s[b * dpad + i], ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
      if (lane == 0) qinv_s[j0 + b] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
    }
    __syncthreads();
    stamp(1);
    // work units (query, chunk of `upc` candidates) dealt round-robin over the 12 warps
    const int myu = lane < nqg ? (cnt_s[j0 + lane] + upc - 1) / upc : 0;
    int incl = myu;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    for (int u = warp; u < total; u += kTailWarps) {
      const int jl = __ffs(__ballot_sync(FULL_MASK, incl > u)) - 1;
      const int c = u - __shfl_sync(FULL_MASK, incl - myu, jl);
      const int j = j0 + jl;
      const unsigned long long* cand = p.cand + (static_cast<size_t>(j) * p.s_total + my_slice) * p.cap;
      if (tp.bf16) rescore_unit<true, METRIC == kL2>(tp, q_s + jl * dpad, qinv_s[j], cosine, cand, cnt_s[j], c * upc, p.inv_norm,
                                                     p.n_rows, L_s[j], tp.fin_count + j, tp.fin_keys + static_cast<size_t>(j) * kFinalCap, lane);
      else rescore_unit<false, METRIC == kL2>(tp, q_s + jl * dpad, qinv_s[j], cosine, cand, cnt_s[j], c * upc, p.inv_norm,
                                              p.n_rows, L_s[j], tp.fin_count + j, tp.fin_keys + static_cast<size_t>(j) * kFinalCap, lane);
    }
    if (j0 + QG < B) __syncthreads();   // q_s is rewritten by the next group
  }
  stamp(2);
  __threadfence();
  __syncthreads();
  stamp(4);
  if (tid == 0) flag_s[0] = (atomicAdd(tp.ticket, 1u) == static_cast<unsigned int>(p.s_total) - 1u) ? 1 : 0;
  __syncthreads();
  stamp(5);
  if (p.trace && tid == 128 && flag_s[0] == 0)
    printf("TRACE cta %d cand %d entry %llu loop_end %llu tail_in %llu q_ready %llu rescored %llu fenced %llu ticket %llu\n",
           slice, cnt_s[0], t_entry, t_loop - t_entry, ts[0], ts[1], ts[2], ts[4], ts[5]);
  if (flag_s[0] == 0) return;
  // ---- last CTA of the search: per query the few exact keys that reached the bound -> top-k, exchange, emit
  // (consecutive searches overlap on the device; their exchanges and results stay in order)
  if (tid == 0) wait_search_done(tp.done_ctr, tp.done_sn - 1u);
  __syncthreads();
  __threadfence();
  if (tid < B) {
    const int n = static_cast<int>(__ldcg(tp.fin_count + tid));
    int o = *reinterpret_cast<volatile int*>(tp.overflow + tid);
    if (n > kFinalCap) {   // only without any bound (fewer than k live rows offered) on a huge candidate set
      tp.overflow[tid] = 1;
      o = 1;
    }
    flag_s[2 + tid] = o;
    flag_s[2 + SQ + tid] = n;
  }
  __syncthreads();
  if (tid == 0) {
    int any = 0;
    for (int j = 0; j < B; ++j) any |= flag_s[2 + j];
    flag_s[1] = any;
  }
  __syncthreads();
  stamp(6);
  const bool xchg = tp.xchg.world > 1;
  if (xchg && flag_s[1]) return;   // a query overflowed: the flag-gated K1 launch redoes the whole collective search
  for (int j = 0; j < B; ++j) {
    const bool over = flag_s[2 + j] != 0;   // CTA-uniform
    if (!over) final_topk<S>(tp.fin_keys + static_cast<size_t>(j) * kFinalCap, flag_s[2 + SQ + j], scratch, final_list, k,
                             warp, lane, kTailWarps);
    stamp(7);
    if (warp == 0) {
      __syncwarp();
      if (over) {
        if (tp.counts_out && lane == 0) tp.counts_out[j] = 0;
      } else if (xchg) {
        scan::xchg_push(tp.xchg, j, final_list, k, lane);
      } else {
        scan::emit_list(final_list, k, lane, tp.keys_out ? tp.keys_out + static_cast<size_t>(j) * k : nullptr,
                        tp.scores_out ? tp.scores_out + static_cast<size_t>(j) * k : nullptr,
                        tp.gids_out ? tp.gids_out + static_cast<size_t>(j) * k : nullptr, tp.counts_out ? tp.counts_out + j : nullptr);
      }
    }
    if (j + 1 < B) __syncthreads();
  }
  if (xchg && warp == 0) {
    const bool ok = scan::xchg_publish_wait(tp.xchg, lane);
    for (int j = 0; j < B; ++j)
      scan::xchg_merge_emit(tp.xchg, j, final_list, k, ok, lane, tp.keys_out ? tp.keys_out + static_cast<size_t>(j) * k : nullptr,
                            tp.scores_out ? tp.scores_out + static_cast<size_t>(j) * k : nullptr,
                            tp.gids_out ? tp.gids_out + static_cast<size_t>(j) * k : nullptr, tp.counts_out ? tp.counts_out + j : nullptr);
  }
  stamp(8);
  if (p.trace && tid == 0)
    printf("TRACE LAST cta %d cand %d entry %llu loop_end %llu tail_in %llu q_ready %llu rescored %llu fenced %llu ticket %llu flags %llu merged %llu done %llu keys %d\n",
           slice, cnt_s[0], t_entry, t_loop - t_entry, ts[0], ts[1], ts[2], ts[4], ts[5], ts[6], ts[7], ts[8], flag_s[2 + SQ]);
}

template <int METRIC, bool I8>
__global__ void __launch_bounds__(kThreads, 1)
gemm_filter_small_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
                         const __grid_constant__ FilterParams p, const __grid_constant__ SmallTail tp) {
  constexpr uint32_t STAGE_BYTES = I8 ? S8_STAGE_BYTES : S_STAGE_BYTES;   // X tile + query operand
  constexpr uint32_t TMEM_COLS_S = I8 ? S8_TMEM_COLS : S_TMEM_COLS;
  constexpr int NQ = I8 ? 2 * SQ : SQ;                                     // accumulator columns per sub-tile
  constexpr int KB_ELEMS = I8 ? 128 : BK;                                  // dims per 128-byte K block
  constexpr int NST = I8 ? S8_STAGES : S_STAGES;                           // ring depth
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stage_base = smem;
  unsigned int* Lq = reinterpret_cast<unsigned int*>(smem + NST * STAGE_BYTES);  // [SQ] mono(best known bound)
  float* Tq = reinterpret_cast<float*>(Lq + SQ);                                        // [SQ] fast-test threshold
  float* q_inv_s = Tq + SQ;
  float* q_nrm_s = q_inv_s + SQ;
  float* q_sq_s = q_nrm_s + SQ;
  float* q_A_s = q_sq_s + SQ;                                                           // [SQ] QueryBound::A
  float* q_C_s = q_A_s + SQ;                                                            // [SQ] QueryBound::C
  float* q_fA_s = q_C_s + SQ;                                                           // [SQ] fast-test form of A
  unsigned int* wcount = reinterpret_cast<unsigned int*>(q_fA_s + SQ);                  // [SQ] candidates of this CTA per query
  uint64_t* bars = reinterpret_cast<uint64_t*>(wcount + SQ);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + S_STAGES;
  uint64_t* tmem_full = bars + 2 * S_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* sched_full = tmem_empty + 2;     // [4] tile-index ring: producer -> MMA warp + epilogue warps
  uint64_t* sched_empty = sched_full + 4;    // [4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sched_empty + 4);
  int* tile_ring = reinterpret_cast<int*>(tmem_ptr + 4);   // [4] tile index per ring slot, -1 = no more tiles
  unsigned int* tile_ts = reinterpret_cast<unsigned int*>(tile_ring + 4);   // [48] trace: ns at which tile #it was ready
  unsigned int* offer = tile_ts + 48;        // [8][SQ] mailbox: best lower bound each epilogue warp has seen per query
  unsigned int* epi_done = offer + 8 * SQ;   // epilogue warps that have finished their tiles
  float* q_s1_s = reinterpret_cast<float*>(epi_done + 4);   // [SQ] int8 queries: scale of the first digit ...
  float* q_s2_s = q_s1_s + SQ;                               // [SQ] ... and of the second

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.y;
  const int k = p.k;
  const float NEG_INF = __int_as_float(0xff800000);
  const unsigned long long t_entry = p.trace ? gtime_ns() : 0ull;

  // programmatic dependent launch: the flag-gated K1 launch behind us may become resident as SMs free up; our
  // own set-up below overlaps the tail of the launch before us (prep), whose results we wait for right after
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_q);
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tmem_full + a), 1);
      mbar_init(smem_u32(tmem_empty + a), 8);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(sched_full + s), 1);
      mbar_init(smem_u32(sched_empty + s), 9);   // MMA warp + 8 epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), TMEM_COLS_S);
  if (p.prep_count != nullptr) {
    // overlap mode: wait for OUR prep by count (it ran early, next to the previous search's last CTAs), not for the
    // launches before it -- this is what lets the first tiles stream while the previous search finishes its tail
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      unsigned int v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.prep_count) : "memory");
        if (static_cast<int>(v - p.prep_target) >= 0) break;
        if (clock64() - t0 > 6000000000ll) break;
        __nanosleep(100);
      } while (true);
      asm volatile("fence.proxy.async;" ::: "memory");   // the query tile is read through TMA (async proxy)
    }
    __syncthreads();
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (warp == 3) {
    for (int j = lane; j < SQ; j += 32) {
      const bool v = j < p.B;
      const QueryBound b = make_query_bound<METRIC>(v ? p.q_inv[j] : 0.0f, v ? p.q_nrm[j] : 0.0f, v ? p.q_sq[j] : 0.0f,
                                                    v ? p.q_bn[j] : 0.0f, v ? p.q_tn[j] : 0.0f, p.acc_rel);
      q_inv_s[j] = b.qinv;
      q_nrm_s[j] = b.qnrm;
      q_sq_s[j] = b.qsq;
      q_A_s[j] = b.A;
      q_C_s[j] = b.C;
      q_fA_s[j] = (METRIC == kCosine) ? b.A * b.qnrm : b.A;
      q_s1_s[j] = (I8 && v) ? p.q_s1[j] : 0.0f;
      q_s2_s[j] = (I8 && v) ? p.q_s2[j] : 0.0f;
      // the caller's score floor is a valid lower bound of every RETURNED score from the start
      float t0 = NEG_INF;
      unsigned int l0 = 0u;
      if (tp.min_score > NEG_INF) {
        l0 = mono_u32(tp.min_score);
        if (METRIC == kCosine) {
          const float t = (tp.min_score - b.C) * b.qnrm;
          t0 = t - 1e-6f * (fabsf(t) + b.qnrm);
        } else if (METRIC == kL2) {
          t0 = (tp.min_score + b.qsq * (1.0f - p.c_l2)) - 1e-6f * (fabsf(tp.min_score) + 2.0f * b.qsq);
        } else {
          t0 = tp.min_score - 1e-6f * fabsf(tp.min_score);
        }
      }
      Lq[j] = l0;
      Tq[j] = t0;
    }
    for (int i = lane; i < SQ; i += 32) wcount[i] = 0u;
    for (int i = lane; i < 48; i += 32) tile_ts[i] = 0u;
    for (int i = lane; i < 8 * SQ + 1; i += 32) offer[i] = 0u;   // (+ epi_done)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // DYNAMIC TILE SCHEDULE.  Tiles are 256 consecutive rows; the producer draws the next tile index from a
  // per-launch atomic counter and hands it to the MMA warp and the epilogue warps through a 4-slot ring in
  // shared memory (mbarrier pipeline: sched_full / sched_empty).  A static round-robin split left the CTAs
  // finishing up to 54 us apart on a 281 us launch (SMs do not get equal shares of the HBM stream); with the
  // counter every CTA ends within one tile (~8 us) of the others.
  // consumers: wait for the ring slot of tile number `sq` of this CTA, read it, release the slot
  auto fetch_tile = [&](int sq) -> int {
    const int slot = sq & 3;
    mbar_wait(smem_u32(sched_full + slot), (static_cast<uint32_t>(sq) >> 2) & 1u);
    const int t = *reinterpret_cast<volatile int*>(tile_ring + slot);
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(sched_empty + slot));
    return t;
  };

  const float c_l2 = p.c_l2;
  auto fast_thr = [&](float l, int j) -> float {
    if (METRIC == kCosine) {
      const float t = (l - q_C_s[j]) * q_nrm_s[j];
      return t - 1e-6f * (fabsf(t) + q_nrm_s[j]);
    }
    if (METRIC == kL2) return (l + q_sq_s[j] * (1.0f - c_l2)) - 1e-6f * (fabsf(l) + 2.0f * q_sq_s[j]);
    return l - 1e-6f * fabsf(l);
  };
  auto raise_bound = [&](int j, float nl) {   // any thread: publish a better bound of query j to the CTA
    const unsigned int m = mono_u32(nl);
    if (atomicMax(Lq + j, m) < m) Tq[j] = fast_thr(nl, j);
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    // the draw of tile sq+1 (a global atomic: a microsecond or two under a saturated HBM stream) is issued
    // before the loads of tile sq and only consumed after them, so the producer never sits on its latency
    auto publish = [&](int sq, int t) {
      const int slot = sq & 3;
      mbar_wait(smem_u32(sched_empty + slot), ((static_cast<uint32_t>(sq) >> 2) & 1u) ^ 1u);
      if (lane == 0) {
        tile_ring[slot] = t;
        mbar_arrive(smem_u32(sched_full + slot));
      }
      __syncwarp();
    };
    auto draw = [&]() -> int {   // lane 0 holds the raw ticket
      return lane == 0 ? static_cast<int>(atomicAdd(p.tile_ctr, 1u)) : 0;
    };
    auto settle = [&](int raw) -> int {
      const int t = __shfl_sync(FULL_MASK, raw, 0);
      return t < p.n_tiles ? t : -1;
    };
    int t_cur = settle(draw());
    publish(0, t_cur);
    for (int sq = 0; t_cur >= 0; ++sq) {
      const int raw_nxt = draw();
      WDBX_ASSERT(t_cur >= 0 && t_cur < p.n_tiles && stage >= 0 && stage < NST);
      const int row0 = t_cur * BN;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
        const uint32_t sb = smem_u32(stage_base) + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint32_t bar = smem_u32(full_bar) + static_cast<uint32_t>(stage) * 8u;
        if (elect_one()) {
          mbar_expect_tx(bar, STAGE_BYTES);
          tma_load_2d(sb, &tm_x, kb * KB_ELEMS, row0, bar);
          tma_load_2d(sb + SX_BYTES, &tm_q, kb * KB_ELEMS, 0, bar);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      t_cur = settle(raw_nxt);
      publish(sq + 1, t_cur);
    }
  } else if (warp == 1) {
    // M = 128 X rows; N = 16 queries (BF16 x BF16 -> F32) or 16 queries x 2 int8 digits (S8 x S8 -> S32)
    const uint32_t idesc = I8 ? make_idesc_i8(128, NQ) : make_idesc(128, SQ, 1u);
    int stage = 0;
    uint32_t phase = 0;
    for (int sq = 0;; ++sq) {
      if (fetch_tile(sq) < 0) break;
      const int a = sq & 1;
      mbar_wait(smem_u32(tmem_empty + a), ((static_cast<uint32_t>(sq) >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * 2 * NQ);
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(stage_base) + static_cast<uint32_t>(stage) * STAGE_BYTES;
        const uint64_t d_x0 = make_desc_kmajor(sb, 128, 2);
        const uint64_t d_x1 = make_desc_kmajor(sb + SX_BYTES / 2, 128, 2);
        const uint64_t d_q = make_desc_kmajor(sb + SX_BYTES, 128, 2);
        const uint32_t ebar = smem_u32(empty_bar) + static_cast<uint32_t>(stage) * 8u;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = static_cast<uint64_t>((kk * 16 * 2) >> 4);
            if (I8) {   // 32 bytes = 32 int8 dims per instruction: the same descriptor walk as 16 bf16 dims
              umma_i8(d_tmem, d_x0 + adv, d_q + adv, idesc, (kb | kk) ? 1u : 0u);
              umma_i8(d_tmem + NQ, d_x1 + adv, d_q + adv, idesc, (kb | kk) ? 1u : 0u);
            } else {
              umma_f16(d_tmem, d_x0 + adv, d_q + adv, idesc, (kb | kk) ? 1u : 0u);
              umma_f16(d_tmem + NQ, d_x1 + adv, d_q + adv, idesc, (kb | kk) ? 1u : 0u);
            }
          }
          umma_commit(ebar);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(smem_u32(tmem_full + a));
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===== BOUND SERVER.  The epilogue warps never touch the shared lower-bound list in global memory themselves
    // (1184 warps compare-and-swapping one cache line at the first tile stalled every epilogue -- and with it the
    // accumulators, the MMAs and the TMA ring -- for ~35 us of a 300 us launch): they drop their best lower bound per
    // query into a shared-memory mailbox and move on; this otherwise idle warp carries the offers into the list
    // (list_multi_insert: all of a query's offers in one or two L2 round trips), picks up the bounds other CTAs
    // have published, and raises the CTA's thresholds.
    const unsigned int rot = static_cast<unsigned int>(blockIdx.y) * 5u + 1u;   // spreads the CTAs over equal slots
    // one sweep = ONE insert round for every query (four queries side by side while k <= 32), results published after
    // every round: a first -- weak -- bound reaches the epilogue after a few microseconds and tightens from there
    auto serve = [&](auto gl_tag) {
      constexpr int GL = decltype(gl_tag)::value;
      constexpr int QPR = 32 / GL;              // queries per round
      constexpr int NS = SQ / QPR;              // rounds per sweep at most
      unsigned int pend[NS];                    // this lane's pending offer per round slot
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) pend[sidx] = 0u;
      for (;;) {
        if (*reinterpret_cast<volatile unsigned int*>(epi_done) >= 8u) break;   // late offers would only polish the bound
        bool work = false;
#pragma unroll
        for (int sidx = 0; sidx < NS; ++sidx) {
          if (sidx * QPR < p.B) {               // warp-uniform
            const int grp = lane / GL, gl = lane & (GL - 1);
            const int j = sidx * QPR + grp;
            const bool qv = j < p.B;
            if (qv && gl < 8) pend[sidx] = max(pend[sidx], atomicExch(offer + gl * SQ + j, 0u));
            const unsigned int cur = qv ? *reinterpret_cast<volatile unsigned int*>(Lq + j) : 0u;
            if (pend[sidx] <= cur) pend[sidx] = 0u;   // cannot beat the list's minimum any more
            if (__any_sync(FULL_MASK, pend[sidx] != 0u)) {
              work = true;
              const unsigned int nl = list_multi_insert<GL>(p.lower_list + static_cast<size_t>(qv ? j : 0) * kMaxKFilter,
                                                            p.lower_glob + (qv ? j : 0), qv ? k : 0, pend[sidx], lane, rot, 1);
              if (qv && gl == 0 && nl > cur) raise_bound(j, unmono_f32(nl));
            }
          }
        }
        if (lane < p.B) {
          const unsigned int g = __ldcg(p.lower_glob + lane);
          if (g > *reinterpret_cast<volatile unsigned int*>(Lq + lane)) raise_bound(lane, unmono_f32(g));
        }
        // PARTITION MAXIMA: slot s holds the best lower bound among the rows of the warps with number = s (mod k) --
        // k disjoint row sets, so the smallest slot is a lower bound of the exact k-th best score too.  Looser than the
        // list (it sits near rank k ln k) but every update is ONE fire-and-forget atomicMax from the epilogue: it is
        // there one round trip after the first tile, while the exact list takes ~10 compare-and-swap rounds to settle.
        for (int j0 = 0; j0 < p.B; j0 += 4) {
          uint4 v4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v4[u] = (j0 + u < p.B && 4 * lane < k)
                        ? __ldcg(reinterpret_cast<const uint4*>(p.part_max + static_cast<size_t>(j0 + u) * kMaxKFilter + 4 * lane))
                        : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j0 + u < p.B) {   // warp-uniform
              unsigned int vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (4 * lane + jj >= k) vv[jj] = 0xFFFFFFFFu;
              const unsigned int gm = __reduce_min_sync(FULL_MASK, min(min(vv[0], vv[1]), min(vv[2], vv[3])));
              if (lane == 0 && gm > 0x007FFFFFu && gm != 0xFFFFFFFFu &&
                  gm > *reinterpret_cast<volatile unsigned int*>(Lq + j0 + u))
                raise_bound(j0 + u, unmono_f32(gm));
            }
          }
        }
        __syncwarp();
        if (!work) __nanosleep(300);
      }
    };
    if (k <= 32) serve(std::integral_constant<int, 8>{});
    else serve(std::integral_constant<int, 32>{});
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps, one X row of the tile per thread (TMEM lane = row inside the 128-row sub-tile)
    const int ew = warp - 4;
    const int quarter = warp & 3, sub = ew >> 2;
    const int row_in_tile = sub * 128 + quarter * 32 + lane;
    const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(sub * NQ);
    // fast test per (row, query), see gemm_filter_kernel (same inequalities, thresholds with 1e-6 slack); admitted
    // pairs are decided by the reference form (bound_eval) below.  L / T live in shared memory per query and are
    // shared by the whole CTA.
    auto qbound = [&](int j) -> QueryBound {
      QueryBound b;
      b.A = q_A_s[j]; b.C = q_C_s[j]; b.qinv = q_inv_s[j]; b.qnrm = q_nrm_s[j]; b.qsq = q_sq_s[j];
      return b;
    };
    auto load_row = [&](int t, float& inx, float& sq, float& rr, float& sx) {
      const long long row = static_cast<long long>(t) * BN + row_in_tile;
      inx = 0.0f;
      sq = 0.0f;
      rr = 0.0f;
      sx = 0.0f;
      if (t >= 0 && row < p.n_rows) {
        if (METRIC == kCosine) inx = __ldg(p.inv_norm + row);
        else sq = __ldg(p.sqnorm + row);
        if (p.rres != nullptr) rr = __ldg(p.rres + row);
        if (I8) sx = __ldg(p.rowscale + row);
      }
    };
    // the rows this warp ever sees form one of k partitions of the store (warp number modulo k)
    const int pm_slot = static_cast<int>((static_cast<unsigned int>(p.slice_base + slice) * 8u + static_cast<unsigned int>(ew)) %
                                         static_cast<unsigned int>(k));
    int t_nxt = fetch_tile(0);
    const int t_first = t_nxt;   // the CTA's first tile is parked (decided last)
    float inx_next, sq_next, rr_next, sx_next;
    load_row(t_nxt, inx_next, sq_next, rr_next, sx_next);
    uint32_t r0[SQ];   // the parked accumulators of this thread's row of the first tile
#pragma unroll
    for (int j = 0; j < SQ; ++j) r0[j] = 0u;
    // iterations 0, 1, ... are the CTA's tiles in the order it drew them; one more iteration decides the parked tile
    for (int it = 0; t_first >= 0; ++it) {
      const bool parked = t_nxt < 0;
      const int t = parked ? t_first : t_nxt;
      const int a = it & 1;
      WDBX_ASSERT(t >= 0 && t < p.n_tiles);
      WDBX_ASSERT(static_cast<uint32_t>(a * 2 * NQ + sub * NQ + NQ) <= TMEM_COLS_S);
      const long long row = static_cast<long long>(t) * BN + row_in_tile;
      const bool row_valid = row < p.n_rows;
      const float inx = inx_next, sq = sq_next, rr = rr_next, sx = sx_next;
      // per-row terms of the fast test: cosine fr = rho_x; ip fe = |x|, fr = |r|; l2 fe = 2|x|, fr = 2|r|, sqs = shrunk |x|^2
      const float xn = sqrtf(sq) * 1.00001f;
      const float fe = (METRIC == kL2) ? 2.0f * xn : xn;
      const float fr = (METRIC == kCosine) ? rr * inx : ((METRIC == kL2) ? 2.0f * rr : rr);
      const float sqs = sq * ((1.0f - c_l2) * (1.0f - 1e-6f));
      if (it == 1) {
        // the offers of the parked tile are on their way through the bound server: give the first bound a moment
        // (shared-memory polling, bounded) rather than appending this whole tile against "no bound yet"
        for (int spin = 0; spin < 400; ++spin) {
          const bool have = lane >= p.B || *reinterpret_cast<volatile unsigned int*>(Lq + lane) > 0x007FFFFFu;
          if (__all_sync(FULL_MASK, have)) break;
          __nanosleep(100);
        }
      }
      if (!parked) {
        t_nxt = fetch_tile(it + 1);
        load_row(t_nxt >= 0 ? t_nxt : t_first, inx_next, sq_next, rr_next, sx_next);
      }
      uint32_t r[SQ];
      if (!parked) {
        mbar_wait(smem_u32(tmem_full + a), (static_cast<uint32_t>(it) >> 1) & 1u);
        if (p.trace && threadIdx.x == 128 && it < 48) tile_ts[it] = static_cast<unsigned int>(gtime_ns() - t_entry);
        tc_fence_after();
        __syncwarp();
        if (I8) {
          // exact integer partial dots of the two query digits -> the dot-product estimate in the row's real units:
          // d = sx * (s1 * A1 + s2 * A2); from here on the epilogue is the same as for bf16 operands
          uint32_t ri[32];
          tmem_ld32(taddr0 + static_cast<uint32_t>(a * 2 * NQ), ri);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < SQ; ++j) {
            const float a1 = static_cast<float>(static_cast<int>(ri[j])), a2 = static_cast<float>(static_cast<int>(ri[SQ + j]));
            r[j] = __float_as_uint(sx * fmaf(a2, q_s2_s[j], a1 * q_s1_s[j]));
          }
        } else {
          tmem_ld16(taddr0 + static_cast<uint32_t>(a * 2 * NQ), r);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tmem_empty + a));   // the accumulator is in registers: free it now
        if (it == 0) {
#pragma unroll
          for (int j = 0; j < SQ; ++j) r0[j] = r[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < SQ; ++j) r[j] = r0[j];
      }
      // fast test: one multiply (or fma) + compare per (row, query); `any` = queries with an admitted row
      unsigned any = 0u;
#pragma unroll
      for (int j = 0; j < SQ; ++j) {
        if (j < p.B) {   // uniform
          const float d = __uint_as_float(r[j]);
          float u;
          if (METRIC == kCosine) u = fmaf(fr, q_fA_s[j], d * inx);
          else if (METRIC == kL2) u = fmaf(2.0f, d, fmaf(fr, q_fA_s[j], fe * q_C_s[j]) - sqs);
          else u = d + fmaf(fr, q_fA_s[j], fe * q_C_s[j]);
          if (__any_sync(FULL_MASK, row_valid && !(u < Tq[j]))) any |= 1u << j;
        }
      }
      if (any) {
        // rare path (warp-uniform).  Reference form of the bound per admitted query; the row with the best
        // lower bound of the warp is offered to the shared list -- one lane per QUERY inserts, all queries
        // in parallel -- then the rows that still pass against the updated bound are appended
        // (warp-aggregated, one shared-memory atomic per warp and query).
        bool dead = false;
        if (row_valid) {
          if (p.tomb != nullptr) dead = (__ldg(p.tomb + (row >> 5)) >> (row & 31)) & 1u;
          if (tp.allow != nullptr) dead = dead || !((__ldg(tp.allow + (row >> 5)) >> (row & 31)) & 1u);
        }
        if (!parked) {
          float my_best = NEG_INF;       // lane j: best lower bound of this warp's rows for query j
          unsigned rem = any;
          while (rem) {
            const int j = __ffs(rem) - 1;
            rem &= rem - 1;
            float d = 0.0f;
#pragma unroll
            for (int jj = 0; jj < SQ; ++jj) d = (jj == j) ? __uint_as_float(r[jj]) : d;   // r[] stays in registers
            float sv, ej;
            bound_eval<METRIC>(d, inx, sq, rr, qbound(j), c_l2, sv, ej);
            const float lo = (row_valid && !dead) ? sv - ej : NEG_INF;   // NaN: mono_u32 ranks it as -inf
            const unsigned int wbest = __reduce_max_sync(FULL_MASK, mono_u32(lo));
            if (lane == j) my_best = unmono_f32(max(wbest, 0x007FFFFFu));
          }
          if (lane < p.B && ((any >> lane) & 1u)) {
            // hand the warp's best row of this tile to the bound server (never block on global memory here)
            const unsigned int m = mono_u32(my_best);
            if (m > max(*reinterpret_cast<volatile unsigned int*>(Lq + lane), 0x007FFFFFu)) {
              atomicMax(offer + ew * SQ + lane, m);
              // ... and, fire-and-forget, into this warp's slot of the PARTITION MAXIMA (one reduction, no retry)
              atomicMax(p.part_max + static_cast<size_t>(lane) * kMaxKFilter + pm_slot, m);
            }
          }
          __syncwarp();
        }
        if (it != 0) {   // tile 0 only offers; it is decided in the parked iteration against the settled bound
          unsigned rem = any;
          while (rem) {
            const int j = __ffs(rem) - 1;
            rem &= rem - 1;
            float d = 0.0f;
#pragma unroll
            for (int jj = 0; jj < SQ; ++jj) d = (jj == j) ? __uint_as_float(r[jj]) : d;
            float sv, ej;
            bound_eval<METRIC>(d, inx, sq, rr, qbound(j), c_l2, sv, ej);
            const float up = sv + ej;   // upper bound of the exact score (reference form)
            const float L = unmono_f32(max(*reinterpret_cast<volatile unsigned int*>(Lq + j), 0x007FFFFFu));
            const bool take = row_valid && !dead && !(up < L);   // !(x < L) keeps NaN (the refine ranks it like K1)
            const unsigned tm = __ballot_sync(FULL_MASK, take);
            if (tm) {
              unsigned int base = 0;
              if (lane == 0) base = atomicAdd(wcount + j, static_cast<unsigned int>(__popc(tm)));   // shared memory
              base = __shfl_sync(FULL_MASK, base, 0);
              if (take) {
                const unsigned int idx = base + __popc(tm & ((1u << lane) - 1u));
                if (idx < static_cast<unsigned int>(p.cap)) {
                  const size_t region = static_cast<size_t>(j) * p.s_total + static_cast<size_t>(p.slice_base + slice);
                  WDBX_ASSERT(j < p.B && region < static_cast<size_t>(p.B) * p.s_total && row >= 0 && row < p.n_rows);
                  p.cand[region * p.cap + idx] = (static_cast<unsigned long long>(p.seg) << 32) | static_cast<unsigned long long>(row);
                }
              }
            }
          }
        }
      }
      if (parked) break;
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      atomicAdd(epi_done, 1u);
    }
  }

  const unsigned long long t_loop = p.trace ? gtime_ns() : 0ull;
  tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 128 && (slice % 16) == 0)   // a sample of CTAs, after the loop: does not disturb it
    for (int i = 0; i + 8 <= 48; i += 8)
      printf("TILES cta %d from %d : %u %u %u %u %u %u %u %u\n", slice, i, tile_ts[i], tile_ts[i + 1], tile_ts[i + 2],
             tile_ts[i + 3], tile_ts[i + 4], tile_ts[i + 5], tile_ts[i + 6], tile_ts[i + 7]);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_S);
  }
  // the TMA ring is idle (every load was consumed by an MMA that an epilogue warp has waited for): reuse it
  constexpr size_t kRing = static_cast<size_t>(NST) * STAGE_BYTES;
  if (k <= 32) small_tail<METRIC, 1>(p, tp, stage_base, kRing, wcount, Lq, slice, t_entry, t_loop);
  else small_tail<METRIC, 4>(p, tp, stage_base, kRing, wcount, Lq, slice, t_entry, t_loop);
}

constexpr size_t kFilterSmallSmem = 1024 + static_cast<size_t>(S_STAGES) * S_STAGE_BYTES + (8 * SQ + SQ) * 4 +
                                    (2 * S_STAGES + 12) * 8 + 16 + 16 + 48 * 4 + (8 * SQ + 4) * 4 + 2 * SQ * 4;
constexpr size_t kFilterSmallSmemI8 = kFilterSmallSmem - static_cast<size_t>(S_STAGES) * S_STAGE_BYTES +
                                      static_cast<size_t>(S8_STAGES) * S8_STAGE_BYTES;

constexpr size_t filter_smem(int ncta) {
  const size_t stages = ncta == 2 ? STAGES_PAIR : STAGES_SINGLE;
  const size_t stage_bytes = static_cast<size_t>(BN / ncta) * BK * 2 + Q_TILE_BYTES;
  return 1024 + stages * stage_bytes + 6 * BN * 4 + (2 * stages + 4) * 8 + 16;
}

// ---------------------------------------------------------------- refine kernel
// gridDim.y CTAs per query (one when the batch alone fills the GPU; more for small batches, each taking a
// share of the candidate regions and writing a partial list that K3 merges).  Candidates are re-scored from the stored rows with K1's arithmetic: lanes-per-row
// lpr, chunk c = lig + j*lpr accumulated in j order with the x,y,z,w fmaf chain, xor-butterfly over
// lpr lanes, then the same score formula -- so keys are bit-identical to scan_topk_kernel's.
struct RefineParams {
  SegDesc seg[kMaxSeg];
  const float* q;                   // [B][dim]
  const unsigned long long* cand;   // [B][s_total][cap]
  const unsigned int* cand_count;   // [B][s_total]
  int s_total;
  int* overflow;                    // [B] set to 1 when the candidate list overflowed
  uint64_t* part;                   // [B][gridDim.y][k] partial lists when gridDim.y > 1
  unsigned int* tickets;            // [B] zeroed per search: the last CTA of a query merges the partial lists
  int sub;                          // warps sharing one candidate region (work units per region)
  int B, dim, dpad, row_bytes, cpr, lpr_log2, nch, k, metric, cap;
  uint64_t* keys_out;
  float* scores_out;
  long long* gids_out;
  int* counts_out;
};

template <bool BF16, bool L2, int S>
__global__ void __launch_bounds__(256, 2) refine_topk_kernel(const __grid_constant__ RefineParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int q = blockIdx.x;
  const int k = p.k, dpad = p.dpad;
  float* q_s = reinterpret_cast<float*>(smem);                                   // [dpad]
  float* misc = reinterpret_cast<float*>(smem + scan::align128(static_cast<size_t>(dpad) * 4));
  uint64_t* scratch = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(misc) + 128);  // [nwarps][32 * S]
  uint64_t* final_list = scratch + static_cast<size_t>(nwarps) * 32 * S;                            // [32 * S]

  // any (query, slice) region that overflowed => adversarial data: let K1 redo this query exactly
  {
    int over = 0;
    for (int sl = tid; sl < p.s_total; sl += blockDim.x)
      over |= p.cand_count[static_cast<size_t>(q) * p.s_total + sl] > static_cast<unsigned int>(p.cap);
    if (__syncthreads_or(over)) {
      if (tid == 0 && blockIdx.y == 0) {
        p.overflow[q] = 1;
        if (p.counts_out) p.counts_out[q] = 0;
      }
      return;
    }
  }
  for (int i = tid; i < dpad; i += blockDim.x) q_s[i] = (i < p.dim) ? __ldg(p.q + static_cast<size_t>(q) * p.dim + i) : 0.0f;
  __syncthreads();
  if (warp == 0) {  // 1/|q| exactly as K1 computes it
    float ss = 0.0f;
    for (int i = lane; i < dpad; i += 32) ss = fmaf(q_s[i], q_s[i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) misc[0] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  __syncthreads();
  const float qinv = misc[0];
  const bool cosine = p.metric == kCosine;
  const int lpr_log2 = p.lpr_log2, lpr = 1 << lpr_log2, G = 32 >> lpr_log2;
  const int g = lane >> lpr_log2, lig = lane & (lpr - 1);
  // k <= 32 * S: a best-first list is S keys per lane.  Scored keys are parked one per lane (`batch`); a
  // full batch is sorted (bitonic network over the lanes) and merged into `acc` in registers.
  uint64_t acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0ull;
  uint64_t batch = 0ull;
  int nb = 0;
  auto flush = [&]() {
    uint64_t b4[S];
#pragma unroll
    for (int s = 0; s < S; ++s) b4[s] = 0ull;
    b4[0] = scan::warp_sort32_desc(batch, lane);
    scan::warp_merge<S>(acc, b4, lane);
    batch = 0ull;
    nb = 0;
  };
  // Work unit = (candidate region, SUB-th part of it): a region's candidates are dealt round-robin to SUB
  // warps, so the ~300 regions of a small batch spread over every warp of every CTA of the query.
  // CU candidates per lane group are in flight at once: entry -> row chunks + norm + id -> score.
  constexpr int CU = 2;
  const int SUB = p.sub;
  const int n_units = p.s_total * SUB;
  for (int unit = blockIdx.y * nwarps + warp; unit < n_units; unit += gridDim.y * nwarps) {
   const int sl = unit / SUB, sub = unit - sl * SUB;
   const int cnt = static_cast<int>(p.cand_count[static_cast<size_t>(q) * p.s_total + sl]);
   const unsigned long long* cand = p.cand + (static_cast<size_t>(q) * p.s_total + sl) * p.cap;
   for (int base = sub * G * CU; base < cnt; base += SUB * G * CU) {
    bool have[CU];
    long long row[CU];
    int sg[CU];
    const unsigned char* rp[CU];
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      const int ci = base + u * G + g;
      have[u] = ci < cnt;
      const unsigned long long ent = have[u] ? cand[ci] : cand[0];
      sg[u] = static_cast<int>(ent >> 32);
      row[u] = static_cast<long long>(ent & 0xFFFFFFFFull);
      WDBX_ASSERT(!have[u] || (sg[u] >= 0 && sg[u] < kMaxSeg && row[u] < p.seg[sg[u]].n_rows));
      rp[u] = p.seg[sg[u]].rows + static_cast<size_t>(row[u]) * p.row_bytes;
    }
    float inx[CU];
    uint32_t gid[CU];
#pragma unroll
    for (int u = 0; u < CU; ++u) {   // requested together with the row chunks, consumed after the reduction
      inx[u] = cosine ? __ldg(p.seg[sg[u]].inv_norm + row[u]) : 1.0f;
      gid[u] = __ldg(p.seg[sg[u]].gids + row[u]);
    }
    float dot[CU];
#pragma unroll
    for (int u = 0; u < CU; ++u) dot[u] = 0.0f;
    // the row chunks of a lane are requested 4 x CU at a time before the first one is consumed (one dependent
    // DRAM round trip per chunk made the refine latency-bound); per candidate the fma order is K1's
    for (int j0 = 0; j0 < p.nch; j0 += 4) {
      uint4 raw[CU][4];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int c = lig + ((j0 + v) << lpr_log2);
          raw[u][v] = (j0 + v < p.nch && c < p.cpr) ? __ldg(reinterpret_cast<const uint4*>(rp[u] + c * 16))
                                                    : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        float acc1 = dot[u];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int c = lig + ((j0 + v) << lpr_log2);
          if (j0 + v < p.nch && c < p.cpr) {
            if (!BF16) {
              const float4 x = make_float4(__uint_as_float(raw[u][v].x), __uint_as_float(raw[u][v].y),
                                           __uint_as_float(raw[u][v].z), __uint_as_float(raw[u][v].w));
              const float4 qv = lds128(q_s + c * 4);
              if (L2) {
                const float d0 = x.x - qv.x, d1 = x.y - qv.y, d2 = x.z - qv.z, d3 = x.w - qv.w;
                acc1 = fmaf(d0, d0, acc1); acc1 = fmaf(d1, d1, acc1); acc1 = fmaf(d2, d2, acc1); acc1 = fmaf(d3, d3, acc1);
              } else {
                acc1 = fmaf(x.x, qv.x, acc1); acc1 = fmaf(x.y, qv.y, acc1); acc1 = fmaf(x.z, qv.z, acc1); acc1 = fmaf(x.w, qv.w, acc1);
              }
            } else {
              const uint32_t w[4] = {raw[u][v].x, raw[u][v].y, raw[u][v].z, raw[u][v].w};
              const float4 qa = lds128(q_s + c * 8);
              const float4 qb = lds128(q_s + c * 8 + 4);
              const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xFFFF0000u);
                if (L2) {
                  const float d0 = lo - qq[2 * i], d1 = hi - qq[2 * i + 1];
                  acc1 = fmaf(d0, d0, acc1); acc1 = fmaf(d1, d1, acc1);
                } else {
                  acc1 = fmaf(lo, qq[2 * i], acc1); acc1 = fmaf(hi, qq[2 * i + 1], acc1);
                }
              }
            }
          }
        }
        dot[u] = acc1;
      }
    }
#pragma unroll
    for (int u = 0; u < CU; ++u) {
      float d = dot[u];
      for (int o = lpr >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(FULL_MASK, d, o);
      float sc = d;
      if (L2) sc = -sc;
      else if (cosine) sc = sc * inx[u] * qinv;
      sc = (sc != sc) ? __int_as_float(0xff800000) : sc;
      const uint64_t key = have[u] ? pack_key(sc, gid[u]) : 0ull;
      unsigned m = __ballot_sync(FULL_MASK, lig == 0 && have[u]);
      while (m) {
        const int src_lane = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t kk = __shfl_sync(FULL_MASK, key, src_lane);
        if (lane == nb) batch = kk;
        if (++nb == 32) flush();
      }
    }
   }
  }
  if (nb > 0) flush();
  // CTA list: binary tree over the warps (registers + shared memory), then -- several CTAs per query --
  // partial list -> global and the last CTA of this query (atomic ticket) folds all of them the same way
  scan::block_tree_merge<S>(acc, scratch, warp, lane, nwarps);
  if (gridDim.y > 1) {
    if (warp == 0) {
      scan::store_list<S>(p.part + (static_cast<size_t>(q) * gridDim.y + blockIdx.y) * k, acc, k, lane);
      __threadfence();
      __syncwarp();
      if (lane == 0) misc[1] = __uint_as_float(atomicAdd(p.tickets + q, 1u));
    }
    __syncthreads();
    if (__float_as_uint(misc[1]) != gridDim.y - 1) return;
    __threadfence();
    scan::grid_merge_fast<S>(p.part + static_cast<size_t>(q) * gridDim.y * k, static_cast<int>(gridDim.y), scratch,
                             final_list, k, warp, lane, nwarps);
  } else if (warp == 0) {
    scan::store_list<S>(final_list, acc, k, lane);
  }
  if (warp == 0) {
    __syncwarp();
    scan::emit_list(final_list, k, lane, p.keys_out ? p.keys_out + static_cast<size_t>(q) * k : nullptr,
                    p.scores_out ? p.scores_out + static_cast<size_t>(q) * k : nullptr,
                    p.gids_out ? p.gids_out + static_cast<size_t>(q) * k : nullptr, p.counts_out ? p.counts_out + q : nullptr);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_bf16() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// [rows, ld] bf16 row-major, logical width `cols`; box = 64 x box_rows, 128B swizzle, OOB -> 0
bool encode_map_bf16(CUtensorMap* map, const void* base, long long rows, int cols, int ld, int box_rows) {
  auto fn = get_encode_bf16();
  if (!fn) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows > 0 ? rows : 1)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, ld] int8 row-major, logical width `cols`; box = 128 bytes x box_rows, 128B swizzle, OOB -> 0
bool encode_map_i8(CUtensorMap* map, const void* base, long long rows, int cols, int ld, int box_rows) {
  auto fn = get_encode_bf16();
  if (!fn) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows > 0 ? rows : 1)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld)};
  cuuint32_t box[2] = {128, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// CTA pairs (cta_group::2) halve the L2 -> shared-memory traffic of the X operand but couple the two
// epilogues; measured on B200 (10M x 768): faster for 2-4 query blocks (B 129..512: 72k vs 64k QPS at
// B=256), equal at 8 blocks, slower beyond (power-capped either way).  WDBX_B200_FILTER_PAIR=0 / 1
// forces never / whenever the block count is even.
bool pair_mode(int n_qblocks) {
  static const int forced = [] {
    const char* v = getenv("WDBX_B200_FILTER_PAIR");
    return v ? (v[0] == '0' ? 0 : 1) : -1;
  }();
  if ((n_qblocks & 1) != 0 || forced == 0) return false;
  return forced == 1 || n_qblocks <= 4;
}

// WDBX_B200_FILTER_SMALL=0 routes small batches through the 128-query kernel (A/B comparisons)
bool small_batch_mode(int B) {
  static const bool on = [] {
    const char* v = getenv("WDBX_B200_FILTER_SMALL");
    return !(v && v[0] == '0');
  }();
  return on && B <= SQ;
}

}  // namespace

// candidate regions per (query, row slice): one per epilogue warp in the small-batch kernel, one per column
// half in the 128-query kernel
int filter_regions_per_slice(int B) { return small_batch_mode(B) ? kSmallWarpRegions : 2; }

int filter_max_k() { return kMaxKFilter; }

int filter_final_cap() { return kFinalCap; }

int filter_ld16(int dim) { return (dim + 7) / 8 * 8; }

// int8 shadow / query operand: row pitch in bytes (a multiple of 16 for TMA)
int filter_ld8(int dim) { return (dim + 15) / 16 * 16; }

// workspace layout (bytes): qb16 [Bpad][ld16] bf16 | q_inv, q_nrm, q_sq, q_bn, q_tn [Bpad] f32 each; Bpad = B rounded up to 128
static int filter_bpad(int B) { return (B + BM - 1) / BM * BM; }
size_t filter_query_workspace_bytes(int B, int dim) {
  const size_t bp = filter_bpad(B);
  return (bp * filter_ld16(dim) * 2 + 15) / 16 * 16 + 5 * bp * 4;
}

// accumulation terms of the error bound (see "error bound" above), relative to |x||q|: tensor-core fp32
// accumulation (dim * 2^-23 * 1.01, truncation allowed) + K1's fp32 dot and score formula + slack
float filter_acc_rel(int dim, int dpad) {
  return static_cast<float>(dim) * 1.21e-7f + (static_cast<float>(dpad) / 32.0f + 40.0f) * 6e-8f + 1e-6f;
}
// l2: relative to |x|^2 + |q|^2 -- rounding of the stored |x|^2 and of |q|^2 (one lane sums dpad/32 squares, then
// a 5-level butterfly), the three additions of the expanded form, K1's direct sum of squared differences
// (2 x (fma chain + butterfly + the rounding of each difference)) and slack
float filter_c_l2(int dpad) {
  return (3.0f * static_cast<float>(dpad) / 32.0f + 64.0f) * 6e-8f + 1e-6f;
}

int filter_slices_for(long long n_rows, int B, int sm_count) {
  const int n_qblocks = (B + BM - 1) / BM;
  const long long n_tiles = (n_rows + BN - 1) / BN;
  long long s = sm_count / n_qblocks;
  if (s < 1) s = 1;
  if (s > n_tiles) s = n_tiles;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

cudaError_t launch_shadow_rows(const float* rows, long long n, int dpad, int ld16, void* dst, float* rres,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  shadow_rows_kernel<<<static_cast<unsigned>((n + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      rows, n, dpad, ld16, static_cast<__nv_bfloat16*>(dst), rres);
  return cudaGetLastError();
}

cudaError_t launch_shadow8_rows(const void* rows, bool src_bf16, long long n, int dpad, int ld8, void* dst, float* sx,
                                float* rres, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  if (src_bf16)
    shadow8_rows_kernel<true><<<grid, wpb * 32, 0, stream>>>(rows, n, dpad, ld8, static_cast<signed char*>(dst), sx, rres);
  else
    shadow8_rows_kernel<false><<<grid, wpb * 32, 0, stream>>>(rows, n, dpad, ld8, static_cast<signed char*>(dst), sx, rres);
  return cudaGetLastError();
}

cudaError_t launch_prep_queries(const float* q, int B, int dim, void* workspace, unsigned int* zero, size_t n_zero,
                                bool small, const unsigned int* done_ctr, unsigned int wait_sn, bool overlap, bool pdl,
                                unsigned int* prep_count, unsigned int* prep_ctas, bool i8, cudaStream_t stream) {
  if (i8) {
    // int8 small-batch operand: [32][ld8] digits | q_inv, q_nrm, q_sq, q_bn, q_tn, s1, s2 [16 each]
    const int ld8 = filter_ld8(dim);
    signed char* qi = static_cast<signed char*>(workspace);
    float* f = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + static_cast<size_t>(2 * SQ_I8) * ld8);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(SQ_I8, 1, 1);      // one-warp CTAs (see below)
    cfg.blockDim = dim3(32, 1, 1);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int pdl_wait = overlap ? 0 : 1;
    if (prep_ctas) *prep_ctas = cfg.gridDim.x;
    return cudaLaunchKernelEx(&cfg, prep_queries_i8_kernel, q, B, dim, ld8, qi, f, f + SQ_I8, f + 2 * SQ_I8, f + 3 * SQ_I8,
                              f + 4 * SQ_I8, f + 5 * SQ_I8, f + 6 * SQ_I8, zero, n_zero, done_ctr, wait_sn, pdl_wait, prep_count);
  }
  const int ld = filter_ld16(dim);
  const int bp = filter_bpad(B);           // layout of the workspace (norm arrays are [bp])
  const int rows = small ? SQ : bp;        // rows the filter's TMA box can touch: 16 for the small-batch kernel
  __nv_bfloat16* qb = static_cast<__nv_bfloat16*>(workspace);
  float* f = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + (static_cast<size_t>(bp) * ld * 2 + 15) / 16 * 16);
  // small-batch path: one-warp CTAs (32 registers x 32 threads) find room on an SM that a filter CTA and a gated K1
  // CTA already share, so the prep of the next search can run while the previous search finishes (engine.cu)
  const int wpb = small ? 1 : 8;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((rows + wpb - 1) / wpb, 1, 1);
  cfg.blockDim = dim3(wpb * 32, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  // overlap: the search orders itself behind search wait_sn through done_ctr only; otherwise it waits for the stream
  const int pdl_wait = overlap ? 0 : 1;
  if (prep_ctas) *prep_ctas = cfg.gridDim.x;
  return cudaLaunchKernelEx(&cfg, prep_queries_kernel, q, B, rows, dim, ld, qb, f, f + bp, f + 2 * bp, f + 3 * bp, f + 4 * bp,
                            zero, n_zero, done_ctr, wait_sn, pdl_wait, prep_count);
}

bool filter_fused_tail(int B) { return small_batch_mode(B); }

namespace {
__global__ void publish_done_kernel(unsigned int* done_ctr, unsigned int sn) { atomicMax(done_ctr, sn); }
}  // namespace

// A numbered search that could not be launched completely must still be marked finished, or the searches behind it
// would each wait (bounded, but seconds) for a number that never comes.
cudaError_t launch_publish_done(unsigned int* done_ctr, unsigned int sn, cudaStream_t stream) {
  publish_done_kernel<<<1, 1, 0, stream>>>(done_ctr, sn);
  return cudaGetLastError();
}

// One launch per segment.  xb = bf16 matrix [n_rows][ld_x] (shadow, or the stored rows of a bf16 engine).
cudaError_t launch_gemm_filter(const void* xb, int ld_x, const float* rres, const SegDesc& seg, int seg_index, int dim,
                               const void* workspace, int B, int k, int metric, float acc_rel, float c_l2, int n_slices,
                               unsigned long long* cand,
                               unsigned int* cand_count, unsigned int* lower_glob, unsigned int* lower_list, int cap,
                               int slice_base, int s_total, const FilterTail* tail, bool pdl, cudaStream_t stream) {
  if (seg.n_rows <= 0 || n_slices <= 0) return cudaSuccess;
  if (tail != nullptr && tail->tile_ctr == nullptr) return cudaErrorInvalidValue;
  const bool i8 = tail != nullptr && tail->rowscale != nullptr;
  static std::atomic<unsigned long long> attr_done{0ull};
  const cudaError_t attr_err = once_per_device(attr_done, [&] {
    cudaError_t err = cudaSuccess;
    auto set = [&](auto kern, size_t bytes) {
      if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    };
    set(gemm_filter_kernel<kCosine, 1>, filter_smem(1)); set(gemm_filter_kernel<kIP, 1>, filter_smem(1));
    set(gemm_filter_kernel<kL2, 1>, filter_smem(1));
    set(gemm_filter_kernel<kCosine, 2>, filter_smem(2)); set(gemm_filter_kernel<kIP, 2>, filter_smem(2));
    set(gemm_filter_kernel<kL2, 2>, filter_smem(2));
    set(gemm_filter_small_kernel<kCosine, false>, kFilterSmallSmem); set(gemm_filter_small_kernel<kIP, false>, kFilterSmallSmem);
    set(gemm_filter_small_kernel<kL2, false>, kFilterSmallSmem);
    set(gemm_filter_small_kernel<kCosine, true>, kFilterSmallSmemI8); set(gemm_filter_small_kernel<kIP, true>, kFilterSmallSmemI8);
    set(gemm_filter_small_kernel<kL2, true>, kFilterSmallSmemI8);
    return err;
  });
  if (attr_err != cudaSuccess) return attr_err;
  const int ld = filter_ld16(dim);
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  const int bp = filter_bpad(B);
  const float* f = reinterpret_cast<const float*>(ws + (static_cast<size_t>(bp) * ld * 2 + 15) / 16 * 16);
  // CTA pairs need an even number of query blocks (the pair = two adjacent blocks on one row slice)
  const int n_qblocks = (B + BM - 1) / BM;
  const bool small = small_batch_mode(B);
  if (small && tail == nullptr) return cudaErrorInvalidValue;
  const int ncta = (!small && pair_mode(n_qblocks)) ? 2 : 1;
  CUtensorMap tm_x, tm_q;
  if (i8) {
    if (!encode_map_i8(&tm_x, xb, seg.n_rows, dim, ld_x, BN) || !encode_map_i8(&tm_q, ws, 2 * SQ_I8, dim, filter_ld8(dim), 2 * SQ_I8))
      return cudaErrorInvalidValue;
  } else if (!encode_map_bf16(&tm_x, xb, seg.n_rows, dim, ld_x, BN / ncta) ||
             !encode_map_bf16(&tm_q, ws, bp, dim, ld, small ? SQ : BM)) {
    return cudaErrorInvalidValue;
  }
  FilterParams p;
  p.inv_norm = seg.inv_norm;
  p.sqnorm = seg.sqnorm;
  p.tomb = seg.tomb;
  p.q_inv = f;
  p.q_nrm = f + bp;
  p.q_sq = f + 2 * bp;
  p.q_bn = f + 3 * bp;
  p.q_tn = f + 4 * bp;
  p.rres = rres;
  p.n_rows = seg.n_rows;
  p.B = B;
  p.k = k;
  p.n_kblocks = i8 ? (dim + 127) / 128 : (dim + BK - 1) / BK;
  p.rowscale = i8 ? tail->rowscale : nullptr;
  p.q_s1 = nullptr;
  p.q_s2 = nullptr;
  if (i8) {   // the int8 prep's workspace layout (launch_prep_queries)
    const float* fi = reinterpret_cast<const float*>(ws + static_cast<size_t>(2 * SQ_I8) * filter_ld8(dim));
    p.q_inv = fi;
    p.q_nrm = fi + SQ_I8;
    p.q_sq = fi + 2 * SQ_I8;
    p.q_bn = fi + 3 * SQ_I8;
    p.q_tn = fi + 4 * SQ_I8;
    p.q_s1 = fi + 5 * SQ_I8;
    p.q_s2 = fi + 6 * SQ_I8;
  }
  p.n_tiles = static_cast<int>((seg.n_rows + BN - 1) / BN);
  p.n_slices = n_slices;
  p.seg = seg_index;
  p.acc_rel = acc_rel;
  p.c_l2 = c_l2;
  p.cand = cand;
  p.cand_count = cand_count;
  p.lower_glob = lower_glob;
  p.lower_list = lower_list;
  p.part_max = tail != nullptr ? tail->part_max : nullptr;
  p.cap = cap;
  p.slice_base = slice_base;
  p.s_total = s_total;
  p.tile_ctr = tail != nullptr ? tail->tile_ctr + seg_index : nullptr;
  p.prep_count = tail != nullptr ? tail->prep_count : nullptr;
  p.prep_target = tail != nullptr ? tail->prep_target : 0u;
  {
    const char* tv = getenv("WDBX_B200_FILTER_TRACE");
    p.trace = (tv && tv[0] == '1') ? 1 : 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_qblocks, n_slices, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = small ? (i8 ? kFilterSmallSmemI8 : kFilterSmallSmem) : filter_smem(ncta);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  if (small) {
    SmallTail tp;
    memset(&tp, 0, sizeof(tp));
    tp.q = tail->q;
    tp.rows = seg.rows;
    tp.gids = seg.gids;
    tp.allow = seg.allow;
    tp.dim = dim;
    tp.dpad = tail->dpad;
    tp.row_bytes = tail->dpad * tail->elem_bytes;
    tp.cpr = tp.row_bytes / 16;
    tp.lpr_log2 = tail->lpr_log2;
    tp.nch = tail->nch;
    tp.bf16 = tail->elem_bytes == 2 ? 1 : 0;
    tp.min_score = tail->min_score;
    tp.overflow = tail->overflow;
    tp.fin_keys = tail->fin_keys;
    tp.fin_count = tail->fin_count;
    tp.ticket = tail->ticket;
    tp.xchg = tail->xchg;
    tp.done_ctr = tail->done_ctr;
    tp.done_sn = tail->done_sn;
    tp.keys_out = tail->keys_out;
    tp.scores_out = tail->scores_out;
    tp.gids_out = tail->gids_out;
    tp.counts_out = tail->counts_out;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = pdl ? 1 : 0;
    auto go = [&](auto kern) { return cudaLaunchKernelEx(&cfg, kern, tm_x, tm_q, p, tp); };
    if (i8) {
      if (metric == kCosine) return go(gemm_filter_small_kernel<kCosine, true>);
      if (metric == kL2) return go(gemm_filter_small_kernel<kL2, true>);
      return go(gemm_filter_small_kernel<kIP, true>);
    }
    if (metric == kCosine) return go(gemm_filter_small_kernel<kCosine, false>);
    if (metric == kL2) return go(gemm_filter_small_kernel<kL2, false>);
    return go(gemm_filter_small_kernel<kIP, false>);
  }
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.numAttrs = ncta == 2 ? 1 : 0;
  auto go = [&](auto kern) { return cudaLaunchKernelEx(&cfg, kern, tm_x, tm_q, p); };
  if (ncta == 2) {
    if (metric == kCosine) return go(gemm_filter_kernel<kCosine, 2>);
    if (metric == kL2) return go(gemm_filter_kernel<kL2, 2>);
    return go(gemm_filter_kernel<kIP, 2>);
  }
  if (metric == kCosine) return go(gemm_filter_kernel<kCosine, 1>);
  if (metric == kL2) return go(gemm_filter_kernel<kL2, 1>);
  return go(gemm_filter_kernel<kIP, 1>);
}

static int warps_per_cta_refine() { return 8; }

cudaError_t launch_refine_topk(const SegDesc* segs, int n_seg, const float* q, int B, int dim, int dpad, int elem_bytes,
                               int lpr_log2, int nch, int k, int metric, const unsigned long long* cand,
                               const unsigned int* cand_count, int cap, int s_total, int* overflow, int ctas_per_query,
                               uint64_t* part, unsigned int* tickets, uint64_t* keys_out, float* scores_out,
                               long long* gids_out, int* counts_out, cudaStream_t stream) {
  RefineParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < n_seg; ++s) p.seg[s] = segs[s];
  p.q = q;
  p.cand = cand;
  p.cand_count = cand_count;
  p.overflow = overflow;
  p.part = part;
  p.tickets = tickets;
  {
    // enough work units that every warp of the launch gets one, whatever the number of regions
    const int total_warps = warps_per_cta_refine() * (ctas_per_query > 1 ? ctas_per_query : 1);
    int sub = (total_warps + s_total - 1) / (s_total > 0 ? s_total : 1);
    p.sub = sub < 4 ? 4 : (sub > 32 ? 32 : sub);
  }
  p.B = B;
  p.dim = dim;
  p.dpad = dpad;
  p.row_bytes = dpad * elem_bytes;
  p.cpr = p.row_bytes / 16;
  p.lpr_log2 = lpr_log2;
  p.nch = nch;
  p.k = k;
  p.metric = metric;
  p.cap = cap;
  p.s_total = s_total;
  p.keys_out = keys_out;
  p.scores_out = scores_out;
  p.gids_out = gids_out;
  p.counts_out = counts_out;
  const int warps = warps_per_cta_refine();
  const int S = k <= 32 ? 1 : 4;
  const size_t smem = ((static_cast<size_t>(dpad) * 4 + 127) & ~static_cast<size_t>(127)) + 128 +
                      (static_cast<size_t>(warps) * 32 + 32) * S * 8;
  const bool bf16 = elem_bytes == 2, l2 = metric == kL2;
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    kern<<<dim3(B, ctas_per_query > 1 ? ctas_per_query : 1, 1), warps * 32, smem, stream>>>(p);
  };
#define WDBX_REFINE(SS)                                                                                   \
  do {                                                                                                    \
    if (bf16) { if (l2) go(refine_topk_kernel<true, true, SS>); else go(refine_topk_kernel<true, false, SS>); }   \
    else { if (l2) go(refine_topk_kernel<false, true, SS>); else go(refine_topk_kernel<false, false, SS>); }      \
  } while (0)
  if (S == 1) WDBX_REFINE(1);
  else WDBX_REFINE(4);
#undef WDBX_REFINE
  return cudaGetLastError();
}

}  // namespace wdbx
