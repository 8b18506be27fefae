// scan_topk.cu -- K1: HBM-bound streaming exact scan with a fused register top-k (sm_100a).
//
// Replaces the hot loop of FaissIndex.search (wdbx/core/indexing.py:1013, IndexFlatIP.search:
// N*D fp32 multiply-adds + a heap) and, through the fused last-block merge, the per-shard
// loop + sort of VectorStore.search (wdbx/core/vector_store.py:323-330).
//
// Design (one CTA per SM, persistent):
//   * every warp owns a private ring of `stages` shared-memory buffers and streams its own
//     tiles of `tile_rows` consecutive rows with 1-D TMA bulk copies (cp.async.bulk, SASS
//     UBLKCP) completing on the warp's own mbarriers -- no producer warp, no CTA-wide barrier
//     in the steady state; tiles are dealt round-robin over all warps of the grid so the
//     chip sweeps the matrix sequentially;
//   * a tile is read from shared memory with 128-bit loads; `lanes per row` (power of two)
//     lanes cooperate on one row and each lane group keeps U rows in flight so the query
//     chunk read from shared memory is amortised over U rows;
//   * scores never leave registers: each warp keeps a sorted top-k list distributed over its
//     lanes (WarpTopK) and rejects almost every row with one compare against the cached k-th
//     key; gid and tombstone bit are only fetched for the rare candidates;
//   * the warps of a CTA merge through shared memory, the CTA writes k keys, and the last CTA
//     to finish (atomic ticket) merges all CTA lists and emits keys / scores / gids / count.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <mutex>

namespace wdbx {

namespace {

constexpr int kMaxWarps = 16;

__device__ __forceinline__ size_t align128(size_t x) { return (x + 127) & ~static_cast<size_t>(127); }

// Fold `count` keys at src (chunks of 32, chunk index first, first+stride, ...) into list M.
template <int KS, bool GLOBAL>
__device__ __forceinline__ void absorb_keys(WarpTopK<KS>& M, const uint64_t* src, int count, int first, int stride,
                                            int k, int lane) {
  auto load = [&](int c) -> uint64_t {
    const int idx = c * 32 + lane;
    if (idx >= count) return 0ull;
    if (GLOBAL) return __ldcg(reinterpret_cast<const unsigned long long*>(src + idx));
    return src[idx];
  };
  const int nchunks = (count + 31) >> 5;
  int c = first;
  uint64_t cur = (c < nchunks) ? load(c) : 0ull;
  while (c < nchunks) {
    const int cn = c + stride;
    const uint64_t nxt = (cn < nchunks) ? load(cn) : 0ull;
    unsigned m = __ballot_sync(FULL_MASK, cur > M.thr);
    while (m) {
      const int src_lane = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t kk = __shfl_sync(FULL_MASK, cur, src_lane);
      M.offer(kk, k, lane);
    }
    cur = nxt;
    c = cn;
  }
}

struct TileLoc {
  int seg;
  long long row0;
  int nrows;
};

__device__ __forceinline__ TileLoc locate_tile(const ScanParams& p, long long t, int& cursor) {
  while (t >= p.tile_end[cursor]) ++cursor;  // t < total_tiles guaranteed by the caller
  const long long tb = cursor ? p.tile_end[cursor - 1] : 0ll;
  TileLoc L;
  L.seg = cursor;
  L.row0 = (t - tb) * p.tile_rows;
  const long long rem = p.seg[cursor].n_rows - L.row0;
  L.nrows = rem < p.tile_rows ? static_cast<int>(rem) : p.tile_rows;
  return L;
}

template <int U, int KS, bool BF16, bool L2>
__global__ void __launch_bounds__((KS > 4 ? 8 : kMaxWarps) * 32, 1) scan_topk_kernel(const __grid_constant__ ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int qi = blockIdx.y;  // one query per grid.y slice
  const int k = p.k;

  // ---- shared memory carve-up
  float* q_s = reinterpret_cast<float*>(smem);  // [dpad], zero padded
  size_t off = align128(static_cast<size_t>(p.dpad) * 4);
  float* misc_s = reinterpret_cast<float*>(smem + off);  // [0] = 1/|q|
  int* flag_s = reinterpret_cast<int*>(smem + off + 16);
  off += 128;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + off);
  off += align128(static_cast<size_t>(nwarps) * p.stages * 8);
  unsigned char* stage_area = smem + off;  // nwarps * stages * stage_bytes, reused for the merges

  const float* qg = p.q + static_cast<size_t>(qi) * p.dim;
  for (int i = tid; i < p.dpad; i += blockDim.x) q_s[i] = (i < p.dim) ? __ldg(qg + i) : 0.0f;
  if (lane == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(smem_u32(mbar + warp * p.stages + s), 1);
  }
  fence_mbar_init();
  __syncthreads();
  if (warp == 0) {
    float ss = 0.0f;
    for (int i = lane; i < p.dpad; i += 32) ss = fmaf(q_s[i], q_s[i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) misc_s[0] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  __syncthreads();
  const bool cosine = (p.metric == kCosine);
  const float qinv = misc_s[0];

  // ---- lane geometry
  const int lpr_log2 = p.lpr_log2;
  const int lpr = 1 << lpr_log2;
  const int G = 32 >> lpr_log2;  // row groups per warp
  const int g = lane >> lpr_log2;
  const int lig = lane & (lpr - 1);
  const int nch = p.nch, cpr = p.cpr, row_bytes = p.row_bytes;

  WarpTopK<KS> L;
  L.reset();
  float thr_f = __int_as_float(0xff800000);  // -inf while the list is not full

  const long long gw = static_cast<long long>(blockIdx.x) * nwarps + warp;
  const long long GW = static_cast<long long>(gridDim.x) * nwarps;
  const long long total = p.total_tiles;
  unsigned char* my_stage = stage_area + static_cast<size_t>(warp) * p.stages * p.stage_bytes;
  const uint32_t my_bar = smem_u32(mbar + warp * p.stages);
  const uint64_t pol = policy_evict_first();
  const bool use_hint = p.evict_first != 0;

  int cur_issue = 0;
  long long t_issue = gw;
  auto issue = [&](int stage) {
    const TileLoc T = locate_tile(p, t_issue, cur_issue);
    if (lane == 0) {
      const uint32_t bytes = static_cast<uint32_t>(T.nrows) * row_bytes;
      const uint32_t bar = my_bar + stage * 8;
      const unsigned char* src = p.seg[T.seg].rows + static_cast<size_t>(T.row0) * row_bytes;
      mbar_expect_tx(bar, bytes);
      if (use_hint) bulk_g2s_hint(smem_u32(my_stage + static_cast<size_t>(stage) * p.stage_bytes), src, bytes, bar, pol);
      else bulk_g2s(smem_u32(my_stage + static_cast<size_t>(stage) * p.stage_bytes), src, bytes, bar);
    }
    t_issue += GW;
  };
  for (int s = 0; s < p.stages; ++s)
    if (t_issue < total) issue(s);

  // compute-side cursor runs one tile ahead so the inv-norm loads of tile i+1 are in flight
  // while tile i is processed
  int cur_next = 0;
  TileLoc Tn;
  float inr_n[U];
#pragma unroll
  for (int u = 0; u < U; ++u) inr_n[u] = 0.0f;
  auto prefetch = [&](long long t) {
    Tn = locate_tile(p, t, cur_next);
    if (cosine) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = u * G + g;
        inr_n[u] = (r < Tn.nrows) ? __ldg(p.seg[Tn.seg].inv_norm + Tn.row0 + r) : 0.0f;
      }
    }
  };
  if (gw < total) prefetch(gw);

  uint32_t it = 0;
  int stage = 0;
  uint32_t parity = 0;
  for (long long t = gw; t < total; t += GW, ++it) {
    const TileLoc T = Tn;
    float inr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) inr[u] = inr_n[u];
    if (t + GW < total) prefetch(t + GW);

    mbar_wait(my_bar + stage * 8, parity);
    const unsigned char* sb = my_stage + static_cast<size_t>(stage) * p.stage_bytes;

    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0f;

#pragma unroll 2
    for (int j = 0; j < nch; ++j) {
      const int c = lig + (j << lpr_log2);
      if (c < cpr) {
        if (!BF16) {
          const float4 qv = lds128(q_s + c * 4);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float4 x = lds128(sb + static_cast<size_t>(u * G + g) * row_bytes + c * 16);
            if (L2) {
              const float d0 = x.x - qv.x, d1 = x.y - qv.y, d2 = x.z - qv.z, d3 = x.w - qv.w;
              acc[u] = fmaf(d0, d0, acc[u]);
              acc[u] = fmaf(d1, d1, acc[u]);
              acc[u] = fmaf(d2, d2, acc[u]);
              acc[u] = fmaf(d3, d3, acc[u]);
            } else {
              acc[u] = fmaf(x.x, qv.x, acc[u]);
              acc[u] = fmaf(x.y, qv.y, acc[u]);
              acc[u] = fmaf(x.z, qv.z, acc[u]);
              acc[u] = fmaf(x.w, qv.w, acc[u]);
            }
          }
        } else {
          const float4 qa = lds128(q_s + c * 8);
          const float4 qb = lds128(q_s + c * 8 + 4);
          const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint4 raw = *reinterpret_cast<const uint4*>(sb + static_cast<size_t>(u * G + g) * row_bytes + c * 16);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float lo = __uint_as_float(w[i] << 16);
              const float hi = __uint_as_float(w[i] & 0xFFFF0000u);
              if (L2) {
                const float d0 = lo - qq[2 * i], d1 = hi - qq[2 * i + 1];
                acc[u] = fmaf(d0, d0, acc[u]);
                acc[u] = fmaf(d1, d1, acc[u]);
              } else {
                acc[u] = fmaf(lo, qq[2 * i], acc[u]);
                acc[u] = fmaf(hi, qq[2 * i + 1], acc[u]);
              }
            }
          }
        }
      }
    }
    // reduce over the lanes of a row group
    for (int o = lpr >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] += __shfl_xor_sync(FULL_MASK, acc[u], o);
    }

#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = acc[u];
      if (L2) s = -s;
      else if (cosine) s = s * inr[u] * qinv;
      s = (s != s) ? __int_as_float(0xff800000) : s;
      const int r = u * G + g;
      const bool pass = (lig == 0) && (r < T.nrows) && (s >= thr_f);
      unsigned m = __ballot_sync(FULL_MASK, pass);
      while (m) {
        const int src_lane = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(FULL_MASK, s, src_lane);
        const long long row = T.row0 + u * G + (src_lane >> lpr_log2);
        const uint32_t* tomb = p.seg[T.seg].tomb;
        if (tomb != nullptr && ((__ldg(tomb + (row >> 5)) >> (row & 31)) & 1u)) continue;
        const uint32_t gid = __ldg(p.seg[T.seg].gids + row);
        const uint64_t key = pack_key(sv, gid);
        if (key > L.thr) {
          L.insert(key, k, lane);
          thr_f = L.thr ? key_score(L.thr) : __int_as_float(0xff800000);
        }
      }
    }

    __syncwarp();
    if (t_issue < total) issue(stage);
    if (++stage == p.stages) {
      stage = 0;
      parity ^= 1u;
    }
  }

  // ---- CTA merge: every warp publishes its list, warp 0 folds them
  __syncthreads();
  uint64_t* mlist = reinterpret_cast<uint64_t*>(stage_area);  // [nwarps][k]
  L.store(mlist + static_cast<size_t>(warp) * k, k, lane);
  __syncthreads();
  uint64_t* cand_q = p.cand + static_cast<size_t>(qi) * gridDim.x * k;
  if (warp == 0) {
    WarpTopK<KS> M;
    M.reset();
    absorb_keys<KS, false>(M, mlist, nwarps * k, 0, 1, k, lane);
    M.store(cand_q + static_cast<size_t>(blockIdx.x) * k, k, lane);
  }

  // ---- last CTA to finish merges all CTA lists of this query
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(p.counters + qi, 1u);
    *flag_s = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*flag_s == 0) return;
  __threadfence();
  {
    WarpTopK<KS> M;
    M.reset();
    absorb_keys<KS, true>(M, cand_q, static_cast<int>(gridDim.x) * k, warp, nwarps, k, lane);
    M.store(mlist + static_cast<size_t>(warp) * k, k, lane);
  }
  __syncthreads();
  if (warp == 0) {
    WarpTopK<KS> M;
    M.reset();
    absorb_keys<KS, false>(M, mlist, nwarps * k, 0, 1, k, lane);
    emit_outputs<KS>(M, k, lane, p.keys_out ? p.keys_out + static_cast<size_t>(qi) * k : nullptr,
                     p.scores_out ? p.scores_out + static_cast<size_t>(qi) * k : nullptr,
                     p.gids_out ? p.gids_out + static_cast<size_t>(qi) * k : nullptr,
                     p.counts_out ? p.counts_out + qi : nullptr);
    if (lane == 0) p.counters[qi] = 0u;  // re-arm for the next launch
  }
}

template <int U, int KS, bool BF16, bool L2>
cudaError_t launch_one(const ScanParams& p, const ScanPlan& plan, cudaStream_t stream) {
  auto kern = scan_topk_kernel<U, KS, BF16, L2>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  dim3 grid(plan.grid, p.B, 1), block(plan.warps * 32, 1, 1);
  kern<<<grid, block, plan.smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

template <int U, int KS>
cudaError_t launch_uk(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  const bool l2 = p.metric == kL2;
  if (bf16) return l2 ? launch_one<U, KS, true, true>(p, plan, stream) : launch_one<U, KS, true, false>(p, plan, stream);
  return l2 ? launch_one<U, KS, false, true>(p, plan, stream) : launch_one<U, KS, false, false>(p, plan, stream);
}

template <int U>
cudaError_t launch_u(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  if (plan.KS == 4) return launch_uk<U, 4>(p, plan, bf16, stream);
  return launch_uk<U, 32>(p, plan, bf16, stream);
}

}  // namespace

int scan_plan(int dim, int dpad, int elem_bytes, int k, int sm_count, const ScanTuning& tune, ScanPlan* plan) {
  if (dim <= 0 || dpad < dim || k <= 0 || k > kMaxK) return -1;
  const int row_bytes = dpad * elem_bytes;
  if (row_bytes % 16 != 0) return -1;
  const int cpr = row_bytes / 16;
  int lpr = 32;
  while (lpr > 1 && cpr < 3 * lpr) lpr >>= 1;
  int lpr_log2 = 0;
  while ((1 << lpr_log2) < lpr) ++lpr_log2;
  const int G = 32 / lpr;
  const int max_warps = (k <= 128) ? kMaxWarps : 8;  // the 1024-entry register list needs 255 regs/thread
  int warps = tune.warps > 0 ? std::min(tune.warps, max_warps) : 8;
  int stages = tune.stages > 0 ? std::min(tune.stages, 8) : 2;
  // budget for the stage area
  const size_t fixed = ((static_cast<size_t>(dpad) * 4 + 127) & ~static_cast<size_t>(127)) + 128 +
                       ((static_cast<size_t>(warps) * stages * 8 + 127) & ~static_cast<size_t>(127));
  const size_t budget = 200 * 1024;
  int U = tune.rows_unroll > 0 ? tune.rows_unroll : 4;
  if (U != 1 && U != 2 && U != 4) U = 4;
  auto stage_bytes_for = [&](int u) { return static_cast<size_t>(u) * G * row_bytes; };
  while (U > 1 && fixed + stage_bytes_for(U) * warps * stages > budget) U >>= 1;
  while (warps > 1 && fixed + stage_bytes_for(U) * warps * stages > budget) warps >>= 1;
  while (stages > 1 && fixed + stage_bytes_for(U) * warps * stages > budget) --stages;
  if (fixed + stage_bytes_for(U) * warps * stages > budget) return -4;  // row too large
  plan->lpr_log2 = lpr_log2;
  plan->nch = (cpr + lpr - 1) / lpr;
  plan->U = U;
  plan->KS = (k <= 128) ? 4 : 32;
  plan->tile_rows = U * G;
  plan->stage_bytes = static_cast<int>(stage_bytes_for(U));
  plan->stages = stages;
  plan->warps = warps;
  plan->grid = tune.grid > 0 ? tune.grid : sm_count;
  const size_t stage_area = std::max(stage_bytes_for(U) * warps * stages, static_cast<size_t>(warps) * k * 8);
  plan->smem_bytes = fixed + stage_area;
  plan->queries_per_block = 1;
  if (plan->smem_bytes > 227 * 1024) return -4;
  return 0;
}

cudaError_t launch_scan_topk(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  switch (plan.U) {
    case 1: return launch_u<1>(p, plan, bf16, stream);
    case 2: return launch_u<2>(p, plan, bf16, stream);
    default: return launch_u<4>(p, plan, bf16, stream);
  }
}

}  // namespace wdbx
