// scan_topk_kernel.cuh -- K1: HBM-bound streaming exact scan with a fused register top-k (sm_100a).
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
//     lanes cooperate on one row and each lane group keeps U rows in flight, scored against QB
//     queries at once (QB x U register block: the row chunk and the query chunk read from
//     shared memory are each reused QB resp. U times);
//   * scores never leave registers: per query each warp keeps a sorted top-k list distributed
//     over its lanes (WarpTopK) and rejects almost every row with one compare against the
//     cached k-th score; gid and tombstone bit are only fetched for the rare candidates;
//   * the warps of a CTA merge through shared memory, the CTA writes k keys per query, and the
//     last CTA to finish (atomic ticket) merges all CTA lists and emits keys/scores/gids/count.
#pragma once
#include "common.cuh"
#include "kernels.h"

#include <mutex>

namespace wdbx {
namespace scan {

constexpr int kMaxWarps = 16;

__device__ __forceinline__ size_t align128(size_t x) { return (x + 127) & ~static_cast<size_t>(127); }

// ---- warp-owned sorted top-k list in SHARED memory (descending, k entries, 0 = empty).
// Lists live in shared memory (not registers) so the list of any query can be addressed with a
// run-time index: one candidate path serves all QB x U values of a tile, and k up to 1024 costs no
// registers.  Inserts are rare (expected ~k*ln(rows_per_warp/k) per warp), the streaming loop
// only compares against a cached k-th score.
// Precondition: key > list[k-1].  Returns the new k-th key.  All lanes must call (warp-uniform).
__device__ __forceinline__ uint64_t list_insert(uint64_t* list, int k, uint64_t key, int lane) {
  for (int base = ((k - 1) >> 5) << 5; base >= 0; base -= 32) {
    const int idx = base + lane;
    const uint64_t v = idx < k ? list[idx] : 0ull;
    const bool gt = v > key;  // sorted: the greater entries are a prefix of the chunk
    const int c = __popc(__ballot_sync(FULL_MASK, gt));
    __syncwarp();
    if (!gt && idx + 1 < k) list[idx + 1] = v;  // shift the tail right by one
    if (c > 0 || base == 0) {
      if (lane == 0) list[base + c] = key;       // base + c < k because key > list[k-1]
      break;
    }
  }
  __syncwarp();
  return list[k - 1];
}

// Fold `count` keys at src (chunks of 32, chunk index first, first+stride, ...) into `list`.
template <bool GLOBAL>
__device__ __forceinline__ void absorb_keys(uint64_t* list, int k, uint64_t& thr, const uint64_t* src, int count,
                                            int first, int stride, int lane) {
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
    unsigned m = __ballot_sync(FULL_MASK, cur > thr);
    while (m) {
      const int src_lane = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t kk = __shfl_sync(FULL_MASK, cur, src_lane);
      if (kk > thr) thr = list_insert(list, k, kk, lane);
    }
    cur = nxt;
    c = cn;
  }
}

__device__ __forceinline__ void list_clear(uint64_t* list, int k, int lane) {
  for (int i = lane; i < k; i += 32) list[i] = 0ull;
  __syncwarp();
}

// Decode a finished list into the user-visible outputs of one query.
__device__ __forceinline__ void emit_list(const uint64_t* list, int k, int lane, uint64_t* keys_out, float* scores_out,
                                          long long* gids_out, int* count_out) {
  int cnt = 0;
  for (int base = 0; base < k; base += 32) {
    const int idx = base + lane;
    const uint64_t key = idx < k ? list[idx] : 0ull;
    cnt += __popc(__ballot_sync(FULL_MASK, key != 0ull));
    if (idx < k) {
      if (keys_out) keys_out[idx] = key;
      if (scores_out) scores_out[idx] = key ? key_score(key) : __int_as_float(0xff800000);
      if (gids_out) gids_out[idx] = key ? static_cast<long long>(key_gid(key)) : -1ll;
    }
  }
  if (count_out && lane == 0) *count_out = cnt;
}

// ---- k <= 32: a list is one key per lane (descending, 0 = empty), merged in registers.
// c[i] = max(a[i], b[31-i]) holds the 32 largest keys of the union and is bitonic; five
// compare-exchange stages sort it descending.
__device__ __forceinline__ uint64_t warp_merge32(uint64_t a, uint64_t b, int lane) {
  const uint64_t br = __shfl_sync(FULL_MASK, b, 31 - lane);
  uint64_t c = a > br ? a : br;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const uint64_t o = __shfl_xor_sync(FULL_MASK, c, s);
    const bool keep_max = (lane & s) == 0;
    c = keep_max ? (c > o ? c : o) : (c < o ? c : o);
  }
  return c;
}

// full bitonic sort (descending) of one key per lane
__device__ __forceinline__ uint64_t warp_sort32_desc(uint64_t v, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const uint64_t o = __shfl_xor_sync(FULL_MASK, v, j);
      const bool desc_block = (lane & k2) == 0;   // the last stage (k2 == 32) is one descending block
      const bool low_lane = (lane & j) == 0;
      const bool keep_max = desc_block == low_lane;
      v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
    }
  }
  return v;
}

// ---- k <= 32*S: a list is S keys per lane, element e = slot*32 + lane, sorted descending.
// Same bitonic scheme on 32*S elements: combine with the reversed second list, compare-exchange
// across slots (strides 32*S/2 .. 32), then across lanes (16 .. 1).
template <int S>
__device__ __forceinline__ void warp_merge(uint64_t (&a)[S], const uint64_t (&b)[S], int lane) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const uint64_t br = __shfl_sync(FULL_MASK, b[S - 1 - s], 31 - lane);
    a[s] = a[s] > br ? a[s] : br;
  }
#pragma unroll
  for (int st = S / 2; st >= 1; st >>= 1) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if ((s & st) == 0) {
        const uint64_t hi = a[s] > a[s + st] ? a[s] : a[s + st];
        const uint64_t lo = a[s] > a[s + st] ? a[s + st] : a[s];
        a[s] = hi;
        a[s + st] = lo;
      }
    }
  }
#pragma unroll
  for (int st = 16; st > 0; st >>= 1) {
    const bool keep_max = (lane & st) == 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const uint64_t o = __shfl_xor_sync(FULL_MASK, a[s], st);
      a[s] = keep_max ? (a[s] > o ? a[s] : o) : (a[s] < o ? a[s] : o);
    }
  }
}

template <int S, bool GLOBAL>
__device__ __forceinline__ void load_list(uint64_t (&r)[S], const uint64_t* src, int k, int lane, bool valid = true) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int e = s * 32 + lane;
    if (!valid || e >= k) r[s] = 0ull;
    else r[s] = GLOBAL ? __ldcg(reinterpret_cast<const unsigned long long*>(src + e)) : src[e];
  }
}

template <int S>
__device__ __forceinline__ void store_list(uint64_t* dst, const uint64_t (&r)[S], int k, int lane) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int e = s * 32 + lane;
    if (e < k) dst[e] = r[s];
  }
}

// Binary tree over the warps of the CTA through shared memory ([nwarps][32*S] keys); every thread of
// the CTA must call it; the result is valid in warp 0.
template <int S>
__device__ __forceinline__ void block_tree_merge(uint64_t (&mine)[S], uint64_t* scratch, int warp, int lane, int nwarps) {
  for (int stride = 1; stride < nwarps; stride <<= 1) {
#pragma unroll
    for (int s = 0; s < S; ++s) scratch[(warp * S + s) * 32 + lane] = mine[s];
    __syncthreads();
    if ((warp & (2 * stride - 1)) == 0 && warp + stride < nwarps) {
      uint64_t other[S];
#pragma unroll
      for (int s = 0; s < S; ++s) other[s] = scratch[((warp + stride) * S + s) * 32 + lane];
      warp_merge<S>(mine, other, lane);
    }
    __syncthreads();
  }
}

// fast epilogue pieces (k <= 32*S), shared by the CTA merge, the last-CTA merge and the exchange merge
template <int S>
__device__ __forceinline__ void cta_merge_fast(const uint64_t* my_list_b, uint64_t* scratch, uint64_t* out_global, int k,
                                               int warp, int lane, int nwarps) {
  uint64_t mine[S];
  load_list<S, false>(mine, my_list_b, k, lane);
  block_tree_merge<S>(mine, scratch, warp, lane, nwarps);
  if (warp == 0) store_list<S>(out_global, mine, k, lane);
}

template <int S>
__device__ __forceinline__ void grid_merge_fast(const uint64_t* cand_q, int n_lists, uint64_t* scratch, uint64_t* final_list,
                                                int k, int warp, int lane, int nwarps) {
  uint64_t acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0ull;
  constexpr int INFLIGHT = S == 1 ? 4 : 2;  // CTA lists in flight per warp
  for (int c0 = warp; c0 < n_lists; c0 += INFLIGHT * nwarps) {
    uint64_t v[INFLIGHT][S];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) {
      const int c = c0 + j * nwarps;
      load_list<S, true>(v[j], cand_q + static_cast<size_t>(c < n_lists ? c : 0) * k, k, lane, c < n_lists);
    }
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) warp_merge<S>(acc, v[j], lane);
  }
  block_tree_merge<S>(acc, scratch, warp, lane, nwarps);
  if (warp == 0) store_list<S>(final_list, acc, k, lane);
}

template <int S>
__device__ __forceinline__ void peers_merge_fast(const uint64_t* my_xbuf_slot_b, int world, uint64_t* final_list, int k,
                                                 int lane) {
  // my_xbuf_slot_b: list of peer r for this query at my_xbuf_slot_b + r * kXchgMaxB * kXchgMaxK
  uint64_t acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0ull;
  for (int r0 = 0; r0 < world; r0 += 2) {
    uint64_t v[2][S];
#pragma unroll
    for (int j = 0; j < 2; ++j)
      load_list<S, true>(v[j], my_xbuf_slot_b + static_cast<size_t>(r0 + j < world ? r0 + j : 0) * kXchgMaxB * kXchgMaxK, k,
                         lane, r0 + j < world);
#pragma unroll
    for (int j = 0; j < 2; ++j) warp_merge<S>(acc, v[j], lane);
  }
  store_list<S>(final_list, acc, k, lane);
  __syncwarp();
}

// ---- NVLink key exchange, stand-alone pieces (same buffers / flags / sequence protocol as the exchange
// fused into this file's kernel; used by the small-batch filter kernel's last CTA and by exchange.cu).
// Warp-level: every rank stores its k best keys of query b into EVERY rank's peer-mapped buffer ...
__device__ __forceinline__ void xchg_push(const XchgCtx& x, int b, const uint64_t* list, int k, int lane) {
  for (int r = 0; r < x.world; ++r) {
    uint64_t* dst = x.peer[r] + ((static_cast<size_t>(x.slot) * kMaxPeers + x.rank) * kXchgMaxB + b) * kXchgMaxK;
    for (int i = lane; i < k; i += 32) dst[i] = list[i];
  }
}
// ... then ONE warp (after every push of this rank is ordered before it: same warp or a CTA barrier)
// release-stores the sequence number into the peers' flag slots and acquire-spins on its own flags
// (bounded ~3 s: a missing peer must not hang the GPU).  Returns false on timeout.
__device__ __forceinline__ bool xchg_publish_wait(const XchgCtx& x, int lane) {
  __threadfence_system();
  __syncwarp();
  if (lane < x.world) {
    unsigned int* flag = reinterpret_cast<unsigned int*>(x.peer[lane] + kXchgKeyCount) + x.slot * kMaxPeers + x.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(x.seq) : "memory");
  }
  const unsigned int* my_flags = reinterpret_cast<const unsigned int*>(x.peer[x.rank] + kXchgKeyCount) + x.slot * kMaxPeers;
  bool ok = true;
  if (lane < x.world) {
    const long long t0 = clock64();
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + lane) : "memory");
      if (v == x.seq) break;
      if (clock64() - t0 > 6000000000ll) { ok = false; break; }
      __nanosleep(100);
    } while (true);
  }
  ok = __all_sync(FULL_MASK, ok);
  __threadfence_system();
  return ok;
}
// ... and merges the G lists of query b that arrived in its own buffer (k <= kXchgMaxK) and emits them.
__device__ __forceinline__ void xchg_merge_emit(const XchgCtx& x, int b, uint64_t* final_list, int k, bool ok, int lane,
                                                uint64_t* keys_out, float* scores_out, long long* gids_out, int* count_out) {
  list_clear(final_list, k, lane);
  if (ok) {
    const uint64_t* base = x.peer[x.rank] + (static_cast<size_t>(x.slot) * kMaxPeers * kXchgMaxB + b) * kXchgMaxK;
    if (k <= 32) peers_merge_fast<1>(base, x.world, final_list, k, lane);
    else peers_merge_fast<4>(base, x.world, final_list, k, lane);
  }
  emit_list(final_list, k, lane, keys_out, scores_out, gids_out, count_out);
  if (!ok && count_out && lane == 0) *count_out = -1;  // exchange timed out
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

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
__host__ __device__ constexpr int max_threads(int QB, int U) { return (QB * U > 8 ? 8 : kMaxWarps) * 32; }

template <int QB, int U, bool BF16, bool L2>
__global__ void __launch_bounds__(max_threads(QB, U), 1) scan_topk_kernel(const __grid_constant__ ScanParams p) {
  constexpr int V = QB * U;          // values (query x row) per lane group and tile
  constexpr int VLOG = ilog2(V);
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int qbase = blockIdx.y * QB;  // QB queries per grid.y slice
  const int nq = (p.B - qbase) < QB ? (p.B - qbase) : QB;
  const int k = p.k;
  const int dpad = p.dpad;
  // let the next launch on this stream start as soon as SMs free up (no-op without the PDL attribute)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.only_flag != nullptr) {  // uniform over the whole query block (all CTAs of this blockIdx.y)
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the flags are written by the launch before us
    bool any = false;
    for (int b = 0; b < nq; ++b) any = any || (p.only_flag[qbase + b] != 0);
    if (!any) {
      // nothing to redo for this query block: its first CTA reports the block as finished
      if (p.done_ctr != nullptr && blockIdx.x == 0 && tid == 0) {
        __threadfence();
        if (atomicAdd(p.done_blocks, 1u) == gridDim.y - 1) atomicMax(p.done_ctr, p.done_sn);
      }
      return;
    }
  }

  // ---- shared memory carve-up (mirrors scan_plan)
  float* q_s = reinterpret_cast<float*>(smem);  // [QB][dpad], zero padded
  size_t off = align128(static_cast<size_t>(QB) * dpad * 4);
  float* qinv_s = reinterpret_cast<float*>(smem + off);  // [QB] 1/|q|
  int* flag_s = reinterpret_cast<int*>(smem + off + 64);
  off += 128;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + off);
  off += align128(static_cast<size_t>(nwarps) * p.stages * 8);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + off);  // [nwarps][QB][k] running top-k lists
  off += align128(static_cast<size_t>(nwarps) * QB * k * 8);
  unsigned char* stage_area = smem + off;  // nwarps * stages * stage_bytes, reused as merge scratch

  // every warp arms its own barriers and starts its TMA pipeline right away; the query staging
  // below overlaps the first bulk copies
  if (lane == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(smem_u32(mbar + warp * p.stages + s), 1);
  }
  fence_mbar_init();
  __syncwarp();
  const bool cosine = (p.metric == kCosine);

  // ---- lane geometry.  lanes-per-row lpr >= V (scan_plan guarantees it): after the reduction every
  // lane group has spread its V sums over its lanes, lane `lig` holding value v = lig >> (lpr_log2 - VLOG)
  const int lpr_log2 = p.lpr_log2;
  const int lpr = 1 << lpr_log2;
  const int G = 32 >> lpr_log2;  // row groups per warp
  const int g = lane >> lpr_log2;
  const int lig = lane & (lpr - 1);
  const int nch = p.nch, cpr = p.cpr, row_bytes = p.row_bytes;
  const int my_v = lig >> (lpr_log2 - VLOG);
  const int my_b = my_v / U, my_u = my_v % U;
  const bool leader = (lig & ((lpr >> VLOG) - 1)) == 0;
  const int my_r = my_u * G + g;  // row (inside the tile) of the value this lane ends up with
  // k-th score of list my_b; while the list is not full: the caller's score floor (threshold
  // push-down, -inf by default)
  float my_thr_f = p.min_score;
  uint64_t* my_lists = lists + static_cast<size_t>(warp) * QB * k;

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

  // ---- stage the queries (zero padded), clear the running lists, 1/|q|
  for (int i = tid; i < QB * dpad; i += blockDim.x) {
    const int b = i / dpad, c = i - b * dpad;
    q_s[i] = (b < nq && c < p.dim) ? __ldg(p.q + static_cast<size_t>(qbase + b) * p.dim + c) : 0.0f;
  }
  for (int i = tid; i < nwarps * QB * k; i += blockDim.x) lists[i] = 0ull;
  __syncthreads();
  for (int b = warp; b < QB; b += nwarps) {
    float ss = 0.0f;
    for (int i = lane; i < dpad; i += 32) ss = fmaf(q_s[b * dpad + i], q_s[b * dpad + i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
    if (lane == 0) qinv_s[b] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  __syncthreads();
  const float my_qinv = qinv_s[my_b];

  // compute-side cursor runs one tile ahead so the inv-norm load of tile i+1 is in flight while
  // tile i is processed
  int cur_next = 0;
  TileLoc Tn;
  Tn.seg = 0; Tn.row0 = 0; Tn.nrows = 0;
  float inr_n = 0.0f;
  auto prefetch = [&](long long t) {
    Tn = locate_tile(p, t, cur_next);
    if (cosine) inr_n = (my_r < Tn.nrows) ? __ldg(p.seg[Tn.seg].inv_norm + Tn.row0 + my_r) : 0.0f;
  };
  if (gw < total) prefetch(gw);

  int stage = 0;
  uint32_t parity = 0;
  for (long long t = gw; t < total; t += GW) {
    const TileLoc T = Tn;
    const float inr = inr_n;
    if (t + GW < total) prefetch(t + GW);

    WDBX_ASSERT(stage >= 0 && stage < p.stages && T.nrows > 0 && T.nrows <= p.tile_rows && T.row0 + T.nrows <= p.seg[T.seg].n_rows);
    mbar_wait(my_bar + stage * 8, parity);
    const unsigned char* sb = my_stage + static_cast<size_t>(stage) * p.stage_bytes;

    float acc[V];  // acc[b * U + u]
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0f;

#pragma unroll 2
    for (int j = 0; j < nch; ++j) {
      const int c = lig + (j << lpr_log2);
      if (c < cpr) {
        if (!BF16) {
          float4 x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) x[u] = lds128(sb + static_cast<size_t>(u * G + g) * row_bytes + c * 16);
#pragma unroll
          for (int b = 0; b < QB; ++b) {
            const float4 qv = lds128(q_s + b * dpad + c * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              float a = acc[b * U + u];
              if (L2) {
                const float d0 = x[u].x - qv.x, d1 = x[u].y - qv.y, d2 = x[u].z - qv.z, d3 = x[u].w - qv.w;
                a = fmaf(d0, d0, a);
                a = fmaf(d1, d1, a);
                a = fmaf(d2, d2, a);
                a = fmaf(d3, d3, a);
              } else {
                a = fmaf(x[u].x, qv.x, a);
                a = fmaf(x[u].y, qv.y, a);
                a = fmaf(x[u].z, qv.z, a);
                a = fmaf(x[u].w, qv.w, a);
              }
              acc[b * U + u] = a;
            }
          }
        } else {
          float xf[U][8];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint4 raw = *reinterpret_cast<const uint4*>(sb + static_cast<size_t>(u * G + g) * row_bytes + c * 16);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              xf[u][2 * i] = __uint_as_float(w[i] << 16);
              xf[u][2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
            }
          }
#pragma unroll
          for (int b = 0; b < QB; ++b) {
            const float4 qa = lds128(q_s + b * dpad + c * 8);
            const float4 qb = lds128(q_s + b * dpad + c * 8 + 4);
            const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int u = 0; u < U; ++u) {
              float a = acc[b * U + u];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (L2) {
                  const float d = xf[u][i] - qq[i];
                  a = fmaf(d, d, a);
                } else {
                  a = fmaf(xf[u][i], qq[i], a);
                }
              }
              acc[b * U + u] = a;
            }
          }
        }
      }
    }

    // ---- reduce over the lpr lanes of a row group.  Recursive halving: in round r the lanes whose
    // bit o is clear keep the lower half of the values, the others the upper half, so V sums cost
    // V-1 shuffles (not V*log2(lpr)) and end up one per lane.  Pairing order (xor 16, 8, 4, 2, 1)
    // is the plain butterfly's, so every sum is bit-identical for every QB / U.
#pragma unroll
    for (int r = 0; r < VLOG; ++r) {
      const int o = lpr >> (r + 1);
      const bool hb = (lane & o) != 0;
      const int half = (V >> r) >> 1;
#pragma unroll
      for (int i = 0; i < (V >> 1); ++i) {
        if (i < half) {
          const float send = hb ? acc[i] : acc[i + half];
          const float keep = hb ? acc[i + half] : acc[i];
          acc[i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
        }
      }
    }
    for (int o = lpr >> (VLOG + 1); o > 0; o >>= 1) acc[0] += __shfl_xor_sync(FULL_MASK, acc[0], o);

    // ---- one candidate test per lane: its value is (query my_b, row my_r)
    float s = acc[0];
    if (L2) s = -s;
    else if (cosine) s = s * inr * my_qinv;
    s = (s != s) ? __int_as_float(0xff800000) : s;
    if (p.all_keys != nullptr) {
      // LARGE-k mode (k > 128, select_topk.cu): no running lists -- the ranking key of EVERY (query, row) goes to
      // HBM (8 bytes per row and query next to the row's hundreds of bytes; 0 = dead / filtered / below the floor) and
      // a radix select finds the k best afterwards
      if (leader && my_r < T.nrows && my_b < nq) {
        const long long row = T.row0 + my_r;
        bool ok = s >= p.min_score;
        const uint32_t* tomb = p.seg[T.seg].tomb;
        if (tomb != nullptr && ((__ldg(tomb + (row >> 5)) >> (row & 31)) & 1u)) ok = false;
        const uint32_t* allow = p.seg[T.seg].allow;
        if (allow != nullptr && !((__ldg(allow + (row >> 5)) >> (row & 31)) & 1u)) ok = false;
        const uint32_t gid = __ldg(p.seg[T.seg].gids + row);
        WDBX_ASSERT(p.seg_row_base[T.seg] + row < p.all_rows && qbase + my_b < p.B);
        p.all_keys[static_cast<size_t>(qbase + my_b) * p.all_rows + p.seg_row_base[T.seg] + row] = ok ? pack_key(s, gid) : 0ull;
      }
      __syncwarp();
      if (t_issue < total) issue(stage);
      if (++stage == p.stages) {
        stage = 0;
        parity ^= 1u;
      }
      continue;
    }
    const bool pass = leader && (my_r < T.nrows) && (my_b < nq) && (s >= my_thr_f);
    unsigned m = __ballot_sync(FULL_MASK, pass);
    while (m) {
      const int src_lane = __ffs(m) - 1;
      m &= m - 1;
      const float sv = __shfl_sync(FULL_MASK, s, src_lane);
      const int sb_ = __shfl_sync(FULL_MASK, my_b, src_lane);
      const long long row = T.row0 + __shfl_sync(FULL_MASK, my_r, src_lane);
      const uint32_t* tomb = p.seg[T.seg].tomb;
      if (tomb != nullptr && ((__ldg(tomb + (row >> 5)) >> (row & 31)) & 1u)) continue;
      const uint32_t* allow = p.seg[T.seg].allow;  // opt-in metadata pre-filter: 1 = row may be returned
      if (allow != nullptr && !((__ldg(allow + (row >> 5)) >> (row & 31)) & 1u)) continue;
      const uint32_t gid = __ldg(p.seg[T.seg].gids + row);
      const uint64_t key = pack_key(sv, gid);
      WDBX_ASSERT(sb_ >= 0 && sb_ < QB && row >= 0 && row < p.seg[T.seg].n_rows && T.seg < p.n_seg);
      uint64_t* list = my_lists + sb_ * k;
      if (key > list[k - 1]) {
        const uint64_t nthr = list_insert(list, k, key, lane);
        if (my_b == sb_) my_thr_f = nthr ? fmaxf(key_score(nthr), p.min_score) : p.min_score;
      }
    }

    __syncwarp();
    if (t_issue < total) issue(stage);
    if (++stage == p.stages) {
      stage = 0;
      parity ^= 1u;
    }
  }

  // Programmatic dependent launch: everything above only READS the store and the queries, so it may
  // overlap the merge tail of the previous search on this stream; from here on we touch the
  // workspace / outputs / exchange buffers that the previous grid may still be using.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.all_keys != nullptr) return;   // large-k mode: the select kernels behind us do the rest

  // ---- CTA merge: fold the nwarps lists of every query and publish k keys per query
  __syncthreads();
  uint64_t* scratch = reinterpret_cast<uint64_t*>(stage_area);  // [nwarps][max(k,32)] (stage buffers are idle now)
  const bool small_k = k <= 128;  // lists fit 1 or 4 keys per lane: register bitonic merges + a warp tree
  if (small_k) {
    for (int b = 0; b < nq; ++b) {
      uint64_t* out = p.cand + (static_cast<size_t>(qbase + b) * gridDim.x + blockIdx.x) * k;
      if (k <= 32) cta_merge_fast<1>(my_lists + b * k, scratch, out, k, warp, lane, nwarps);
      else cta_merge_fast<4>(my_lists + b * k, scratch, out, k, warp, lane, nwarps);
    }
  } else {
    for (int b = warp; b < nq; b += nwarps) {
      uint64_t* dst = scratch + static_cast<size_t>(warp) * k;
      list_clear(dst, k, lane);
      uint64_t thr = 0ull;
      for (int w = 0; w < nwarps; ++w)
        absorb_keys<false>(dst, k, thr, lists + (static_cast<size_t>(w) * QB + b) * k, k, 0, 1, lane);
      uint64_t* out = p.cand + (static_cast<size_t>(qbase + b) * gridDim.x + blockIdx.x) * k;
      for (int i = lane; i < k; i += 32) out[i] = dst[i];
    }
  }

  // ---- last CTA to finish merges all CTA lists of these queries
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(p.counters + blockIdx.y, 1u);
    *flag_s = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*flag_s == 0) return;
  __threadfence();
  uint64_t* final_list = lists;  // the running lists are dead: reuse the first k slots
  for (int b = 0; b < nq; ++b) {
    const uint64_t* cand_q = p.cand + static_cast<size_t>(qbase + b) * gridDim.x * k;
    if (small_k) {
      if (k <= 32) grid_merge_fast<1>(cand_q, static_cast<int>(gridDim.x), scratch, final_list, k, warp, lane, nwarps);
      else grid_merge_fast<4>(cand_q, static_cast<int>(gridDim.x), scratch, final_list, k, warp, lane, nwarps);
    } else {
      uint64_t* dst = scratch + static_cast<size_t>(warp) * k;
      list_clear(dst, k, lane);
      uint64_t thr = 0ull;
      absorb_keys<true>(dst, k, thr, cand_q, static_cast<int>(gridDim.x) * k, warp, nwarps, lane);
    }
    __syncthreads();
    if (warp == 0) {
      const int qi = qbase + b;
      if (!small_k) {
        list_clear(final_list, k, lane);
        uint64_t thr = 0ull;
        absorb_keys<false>(final_list, k, thr, scratch, nwarps * k, 0, 1, lane);
      }
      __syncwarp();
      if (p.xchg_world > 1) {
        // push this rank's list for query b into every rank's exchange buffer (NVLink P2P stores)
        for (int r = 0; r < p.xchg_world; ++r) {
          uint64_t* dst = p.xchg_peer[r] +
                          ((static_cast<size_t>(p.xchg_slot) * kMaxPeers + p.xchg_rank) * kXchgMaxB + b) * kXchgMaxK;
          for (int i = lane; i < k; i += 32) dst[i] = final_list[i];
        }
      } else {
        emit_list(final_list, k, lane, p.keys_out ? p.keys_out + static_cast<size_t>(qi) * k : nullptr,
                  p.scores_out ? p.scores_out + static_cast<size_t>(qi) * k : nullptr,
                  p.gids_out ? p.gids_out + static_cast<size_t>(qi) * k : nullptr,
                  p.counts_out ? p.counts_out + qi : nullptr);
      }
    }
    __syncthreads();
  }
  if (p.xchg_world > 1 && warp == 0) {
    // publish: release-store the sequence number into every peer's flag slot for this rank
    __threadfence_system();
    __syncwarp();
    if (lane < p.xchg_world) {
      unsigned int* flag = reinterpret_cast<unsigned int*>(p.xchg_peer[lane] + kXchgKeyCount) +
                           p.xchg_slot * kMaxPeers + p.xchg_rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(p.xchg_seq) : "memory");
    }
    // wait for every peer's push into OUR buffer (bounded spin: a missing peer must not hang the GPU)
    const unsigned int* my_flags = reinterpret_cast<const unsigned int*>(p.xchg_peer[p.xchg_rank] + kXchgKeyCount) +
                                   p.xchg_slot * kMaxPeers;
    bool ok = true;
    if (lane < p.xchg_world) {
      const long long t0 = clock64();
      unsigned int v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + lane) : "memory");
        if (v == p.xchg_seq) break;
        if (clock64() - t0 > 6000000000ll) { ok = false; break; }  // ~3 s
        __nanosleep(100);
      } while (true);
    }
    ok = __all_sync(FULL_MASK, ok);
    __threadfence_system();
    for (int b = 0; b < nq; ++b) {
      const int qi = qbase + b;
      list_clear(final_list, k, lane);
      uint64_t thr = 0ull;
      if (ok && small_k) {
        const uint64_t* base = p.xchg_peer[p.xchg_rank] +
                               (static_cast<size_t>(p.xchg_slot) * kMaxPeers * kXchgMaxB + b) * kXchgMaxK;
        if (k <= 32) peers_merge_fast<1>(base, p.xchg_world, final_list, k, lane);
        else peers_merge_fast<4>(base, p.xchg_world, final_list, k, lane);
      } else if (ok) {
        for (int r = 0; r < p.xchg_world; ++r) {
          const uint64_t* src = p.xchg_peer[p.xchg_rank] +
                                ((static_cast<size_t>(p.xchg_slot) * kMaxPeers + r) * kXchgMaxB + b) * kXchgMaxK;
          absorb_keys<true>(final_list, k, thr, src, k, 0, 1, lane);
        }
      }
      emit_list(final_list, k, lane, p.keys_out ? p.keys_out + static_cast<size_t>(qi) * k : nullptr,
                p.scores_out ? p.scores_out + static_cast<size_t>(qi) * k : nullptr,
                p.gids_out ? p.gids_out + static_cast<size_t>(qi) * k : nullptr,
                p.counts_out ? p.counts_out + qi : nullptr);
      if (!ok && p.counts_out && lane == 0) p.counts_out[qi] = -1;  // exchange timed out
    }
  }
  if (tid == 0) p.counters[blockIdx.y] = 0u;  // re-arm for the next launch
  if (p.done_ctr != nullptr && tid == 0) {    // (only the last CTA of the query block gets here)
    __threadfence();
    if (atomicAdd(p.done_blocks, 1u) == gridDim.y - 1) atomicMax(p.done_ctr, p.done_sn);
  }
}

template <int QB, int U, bool BF16, bool L2>
cudaError_t launch_one(const ScanParams& p, const ScanPlan& plan, cudaStream_t stream) {
  auto kern = scan_topk_kernel<QB, U, BF16, L2>;
  static std::atomic<unsigned long long> attr_done{0ull};
  const cudaError_t attr_err = once_per_device(attr_done, [&] {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  dim3 grid(plan.grid, (p.B + QB - 1) / QB, 1), block(plan.warps * 32, 1, 1);
  if (!plan.pdl) {
    kern<<<grid, block, plan.smem_bytes, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

// One translation unit per QB instantiates its kernels through this dispatcher.
template <int QB>
cudaError_t launch_qb(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  const bool l2 = p.metric == kL2;
#define WDBX_DISPATCH(UU)                                                                      \
  do {                                                                                         \
    if (bf16) return l2 ? launch_one<QB, UU, true, true>(p, plan, stream)                      \
                        : launch_one<QB, UU, true, false>(p, plan, stream);                    \
    return l2 ? launch_one<QB, UU, false, true>(p, plan, stream)                               \
              : launch_one<QB, UU, false, false>(p, plan, stream);                             \
  } while (0)
  if (plan.U == 1) WDBX_DISPATCH(1);
  if (plan.U == 2) WDBX_DISPATCH(2);
  WDBX_DISPATCH(4);
#undef WDBX_DISPATCH
}

}  // namespace scan
}  // namespace wdbx
