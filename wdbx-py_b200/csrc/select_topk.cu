// select_topk.cu -- exact top-k for LARGE k (128 < k <= 1024) by radix select (sm_100a).
//
// The reference asks FAISS / hnswlib for k = 1000 neighbours from its visualisation layer
// (wdbx/utils/visualization.py:493-498, :742-746 -> VectorStore.search limit=1000 -> IndexFlatIP.search,
// wdbx/core/indexing.py:1013, whose heap keeps k entries).  Running top-k lists stop paying off there: with
// k = 1000 a warp's share of a 1M-row store (a few hundred rows) never even fills its list, so every row is a
// sorted insert.  Instead the scan kernel (scan_topk_kernel, `all_keys` mode) writes the 8-byte ranking key of
// every (query, row) to HBM -- 8 bytes next to the 1.5 KB of a 384-d fp32 row, +0.5 % traffic -- and this file
// selects the k largest keys exactly:
//   1. radix_pass_kernel x 8: most-significant-byte-first histogram over the keys that match the prefix found
//      so far; the last CTA of a pass (atomic ticket) picks the bucket that holds the k-th largest key and
//      narrows the prefix.  Keys are unique (the low 32 bits are the row's global id), so after 8 passes the
//      prefix IS the k-th largest key.  The key array of one query (8 MB for 1M rows) stays in the 126 MB L2.
//   2. collect_kernel: keys >= that threshold -> a k-entry buffer (exactly min(k, live rows) of them);
//   3. sort_emit_kernel: one CTA per query sorts the <= 1024 keys (bitonic network in shared memory) and emits
//      keys / scores / gids / count -- "score desc, gid asc", the same order every other path returns.
#include "scan_topk_kernel.cuh"

namespace wdbx {

namespace {

constexpr int kSelThreads = 256;
constexpr int kSelMaxK = 1024;

// per-query select state in the zero-initialised workspace (32-bit words):
//   [0..255] histogram | [256] ticket | [257] k still to find inside the prefix | [258,259] prefix (lo, hi)
//   | [260] collected count
constexpr int kSelWords = 264;

__global__ void select_init_kernel(unsigned int* ws, int B, int k) {
  const int q = blockIdx.x;
  unsigned int* st = ws + static_cast<size_t>(q) * kSelWords;
  for (int i = threadIdx.x; i < kSelWords; i += blockDim.x) st[i] = (i == 257) ? static_cast<unsigned int>(k) : 0u;
}

__global__ void __launch_bounds__(kSelThreads) radix_pass_kernel(const uint64_t* __restrict__ keys, long long n, unsigned int* ws,
                                                                 int shift) {
  __shared__ unsigned int hist[256];
  __shared__ int last;
  const int q = blockIdx.y;
  unsigned int* st = ws + static_cast<size_t>(q) * kSelWords;
  const uint64_t prefix = (static_cast<uint64_t>(st[259]) << 32) | st[258];
  const uint64_t mask = shift == 56 ? 0ull : (~0ull << (shift + 8));   // bytes above `shift` are decided
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const uint64_t* kq = keys + static_cast<size_t>(q) * n;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count; lanes that hit the same bucket are aggregated (scores share their leading bytes, so the
  // first passes would otherwise serialise 32 shared-memory atomics on one address)
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x + (threadIdx.x & ~31); base < n;
       base += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = base + lane;
    const uint64_t key = i < n ? __ldcg(reinterpret_cast<const unsigned long long*>(kq + i)) : 0ull;
    const bool valid = i < n && (key & mask) == prefix;
    const unsigned vm = __ballot_sync(FULL_MASK, valid);
    if (valid) {
      const unsigned int b = static_cast<unsigned int>((key >> shift) & 0xFFull);
      const unsigned peers = __match_any_sync(vm, b);
      if (lane == __ffs(peers) - 1) atomicAdd(hist + b, static_cast<unsigned int>(__popc(peers)));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (hist[i]) atomicAdd(st + i, hist[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(st + 256, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // last CTA of the pass: the bucket that contains the k-th largest matching key (buckets scanned from the top)
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned int k_rem = st[257];
    unsigned int above = 0;
    int found = -1;
    for (int base = 255; base >= 0 && found < 0; base -= 32) {   // lane l looks at bucket base - l
      const int b = base - lane;
      const unsigned int c = __ldcg(st + b);
      // inclusive sum over higher buckets of this chunk (lane 0 = highest)
      unsigned int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      const bool hit = above + incl >= k_rem;
      const unsigned m = __ballot_sync(FULL_MASK, hit);
      if (m) {
        const int l = __ffs(m) - 1;
        found = base - l;
        const unsigned int excl = __shfl_sync(FULL_MASK, incl - c, l);
        k_rem -= above + excl;
      } else {
        above += __shfl_sync(FULL_MASK, incl, 31);
      }
    }
    if (found < 0) found = 0;   // fewer than k keys match: everything down to the smallest bucket is selected
    __syncwarp();
    for (int i = lane; i < 256; i += 32) st[i] = 0u;
    if (lane == 0) {
      const uint64_t np = prefix | (static_cast<uint64_t>(found) << shift);
      st[258] = static_cast<unsigned int>(np);
      st[259] = static_cast<unsigned int>(np >> 32);
      st[257] = k_rem;
      st[256] = 0u;
    }
  }
}

__global__ void __launch_bounds__(kSelThreads) collect_kernel(const uint64_t* __restrict__ keys, long long n, unsigned int* ws, int k,
                                                              uint64_t* __restrict__ sel) {
  const int q = blockIdx.y;
  unsigned int* st = ws + static_cast<size_t>(q) * kSelWords;
  uint64_t thr = (static_cast<uint64_t>(st[259]) << 32) | st[258];
  if (thr == 0ull) thr = 1ull;   // fewer live rows than k: every non-empty key
  const uint64_t* kq = keys + static_cast<size_t>(q) * n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint64_t key = __ldcg(reinterpret_cast<const unsigned long long*>(kq + i));
    if (key >= thr) {
      const unsigned int pos = atomicAdd(st + 260, 1u);
      WDBX_ASSERT(pos < static_cast<unsigned int>(k) || thr == 1ull);   // exactly k keys reach a real threshold
      if (pos < static_cast<unsigned int>(k)) sel[static_cast<size_t>(q) * kSelMaxK + pos] = key;
    }
  }
}

__global__ void __launch_bounds__(512) sort_emit_kernel(const uint64_t* __restrict__ sel, const unsigned int* ws, int k,
                                                        uint64_t* keys_out, float* scores_out, long long* gids_out,
                                                        int* counts_out) {
  __shared__ uint64_t s[kSelMaxK];
  const int q = blockIdx.x;
  const unsigned int cnt_raw = ws[static_cast<size_t>(q) * kSelWords + 260];
  const int cnt = cnt_raw < static_cast<unsigned int>(k) ? static_cast<int>(cnt_raw) : k;
  for (int i = threadIdx.x; i < kSelMaxK; i += blockDim.x) s[i] = i < cnt ? sel[static_cast<size_t>(q) * kSelMaxK + i] : 0ull;
  __syncthreads();
  // bitonic sort, descending, 1024 elements / 512 threads
  for (int k2 = 2; k2 <= kSelMaxK; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSelMaxK / 2; t += blockDim.x) {
        const int i = ((t / j) * 2 * j) + (t % j);   // lower index of the pair
        const int p = i + j;
        const bool desc = (i & k2) == 0;
        const uint64_t a = s[i], b = s[p];
        if ((a < b) == desc) { s[i] = b; s[p] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = s[i];
    if (keys_out) keys_out[static_cast<size_t>(q) * k + i] = key;
    if (scores_out) scores_out[static_cast<size_t>(q) * k + i] = key ? key_score(key) : __int_as_float(0xff800000);
    if (gids_out) gids_out[static_cast<size_t>(q) * k + i] = key ? static_cast<long long>(key_gid(key)) : -1ll;
  }
  if (counts_out && threadIdx.x == 0) counts_out[q] = cnt;
}

}  // namespace

size_t select_workspace_words(int B) { return static_cast<size_t>(B) * kSelWords; }

cudaError_t launch_select_topk(const uint64_t* all_keys, long long n_keys, int B, int k, unsigned int* workspace,
                               uint64_t* sel_keys, int sm_count, uint64_t* keys_out, float* scores_out, long long* gids_out,
                               int* counts_out, cudaStream_t stream) {
  if (B < 1 || k < 1 || k > kSelMaxK) return cudaErrorInvalidValue;
  select_init_kernel<<<B, 64, 0, stream>>>(workspace, B, k);
  long long want = (n_keys + kSelThreads * 8 - 1) / (kSelThreads * 8);
  int blocks = static_cast<int>(std::min<long long>(std::max<long long>(want, 1), 4ll * sm_count));
  if (n_keys > 0) {
    for (int shift = 56; shift >= 0; shift -= 8)
      radix_pass_kernel<<<dim3(blocks, B), kSelThreads, 0, stream>>>(all_keys, n_keys, workspace, shift);
    collect_kernel<<<dim3(blocks, B), kSelThreads, 0, stream>>>(all_keys, n_keys, workspace, k, sel_keys);
  }
  sort_emit_kernel<<<B, 512, 0, stream>>>(sel_keys, workspace, k, keys_out, scores_out, gids_out, counts_out);
  return cudaGetLastError();
}

}  // namespace wdbx
