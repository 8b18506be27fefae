// exchange.cu -- stand-alone cross-GPU key exchange + merge over NVLink peer memory (sm_100a).
//
// Same protocol as the exchange fused into K1's last CTA (scan_topk_kernel.cuh): every rank stores its
// k best keys per query into EVERY rank's peer-mapped exchange buffer (P2P stores), release-stores the
// collective sequence number into the peers' flag slots, acquire-spins (bounded) on its own flags and
// merges the G lists itself.  Used after the bf16-filter path (K2b), whose final list is produced by
// the refine kernel; replaces the host-side concat + sort of VectorStore.search
// (wdbx/core/vector_store.py:323-330) across GPUs without an NCCL call.
#include "scan_topk_kernel.cuh"

namespace wdbx {

namespace {

struct XchgParams {
  uint64_t* peer[kMaxPeers];
  int world, rank, slot;
  unsigned int seq;
  const uint64_t* keys_in;  // [B][k] this rank's best-first lists
  int B, k;
  uint64_t* keys_out;
  float* scores_out;
  long long* gids_out;
  int* counts_out;
};

__global__ void __launch_bounds__(32 * kXchgMaxB, 1) exchange_merge_kernel(const __grid_constant__ XchgParams p) {
  __shared__ uint64_t lists[kXchgMaxB][kXchgMaxK];
  __shared__ int ok_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k;
  if (warp < p.B) {
    for (int r = 0; r < p.world; ++r) {
      uint64_t* dst = p.peer[r] + ((static_cast<size_t>(p.slot) * kMaxPeers + p.rank) * kXchgMaxB + warp) * kXchgMaxK;
      for (int i = lane; i < k; i += 32) dst[i] = p.keys_in[static_cast<size_t>(warp) * k + i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (warp == 0) {
    if (lane < p.world) {
      unsigned int* flag = reinterpret_cast<unsigned int*>(p.peer[lane] + kXchgKeyCount) + p.slot * kMaxPeers + p.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(p.seq) : "memory");
    }
    const unsigned int* my_flags = reinterpret_cast<const unsigned int*>(p.peer[p.rank] + kXchgKeyCount) + p.slot * kMaxPeers;
    bool ok = true;
    if (lane < p.world) {
      const long long t0 = clock64();
      unsigned int v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + lane) : "memory");
        if (v == p.seq) break;
        if (clock64() - t0 > 6000000000ll) { ok = false; break; }  // ~3 s: a missing peer must not hang the GPU
        __nanosleep(100);
      } while (true);
    }
    ok = __all_sync(FULL_MASK, ok);
    if (lane == 0) ok_s = ok ? 1 : 0;
    __threadfence_system();
  }
  __syncthreads();
  if (warp >= p.B) return;
  uint64_t* final_list = lists[warp];
  scan::list_clear(final_list, k, lane);
  if (ok_s) {
    const uint64_t* base = p.peer[p.rank] + (static_cast<size_t>(p.slot) * kMaxPeers * kXchgMaxB + warp) * kXchgMaxK;
    if (k <= 32) scan::peers_merge_fast<1>(base, p.world, final_list, k, lane);
    else scan::peers_merge_fast<4>(base, p.world, final_list, k, lane);
  }
  scan::emit_list(final_list, k, lane, p.keys_out ? p.keys_out + static_cast<size_t>(warp) * k : nullptr,
                  p.scores_out ? p.scores_out + static_cast<size_t>(warp) * k : nullptr,
                  p.gids_out ? p.gids_out + static_cast<size_t>(warp) * k : nullptr, p.counts_out ? p.counts_out + warp : nullptr);
  if (!ok_s && p.counts_out && lane == 0) p.counts_out[warp] = -1;  // exchange timed out
}

}  // namespace

cudaError_t launch_exchange_merge(uint64_t* const* peers, int world, int rank, unsigned int seq, const uint64_t* keys_in,
                                  int B, int k, uint64_t* keys_out, float* scores_out, long long* gids_out,
                                  int* counts_out, cudaStream_t stream) {
  if (B < 1 || B > kXchgMaxB || k < 1 || k > kXchgMaxK || world < 2 || world > kMaxPeers) return cudaErrorInvalidValue;
  XchgParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) p.peer[r] = peers[r];
  p.world = world;
  p.rank = rank;
  p.slot = static_cast<int>(seq & 1u);
  p.seq = seq;
  p.keys_in = keys_in;
  p.B = B;
  p.k = k;
  p.keys_out = keys_out;
  p.scores_out = scores_out;
  p.gids_out = gids_out;
  p.counts_out = counts_out;
  exchange_merge_kernel<<<1, 32 * kXchgMaxB, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace wdbx
