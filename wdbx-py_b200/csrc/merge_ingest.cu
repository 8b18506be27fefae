// merge_ingest.cu -- K3 (k-way merge of best-first key lists) and K4 (ingest: store rows + norms).
#include "common.cuh"
#include "kernels.h"

namespace wdbx {

namespace {

// ------------------------------------------------------------------ K3
// One warp per query.  keys: [G][B][k] best-first lists (0 = empty slot), e.g. the NCCL
// all-gather of every rank's local top-k.  Replaces the concat + sort + [:limit] of
// VectorStore.search (wdbx/core/vector_store.py:324-330, :345) across GPUs.
template <int KS>
__global__ void merge_topk_kernel(const uint64_t* __restrict__ keys, int G, int B, int k, uint64_t* keys_out,
                                  float* scores_out, long long* gids_out, int* counts_out) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= B) return;
  WarpTopK<KS> M;
  M.reset();
  for (int g = 0; g < G; ++g) {
    const uint64_t* src = keys + (static_cast<size_t>(g) * B + q) * k;
    for (int base = 0; base < k; base += 32) {
      const int idx = base + lane;
      const uint64_t key = idx < k ? src[idx] : 0ull;
      unsigned m = __ballot_sync(FULL_MASK, key > M.thr);
      while (m) {
        const int sl = __ffs(m) - 1;
        m &= m - 1;
        M.offer(__shfl_sync(FULL_MASK, key, sl), k, lane);
      }
    }
  }
  emit_outputs<KS>(M, k, lane, keys_out ? keys_out + static_cast<size_t>(q) * k : nullptr,
                   scores_out ? scores_out + static_cast<size_t>(q) * k : nullptr,
                   gids_out ? gids_out + static_cast<size_t>(q) * k : nullptr, counts_out ? counts_out + q : nullptr);
}

// ------------------------------------------------------------------ K4
// One warp per row: copy (optionally round to bf16, RNE) into the padded stored layout and
// compute |x|^2 and 1/|x| of the STORED values.  Replaces the normalise + index.add of
// FaissIndex.add / batch_add (wdbx/core/indexing.py:886-890, :937-950); rows are kept raw so
// that ip / l2 can be served from the same matrix, the norm is applied at scan time.
template <bool BF16>
__global__ void append_rows_kernel(const float* __restrict__ src, long long n, int dim, int dpad,
                                   unsigned char* __restrict__ dst_rows, float* __restrict__ inv_norm,
                                   float* __restrict__ sqnorm, uint32_t* __restrict__ gids_dst,
                                   const uint32_t* __restrict__ gids_src, uint32_t gid_base) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* s = src + row * dim;
  float ss = 0.0f;
  if (BF16) {
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst_rows) + row * dpad;
    for (int c = lane; c < dpad; c += 32) {
      const float v = c < dim ? s[c] : 0.0f;
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      d[c] = b;
      const float r = __bfloat162float(b);
      ss = fmaf(r, r, ss);
    }
  } else {
    float* d = reinterpret_cast<float*>(dst_rows) + row * dpad;
    for (int c = lane; c < dpad; c += 32) {
      const float v = c < dim ? s[c] : 0.0f;
      d[c] = v;
      ss = fmaf(v, v, ss);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
  if (lane == 0) {
    sqnorm[row] = ss;
    inv_norm[row] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
    gids_dst[row] = gids_src ? gids_src[row] : gid_base + static_cast<uint32_t>(row);
  }
}

// stored rows -> dense fp32 [n, dim] (VectorStore.get / persistence); one CTA per row
template <bool BF16>
__global__ void export_rows_kernel(const unsigned char* __restrict__ rows, int dim, int row_bytes,
                                   float* __restrict__ dst) {
  const unsigned char* row = rows + static_cast<size_t>(blockIdx.x) * row_bytes;
  float* out = dst + static_cast<size_t>(blockIdx.x) * dim;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    out[c] = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[c])
                  : reinterpret_cast<const float*>(row)[c];
  }
}

}  // namespace

cudaError_t launch_merge_topk(const uint64_t* keys, int G, int B, int k, uint64_t* keys_out, float* scores_out,
                              long long* gids_out, int* counts_out, cudaStream_t stream) {
  const int wpb = 4;
  dim3 grid((B + wpb - 1) / wpb), block(wpb * 32);
  if (k <= 128) merge_topk_kernel<4><<<grid, block, 0, stream>>>(keys, G, B, k, keys_out, scores_out, gids_out, counts_out);
  else merge_topk_kernel<32><<<grid, block, 0, stream>>>(keys, G, B, k, keys_out, scores_out, gids_out, counts_out);
  return cudaGetLastError();
}

cudaError_t launch_append_rows(const float* src, long long n, int dim, int dpad, bool bf16, unsigned char* dst_rows,
                               float* inv_norm, float* sqnorm, uint32_t* gids_dst, const uint32_t* gids_src,
                               uint32_t gid_base, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int wpb = 8;
  dim3 grid(static_cast<unsigned>((n + wpb - 1) / wpb)), block(wpb * 32);
  if (bf16) append_rows_kernel<true><<<grid, block, 0, stream>>>(src, n, dim, dpad, dst_rows, inv_norm, sqnorm, gids_dst, gids_src, gid_base);
  else append_rows_kernel<false><<<grid, block, 0, stream>>>(src, n, dim, dpad, dst_rows, inv_norm, sqnorm, gids_dst, gids_src, gid_base);
  return cudaGetLastError();
}

cudaError_t launch_export_rows(const unsigned char* rows, long long n, int dim, int row_bytes, bool bf16, float* dst,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const dim3 grid(static_cast<unsigned>(n)), block(128);
  if (bf16) export_rows_kernel<true><<<grid, block, 0, stream>>>(rows, dim, row_bytes, dst);
  else export_rows_kernel<false><<<grid, block, 0, stream>>>(rows, dim, row_bytes, dst);
  return cudaGetLastError();
}

}  // namespace wdbx
