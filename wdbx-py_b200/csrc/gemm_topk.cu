// gemm_topk.cu -- K2: tensor-core regime of the exact search (large query batches), sm_100a only.
//
// S = Q . X^T as a tcgen05 GEMM tile with the top-k fused into the TMEM epilogue, so the
// B x N score matrix (41 GB for 1024 queries x 10M rows) never exists.  Replaces, for a batch of
// queries, what the reference does one query at a time in FaissIndex.search
// (wdbx/core/indexing.py:1002-1024: normalise, IndexFlatIP.search, top-k) -- the reference has
// no batch entry point (wdbx/core/wdbx.py:303-336 takes one vector); this is the additive
// search_batch path of SURVEY.md section 8b.
//
// fp32-grade accuracy on tensor cores: 3xTF32 split.  x = x_hi + x_lo with
// x_hi = tf32(x) (round to nearest), x_lo = tf32(x - x_hi); same for q.  Per K block three
// kind::tf32 MMAs accumulate q_hi.x_hi + q_lo.x_hi + q_hi.x_lo in fp32 TMEM (the dropped
// q_lo.x_lo term is ~2^-22 relative).  X is split ON THE FLY: TMA lands the raw fp32 tile in shared
// memory, four converter warps rewrite it in place as x_hi and write x_lo to a sibling buffer
// (an elementwise pass, so the 128B swizzle is irrelevant), then the MMA warp consumes both.  The
// queries are split once per batch by a small pre-kernel.
//
// CTA = 12 warps: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 idle | 4-7 epilogue (one
// query = one TMEM lane = one thread, running top-k list in shared memory) | 8-11 converters.
// Tile: 128 queries (M) x 256 rows (N) x 16 dims (K block, one 64-byte swizzle atom), 4 stages;
// two accumulator tiles in TMEM (512 columns) so the epilogue of tile t overlaps the MMAs of t+1.
// Grid: (query blocks) x (row slices); every CTA sweeps the row tiles of its slice for its query
// block and writes k keys per query; K3 (merge_topk) folds the slices.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace wdbx {

namespace {

using namespace tc05;

constexpr int BM = 128;       // queries per CTA (TMEM lanes)
constexpr int BN = 256;       // database rows per accumulator tile (TMEM columns)
constexpr int BK = 16;        // fp32 elements per K block = 64 bytes = one SWIZZLE_64B atom row
constexpr int STAGES = 4;     // small stages, deep ring: TMA latency + split + MMA of a stage overlap 3 others
constexpr int kThreads = 384;
constexpr int kMaxKGemm = 16;  // per-thread list length limit (shared memory)

constexpr uint32_t X_TILE_BYTES = BN * BK * 4;  // 16 KB
constexpr uint32_t Q_TILE_BYTES = BM * BK * 4;  // 8 KB
constexpr uint32_t STAGE_BYTES = 2 * X_TILE_BYTES + 2 * Q_TILE_BYTES;  // x, x_lo, q_hi, q_lo = 48 KB
constexpr uint32_t TMEM_COLS = 2 * BN;          // two accumulator tiles: epilogue of tile t overlaps MMAs of t+1

struct GemmParams {
  const float* inv_norm;   // [n_rows]
  const float* sqnorm;     // [n_rows]
  const uint32_t* gids;    // [n_rows]
  const uint32_t* tomb;    // bitmap or NULL
  const float* q_inv;      // [B] 1/|q|
  const float* q_sq;       // [B] |q|^2
  long long n_rows;
  int B, k, metric;
  int n_kblocks;           // ceil(dim / 32)
  int n_tiles;             // ceil(n_rows / BN)
  int n_slices;            // gridDim.y
  uint64_t* out_lists;     // [total_slices][B][k]
  int slice_base;          // first slice index of this launch (one launch per segment)
};

// K-major, 64-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (unused for one swizzle atom along K) | SBO = 8 rows x 64 B between
// 8-row groups | version 1 (Blackwell) | layout SWIZZLE_64B (= 4).
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * BK * 4) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major.
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                              // c_format = F32
  d |= 2u << 7;                              // a_format = TF32
  d |= 2u << 10;                             // b_format = TF32
  d |= static_cast<uint32_t>(N >> 3) << 17;  // n_dim
  d |= static_cast<uint32_t>(M >> 4) << 24;  // m_dim
  return d;
}

__device__ __forceinline__ uint32_t tf32_rna(uint32_t bits) {  // round fp32 bits to the nearest tf32
  return (bits + 0x1000u) & 0xFFFFE000u;
}

// ---------------------------------------------------------------- query split pre-kernel
// q -> q_hi (tf32-rounded), q_lo (tf32 of the remainder), 1/|q| and |q|^2; rows padded to `ld` floats.
__global__ void split_queries_kernel(const float* __restrict__ q, int B, int dim, int ld, float* __restrict__ q_hi,
                                     float* __restrict__ q_lo, float* __restrict__ q_inv, float* __restrict__ q_sq) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float ss = 0.0f;
  for (int c = lane; c < ld; c += 32) {
    const float v = c < dim ? q[static_cast<size_t>(b) * dim + c] : 0.0f;
    const float hi = __uint_as_float(tf32_rna(__float_as_uint(v)));
    const float lo = __uint_as_float(tf32_rna(__float_as_uint(v - hi)));
    q_hi[static_cast<size_t>(b) * ld + c] = hi;
    q_lo[static_cast<size_t>(b) * ld + c] = lo;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, o);
  if (lane == 0) {
    q_sq[b] = ss;
    q_inv[b] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
}

// ---------------------------------------------------------------- the GEMM + top-k kernel
// Insert `key` into thread `et`'s sorted list (column-major [k][BM] so lanes never bank-conflict).
__device__ __forceinline__ float list_push(uint64_t* lists, int k, int et, uint64_t key) {
  int i = k - 1;
  while (i > 0) {
    const uint64_t prev = lists[static_cast<size_t>(i - 1) * BM + et];
    if (prev >= key) break;
    lists[static_cast<size_t>(i) * BM + et] = prev;
    --i;
  }
  lists[static_cast<size_t>(i) * BM + et] = key;
  const uint64_t last = lists[static_cast<size_t>(k - 1) * BM + et];
  return last ? key_score(last) : __int_as_float(0xff800000);
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_qhi,
                 const __grid_constant__ CUtensorMap tm_qlo, const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve-up: stages (1024-aligned) | per-thread lists [k][128] u64 | column scales [2][BN] | barriers | tmem ptr
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // [k][BM]
  float* colscale = reinterpret_cast<float*>(lists + static_cast<size_t>(kMaxKGemm) * BM);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(colscale + 2 * BN);
  uint64_t* full_bar = bars;                // [STAGES] TMA landed
  uint64_t* conv_bar = bars + STAGES;       // [STAGES] x_hi / x_lo written
  uint64_t* empty_bar = bars + 2 * STAGES;  // [STAGES] MMAs of the stage retired
  uint64_t* tmem_full = bars + 3 * STAGES;  // [2] accumulator tile complete
  uint64_t* tmem_empty = tmem_full + 2;     // [2] epilogue drained the accumulator
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x;      // query block
  const int slice = blockIdx.y;   // row slice
  const int k = p.k;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_qhi);
    prefetch_tmap(&tm_qlo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(conv_bar + s), 4);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tmem_full + a), 1);
      mbar_init(smem_u32(tmem_empty + a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int my_tiles = (p.n_tiles - slice + p.n_slices - 1) / p.n_slices;  // tiles slice, slice + n_slices, ...

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int row0 = (slice + t * p.n_slices) * BN;
        for (int kb = 0; kb < p.n_kblocks; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
          unsigned char* sb = stage_base + static_cast<size_t>(stage) * STAGE_BYTES;
          const uint32_t bar = smem_u32(full_bar + stage);
          mbar_expect_tx(bar, X_TILE_BYTES + 2 * Q_TILE_BYTES);
          tma_load_2d(smem_u32(sb), &tm_x, kb * BK, row0, bar);
          tma_load_2d(smem_u32(sb + 2 * X_TILE_BYTES), &tm_qhi, kb * BK, qb * BM, bar);
          tma_load_2d(smem_u32(sb + 2 * X_TILE_BYTES + Q_TILE_BYTES), &tm_qlo, kb * BK, qb * BM, bar);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread); accumulator tile t lives in TMEM columns [(t&1)*BN, +BN)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int a = t & 1;
        const uint32_t aphase = (static_cast<uint32_t>(t) >> 1) & 1u;
        mbar_wait(smem_u32(tmem_empty + a), aphase ^ 1u);  // epilogue drained tile t-2
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * BN);
        for (int kb = 0; kb < p.n_kblocks; ++kb) {
          mbar_wait(smem_u32(conv_bar + stage), phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(stage_base + static_cast<size_t>(stage) * STAGE_BYTES);
          const uint64_t d_xhi = make_desc_sw64(sb);
          const uint64_t d_xlo = make_desc_sw64(sb + X_TILE_BYTES);
          const uint64_t d_qhi = make_desc_sw64(sb + 2 * X_TILE_BYTES);
          const uint64_t d_qlo = make_desc_sw64(sb + 2 * X_TILE_BYTES + Q_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = static_cast<uint64_t>((kk * 8 * 4) >> 4);  // 32 bytes per K=8 step
            // small terms first, then the dominant hi.hi product
            umma_tf32(d_tmem, d_qlo + adv, d_xhi + adv, idesc, (kb | kk) ? 1u : 0u);
            umma_tf32(d_tmem, d_qhi + adv, d_xlo + adv, idesc, 1u);
            umma_tf32(d_tmem, d_qhi + adv, d_xhi + adv, idesc, 1u);
          }
          umma_commit(smem_u32(empty_bar + stage));  // frees the stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(tmem_full + a));  // accumulator tile t complete
      }
    }
  } else if (warp >= 8) {
    // ===== converters: x -> x_hi (in place) and x_lo (sibling buffer)
    const int ct = threadIdx.x - 256;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    const int total = my_tiles * p.n_kblocks;
    for (int it = 0; it < total; ++it) {
      mbar_wait(smem_u32(full_bar + stage), phase);
      uint4* xs = reinterpret_cast<uint4*>(stage_base + static_cast<size_t>(stage) * STAGE_BYTES);
      uint4* xl = reinterpret_cast<uint4*>(stage_base + static_cast<size_t>(stage) * STAGE_BYTES + X_TILE_BYTES);
#pragma unroll
      for (int i = 0; i < static_cast<int>(X_TILE_BYTES / 16 / 128); ++i) {
        const int c = ct + i * 128;
        uint4 v = xs[c];
        uint4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)));
        l.y = tf32_rna(__float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)));
        l.z = tf32_rna(__float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)));
        l.w = tf32_rna(__float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)));
        xs[c] = h;
        xl[c] = l;
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(conv_bar + stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = query row (TMEM lane), running top-k in shared memory.
    // Overlaps the MMAs of the next tile (two accumulator tiles in TMEM).
    const int et = threadIdx.x - 128;           // 0..127 == TMEM lane == query inside the block
    const int q = qb * BM + et;
    const bool q_valid = q < p.B;
    const float qinv = q_valid ? p.q_inv[q] : 0.0f;
    const float qsq = q_valid ? p.q_sq[q] : 0.0f;
    for (int i = 0; i < k; ++i) lists[static_cast<size_t>(i) * BM + et] = 0ull;
    float thr = __int_as_float(0xff800000);     // -inf until the list is full
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int t = 0; t < my_tiles; ++t) {
      const int a = t & 1;
      const long long row0 = static_cast<long long>(slice + t * p.n_slices) * BN;
      const long long rem = p.n_rows - row0;
      const int valid = rem < BN ? static_cast<int>(rem) : BN;
      float* cs = colscale + a * BN;
      // per-column scale of this tile (cosine: 1/|x|, l2: |x|^2); cs[a] was last read for tile t-2
      if (METRIC != kIP) {
        for (int c = et; c < BN; c += 128) {
          float v = 0.0f;
          if (c < valid) v = (METRIC == kCosine) ? __ldg(p.inv_norm + row0 + c) : __ldg(p.sqnorm + row0 + c);
          cs[c] = v;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(smem_u32(tmem_full + a), (static_cast<uint32_t>(t) >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(lane_base + static_cast<uint32_t>(a * BN + c0), r);
        tmem_ld_wait();
        if (c0 < valid && q_valid) {
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (METRIC != kIP) c4 = *reinterpret_cast<const float4*>(cs + c0 + j4);  // warp-wide broadcast
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 + jj;
              const float d = __uint_as_float(r[j]);
              float s;
              if (METRIC == kCosine) s = d * cc[jj] * qinv;
              else if (METRIC == kL2) s = -((cc[jj] - 2.0f * d) + qsq);
              else s = d;
              // !(s < thr) also lets NaN through; the slow path ranks it as -inf like K1 does
              if (!(s < thr) && (c0 + j) < valid) {
                s = (s != s) ? __int_as_float(0xff800000) : s;
                const long long row = row0 + c0 + j;
                bool dead = false;
                if (p.tomb != nullptr) dead = (__ldg(p.tomb + (row >> 5)) >> (row & 31)) & 1u;
                if (!dead) {
                  const uint64_t key = pack_key(s, __ldg(p.gids + row));
                  if (key > lists[static_cast<size_t>(k - 1) * BM + et]) thr = list_push(lists, k, et, key);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty + a));
    }
    if (q_valid) {
      uint64_t* out = p.out_lists + (static_cast<size_t>(p.slice_base + slice) * p.B + q) * k;
      for (int i = 0; i < k; ++i) out[i] = lists[static_cast<size_t>(i) * BM + et];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

constexpr size_t kGemmSmem = 1024 + static_cast<size_t>(STAGES) * STAGE_BYTES + static_cast<size_t>(kMaxKGemm) * BM * 8 +
                             2 * BN * 4 + (3 * STAGES + 4) * 8 + 16;

// ---------------------------------------------------------------- tensor maps (driver entry point, no libcuda link)
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// [rows, ld] fp32 row-major, logical width `cols`; box = BK floats x box_rows, 64B swizzle, OOB -> 0.
bool encode_map(CUtensorMap* map, const void* base, long long rows, int cols, int ld, int box_rows) {
  auto fn = get_encode();
  if (!fn) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows > 0 ? rows : 1)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

size_t gemm_query_workspace_floats(int B, int dim) {
  const int ld = (dim + 3) / 4 * 4;
  return 2 * static_cast<size_t>(B) * ld + 2 * static_cast<size_t>(B);
}

int gemm_max_k() { return kMaxKGemm; }

int gemm_slices_for(long long n_rows, int B, int sm_count) {
  const int n_qblocks = (B + BM - 1) / BM;
  const long long n_tiles = (n_rows + BN - 1) / BN;
  long long s = sm_count / n_qblocks;
  if (s < 1) s = 1;
  if (s > n_tiles) s = n_tiles;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

cudaError_t launch_split_queries(const float* q, int B, int dim, float* workspace, cudaStream_t stream) {
  const int ld = (dim + 3) / 4 * 4;
  float* q_hi = workspace;
  float* q_lo = q_hi + static_cast<size_t>(B) * ld;
  float* q_inv = q_lo + static_cast<size_t>(B) * ld;
  float* q_sq = q_inv + B;
  const int wpb = 8;
  split_queries_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, stream>>>(q, B, dim, ld, q_hi, q_lo, q_inv, q_sq);
  return cudaGetLastError();
}

// One launch per segment.  out_lists: [total_slices][B][k]; this launch fills slices
// [slice_base, slice_base + n_slices).  workspace = output of launch_split_queries.
cudaError_t launch_gemm_topk(const SegDesc& seg, int dim, int dpad, const float* workspace, int B, int k, int metric,
                             int n_slices, int slice_base, uint64_t* out_lists, cudaStream_t stream) {
  if (seg.n_rows <= 0 || n_slices <= 0) return cudaSuccess;
  static std::atomic<unsigned long long> attr_done{0ull};
  const cudaError_t attr_err = once_per_device(attr_done, [&] {
    cudaError_t err = cudaFuncSetAttribute(gemm_topk_kernel<kCosine>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(gemm_topk_kernel<kIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(gemm_topk_kernel<kL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    return err;
  });
  if (attr_err != cudaSuccess) return attr_err;
  const int ld = (dim + 3) / 4 * 4;
  const float* q_hi = workspace;
  const float* q_lo = q_hi + static_cast<size_t>(B) * ld;
  const float* q_inv = q_lo + static_cast<size_t>(B) * ld;
  const float* q_sq = q_inv + B;
  CUtensorMap tm_x, tm_qhi, tm_qlo;
  if (!encode_map(&tm_x, seg.rows, seg.n_rows, dim, dpad, BN) || !encode_map(&tm_qhi, q_hi, B, dim, ld, BM) ||
      !encode_map(&tm_qlo, q_lo, B, dim, ld, BM))
    return cudaErrorInvalidValue;
  GemmParams p;
  p.inv_norm = seg.inv_norm;
  p.sqnorm = seg.sqnorm;
  p.gids = seg.gids;
  p.tomb = seg.tomb;
  p.q_inv = q_inv;
  p.q_sq = q_sq;
  p.n_rows = seg.n_rows;
  p.B = B;
  p.k = k;
  p.metric = metric;
  p.n_kblocks = (dim + BK - 1) / BK;
  p.n_tiles = static_cast<int>((seg.n_rows + BN - 1) / BN);
  p.n_slices = n_slices;
  p.out_lists = out_lists;
  p.slice_base = slice_base;
  dim3 grid((B + BM - 1) / BM, n_slices, 1), block(kThreads, 1, 1);
  if (metric == kCosine) gemm_topk_kernel<kCosine><<<grid, block, kGemmSmem, stream>>>(tm_x, tm_qhi, tm_qlo, p);
  else if (metric == kL2) gemm_topk_kernel<kL2><<<grid, block, kGemmSmem, stream>>>(tm_x, tm_qhi, tm_qlo, p);
  else gemm_topk_kernel<kIP><<<grid, block, kGemmSmem, stream>>>(tm_x, tm_qhi, tm_qlo, p);
  return cudaGetLastError();
}

}  // namespace wdbx
