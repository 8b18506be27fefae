// kernels.h -- host-visible launch interface between engine.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdbx {

constexpr int kMaxSeg = 64;
constexpr int kMaxK = 1024;

enum Metric : int { kCosine = 0, kIP = 1, kL2 = 2 };

// Exchange buffer layout (per rank, peer-mapped): keys [2 slots][kMaxPeers][kXchgMaxB][kXchgMaxK] u64,
// then flags [2 slots][kMaxPeers] u32 (sequence number of the last completed push of that peer).
constexpr int kMaxPeers = 8;
constexpr int kXchgMaxB = 8;
constexpr int kXchgMaxK = 128;
constexpr size_t kXchgKeyCount = 2ull * kMaxPeers * kXchgMaxB * kXchgMaxK;
constexpr size_t kXchgBytes = kXchgKeyCount * 8 + 2 * kMaxPeers * 4 + 64;

// One collective key exchange (world <= 1: none).  peer[r] = base of rank r's exchange buffer (own rank included).
struct XchgCtx {
  uint64_t* peer[kMaxPeers];
  int world, rank, slot;   // slot = seq & 1
  unsigned int seq;        // collective sequence number (same on all ranks)
};

struct SegDesc {
  const unsigned char* rows;  // [n_rows][row_bytes]
  const float* inv_norm;      // [n_rows] 1/|x| (0 for zero rows)
  const float* sqnorm;        // [n_rows] |x|^2 (GEMM path, l2)
  const uint32_t* gids;       // [n_rows] global insertion ids
  const uint32_t* tomb;       // bitmap, 1 = dead; may be NULL when the segment has no tombstones
  const uint32_t* allow;      // optional per-search bitmap, 1 = row may be returned (metadata pre-filter)
  long long n_rows;
};

// Tuning knobs of the streaming scan (K1); 0 = pick automatically.
struct ScanTuning {
  int warps;      // consumer warps per CTA (each owns a private TMA pipeline)
  int stages;     // pipeline depth per warp
  int rows_unroll;  // U: rows held per lane group per tile (1, 2 or 4)
  int grid;       // CTAs (default: one per SM)
  int evict_first;  // -1 auto, 0 off, 1 on
  int queries_per_pass;  // QB: queries scored per streamed row (1, 2, 4, 8); 0 = auto (up to 8)
};

struct ScanParams {
  SegDesc seg[kMaxSeg];
  long long tile_end[kMaxSeg];  // exclusive prefix sum of tiles per segment
  int n_seg;
  long long total_tiles;
  const float* q;  // [B][dim] fp32
  int B;
  int dim;        // logical dimension
  int dpad;       // stored elements per row
  int row_bytes;  // dpad * sizeof(elem)
  int cpr;        // 16-byte chunks per row
  int lpr_log2;   // log2(lanes per row)
  int nch;        // chunks per lane = ceil(cpr / lanes_per_row)
  int tile_rows;  // rows per warp tile = U * (32 >> lpr_log2)
  int stages;
  int stage_bytes;
  int k;
  int metric;
  int evict_first;
  uint64_t* cand;          // [B][grid][k] per-CTA partial lists
  unsigned int* counters;  // [gridDim.y] last-block-done tickets (zero between launches)
  // fused cross-GPU exchange (world > 1, gridDim.y == 1): the last CTA pushes its k keys per query
  // into every peer's exchange buffer over NVLink (P2P stores), waits for the peers' pushes and
  // merges the G lists itself -- no NCCL call, no second kernel.
  uint64_t* xchg_peer[kMaxPeers];  // base of every rank's exchange buffer (own rank included)
  int xchg_world;                  // 0/1 = exchange disabled
  int xchg_rank;
  int xchg_slot;                   // seq & 1
  unsigned int xchg_seq;           // collective sequence number (same on all ranks)
  uint64_t* all_keys;      // large-k mode: [B][all_rows] ranking key of every row (NULL = normal top-k mode)
  long long all_rows;      // rows over all scanned segments
  long long seg_row_base[kMaxSeg];  // first index of segment n inside a query's all_keys row
  float min_score;         // score floor (threshold push-down); -inf = none
  const int* only_flag;    // optional [B]: a query block runs only if one of its queries is flagged
                           // (K2b re-runs queries whose candidate list overflowed); NULL = run all
  // flag-gated launches close a numbered search of the fused filter path: once every query block has finished (or
  // found nothing to do) the search number is published, see "overlapping consecutive searches" in engine.cu
  unsigned int* done_ctr;     // NULL = nothing to publish
  unsigned int* done_blocks;  // zero at launch: query blocks finished so far
  unsigned int done_sn;
  uint64_t* keys_out;      // [B][k] or NULL
  float* scores_out;       // [B][k] or NULL
  long long* gids_out;     // [B][k] or NULL
  int* counts_out;         // [B] or NULL
};

struct ScanPlan {
  int lpr_log2, nch, U, KS, tile_rows, stage_bytes, stages, warps, grid;
  size_t smem_bytes;
  int queries_per_block;
  int pdl;  // launch with programmatic dependent launch: the scan of query i+1 overlaps the merge tail of query i
};

// Fill the derived fields (plan) for the given shape.  Returns 0 or a negative wdbx error code.
int scan_plan(int dim, int dpad, int elem_bytes, int k, int B, int sm_count, const ScanTuning& tune, ScanPlan* plan);

// Launch K1 (+ fused last-block merge).  p.seg / tile_end / totals must be consistent with plan.
cudaError_t launch_scan_topk(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream);

// K2: tcgen05 GEMM + fused top-k (large batches, fp32 storage).  See gemm_topk.cu.
size_t gemm_query_workspace_floats(int B, int dim);
int gemm_max_k();
int gemm_slices_for(long long n_rows, int B, int sm_count);
cudaError_t launch_split_queries(const float* q, int B, int dim, float* workspace, cudaStream_t stream);
cudaError_t launch_gemm_topk(const SegDesc& seg, int dim, int dpad, const float* workspace, int B, int k, int metric,
                             int n_slices, int slice_base, uint64_t* out_lists, cudaStream_t stream);

// K2b: bf16 tensor-core filter + exact fp32 refine (results bit-identical to K1).  See gemm_filter.cu.
int filter_max_k();
int filter_final_cap();
int filter_ld16(int dim);
size_t filter_query_workspace_bytes(int B, int dim);
int filter_slices_for(long long n_rows, int B, int sm_count);
int filter_regions_per_slice(int B);
float filter_acc_rel(int dim, int dpad);
float filter_c_l2(int dpad);
// bf16 shadow rows + per-row upper bound of |x - bf16(x)| (rres)
cudaError_t launch_shadow_rows(const float* rows, long long n, int dpad, int ld16, void* dst, float* rres, cudaStream_t stream);
cudaError_t launch_prep_queries(const float* q, int B, int dim, void* workspace, unsigned int* zero, size_t n_zero,
                                bool small, const unsigned int* done_ctr, unsigned int wait_sn, bool overlap, bool pdl,
                                unsigned int* prep_count, unsigned int* prep_ctas, bool i8, cudaStream_t stream);
int filter_ld8(int dim);
// int8 shadow rows (x ~ sx * xi, xi in [-127, 127]) + per-row scale + upper bound of |x - sx * xi|
cudaError_t launch_shadow8_rows(const void* rows, bool src_bf16, long long n, int dpad, int ld8, void* dst, float* sx,
                                float* rres, cudaStream_t stream);
// Small batches (filter_fused_tail(B)): the filter kernel itself re-scores its candidates, and the last CTA of
// the search merges all CTA lists, runs the cross-GPU exchange and emits -- no refine / exchange launch.
struct FilterTail {
  const float* q;            // [B][dim] fp32 queries
  int dpad, elem_bytes;      // stored row layout
  int lpr_log2, nch;         // K1's lane mapping (scan_plan): the re-scored keys are bit-identical to K1's
  float min_score;           // score floor (-inf = none)
  int* overflow;             // [B] set when a candidate region overflowed (the flag-gated K1 launch re-runs the query)
  uint64_t* fin_keys;        // [B][filter_final_cap()] exact keys that reached the query's bound (final lists)
  unsigned int* fin_count;   // [B] zero at launch
  unsigned int* ticket;      // zero at launch; the CTA that draws s_total - 1 finishes the search
  unsigned int* tile_ctr;    // [kMaxSeg] zero at launch: dynamic tile counter of every segment's launch
  XchgCtx xchg;
  unsigned int* part_max;    // [B][filter_max_k()] zero at launch: partition maxima of the lower bounds
  const float* rowscale;     // int8 shadow: per-row scale (x ~ rowscale * xi); NULL = bf16 operand
  const unsigned int* prep_count;   // overlap mode: the filter waits for its prep by CTA count (NULL: by launch order)
  unsigned int prep_target;
  unsigned int* done_ctr;    // search numbers completed on this workspace (the last CTA waits for done_sn - 1 before it
  unsigned int done_sn;      // exchanges / emits: searches overlap on the device, their results stay ordered)
  uint64_t* keys_out;
  float* scores_out;
  long long* gids_out;
  int* counts_out;
};
bool filter_fused_tail(int B);
cudaError_t launch_publish_done(unsigned int* done_ctr, unsigned int sn, cudaStream_t stream);
cudaError_t launch_gemm_filter(const void* xb, int ld_x, const float* rres, const SegDesc& seg, int seg_index, int dim,
                               const void* workspace, int B, int k, int metric, float acc_rel, float c_l2, int n_slices,
                               unsigned long long* cand,
                               unsigned int* cand_count, unsigned int* lower_glob, unsigned int* lower_list, int cap,
                               int slice_base, int s_total, const FilterTail* tail, bool pdl, cudaStream_t stream);
cudaError_t launch_refine_topk(const SegDesc* segs, int n_seg, const float* q, int B, int dim, int dpad, int elem_bytes,
                               int lpr_log2, int nch, int k, int metric, const unsigned long long* cand,
                               const unsigned int* cand_count, int cap, int s_total, int* overflow, int ctas_per_query,
                               uint64_t* part, unsigned int* tickets, uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out,
                               cudaStream_t stream);

// Large k (128 < k <= 1024): exact top-k of n_keys ranking keys per query by radix select (select_topk.cu).
size_t select_workspace_words(int B);
cudaError_t launch_select_topk(const uint64_t* all_keys, long long n_keys, int B, int k, unsigned int* workspace,
                               uint64_t* sel_keys, int sm_count, uint64_t* keys_out, float* scores_out, long long* gids_out,
                               int* counts_out, cudaStream_t stream);

// K3: merge G best-first lists per query.
cudaError_t launch_merge_topk(const uint64_t* keys, int G, int B, int k, uint64_t* keys_out, float* scores_out,
                              long long* gids_out, int* counts_out, cudaStream_t stream);

// Stand-alone NVLink key exchange + merge (same protocol as the one fused into K1's last CTA); see exchange.cu.
cudaError_t launch_exchange_merge(uint64_t* const* peers, int world, int rank, unsigned int seq, const uint64_t* keys_in,
                                  int B, int k, uint64_t* keys_out, float* scores_out, long long* gids_out,
                                  int* counts_out, cudaStream_t stream);

// K4: ingest rows (fp32 source) into the stored layout + norms.
cudaError_t launch_append_rows(const float* src, long long n, int dim, int dpad, bool bf16, unsigned char* dst_rows,
                               float* inv_norm, float* sqnorm, uint32_t* gids_dst, const uint32_t* gids_src,
                               uint32_t gid_base, cudaStream_t stream);

// stored row -> fp32 (read_row)
cudaError_t launch_export_rows(const unsigned char* rows, long long n, int dim, int row_bytes, bool bf16, float* dst,
                               cudaStream_t stream);

}  // namespace wdbx
