/* wdbx_b200.h -- C ABI of libwdbx_b200.so, the B200 (sm_100a) exact-search engine that sits under
 * WDBX's VectorIndex / VectorStore.search boundary.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository donaldfilimon/wdbx-py).  The reference-side binding is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every call returns WDBX_B200_OK (0) or a negative error code, never throws, never exits;
 *     wdbx_b200_last_error() returns a thread-local message for the last failure;
 *   - one engine owns the row partitions ("segments", one per logical WDBX shard) that live in
 *     the HBM of ONE device.  Several GPUs: either one engine per rank (one process per GPU), which exchange
 *     packed keys on the device over NVLink peer memory (wdbx_b200_search_exchange*) or through an NCCL
 *     all-gather in the host layer that wdbx_b200_merge() reduces -- or ONE process driving 2-8 engines as a
 *     group (wdbx_b200_group_*);
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *     WDBX_B200_ERR_CUDA.
 *
 * Ranking key (shared by the scan kernels, the merge kernel and the cross-GPU exchange):
 *   key = (monotone_u32(score) << 32) | ~gid        -- bigger key = better hit
 *   so "higher score first, equal score => lower global insertion id first" is an integer max.
 *   key 0 is "empty".  NaN scores are ranked as -inf.
 */
#ifndef WDBX_B200_H
#define WDBX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WDBX_B200_ABI_VERSION 1

/* error codes */
#define WDBX_B200_OK 0
#define WDBX_B200_ERR_ARG (-1)    /* bad argument (dimension, k, segment, NULL pointer ...) */
#define WDBX_B200_ERR_CUDA (-2)   /* CUDA runtime failure (no device, launch error, ...) */
#define WDBX_B200_ERR_OOM (-3)    /* device or host allocation failed */
#define WDBX_B200_ERR_LIMIT (-4)  /* unsupported size (dim too large, k > WDBX_B200_MAX_K, ...) */

/* storage element type of the stored embedding matrix */
#define WDBX_B200_F32 0
#define WDBX_B200_BF16 1 /* rows rounded to bf16 (RNE) at ingest; queries and accumulation stay fp32 */

/* similarity; larger score is always better (VectorStore.search sorts descending,
 * wdbx/core/vector_store.py:330) */
#define WDBX_B200_COSINE 0 /* <x,q>/(|x||q|); 0 if either norm is 0 (FaissIndex._normalize_vector, indexing.py:851-856) */
#define WDBX_B200_IP 1     /* <x,q>  (extension; IndexFlatIP without the normalisation) */
#define WDBX_B200_L2 2     /* -sum (x-q)^2  (extension) */

#define WDBX_B200_MAX_SEGMENTS 64
#define WDBX_B200_MAX_K 1024
#define WDBX_B200_ALL_SEGMENTS (-1)  /* one top-k over all segments merged */
#define WDBX_B200_EACH_SEGMENT (-2)  /* one top-k per segment (search_host only) */

typedef struct wdbx_b200_engine wdbx_b200_engine;

typedef struct wdbx_b200_stats {
  int32_t abi_version;
  int32_t device;
  int32_t dim;          /* logical vector dimension */
  int32_t dim_padded;   /* stored row length in elements (16-byte multiple) */
  int32_t dtype;
  int32_t num_segments;
  int32_t sm_count;
  int32_t last_kernel;  /* dominant kernel of the last timed search: 0 none, 1 K1 streaming scan, 2 K2b filter over the
                           bf16 shadow, 3 K2b small-batch filter over the int8 shadow */
  int64_t rows_total;     /* rows appended over all segments (tombstoned rows included) */
  int64_t rows_live;      /* rows_total minus tombstones */
  int64_t capacity_rows;  /* rows that fit without growing */
  int64_t bytes_resident; /* HBM bytes held by the engine */
  int64_t kernel_launches;     /* kernels launched by this engine since creation */
  int64_t searches;            /* search calls served */
  double last_search_ms;       /* device time of the last wdbx_b200_search_host call */
  int64_t seg_rows[WDBX_B200_MAX_SEGMENTS];
  int64_t seg_live[WDBX_B200_MAX_SEGMENTS];
  double last_kernel_ms;       /* duration of that kernel's launches (CUDA events on its stream); 0 unless
                                  wdbx_b200_set_kernel_timing(e, 1) was called */
  int64_t last_candidates;     /* K2b, kernel timing on: rows the filter of the last search handed to the exact
                                  refine, summed over its queries (0 otherwise) */
} wdbx_b200_stats;

/* ABI version of the loaded library (== WDBX_B200_ABI_VERSION). */
int wdbx_b200_version(void);

/* Thread-local text of the last error raised on the calling thread ("" if none). */
const char* wdbx_b200_last_error(void);

/* Number of CUDA devices visible, or a negative error code.  Used by the host layer to fail
 * loudly when enable_gpu is requested on a box without a GPU. */
int wdbx_b200_device_count(void);

/* Create an engine on `device` holding `num_segments` independent row partitions of
 * `dim`-dimensional vectors.
 * Replaces: VectorStore._init_indices (wdbx/core/vector_store.py:111-134), which builds one
 * FaissIndex/HNSWIndex per shard, and FaissIndex._create_index (wdbx/core/indexing.py:709-758)
 * including its dead `use_gpu` branch (indexing.py:742-748). */
int wdbx_b200_create(int device, int dim, int dtype, int num_segments, wdbx_b200_engine** out);

/* Release every device and host resource of the engine.
 * Replaces: FaissIndex.shutdown (indexing.py:845-849) minus persistence. */
void wdbx_b200_destroy(wdbx_b200_engine* e);

/* Make room for `rows` rows in `segment` without further growth (optional; append grows
 * geometrically).  Replaces hnswlib's init_index(max_elements=...) (indexing.py:274-278). */
int wdbx_b200_reserve(wdbx_b200_engine* e, int segment, int64_t rows);

/* Append n rows (row-major [n, dim] fp32, host memory or device memory of the engine's device)
 * to `segment`; stores them (optionally as bf16), computes 1/|x| and |x|^2 per row (kernel K4)
 * and records gids[i] (host array, may be NULL => gid = running row count of the engine) as
 * the global insertion id used for tie-breaking and returned by search.
 * *first_row_out receives the segment-local index of the first appended row.
 * Replaces: FaissIndex.add / batch_add (indexing.py:858-905, :921-968): normalise + index.add. */
int wdbx_b200_append(wdbx_b200_engine* e, int segment, const float* rows, int64_t n,
                     int src_is_device, const uint32_t* gids, int64_t* first_row_out);

/* Overwrite one stored row in place (same gid) and clear its tombstone.
 * Replaces: the duplicate-id branch of HNSWIndex.add (indexing.py:370-375, replace_vector). */
int wdbx_b200_overwrite(wdbx_b200_engine* e, int segment, int64_t row, const float* v_host);

/* Mark a row dead (dead=1) or alive (dead=0).  Dead rows are never returned.
 * Replaces: FaissIndex.remove / HNSWIndex.remove (indexing.py:1050-1074, :525-560), whose rows
 * keep competing in the top-k (SURVEY.md section 8c decision 1). */
int wdbx_b200_tombstone(wdbx_b200_engine* e, int segment, int64_t row, int dead);

/* Drop all rows of one segment (or all with WDBX_B200_ALL_SEGMENTS); capacity is kept.
 * Replaces: FaissIndex.clear (indexing.py:1089-1112). */
int wdbx_b200_clear(wdbx_b200_engine* e, int segment);

/* Copy one stored row back to the host as fp32 (dim floats).
 * Serves VectorStore.get (vector_store.py:579-596) without a host-side copy of every vector. */
int wdbx_b200_read_row(wdbx_b200_engine* e, int segment, int64_t row, float* out_host);

/* Copy n consecutive stored rows back to the host as dense fp32 [n, dim].
 * Serves persistence: VectorStore._save_vectors (vector_store.py:168-176, a pickle of the host
 * dict) becomes "dump the device partition". */
int wdbx_b200_read_rows(wdbx_b200_engine* e, int segment, int64_t row0, int64_t n, float* out_host);

/* Exact top-k of B queries over one segment, or over all segments merged (segment ==
 * WDBX_B200_ALL_SEGMENTS), entirely on the device and asynchronously on `cuda_stream`
 * (a cudaStream_t; NULL = legacy default stream).  q_dev: [B, dim] fp32 on the device.
 * Outputs (device pointers, any may be NULL): keys_out [B,k] packed ranking keys best-first
 * (0 = empty slot), scores_out [B,k] fp32, gids_out [B,k] int64 (-1 = empty), counts_out [B].
 * Scores never go to HBM (k <= 128): the filter path (K2b) streams a 2-byte shadow of the rows through the
 * tensor cores and re-scores the few surviving rows exactly; the streaming scan (K1) keeps running top-k lists
 * in shared memory; both answers are bit-identical.  k > 128 dumps 8-byte keys and radix-selects.  No
 * synchronisation in steady state => CUDA-graph capturable, the first call after ingest included.
 * Replaces: FaissIndex.search (indexing.py:983-1030) -- normalise the query (:1002),
 * IndexFlatIP.search (:1013) -- and, for ALL_SEGMENTS, the per-shard loop + sort of
 * VectorStore.search (vector_store.py:323-330, :345). */
int wdbx_b200_search(wdbx_b200_engine* e, int segment, const float* q_dev, int B, int k,
                     int metric, uint64_t* keys_out, float* scores_out, int64_t* gids_out,
                     int32_t* counts_out, void* cuda_stream);

/* Same search with HOST buffers: copies the queries in, runs the kernels, copies results out
 * and synchronises.  segment >= 0: that segment only (VectorIndex.search on one shard,
 * indexing.py:983-1030); WDBX_B200_ALL_SEGMENTS: one merged top-k ([B, k] outputs);
 * WDBX_B200_EACH_SEGMENT: one top-k per segment ([num_segments, B, k] outputs, counts
 * [num_segments, B]) -- the candidate set VectorStore.search builds before its metadata
 * post-filter (vector_store.py:323-342).  Any output pointer may be NULL.
 * This is the call the reference-facing plugin makes. */
int wdbx_b200_search_host(wdbx_b200_engine* e, int segment, const float* q_host, int B, int k,
                          int metric, float* scores_host, int64_t* gids_host, uint64_t* keys_host,
                          int32_t* counts_host);

/* Opt-in "strictly better than the reference" search: metadata PRE-filter and
 * threshold push-down.  allow_bitmaps: [num_segments] host pointers (NULL entry = all rows of that
 * segment allowed; NULL array = no filter) to bitmaps over the segment's rows, bit = 1 means the row's
 * metadata matches; min_score: rows scoring below it are never kept (-INFINITY = none).  Returns the
 * exact top-k AMONG THE ALLOWED ROWS, i.e. a full k where the reference's post-filter
 * (vector_store.py:333-342, :414-463) returns a truncated list.  Merged over all segments, host buffers. */
int wdbx_b200_search_filtered_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric,
                                   float min_score, const uint32_t* const* allow_bitmaps,
                                   float* scores_host, int64_t* gids_host, int32_t* counts_host);

/* k-way merge of G best-first key lists per query (keys_dev [G, B, k], e.g. the output of an
 * NCCL all-gather of every rank's wdbx_b200_search keys) into one best-first top-k per query.
 * Replaces: the cross-shard concat + sort + [:limit] of VectorStore.search
 * (vector_store.py:324-330, :345) when shards live on different GPUs. */
int wdbx_b200_merge(wdbx_b200_engine* e, const uint64_t* keys_dev, int G, int B, int k,
                    uint64_t* keys_out, float* scores_out, int64_t* gids_out, int32_t* counts_out,
                    void* cuda_stream);

/* Fused cross-GPU merge (one process per GPU, all GPUs of one box).  Instead of "local top-k ->
 * NCCL all-gather -> merge kernel", the scan kernel's last CTA pushes its k keys per query into
 * every peer's exchange buffer with NVLink peer-to-peer stores, waits (bounded) for the peers'
 * pushes and merges the G lists itself: one launch per query batch on every rank, no collective
 * call.  Replaces: the cross-shard concat + sort + [:limit] of VectorStore.search
 * (vector_store.py:324-330, :345) for shards that live on different GPUs.
 *   1. every rank: wdbx_b200_exchange_init -> 64-byte CUDA IPC handle of its exchange buffer;
 *   2. the host layer all-gathers the handles (torch.distributed) and every rank calls
 *      wdbx_b200_exchange_attach with the world x 64 byte array (rank order);
 *   3. wdbx_b200_search_exchange is then a COLLECTIVE: every rank must call it with the same
 *      (B, k, metric) in the same order; every rank receives the merged global top-k.
 * Limits: B <= 8 queries (one pass), k <= 128, world <= 8.  counts_out = -1 signals that a peer
 * did not show up within ~3 s. */
#define WDBX_B200_IPC_HANDLE_BYTES 64
int wdbx_b200_exchange_init(wdbx_b200_engine* e, int rank, int world, void* ipc_handle_out);
int wdbx_b200_exchange_attach(wdbx_b200_engine* e, int world, const void* ipc_handles);
int wdbx_b200_search_exchange(wdbx_b200_engine* e, const float* q_dev, int B, int k, int metric,
                              uint64_t* keys_out, float* scores_out, int64_t* gids_out,
                              int32_t* counts_out, void* cuda_stream);

/* Host-buffer form of wdbx_b200_search_exchange: pinned H2D of the queries -> collective search + on-device
 * exchange -> ONE D2H of the packed result -> synchronise.  Every rank must call it with the same queries.
 * Replaces: VectorStore.search's shard loop + merge (vector_store.py:323-330, :345) when the shards live on
 * several GPUs. */
int wdbx_b200_search_exchange_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric,
                                   float* scores_host, int64_t* gids_host, uint64_t* keys_host,
                                   int32_t* counts_host);

/* wdbx_b200_search_exchange_host with the opt-in pre-filter of wdbx_b200_search_filtered_host: every rank passes the
 * bitmaps of ITS rows (and the same min_score); the result is the exact top-k among the allowed rows of all ranks.
 * Replaces: the metadata post-filter of VectorStore.search (vector_store.py:333-342, :414-463) on a store
 * whose shards live on several GPUs. */
int wdbx_b200_search_exchange_filtered_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric,
                                            float min_score, const uint32_t* const* allow_bitmaps,
                                            float* scores_host, int64_t* gids_host, int32_t* counts_host);

/* SINGLE PROCESS, SEVERAL GPUs.  A group ties n engines (2..8, one per device of one box, same dim / dtype /
 * num_segments) together so that an ordinary process -- the reference's REST server (wdbx/api/server.py:141-152)
 * or CLI (wdbx/cli.py:541) behind WDBX.vector_search -- uses all GPUs without torchrun: peer access is enabled
 * directly, every group search is launched on all devices from the calling thread and the devices merge their
 * top-k over NVLink (the same on-device exchange as wdbx_b200_search_exchange for B <= 8 and k <= 128, peer
 * copies of the packed keys to device 0 + the merge kernel otherwise).  The host layer stripes the rows of every
 * segment over the engines (row n -> engine n % n_engines) with the per-engine calls above.
 * Replaces: ShardManager._allocate_shards (wdbx/core/distributed.py:547-654) + the shard loop and merge of
 * VectorStore.search (vector_store.py:323-330, :345).  The group does not own the engines: destroy it first. */
typedef struct wdbx_b200_group wdbx_b200_group;
int wdbx_b200_group_create(wdbx_b200_engine* const* engines, int n, wdbx_b200_group** out);
void wdbx_b200_group_destroy(wdbx_b200_group* g);

/* Host-buffer search over the whole group; `segment`, outputs and semantics as wdbx_b200_search_host.
 * min_score / allow_bitmaps as wdbx_b200_search_filtered_host (-INFINITY / NULL = unfiltered); allow_bitmaps is
 * engine-major: [n_engines * num_segments] pointers to bitmaps over each engine's OWN rows of each segment. */
int wdbx_b200_group_search_host(wdbx_b200_group* g, int segment, const float* q_host, int B, int k, int metric,
                                float min_score, const uint32_t* const* allow_bitmaps, float* scores_host,
                                int64_t* gids_host, uint64_t* keys_host, int32_t* counts_host);

/* Device-resident group search: q_dev0 [B, dim] fp32 and the outputs live on the FIRST engine's device;
 * asynchronous with respect to the host, ordered after / before `cuda_stream` (a stream of that device). */
int wdbx_b200_group_search(wdbx_b200_group* g, const float* q_dev0, int B, int k, int metric, uint64_t* keys_out,
                           float* scores_out, int64_t* gids_out, int32_t* counts_out, void* cuda_stream);

/* Override the scan kernel's launch geometry (0 / -1 = automatic): consumer warps per CTA,
 * TMA pipeline stages per warp, rows held per lane group (1, 2, 4), CTAs, L2 evict-first hint.
 * Benchmark / profiling hook; the counterpart of the reference's HNSW_EF_SEARCH / FAISS_NPROBE
 * knobs (indexing.py:242-245, :689-690) in the sense of "search-time tuning", results are
 * identical for every setting. */
int wdbx_b200_set_tuning(wdbx_b200_engine* e, int warps, int stages, int rows_unroll, int grid,
                         int evict_first);

/* Change a routing knob of a live engine (the WDBX_B200_* environment variables read at creation):
 * "shadow_min_mb" (-1 = never use the bf16-shadow filter for small batches, else the store size in MiB from which
 * it is used), "gemm_min_batch" (0 = never use the tensor-core path), "gemm_mode", "pdl", "queries_per_pass",
 * "filter_i8" (0 = small batches stream the bf16 shadow instead of the 1-byte one),
 * "overlap" (1 = consecutive device-resident small-batch searches on one stream overlap on the device: the next
 * search streams its first tiles while the previous one finishes its tail; results stay in launch order.  Contract:
 * the query buffer of such a search must not be produced by work enqueued on that stream after the previous search).
 * Results are identical for every setting; benchmark hook (bench.py times the fp32 streaming scan and the
 * filter path on the same resident matrix).  No reference counterpart. */
int wdbx_b200_set_option(wdbx_b200_engine* e, const char* name, long long value);

/* Bracket the dominant kernel of every search (K1 scan, or the K2b filter launches) with CUDA events on
 * the search stream; wdbx_b200_get_stats then reports its duration.  Measurement hook for bench.py's
 * roofline (off by default); no reference counterpart. */
int wdbx_b200_set_kernel_timing(wdbx_b200_engine* e, int enable);

/* Fill *out.  Replaces: FaissIndex.get_stats / size (indexing.py:1161-1183). */
int wdbx_b200_get_stats(wdbx_b200_engine* e, wdbx_b200_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* WDBX_B200_H */
