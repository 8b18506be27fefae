// engine.cu -- C ABI of libwdbx_b200.so (see include/wdbx_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host-side runtime: device-resident segment
// store (rows + norms + gids + tombstones in HBM), growth, per-stream workspaces, the pinned
// host path, and the launch plumbing for kernels K1/K3/K4.
#include "../../include/wdbx_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "kernels.h"

using namespace wdbx;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      cudaGetLastError();                                                                              \
      return fail(_e == cudaErrorMemoryAllocation ? WDBX_B200_ERR_OOM : WDBX_B200_ERR_CUDA, "%s: %s", \
                  #expr, cudaGetErrorString(_e));                                                      \
    }                                                                                                  \
  } while (0)


int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// ---- growable device arrays ----------------------------------------------------------------------------
// A segment's arrays grow IN PLACE: a virtual address range is reserved once (CUDA virtual memory management,
// cuMemAddressReserve) and physical memory is mapped behind it as rows arrive (cuMemCreate + cuMemMap).  The base
// pointer never changes, so growth needs no copy, no transient second buffer and no synchronisation with searches
// in flight (they only read rows that existed when they were launched) -- the reference's analogue is
// IndexFlatIP.add appending in place (wdbx/core/indexing.py:890).  The earlier realloc-and-copy growth held up to
// 2.5x the store for a moment and threw the bf16 shadow away.  If the driver entry points are unavailable the
// buffer falls back to realloc + copy.
struct VmmApi {
  PFN_cuMemAddressReserve_v10020 reserve = nullptr;
  PFN_cuMemAddressFree_v10020 afree = nullptr;
  PFN_cuMemCreate_v10020 create = nullptr;
  PFN_cuMemRelease_v10020 release = nullptr;
  PFN_cuMemMap_v10020 map = nullptr;
  PFN_cuMemUnmap_v10020 unmap = nullptr;
  PFN_cuMemSetAccess_v10020 set_access = nullptr;
  PFN_cuMemGetAllocationGranularity_v10020 granularity = nullptr;
  bool ok = false;
};

const VmmApi& vmm() {
  static VmmApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    if (env_int("WDBX_B200_VMM", 1) == 0) return;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
             *fn != nullptr;
    };
    api.ok = get("cuMemAddressReserve", reinterpret_cast<void**>(&api.reserve)) &&
             get("cuMemAddressFree", reinterpret_cast<void**>(&api.afree)) &&
             get("cuMemCreate", reinterpret_cast<void**>(&api.create)) &&
             get("cuMemRelease", reinterpret_cast<void**>(&api.release)) &&
             get("cuMemMap", reinterpret_cast<void**>(&api.map)) && get("cuMemUnmap", reinterpret_cast<void**>(&api.unmap)) &&
             get("cuMemSetAccess", reinterpret_cast<void**>(&api.set_access)) &&
             get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&api.granularity));
    cudaGetLastError();
  });
  return api;
}

struct GrowBuf {
  unsigned char* ptr = nullptr;   // base (stable once reserved)
  size_t bytes = 0;               // usable bytes (mapped, or allocated in fallback mode)
  // virtual-memory mode
  CUdeviceptr va = 0;
  size_t va_bytes = 0, gran = 0;
  std::vector<CUmemGenericAllocationHandle> handles;
  std::vector<size_t> sizes;
  bool vm = false;
};

void growbuf_free(GrowBuf& b) {
  if (b.vm) {
    const VmmApi& v = vmm();
    size_t off = 0;
    for (size_t i = 0; i < b.handles.size(); ++i) {
      v.unmap(b.va + off, b.sizes[i]);
      v.release(b.handles[i]);
      off += b.sizes[i];
    }
    if (b.va) v.afree(b.va, b.va_bytes);
  } else {
    cudaFree(b.ptr);
  }
  b = GrowBuf();
}

// Make at least `need` bytes usable, preserving the first `keep` bytes.  `max_bytes`: upper bound the array can ever
// reach (sizes the address reservation).  Returns cudaSuccess / cudaErrorMemoryAllocation / another error; on
// failure the buffer is unchanged.  *moved is set when the base pointer changed (fallback mode only: the caller
// must then wait for in-flight readers before freeing the old copy -- done here with a device synchronise).
cudaError_t growbuf_grow(GrowBuf& b, int device, size_t need, size_t keep, size_t max_bytes, cudaStream_t stream) {
  if (need <= b.bytes) return cudaSuccess;
  const VmmApi& v = vmm();
  if (v.ok && (b.vm || b.ptr == nullptr)) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    if (b.va == 0) {
      size_t gran = 0;
      if (v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2u << 20;
      const size_t want = (std::max(max_bytes, need) + gran - 1) / gran * gran;
      CUdeviceptr va = 0;
      if (v.reserve(&va, want, 0, 0, 0) == CUDA_SUCCESS) {
        b.va = va;
        b.va_bytes = want;
        b.gran = gran;
        b.vm = true;
        b.ptr = reinterpret_cast<unsigned char*>(va);
      }
    }
    if (b.vm) {
      if (need > b.va_bytes) return cudaErrorMemoryAllocation;   // beyond the reservation (cannot happen: max_bytes)
      const size_t add = (need - b.bytes + b.gran - 1) / b.gran * b.gran;
      CUmemGenericAllocationHandle h;
      if (v.create(&h, add, &prop, 0) != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
      if (v.map(b.va + b.bytes, add, 0, h, 0) != CUDA_SUCCESS) {
        v.release(h);
        return cudaErrorMemoryAllocation;
      }
      CUmemAccessDesc acc;
      memset(&acc, 0, sizeof(acc));
      acc.location = prop.location;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      if (v.set_access(b.va + b.bytes, add, &acc, 1) != CUDA_SUCCESS) {
        v.unmap(b.va + b.bytes, add);
        v.release(h);
        return cudaErrorUnknown;
      }
      b.handles.push_back(h);
      b.sizes.push_back(add);
      b.bytes += add;
      return cudaSuccess;
    }
  }
  // fallback: realloc + copy; readers of the old buffer (any stream) must finish before it is freed
  unsigned char* n = nullptr;
  cudaError_t err = cudaMalloc(&n, need);
  if (err != cudaSuccess) { cudaGetLastError(); return err; }
  if (keep > 0 && b.ptr) {
    err = cudaMemcpyAsync(n, b.ptr, keep, cudaMemcpyDeviceToDevice, stream);
    if (err != cudaSuccess) { cudaFree(n); return err; }
  }
  err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { cudaFree(n); return err; }
  cudaFree(b.ptr);
  b.ptr = n;
  b.bytes = need;
  return cudaSuccess;
}

struct Segment {
  GrowBuf b_rows, b_inv, b_sq, b_gids, b_tomb, b_shadow, b_rres;   // backing stores of the pointers below
  GrowBuf b_shadow8, b_sc8, b_rres8;
  void* shadow8 = nullptr;   // int8 copy of the fp32 rows (x ~ sc8 * xi) for the small-batch filter kernel
  float* sc8 = nullptr;      // per-row scale
  float* rres8 = nullptr;    // per-row bound of |x - sc8 * xi|
  int64_t shadow8_rows = 0, shadow8_cap = 0;
  unsigned char* rows = nullptr;
  float* inv_norm = nullptr;
  float* sqnorm = nullptr;
  uint32_t* gids = nullptr;
  uint32_t* tomb = nullptr;  // device bitmap, allocated on first tombstone
  uint32_t* allow = nullptr; // device copy of a per-search "allowed rows" bitmap (metadata pre-filter)
  size_t allow_words = 0;
  void* shadow = nullptr;    // bf16 copy of the fp32 rows for the tensor-core filter (K2b), built lazily
  float* rres = nullptr;     // per-row bound of |x - bf16(x)| (same capacity as the shadow): the data-derived filter bound
  int64_t shadow_rows = 0, shadow_cap = 0;
  std::vector<uint32_t> tomb_host;
  int64_t n_rows = 0, cap_rows = 0, n_dead = 0;
};

struct Workspace {
  cudaStream_t stream = nullptr;
  uint64_t* cand = nullptr;
  size_t cand_keys = 0;
  unsigned int* counters = nullptr;
  int n_counters = 0;
  float* qsplit = nullptr;  // K2: q_hi | q_lo | 1/|q| | |q|^2
  size_t qsplit_floats = 0;
  // K2b: bf16 queries + norms | candidate rows [B][cap] | per-query count, shared lower bound, overflow flag
  void* fws = nullptr;
  size_t fws_bytes = 0;
  unsigned long long* fcand = nullptr;
  size_t fcand_n = 0;
  unsigned int* fzero = nullptr;  // per-search zeroed block (counts, flags, tickets, lower-bound lists)
  size_t fzero_n = 0;
  uint64_t* fpart = nullptr;  // refine partial lists [ctas_per_query][B][k]
  size_t fpart_n = 0;
  // large-k select path: every row's key [chunk][rows] | selected keys [chunk][1024] | select state
  uint64_t* sel_all = nullptr;
  size_t sel_all_n = 0;
  uint64_t* sel_keys = nullptr;
  unsigned int* sel_ws = nullptr;
  int sel_chunk = 0;
  // overlapping consecutive searches of the fused small-batch path (see filter_segments)
  unsigned int* done = nullptr;    // device: number of the last such search that has completely finished
  unsigned int fseq = 0;           // host: number of the last such search launched on this stream
  unsigned int prep_target = 0;    // host: prep CTAs launched so far in overlap mode (device count: done[1])
  bool last_fused = false;         // the launch before this one on the stream was such a search
};

}  // namespace

struct wdbx_b200_engine {
  int device = 0, dim = 0, dpad = 0, dtype = 0, elem_bytes = 4, nseg = 1, sm_count = 148;
  Segment seg[kMaxSeg];
  std::mutex mu;  // segments, workspaces, launches
  uint32_t next_gid = 0;
  std::vector<Workspace> ws;
  ScanTuning tune{0, 0, 0, 0, -1, 0};
  // small batches use the bf16-shadow filter from this many stored bytes on (<0: never).  Measured: at 1.5 GB
  // (1M x 384) the filter path already wins (195 vs 222 us per query); below ~1 GB its extra launches cost more
  long long shadow_min_bytes = 1ll << 30;
  int gemm_min_batch = 16;  // B >= this => tcgen05 path (0 = never); measured crossover vs K1 (8 queries/pass) ~ 12-16
  int gemm_mode = 0;        // 0 = bf16 filter + exact refine (K2b), 1 = 3xTF32 with fused top-k (K2)
  int pdl = 1;              // programmatic dependent launch between the launches of a search (WDBX_B200_PDL=0 disables)
  bool host_call = false;   // the search being launched came through a host-buffer entry point (copies around it)
  int overlap = 0;          // consecutive device-resident small-batch searches on one stream may overlap (opt-in, see
                            // filter_segments; WDBX_B200_OVERLAP / wdbx_b200_set_option "overlap")
  bool shadow_warned = false;
  std::vector<void*> retired;   // workspaces outgrown during a stream capture (freed with the engine)
  bool shadow_failed = false;   // the bf16 shadow could not be allocated: fp32 stores are served by K1 only
  int filter_i8 = 1;            // small batches stream a 1-byte (int8) shadow instead of the bf16 one (WDBX_B200_FILTER_I8)
  bool shadow8_failed = false;
  size_t total_mem = 0;         // device memory (sizes the address reservations of the growable arrays)
  cudaStream_t mstream = nullptr;  // mutations
  // staging for host-sourced appends
  float* stage_rows = nullptr;
  size_t stage_rows_bytes = 0;
  uint32_t* stage_gids = nullptr;
  size_t stage_gids_n = 0;
  // host search path
  std::mutex host_mu;
  cudaStream_t hstream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float* hq_pinned = nullptr;
  float* dq = nullptr;
  size_t q_floats = 0;
  unsigned char* hres_pinned = nullptr;
  unsigned char* dres = nullptr;
  size_t res_bytes = 0;
  // fused cross-GPU exchange (peer-mapped buffers, see kernels.h)
  uint64_t* xbuf = nullptr;
  uint64_t* xpeer[kMaxPeers] = {nullptr};
  bool xopened[kMaxPeers] = {false};
  int xworld = 0, xrank = 0;
  unsigned int xseq = 0;
  uint64_t* xkeys = nullptr;  // local [B][k] keys of the filter path before the stand-alone exchange
  // stats
  std::atomic<long long> launches{0}, searches{0};
  double last_search_ms = 0.0;
  // optional timing of the dominant kernel (wdbx_b200_set_kernel_timing)
  bool ktiming = false;
  cudaEvent_t kev0 = nullptr, kev1 = nullptr;
  int last_kernel = 0;      // 1 = K1 scan, 2 = K2b filter
  bool kpending = false;    // events recorded, not read yet
  const unsigned int* last_fcount = nullptr;  // K2b: candidate counts of the last timed search (device) ...
  size_t last_fregions = 0;                   // ... one per (query, region)
};

struct GroupJob {   // one group search, as seen by the launcher threads
  int segment = 0, B = 0, k = 0, metric = 0;
  const float* q_dev0 = nullptr;
  cudaEvent_t ev_in = nullptr;
  float min_score = 0.0f;
  const uint32_t* const* allow = nullptr;
  bool exchange = false;
  uint64_t* out0_keys = nullptr;
  float* out0_scores = nullptr;
  long long* out0_gids = nullptr;
  int* out0_counts = nullptr;
};

struct wdbx_b200_group {
  std::vector<wdbx_b200_engine*> eng;   // engine i = rank i of the exchange; not owned
  std::mutex mu;                        // one group search at a time
  float* hq = nullptr;                  // portable pinned staging: queries in, packed result out
  size_t hq_floats = 0;
  unsigned char* hres = nullptr;
  size_t hres_bytes = 0;
  uint64_t* gkeys = nullptr;            // device 0: [lists][G][B][k] gathered keys (batches beyond the exchange's limits)
  size_t gkeys_n = 0;
  std::vector<cudaEvent_t> ev;          // per engine: "its keys have arrived on device 0"
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;   // device-resident searches: hand-over with the caller's stream
  // launcher threads (one per engine > 0), see group_worker
  std::vector<std::thread> workers;
  std::mutex wmu;
  std::condition_variable wcv;
  std::atomic<uint64_t> gen{0};
  std::atomic<int> done{0}, sleepers{0};
  std::atomic<bool> stop{false};
  GroupJob job;
  std::vector<int> rcs;
  std::vector<std::string> errs;
};

namespace {

void group_worker(wdbx_b200_group* g, int i);

// While the caller's stream is being captured into a CUDA graph, host-side calls such as cudaMalloc (workspace
// growth) or a kernel's first-use set-up are legal only in the thread's RELAXED capture mode: switch for the
// duration of one search call, restore afterwards.
struct CaptureRelax {
  bool active = false;
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  explicit CaptureRelax(cudaStream_t stream) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
      active = cudaThreadExchangeStreamCaptureMode(&mode) == cudaSuccess;
    else
      cudaGetLastError();
  }
  ~CaptureRelax() {
    if (active) cudaThreadExchangeStreamCaptureMode(&mode);
  }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

size_t row_bytes(const wdbx_b200_engine* e) { return static_cast<size_t>(e->dpad) * e->elem_bytes; }

int64_t round_cap(int64_t rows) { return (rows + 127) / 128 * 128; }

void free_segment(Segment& s) {
  cudaFree(s.allow);
  growbuf_free(s.b_shadow);
  growbuf_free(s.b_rres);
  growbuf_free(s.b_shadow8);
  growbuf_free(s.b_sc8);
  growbuf_free(s.b_rres8);
  growbuf_free(s.b_rows);
  growbuf_free(s.b_inv);
  growbuf_free(s.b_sq);
  growbuf_free(s.b_gids);
  growbuf_free(s.b_tomb);
  s = Segment();
}

// rows the segment could ever hold on this device (sizes the address reservations)
int64_t segment_max_rows(const wdbx_b200_engine* e) {
  const size_t per_row = static_cast<size_t>(e->dpad) * e->elem_bytes;
  int64_t by_mem = static_cast<int64_t>(e->total_mem / (per_row ? per_row : 1)) + 1024;
  return std::min<int64_t>(by_mem, 0xFFFFFFF0ll);
}

bool shadow_wanted(const wdbx_b200_engine* e) {
  if (e->gemm_mode == 1 || e->gemm_min_batch <= 0) return false;
  if (e->dtype == WDBX_B200_F32) return !e->shadow_failed;
  // bf16 stores are their own bf16 operand; what they can still gain is the 1-byte shadow of the small-batch kernel
  return e->filter_i8 != 0 && !e->shadow8_failed && e->shadow_min_bytes >= 0;
}

void refresh_pointers(Segment& s) {
  s.rows = s.b_rows.ptr;
  s.inv_norm = reinterpret_cast<float*>(s.b_inv.ptr);
  s.sqnorm = reinterpret_cast<float*>(s.b_sq.ptr);
  s.gids = reinterpret_cast<uint32_t*>(s.b_gids.ptr);
  s.tomb = reinterpret_cast<uint32_t*>(s.b_tomb.ptr);
  s.shadow = s.b_shadow.ptr;
  s.rres = reinterpret_cast<float*>(s.b_rres.ptr);
  s.shadow8 = s.b_shadow8.ptr;
  s.sc8 = reinterpret_cast<float*>(s.b_sc8.ptr);
  s.rres8 = reinterpret_cast<float*>(s.b_rres8.ptr);
}

// bf16 shadow + per-row residual bound for `cap` rows (fp32 stores: the operand of the tensor-core filter).
// Failure is not an error: the engine then serves every search from the stored rows and says so once.
void grow_shadow(wdbx_b200_engine* e, Segment& s, int64_t cap) {
  const int ld16 = filter_ld16(e->dim);
  const int64_t maxr = segment_max_rows(e);
  if (e->dtype == WDBX_B200_F32) {
    const cudaError_t e1 = growbuf_grow(s.b_shadow, e->device, static_cast<size_t>(cap) * ld16 * 2,
                                        static_cast<size_t>(s.shadow_rows) * ld16 * 2, static_cast<size_t>(maxr) * ld16 * 2, e->mstream);
    const cudaError_t e2 = e1 == cudaSuccess ? growbuf_grow(s.b_rres, e->device, static_cast<size_t>(cap) * 4,
                                                            static_cast<size_t>(s.shadow_rows) * 4, static_cast<size_t>(maxr) * 4,
                                                            e->mstream)
                                             : e1;
    refresh_pointers(s);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      cudaGetLastError();
      e->shadow_failed = true;
      if (!e->shadow_warned) {
        e->shadow_warned = true;
        fprintf(stderr, "[wdbx_b200] device %d: no memory for the bf16 shadow of the stored rows; every search is served by "
                        "the fp32 scan (about half the queries/s on large stores)\n", e->device);
      }
      return;
    }
    s.shadow_cap = std::min<int64_t>(static_cast<int64_t>(s.b_shadow.bytes / (static_cast<size_t>(ld16) * 2)),
                                     static_cast<int64_t>(s.b_rres.bytes / 4));
  }
  if (e->filter_i8 && !e->shadow8_failed) {
    // the 1-byte shadow of the small-batch kernel; failing to get it only means that kernel streams the bf16 shadow
    const int ld8 = filter_ld8(e->dim);
    const size_t c = static_cast<size_t>(cap), n8 = static_cast<size_t>(s.shadow8_rows), m = static_cast<size_t>(maxr);
    cudaError_t e8 = growbuf_grow(s.b_shadow8, e->device, c * ld8, n8 * ld8, m * ld8, e->mstream);
    if (e8 == cudaSuccess) e8 = growbuf_grow(s.b_sc8, e->device, c * 4, n8 * 4, m * 4, e->mstream);
    if (e8 == cudaSuccess) e8 = growbuf_grow(s.b_rres8, e->device, c * 4, n8 * 4, m * 4, e->mstream);
    refresh_pointers(s);
    if (e8 != cudaSuccess) {
      cudaGetLastError();
      e->shadow8_failed = true;
    } else {
      s.shadow8_cap = std::min<int64_t>(static_cast<int64_t>(s.b_shadow8.bytes / static_cast<size_t>(ld8)),
                                        std::min<int64_t>(static_cast<int64_t>(s.b_sc8.bytes / 4),
                                                          static_cast<int64_t>(s.b_rres8.bytes / 4)));
    }
  }
}

// build the int8 shadow of rows [r0, r0 + m) of a segment on `stream` (rows must already be stored)
int build_shadow8(wdbx_b200_engine* e, Segment& s, int64_t r0, int64_t m, cudaStream_t stream) {
  const int ld8 = filter_ld8(e->dim);
  CU_TRY(launch_shadow8_rows(s.rows + static_cast<size_t>(r0) * row_bytes(e), e->dtype == WDBX_B200_BF16, m, e->dpad, ld8, static_cast<unsigned char*>(s.shadow8) + static_cast<size_t>(r0) * ld8, s.sc8 + r0,
                             s.rres8 + r0, stream));
  e->launches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

// Grow the segment IN PLACE to hold at least `rows` rows (see GrowBuf).  Caller holds e->mu, device is set.
int ensure_capacity(wdbx_b200_engine* e, Segment& s, int64_t rows) {
  if (rows <= s.cap_rows) return WDBX_B200_OK;
  int64_t cap = std::max<int64_t>(round_cap(rows), 1024);
  // mapping more memory is cheap but not free: grow by at least 1/8 of the current size
  if (s.cap_rows > 0) cap = std::max<int64_t>(cap, round_cap(s.cap_rows + s.cap_rows / 8));
  const int64_t maxr = segment_max_rows(e);
  cap = std::max<int64_t>(round_cap(rows), std::min<int64_t>(cap, maxr));   // over-allocation never exceeds the device
  const size_t rb = row_bytes(e);
  const size_t n = static_cast<size_t>(s.n_rows), c = static_cast<size_t>(cap), m = static_cast<size_t>(std::max(maxr, cap));
  struct Part { GrowBuf* b; size_t per; } parts[4] = {{&s.b_rows, rb}, {&s.b_inv, 4}, {&s.b_sq, 4}, {&s.b_gids, 4}};
  for (const Part& p : parts) {
    const cudaError_t err = growbuf_grow(*p.b, e->device, c * p.per, n * p.per, m * p.per, e->mstream);
    refresh_pointers(s);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(err == cudaErrorMemoryAllocation ? WDBX_B200_ERR_OOM : WDBX_B200_ERR_CUDA,
                  "cannot grow segment to %lld rows of %zu bytes: %s", (long long)cap, rb, cudaGetErrorString(err));
    }
  }
  if (s.tomb) {
    const size_t old_words = s.tomb_host.size(), words = c / 32;
    const cudaError_t err = growbuf_grow(s.b_tomb, e->device, words * 4, old_words * 4, (m / 32 + 1) * 4, e->mstream);
    refresh_pointers(s);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(WDBX_B200_ERR_OOM, "cannot grow the tombstone bitmap: %s", cudaGetErrorString(err));
    }
    s.tomb_host.resize(words, 0u);
    if (words > old_words)
      CU_TRY(cudaMemsetAsync(s.tomb + old_words, 0, (words - old_words) * 4, e->mstream));
  }
  s.cap_rows = cap;
  if (shadow_wanted(e)) grow_shadow(e, s, cap);
  return WDBX_B200_OK;
}

int get_workspace(wdbx_b200_engine* e, cudaStream_t stream, size_t cand_keys, int n_counters, Workspace** out,
                  size_t qsplit_floats = 0) {
  Workspace* w = nullptr;
  for (auto& x : e->ws)
    if (x.stream == stream) { w = &x; break; }
  if (!w) {
    if (e->ws.size() >= 64) return fail(WDBX_B200_ERR_LIMIT, "too many distinct streams (64) used with one engine");
    e->ws.emplace_back();
    w = &e->ws.back();
    w->stream = stream;
  }
  w->last_fused = false;   // (filter_segments re-arms it when a fused small-batch search has been launched completely)
  if (w->cand_keys < cand_keys || w->n_counters < n_counters || w->qsplit_floats < qsplit_floats) {
    // Growing while `stream` is being captured into a CUDA graph is fine: nothing is enqueued here, the old buffers
    // (earlier captured nodes may reference them) are retired instead of freed, and the stream is not synchronised.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    const bool capturing = cs != cudaStreamCaptureStatusNone;
    if (!capturing) CU_TRY(cudaStreamSynchronize(stream));
    auto cudaFree = [&](void* ptr) {   // shadows ::cudaFree inside this block
      if (ptr == nullptr) return;
      if (capturing) e->retired.push_back(ptr);
      else ::cudaFree(ptr);
    };
    if (w->cand_keys < cand_keys) {
      cudaFree(w->cand);
      w->cand = nullptr;
      w->cand_keys = 0;
      CU_TRY(cudaMalloc(&w->cand, cand_keys * 8));
      w->cand_keys = cand_keys;
    }
    if (w->qsplit_floats < qsplit_floats) {
      cudaFree(w->qsplit);
      w->qsplit = nullptr;
      w->qsplit_floats = 0;
      CU_TRY(cudaMalloc(&w->qsplit, qsplit_floats * 4));
      w->qsplit_floats = qsplit_floats;
    }
    if (w->n_counters < n_counters) {
      cudaFree(w->counters);
      w->counters = nullptr;
      w->n_counters = 0;
      const int n = std::max(n_counters, 64);
      CU_TRY(cudaMalloc(&w->counters, static_cast<size_t>(n) * 4));
      CU_TRY(cudaMemsetAsync(w->counters, 0, static_cast<size_t>(n) * 4, stream));
      w->n_counters = n;
    }
  }
  *out = w;
  return WDBX_B200_OK;
}

// Launch one K1 scan over segments [s0, s1).  Caller holds e->mu and has set the device.
int scan_segments(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric,
                  uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream,
                  unsigned int xseq = 0u, const int* only_flag = nullptr, float min_score = -INFINITY,
                  bool use_allow = false, uint64_t* all_keys = nullptr, unsigned int* done_ctr = nullptr,
                  unsigned int* done_blocks = nullptr, unsigned int done_sn = 0u) {
  // xseq != 0: collective search, the last CTA exchanges its lists with the peer ranks under this sequence number
  const bool exchange = xseq != 0u;
  // all_keys != NULL: large-k mode, the kernel writes every row's ranking key instead of keeping lists (plan as for k = 1)
  if (all_keys != nullptr) k = 1;
  ScanPlan plan;
  ScanTuning tune = e->tune;
  if (only_flag != nullptr && done_ctr != nullptr) {
    // flag-gated re-run behind the fused filter kernel: normally every CTA exits at once.  Launched with the smallest
    // footprint the kernel runs with (ONE warp, one stage, one row per lane group; one query per block unless the
    // exchange needs them together) so that its CTAs fit NEXT TO a resident filter CTA -- 384 x 152 registers and 212 KB
    // of shared memory leave room for 3k registers and ~14 KB -- and still leave room for the next search's prep CTAs
    // (32 threads x 32 registers): only then can the launches of the next search begin while this search finishes its
    // tail.  A real re-run is slow in this shape; it only happens on data that defeats the filter.
    tune.warps = 1;
    tune.stages = 1;
    tune.rows_unroll = 1;
    if (!exchange) tune.queries_per_pass = 1;
  }
  const int rc = scan_plan(e->dim, e->dpad, e->elem_bytes, k, B, e->sm_count, tune, &plan);
  if (rc == -4) return fail(WDBX_B200_ERR_LIMIT, "dimension %d too large for the scan kernel's shared-memory stage", e->dim);
  if (rc != 0) return fail(WDBX_B200_ERR_ARG, "invalid scan shape (dim=%d k=%d)", e->dim, k);
  ScanParams p;
  memset(&p, 0, sizeof(p));
  long long tiles = 0, bytes = 0;
  int n = 0;
  for (int s = s0; s < s1; ++s) {
    const Segment& sg = e->seg[s];
    p.seg[n].rows = sg.rows;
    p.seg[n].inv_norm = sg.inv_norm;
    p.seg[n].sqnorm = sg.sqnorm;
    p.seg[n].gids = sg.gids;
    p.seg[n].tomb = sg.n_dead > 0 ? sg.tomb : nullptr;
    p.seg[n].allow = use_allow ? sg.allow : nullptr;
    p.seg[n].n_rows = sg.n_rows;
    p.seg_row_base[n] = n == 0 ? 0 : p.seg_row_base[n - 1] + e->seg[s - 1].n_rows;
    tiles += (sg.n_rows + plan.tile_rows - 1) / plan.tile_rows;
    bytes += sg.n_rows * static_cast<long long>(row_bytes(e));
    p.tile_end[n] = tiles;
    ++n;
  }
  p.n_seg = n;
  p.total_tiles = tiles;
  // small stores: fewer CTAs than SMs would idle warps; keep one CTA per SM but never more CTAs than tiles
  long long want = (tiles + plan.warps - 1) / plan.warps;
  if (want < 1) want = 1;
  if (want < plan.grid) plan.grid = static_cast<int>(want);
  p.q = q_dev;
  p.B = B;
  p.dim = e->dim;
  p.dpad = e->dpad;
  p.row_bytes = static_cast<int>(row_bytes(e));
  p.cpr = p.row_bytes / 16;
  p.lpr_log2 = plan.lpr_log2;
  p.nch = plan.nch;
  p.tile_rows = plan.tile_rows;
  p.stages = plan.stages;
  p.stage_bytes = plan.stage_bytes;
  p.k = k;
  p.metric = metric;
  // streamed-once data should not evict the rest of L2; small stores want to stay resident
  p.evict_first = e->tune.evict_first >= 0 ? e->tune.evict_first : (bytes > (96ll << 20) ? 1 : 0);
  Workspace* w = nullptr;
  const int wrc = get_workspace(e, stream, static_cast<size_t>(B) * plan.grid * k, B, &w);
  if (wrc != WDBX_B200_OK) return wrc;
  p.cand = w->cand;
  p.counters = w->counters;
  p.keys_out = keys_out;
  p.scores_out = scores_out;
  p.gids_out = gids_out;
  p.counts_out = counts_out;
  p.only_flag = only_flag;
  p.min_score = min_score;
  p.done_ctr = done_ctr;
  p.done_blocks = done_blocks;
  p.done_sn = done_sn;
  p.all_keys = all_keys;
  p.all_rows = n > 0 ? p.seg_row_base[n - 1] + e->seg[s1 - 1].n_rows : 0;
  if (exchange) {
    if (e->xworld < 2) return fail(WDBX_B200_ERR_ARG, "exchange not attached (call wdbx_b200_exchange_init/attach first)");
    if (B > plan.queries_per_block || B > kXchgMaxB || k > kXchgMaxK)
      return fail(WDBX_B200_ERR_LIMIT, "fused exchange supports B <= %d queries in one pass and k <= %d", kXchgMaxB, kXchgMaxK);
    for (int r = 0; r < e->xworld; ++r) p.xchg_peer[r] = e->xpeer[r];
    p.xchg_world = e->xworld;
    p.xchg_rank = e->xrank;
    p.xchg_seq = xseq;
    p.xchg_slot = static_cast<int>(xseq & 1u);
  }
  // back-to-back launches: let this one start on SMs that have finished while the last CTA of the previous launch
  // still merges (programmatic dependent launch; the flag-gated re-run waits for its flags inside the kernel)
  plan.pdl = e->pdl ? 1 : 0;
  const bool timed = e->ktiming && only_flag == nullptr;
  if (timed) CU_TRY(cudaEventRecord(e->kev0, stream));
  CU_TRY(launch_scan_topk(p, plan, e->dtype == WDBX_B200_BF16, stream));
  if (timed) {
    CU_TRY(cudaEventRecord(e->kev1, stream));
    e->last_kernel = 1;
    e->kpending = true;
  }
  e->launches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

// K2 path: split the queries once, one tcgen05 GEMM + top-k launch per segment, one K3 merge.
// Caller holds e->mu and has set the device.
int gemm_segments(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric,
                  uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream) {
  int slices[kMaxSeg];
  int total_slices = 0;
  for (int s = s0; s < s1; ++s) {
    slices[s] = e->seg[s].n_rows > 0 ? gemm_slices_for(e->seg[s].n_rows, B, e->sm_count) : 0;
    total_slices += slices[s];
  }
  if (total_slices == 0) total_slices = 1;  // empty store: merge one all-zero list
  Workspace* w = nullptr;
  const int wrc = get_workspace(e, stream, static_cast<size_t>(total_slices) * B * k, 64, &w,
                                gemm_query_workspace_floats(B, e->dim));
  if (wrc != WDBX_B200_OK) return wrc;
  CU_TRY(cudaMemsetAsync(w->cand, 0, static_cast<size_t>(total_slices) * B * k * 8, stream));
  CU_TRY(launch_split_queries(q_dev, B, e->dim, w->qsplit, stream));
  e->launches.fetch_add(1, std::memory_order_relaxed);
  int base = 0;
  for (int s = s0; s < s1; ++s) {
    const Segment& sg = e->seg[s];
    if (slices[s] == 0) continue;
    SegDesc d;
    d.rows = sg.rows;
    d.inv_norm = sg.inv_norm;
    d.sqnorm = sg.sqnorm;
    d.gids = sg.gids;
    d.tomb = sg.n_dead > 0 ? sg.tomb : nullptr;
    d.allow = nullptr;
    d.n_rows = sg.n_rows;
    CU_TRY(launch_gemm_topk(d, e->dim, e->dpad, w->qsplit, B, k, metric, slices[s], base, w->cand, stream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    base += slices[s];
  }
  CU_TRY(launch_merge_topk(w->cand, total_slices, B, k, keys_out, scores_out, gids_out, counts_out, stream));
  e->launches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

// K2b path: (lazily built / extended) bf16 shadow -> prep -> one filter launch per segment (small batches: the
// filter kernel re-scores its own candidates and its last CTA merges / exchanges / emits; larger batches: a
// refine launch) -> flag-gated K1 re-run of the queries whose candidate region overflowed.  Returns
// WDBX_B200_ERR_OOM when the shadow cannot be allocated (the caller then serves from the stored rows).
// xseq != 0 (fused small-batch kernel only): collective search with the on-device key exchange.
// Caller holds e->mu, device is set.
int filter_segments(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric,
                    uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream,
                    unsigned int xseq = 0u, float min_score = -INFINITY, bool use_allow = false) {
  const int ld16 = filter_ld16(e->dim);
  const bool f32 = e->dtype == WDBX_B200_F32;
  const bool fused = filter_fused_tail(B);
  if (!fused && (xseq != 0u || use_allow || min_score > -INFINITY))
    return fail(WDBX_B200_ERR_ARG, "exchange / pre-filter on the filter path need the small-batch kernel");
  long long total_rows = 0;
  for (int s = s0; s < s1; ++s) total_rows += e->seg[s].n_rows;
  if (total_rows == 0)   // nothing to filter: K1 emits the empty lists (and joins the exchange)
    return scan_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, xseq, nullptr,
                         min_score, use_allow);
  bool built = false;
  for (int s = s0; s < s1; ++s) {
    Segment& sg = e->seg[s];
    if (!f32 || sg.n_rows == 0) continue;
    if (!sg.shadow || sg.shadow_cap < sg.n_rows) {
      // normally built by append (eager); here only when the shadow was switched on after the rows arrived
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cs);
      if (cs != cudaStreamCaptureStatusNone)
        return fail(WDBX_B200_ERR_ARG, "the bf16 shadow must be built during stream capture: run one warm-up search first");
      CU_TRY(cudaStreamSynchronize(stream));
      e->shadow_failed = false;
      grow_shadow(e, sg, sg.cap_rows);
      if (e->shadow_failed || sg.shadow_cap < sg.n_rows)
        return fail(WDBX_B200_ERR_OOM, "no device memory for the bf16 shadow of segment %d", s);
      CU_TRY(cudaStreamSynchronize(e->mstream));
    }
    if (sg.shadow_rows < sg.n_rows) {
      CU_TRY(launch_shadow_rows(reinterpret_cast<const float*>(sg.rows + static_cast<size_t>(sg.shadow_rows) * row_bytes(e)),
                                sg.n_rows - sg.shadow_rows, e->dpad, ld16,
                                static_cast<unsigned char*>(sg.shadow) + static_cast<size_t>(sg.shadow_rows) * ld16 * 2,
                                sg.rres + sg.shadow_rows, stream));
      e->launches.fetch_add(1, std::memory_order_relaxed);
      sg.shadow_rows = sg.n_rows;
      built = true;
    }
  }
  // small batches stream the 1-byte shadow when every segment has one (normally built by append)
  bool use_i8 = fused && e->filter_i8 != 0 && !e->shadow8_failed;
  for (int s = s0; s < s1 && use_i8; ++s) {
    Segment& sg = e->seg[s];
    if (sg.n_rows == 0) continue;
    if (!sg.shadow8 || sg.shadow8_cap < sg.n_rows) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cs);
      if (cs != cudaStreamCaptureStatusNone) { use_i8 = false; break; }
      CU_TRY(cudaStreamSynchronize(stream));
      grow_shadow(e, sg, sg.cap_rows);
      CU_TRY(cudaStreamSynchronize(e->mstream));
      if (e->shadow8_failed || !sg.shadow8 || sg.shadow8_cap < sg.n_rows) { use_i8 = false; break; }
    }
    if (sg.shadow8_rows < sg.n_rows) {
      const int brc = build_shadow8(e, sg, sg.shadow8_rows, sg.n_rows - sg.shadow8_rows, stream);
      if (brc != WDBX_B200_OK) return brc;
      sg.shadow8_rows = sg.n_rows;
      built = true;
    }
  }
  if (built) CU_TRY(cudaStreamSynchronize(stream));  // later searches on other streams must see the shadow
  bool prev_fused = false;   // was the previous launch on this stream's workspace a fused small-batch search?
  for (auto& x : e->ws)
    if (x.stream == stream) prev_fused = x.last_fused;
  Workspace* w = nullptr;
  int wrc = get_workspace(e, stream, 0, 0, &w);
  if (wrc != WDBX_B200_OK) return wrc;
  // one private candidate region per (query, row slice); B * slices is ~ (#SMs x 128) whatever B is
  int slices[kMaxSeg];
  int s_total = 0;
  const int rps = filter_regions_per_slice(B);
  for (int s = s0; s < s1; ++s) {
    slices[s] = e->seg[s].n_rows > 0 ? filter_slices_for(e->seg[s].n_rows, B, e->sm_count) : 0;
    s_total += rps * slices[s];  // candidate regions per (query, slice): one per epilogue column half / CTA
  }
  const int cap = rps == 2 ? 512 : 1024;
  const size_t n_regions = static_cast<size_t>(B) * s_total;
  const size_t need_ws = filter_query_workspace_bytes(B, e->dim);
  // refine launch (larger batches): several CTAs per query (each a share of the candidate regions), the last one merges
  int refine_ctas = (2 * e->sm_count + B - 1) / B;   // two 256-thread CTAs per SM
  if (refine_ctas > 8 * s_total) refine_ctas = 8 * s_total;
  if (refine_ctas < 1) refine_ctas = 1;
  // fused tail: the queries' final key lists [B][filter_final_cap()]; refine launch: partial lists [refine_ctas][B][k]
  const size_t need_part = fused ? static_cast<size_t>(B) * filter_final_cap()
                                 : (refine_ctas > 1 ? static_cast<size_t>(refine_ctas) * B * k : 0);
  // one zero-initialised block per search: [n_regions] candidate counts | [B] overflow flags | [B] tickets
  // (fused tail: [0] = the search's ticket) | [B][filter_max_k] shared lower-bound lists | [B] published
  // bounds  (0 = "no bound yet")
  const size_t lk = static_cast<size_t>(filter_max_k());
  const size_t off_over = (n_regions + 15) / 16 * 16;
  const size_t off_ticket = off_over + static_cast<size_t>(B);
  const size_t off_list = (off_ticket + static_cast<size_t>(B) + 15) / 16 * 16;
  const size_t off_glob = off_list + static_cast<size_t>(B) * lk;
  const size_t off_ctr = off_glob + static_cast<size_t>(B);   // [kMaxSeg] dynamic tile counters (fused small-batch kernel)
  const size_t off_fin = off_ctr + static_cast<size_t>(kMaxSeg);   // [B] final-list counts (fused small-batch kernel)
  const size_t off_doneb = off_fin + static_cast<size_t>(B);        // [1] query blocks of the closing launch that are finished
  const size_t off_pm = (off_doneb + 1 + 15) / 16 * 16;             // [B][filter_max_k] partition maxima (fused small-batch kernel)
  const size_t need_zero = off_pm + (fused ? static_cast<size_t>(B) * lk : 0);
  // OVERLAPPING CONSECUTIVE SEARCHES (fused small-batch path).  Every piece of per-search state is double-buffered
  // (parity of the search's number on this stream), the query prep orders itself behind search n-2 -- the previous user
  // of its buffers -- by NUMBER instead of waiting for the launch before it, the flag-gated K1 launch that closes a search
  // is small enough to sit next to a filter CTA, and the filter's last CTA waits for search n-1 before it exchanges /
  // emits.  So prep(n+1) and the first tiles of filter(n+1) run on the SMs that search n has already left while its last
  // CTAs still re-score, merge and exchange: back-to-back searches cost the streaming time plus a few microseconds, not
  // plus the ~25 us tail and two launch gaps.  Results and exchanges stay in launch order.
  // OPT-IN (`overlap`): the prep of search n+1 then no longer waits for the launch before it, so the caller must not
  // produce the QUERY buffer of a device-resident search with work enqueued on the same stream after the previous search
  // (queries staged up front, CUDA-graph loops, bench.py's throughput loop).  Off: same launches, fully ordered.
  const size_t par_mult = fused ? 2 : 1;
  const size_t ws_stride = (need_ws + 255) / 256 * 256;
  const size_t zero_stride = (need_zero + 63) / 64 * 64;
  if (w->fpart_n < need_part * par_mult || w->fws_bytes < ws_stride * par_mult || w->fcand_n < n_regions * cap * par_mult ||
      w->fzero_n < zero_stride * par_mult || (fused && !w->done)) {
    // Growing while `stream` is being captured into a CUDA graph is fine: nothing is enqueued here, the old buffers
    // (earlier captured nodes may reference them) are retired instead of freed, and the stream is not synchronised.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    const bool capturing = cs != cudaStreamCaptureStatusNone;
    if (!capturing) CU_TRY(cudaStreamSynchronize(stream));
    auto cudaFree = [&](void* ptr) {   // shadows ::cudaFree inside this block
      if (ptr == nullptr) return;
      if (capturing) e->retired.push_back(ptr);
      else ::cudaFree(ptr);
    };
    if (w->fws_bytes < ws_stride * par_mult) {
      cudaFree(w->fws); w->fws = nullptr; w->fws_bytes = 0;
      CU_TRY(cudaMalloc(&w->fws, ws_stride * 2));
      w->fws_bytes = ws_stride * 2;
    }
    if (w->fcand_n < n_regions * cap * par_mult) {
      cudaFree(w->fcand); w->fcand = nullptr; w->fcand_n = 0;
      CU_TRY(cudaMalloc(&w->fcand, n_regions * cap * 2 * 8));
      w->fcand_n = n_regions * cap * 2;
    }
    if (w->fpart_n < need_part * par_mult) {
      cudaFree(w->fpart); w->fpart = nullptr; w->fpart_n = 0;
      CU_TRY(cudaMalloc(&w->fpart, std::max<size_t>(need_part * 2, 1) * 8));
      w->fpart_n = need_part * 2;
    }
    if (w->fzero_n < zero_stride * par_mult) {
      cudaFree(w->fzero); w->fzero = nullptr; w->fzero_n = 0;
      CU_TRY(cudaMalloc(&w->fzero, zero_stride * 2 * 4));
      w->fzero_n = zero_stride * 2;
    }
    if (!w->done) {
      CU_TRY(cudaMalloc(&w->done, 64));
      CU_TRY(cudaMemset(w->done, 0, 64));
    }
    prev_fused = false;   // whatever ran before has been waited for (or is ordered by the capture)
  }
  // this search's number and buffers
  unsigned int sn = 0u;
  size_t par = 0;
  if (fused) {
    sn = ++w->fseq;
    if (sn == 0u) sn = w->fseq = 1u;
    par = sn & 1u;
  }
  // from here on the search has a number: if anything below fails, publish it as finished anyway
  struct DoneGuard {
    unsigned int* done;
    unsigned int sn;
    cudaStream_t stream;
    bool armed;
    ~DoneGuard() {
      if (armed && done != nullptr && sn != 0u) {
        launch_publish_done(done, sn, stream);
        cudaGetLastError();
      }
    }
  } done_guard{fused ? w->done : nullptr, sn, stream, true};
  unsigned int* zbase = w->fzero + par * zero_stride;
  unsigned char* wsbase = static_cast<unsigned char*>(w->fws) + par * ws_stride;
  unsigned long long* candbase = w->fcand + par * n_regions * cap;
  uint64_t* partbase = w->fpart + par * need_part;
  unsigned int* fcount = zbase;
  int* foverflow = reinterpret_cast<int*>(zbase + off_over);
  unsigned int* ftickets = zbase + off_ticket;
  unsigned int* lower_list = zbase + off_list;
  unsigned int* lower_glob = zbase + off_glob;
  const bool pdl = e->pdl != 0;
  // (also zeroes this search's state block).  Fused path: ordered by search number; programmatic launch only behind
  // another fused search -- behind anything else the launch waits for the stream as usual
  // (never while the stream is being captured: a replayed graph would find its counts already reached)
  cudaStreamCaptureStatus ocs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &ocs);
  const bool overlap = fused && pdl && e->overlap != 0 && !e->host_call && prev_fused && ocs == cudaStreamCaptureStatusNone;
  unsigned int prep_ctas = 0;
  CU_TRY(launch_prep_queries(q_dev, B, e->dim, wsbase, zbase, need_zero, fused, fused ? w->done : nullptr,
                             sn >= 2u ? sn - 2u : 0u, overlap, pdl, overlap ? w->done + 1 : nullptr, &prep_ctas, use_i8, stream));
  if (overlap) w->prep_target += prep_ctas;
  e->launches.fetch_add(1, std::memory_order_relaxed);
  // the operand-rounding part of the filter's error bound is derived from the data (per-row |x - bf16(x)|,
  // per-query |q - bf16(q)|); these are the accumulation terms on top (gemm_filter.cu, "error bound")
  const float acc_rel = filter_acc_rel(e->dim, e->dpad);
  const float c_l2 = filter_c_l2(e->dpad);
  ScanTuning t1 = e->tune;
  t1.queries_per_pass = 1;
  ScanPlan plan;
  if (scan_plan(e->dim, e->dpad, e->elem_bytes, k, 1, e->sm_count, t1, &plan) != 0)
    return fail(WDBX_B200_ERR_ARG, "invalid scan shape (dim=%d k=%d)", e->dim, k);
  FilterTail tail;
  memset(&tail, 0, sizeof(tail));
  if (fused) {
    tail.q = q_dev;
    tail.dpad = e->dpad;
    tail.elem_bytes = e->elem_bytes;
    tail.lpr_log2 = plan.lpr_log2;
    tail.nch = plan.nch;
    tail.min_score = min_score;
    tail.overflow = foverflow;
    tail.fin_keys = partbase;
    tail.fin_count = zbase + off_fin;
    tail.ticket = ftickets;
    tail.tile_ctr = zbase + off_ctr;
    tail.part_max = zbase + off_pm;
    tail.done_ctr = w->done;
    tail.done_sn = sn;
    tail.prep_count = overlap ? w->done + 1 : nullptr;
    tail.prep_target = w->prep_target;
    if (xseq != 0u) {
      for (int r = 0; r < e->xworld; ++r) tail.xchg.peer[r] = e->xpeer[r];
      tail.xchg.world = e->xworld;
      tail.xchg.rank = e->xrank;
      tail.xchg.seq = xseq;
      tail.xchg.slot = static_cast<int>(xseq & 1u);
    }
    tail.keys_out = keys_out;
    tail.scores_out = scores_out;
    tail.gids_out = gids_out;
    tail.counts_out = counts_out;
  }
  SegDesc descs[kMaxSeg];
  memset(descs, 0, sizeof(descs));
  int slice_base = 0;
  if (e->ktiming) CU_TRY(cudaEventRecord(e->kev0, stream));
  for (int s = s0; s < s1; ++s) {
    const Segment& sg = e->seg[s];
    SegDesc& d = descs[s];
    d.rows = sg.rows;
    d.inv_norm = sg.inv_norm;
    d.sqnorm = sg.sqnorm;
    d.gids = sg.gids;
    d.tomb = sg.n_dead > 0 ? sg.tomb : nullptr;
    d.allow = use_allow ? sg.allow : nullptr;
    d.n_rows = sg.n_rows;
    if (sg.n_rows == 0) continue;
    tail.rowscale = use_i8 ? sg.sc8 : nullptr;
    CU_TRY(launch_gemm_filter(use_i8 ? sg.shadow8 : (f32 ? sg.shadow : static_cast<const void*>(sg.rows)),
                              use_i8 ? filter_ld8(e->dim) : (f32 ? ld16 : e->dpad),
                              use_i8 ? sg.rres8 : (f32 ? sg.rres : nullptr), d, s, e->dim, wsbase, B, k, metric, acc_rel, c_l2, slices[s], candbase,
                              fcount, lower_glob, lower_list, cap, slice_base, s_total, fused ? &tail : nullptr, pdl, stream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    slice_base += slices[s];
  }
  if (e->ktiming) CU_TRY(cudaEventRecord(e->kev1, stream));
  if (!fused) {
    CU_TRY(launch_refine_topk(descs, kMaxSeg, q_dev, B, e->dim, e->dpad, e->elem_bytes, plan.lpr_log2, plan.nch, k, metric,
                              candbase, fcount, cap, s_total, foverflow, refine_ctas, partbase, ftickets, keys_out, scores_out,
                              gids_out, counts_out, stream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
  }
  // exact re-run (K1) of the queries whose candidate list overflowed; exits immediately otherwise.  Collective
  // searches: a flagged query makes the filter's last CTA skip the exchange and K1 redo all B queries with it.
  const int rrc = scan_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream,
                                xseq, foverflow, min_score, use_allow, nullptr, fused ? w->done : nullptr,
                                fused ? zbase + off_doneb : nullptr, sn);
  w->last_fused = fused && rrc == WDBX_B200_OK;
  done_guard.armed = rrc != WDBX_B200_OK;   // the closing launch publishes the number itself
  if (e->ktiming) {
    e->last_kernel = use_i8 ? 3 : 2;
    e->kpending = true;
    e->last_fcount = fcount;
    e->last_fregions = n_regions;
  }
  return rrc;
}

// Large k (128 < k <= 1024): the scan writes every row's ranking key, a radix select picks the k best
// (select_topk.cu).  Queries go through in chunks of up to 8 (one streamed pass of the rows each).
// Caller holds e->mu, device is set.
constexpr int kSelectMinK = 128;

int select_segments(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric, uint64_t* keys_out,
                    float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream, float min_score = -INFINITY,
                    bool use_allow = false) {
  long long rows = 0;
  for (int s = s0; s < s1; ++s) rows += e->seg[s].n_rows;
  if (rows == 0)
    return scan_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, 0u, nullptr,
                         min_score, use_allow);
  ScanPlan plan;
  if (scan_plan(e->dim, e->dpad, e->elem_bytes, 1, B, e->sm_count, e->tune, &plan) != 0)
    return fail(WDBX_B200_ERR_ARG, "invalid scan shape (dim=%d)", e->dim);
  // keys of one chunk of queries: at most 2 GiB of workspace
  int chunk = std::min(B, plan.queries_per_block);
  while (chunk > 1 && static_cast<size_t>(chunk) * rows * 8 > (2ull << 30)) chunk >>= 1;
  Workspace* w = nullptr;
  int rc = get_workspace(e, stream, 0, 0, &w);
  if (rc != WDBX_B200_OK) return rc;
  const size_t need_all = static_cast<size_t>(chunk) * rows;
  if (w->sel_all_n < need_all || w->sel_chunk < chunk) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    const bool capturing = cs != cudaStreamCaptureStatusNone;
    if (!capturing) CU_TRY(cudaStreamSynchronize(stream));
    auto drop = [&](void* ptr) {
      if (ptr == nullptr) return;
      if (capturing) e->retired.push_back(ptr);
      else cudaFree(ptr);
    };
    if (w->sel_all_n < need_all) {
      drop(w->sel_all); w->sel_all = nullptr; w->sel_all_n = 0;
      CU_TRY(cudaMalloc(&w->sel_all, need_all * 8));
      w->sel_all_n = need_all;
    }
    if (w->sel_chunk < chunk) {
      drop(w->sel_keys); drop(w->sel_ws);
      w->sel_keys = nullptr; w->sel_ws = nullptr; w->sel_chunk = 0;
      CU_TRY(cudaMalloc(&w->sel_keys, static_cast<size_t>(chunk) * 1024 * 8));
      CU_TRY(cudaMalloc(&w->sel_ws, select_workspace_words(chunk) * 4));
      w->sel_chunk = chunk;
    }
  }
  ScanTuning saved = e->tune;
  if (chunk < plan.queries_per_block) e->tune.queries_per_pass = chunk;   // one block of `chunk` queries per launch
  for (int b0 = 0; b0 < B && rc == WDBX_B200_OK; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    const size_t o = static_cast<size_t>(b0) * k;
    rc = scan_segments(e, s0, s1, q_dev + static_cast<size_t>(b0) * e->dim, nb, k, metric, nullptr, nullptr, nullptr, nullptr,
                       stream, 0u, nullptr, min_score, use_allow, w->sel_all);
    if (rc != WDBX_B200_OK) break;
    const cudaError_t ce = launch_select_topk(w->sel_all, rows, nb, k, w->sel_ws, w->sel_keys, e->sm_count,
                                              keys_out ? keys_out + o : nullptr, scores_out ? scores_out + o : nullptr,
                                              gids_out ? gids_out + o : nullptr, counts_out ? counts_out + b0 : nullptr, stream);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      rc = fail(WDBX_B200_ERR_CUDA, "select launch: %s", cudaGetErrorString(ce));
    }
    e->launches.fetch_add(11, std::memory_order_relaxed);
  }
  e->tune = saved;
  return rc;
}

// the exact streaming path for any k: running lists (k <= 128) or key dump + radix select (larger k)
int scan_any_k(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric, uint64_t* keys_out,
               float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream, float min_score = -INFINITY,
               bool use_allow = false) {
  if (k > kSelectMinK)
    return select_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, min_score, use_allow);
  return scan_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, 0u, nullptr, min_score,
                       use_allow);
}

// Regime choice.  Tensor path (K2b filter + exact refine) for batches of gemm_min_batch queries or more;
// for SMALLER batches on fp32 storage too once the segments are large (shadow_min_bytes): the filter
// streams the 2-byte shadow rows instead of the 4-byte stored rows, i.e. half the HBM bytes per query, and
// the refine restores the exact fp32 result -- below that size the extra launches cost more than they save.
bool use_gemm(const wdbx_b200_engine* e, int s0, int s1, int B, int k) {
  if (e->gemm_min_batch <= 0) return false;
  if (e->gemm_mode == 1) return B >= e->gemm_min_batch && e->dtype == WDBX_B200_F32 && k <= gemm_max_k();
  if (k > filter_max_k()) return false;
  // 32 < k <= 128: only the small-batch kernel, whose bound server inserts into the wide shared list cooperatively
  if (k > 32 && !filter_fused_tail(B)) return false;
  if (e->dtype == WDBX_B200_F32 && e->shadow_failed) return false;   // no room for the bf16 shadow (said so once)
  if (static_cast<size_t>(e->dpad) * 4 > 160 * 1024) return false;   // the fused tail stages fp32 queries in shared memory
  if (B >= e->gemm_min_batch) return true;
  if (e->shadow_min_bytes < 0) return false;
  // small batches: worth it only where the filter streams fewer bytes than the scan would -- fp32 stores (bf16 or int8
  // shadow) and bf16 stores that carry the int8 shadow
  if (e->dtype != WDBX_B200_F32 && (e->filter_i8 == 0 || e->shadow8_failed)) return false;
  long long bytes = 0;
  for (int s = s0; s < s1; ++s) bytes += e->seg[s].n_rows * static_cast<long long>(row_bytes(e));
  return bytes >= e->shadow_min_bytes;
}

int search_segments(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric,
                    uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream) {
  if (use_gemm(e, s0, s1, B, k)) {
    if (e->gemm_mode == 1)
      return gemm_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream);
    const int rc = filter_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream);
    if (rc != WDBX_B200_ERR_OOM) return rc;
    // no room for the bf16 shadow: keep serving from the stored rows (K1) and stop trying for small batches
    cudaGetLastError();
    e->shadow_min_bytes = -1;
    if (!e->shadow_warned) {
      e->shadow_warned = true;
      fprintf(stderr, "[wdbx_b200] device %d: no memory for the bf16 shadow; small batches are served by the fp32 scan "
                      "(about half the queries/s)\n", e->device);
    }
  }
  return scan_any_k(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream);
}

int check_search_args(wdbx_b200_engine* e, int B, int k, int metric) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  if (B <= 0) return fail(WDBX_B200_ERR_ARG, "B must be positive (got %d)", B);
  if (k <= 0) return fail(WDBX_B200_ERR_ARG, "k must be positive (got %d)", k);
  if (k > WDBX_B200_MAX_K) return fail(WDBX_B200_ERR_LIMIT, "k=%d exceeds WDBX_B200_MAX_K=%d", k, WDBX_B200_MAX_K);
  if (B > 65535) return fail(WDBX_B200_ERR_LIMIT, "B=%d exceeds 65535 queries per call", B);
  if (metric < 0 || metric > 2) return fail(WDBX_B200_ERR_ARG, "unknown metric %d", metric);
  return WDBX_B200_OK;
}


// Collective search of all segments + on-device key exchange with the peer ranks.  Caller holds e->mu.
// Every rank takes the same decisions here (they depend on B, k, dim only -- never on the rank's own rows), and
// the routes a rank may take for one search (fused filter kernel, K1 scan, stand-alone exchange kernel) all
// speak the same exchange protocol under the same sequence number.
int exchange_search_locked(wdbx_b200_engine* e, int s0, int s1, const float* q_dev, int B, int k, int metric,
                           uint64_t* keys_out, float* scores_out, long long* gids_out, int* counts_out, cudaStream_t stream,
                           float min_score = -INFINITY, bool use_allow = false) {
  if (e->xworld < 2) return fail(WDBX_B200_ERR_ARG, "exchange not attached (call wdbx_b200_exchange_init/attach first)");
  if (B > kXchgMaxB || k > kXchgMaxK)
    return fail(WDBX_B200_ERR_LIMIT, "fused exchange supports B <= %d queries and k <= %d", kXchgMaxB, kXchgMaxK);
  // the K1 scan (and the flag-gated re-run behind the filter) exchanges from ONE query block: batches larger
  // than the block the plan allows for this shape (e.g. 1 query for dim 64) run as consecutive collective passes
  ScanPlan plan;
  const int prc = scan_plan(e->dim, e->dpad, e->elem_bytes, k, B, e->sm_count, e->tune, &plan);
  if (prc != 0) return fail(prc == -4 ? WDBX_B200_ERR_LIMIT : WDBX_B200_ERR_ARG, "invalid scan shape (dim=%d k=%d)", e->dim, k);
  if (B > plan.queries_per_block) {
    const int qb = plan.queries_per_block;
    for (int b0 = 0; b0 < B; b0 += qb) {
      const int nb = std::min(qb, B - b0);
      const size_t o = static_cast<size_t>(b0) * k;
      const int rc = exchange_search_locked(e, s0, s1, q_dev + static_cast<size_t>(b0) * e->dim, nb, k, metric,
                                            keys_out ? keys_out + o : nullptr, scores_out ? scores_out + o : nullptr,
                                            gids_out ? gids_out + o : nullptr, counts_out ? counts_out + b0 : nullptr, stream,
                                            min_score, use_allow);
      if (rc != WDBX_B200_OK) return rc;
    }
    return WDBX_B200_OK;
  }
  e->xseq += 1;
  if (e->xseq == 0u) e->xseq = 1u;   // 0 means "no exchange"
  const unsigned int seq = e->xseq;
  int rc;
  // (the regime choice must not depend on this rank's own row count: use_gemm looks at the segments' bytes, which
  // row striping keeps within one row of every other rank's -- and every route speaks the same protocol anyway)
  const bool filtered = use_allow || min_score > -INFINITY;
  if (e->gemm_mode != 1 && use_gemm(e, s0, s1, B, k) && (filter_fused_tail(B) || !filtered)) {
    if (filter_fused_tail(B)) {
      // bf16-filter path, ONE launch: filter + in-kernel refine; its last CTA pushes / awaits / merges the keys
      rc = filter_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, seq,
                           min_score, use_allow);
    } else {
      // 128-query filter kernel (WDBX_B200_FILTER_SMALL=0): local exact top-k, then the stand-alone exchange kernel
      if (!e->xkeys) CU_TRY(cudaMalloc(&e->xkeys, static_cast<size_t>(kXchgMaxB) * kXchgMaxK * 8));
      rc = filter_segments(e, s0, s1, q_dev, B, k, metric, e->xkeys, nullptr, nullptr, nullptr, stream);
      if (rc == WDBX_B200_OK) {
        CU_TRY(launch_exchange_merge(e->xpeer, e->xworld, e->xrank, seq, e->xkeys, B, k, keys_out, scores_out, gids_out,
                                     counts_out, stream));
        e->launches.fetch_add(1, std::memory_order_relaxed);
      }
    }
    if (rc != WDBX_B200_ERR_OOM) return rc;
    cudaGetLastError();
    e->shadow_min_bytes = -1;
    if (!e->shadow_warned) {
      e->shadow_warned = true;
      fprintf(stderr, "[wdbx_b200] device %d: no memory for the bf16 shadow; small batches are served by the fp32 scan "
                      "(about half the queries/s)\n", e->device);
    }
  }
  return scan_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, gids_out, counts_out, stream, seq, nullptr,
                       min_score, use_allow);
}


// ---- host-buffer searches ------------------------------------------------------------------------------
// result block layout (device and pinned host): keys[nres] u64 | gids[nres] i64 | scores[nres] f32 | counts[lists*B] i32
struct ResLayout {
  size_t nres, off_keys, off_gids, off_scores, off_counts, bytes;
  ResLayout(int lists, int B, int k) {
    nres = static_cast<size_t>(lists) * B * k;
    off_keys = 0;
    off_gids = nres * 8;
    off_scores = nres * 16;
    off_counts = nres * 20;
    bytes = off_counts + static_cast<size_t>(lists) * B * 4;
  }
};

// device-side staging of the host path (queries, packed results); pinned mirrors when `pinned`
int ensure_host_buffers(wdbx_b200_engine* e, size_t nq, size_t bytes, bool pinned) {
  if (e->q_floats < nq || (pinned && !e->hq_pinned)) {
    cudaFreeHost(e->hq_pinned); e->hq_pinned = nullptr;
    cudaFree(e->dq); e->dq = nullptr;
    e->q_floats = 0;
    if (pinned) CU_TRY(cudaMallocHost(&e->hq_pinned, nq * 4));
    CU_TRY(cudaMalloc(&e->dq, nq * 4));
    e->q_floats = nq;
  }
  if (e->res_bytes < bytes || (pinned && !e->hres_pinned)) {
    cudaFreeHost(e->hres_pinned); e->hres_pinned = nullptr;
    cudaFree(e->dres); e->dres = nullptr;
    e->res_bytes = 0;
    if (pinned) CU_TRY(cudaMallocHost(&e->hres_pinned, bytes));
    CU_TRY(cudaMalloc(&e->dres, bytes));
    e->res_bytes = bytes;
  }
  return WDBX_B200_OK;
}

// copy the per-segment "allowed rows" bitmaps of one filtered search to the device.  Caller holds e->mu.
int upload_allow(wdbx_b200_engine* e, const uint32_t* const* allow_bitmaps, cudaStream_t st) {
  for (int s = 0; s < e->nseg; ++s) {
    Segment& sg = e->seg[s];
    const size_t words = static_cast<size_t>((sg.n_rows + 31) / 32);
    if (words == 0) continue;
    if (sg.allow_words < words) {
      CU_TRY(cudaStreamSynchronize(st));
      cudaFree(sg.allow);
      sg.allow = nullptr;
      sg.allow_words = 0;
      const size_t cap_words = static_cast<size_t>(sg.cap_rows / 32 + 1);
      CU_TRY(cudaMalloc(&sg.allow, cap_words * 4));
      sg.allow_words = cap_words;
    }
    if (allow_bitmaps[s]) CU_TRY(cudaMemcpyAsync(sg.allow, allow_bitmaps[s], words * 4, cudaMemcpyHostToDevice, st));
    else CU_TRY(cudaMemsetAsync(sg.allow, 0xFF, words * 4, st));  // NULL entry = every row allowed
  }
  return WDBX_B200_OK;
}

// Launch the searches of one host call on `st` (queries already in e->dq): `lists` result lists into e->dres.
// exchange: collective search with the peer engines / ranks.  Takes e->mu.
int launch_host_lists(wdbx_b200_engine* e, int segment, bool exchange, const float* q_dev, int B, int k, int metric,
                      float min_score, const uint32_t* const* allow_bitmaps, const ResLayout& R, cudaStream_t st,
                      uint64_t* keys0 = nullptr, float* scores0 = nullptr, long long* gids0 = nullptr, int* counts0 = nullptr) {
  std::lock_guard<std::mutex> lk(e->mu);
  struct HostCall {   // host-buffer searches are synchronous: never overlapped with their predecessor
    wdbx_b200_engine* e;
    explicit HostCall(wdbx_b200_engine* e_) : e(e_) { e->host_call = true; }
    ~HostCall() { e->host_call = false; }
  } host_call(e);
  const bool per_segment = segment == WDBX_B200_EACH_SEGMENT;
  const int lists = per_segment ? e->nseg : 1;
  const bool use_allow = allow_bitmaps != nullptr;
  const bool filtered = use_allow || min_score > -INFINITY;
  if (use_allow) {
    const int rc = upload_allow(e, allow_bitmaps, st);
    if (rc != WDBX_B200_OK) return rc;
  }
  // outputs: the engine's packed result block, unless the caller wants them somewhere else on this device
  uint64_t* keys = keys0 ? keys0 : reinterpret_cast<uint64_t*>(e->dres + R.off_keys);
  float* scores = scores0 ? scores0 : reinterpret_cast<float*>(e->dres + R.off_scores);
  long long* gids = gids0 ? gids0 : reinterpret_cast<long long*>(e->dres + R.off_gids);
  int* counts = counts0 ? counts0 : reinterpret_cast<int*>(e->dres + R.off_counts);
  for (int l = 0; l < lists; ++l) {
    const size_t o = static_cast<size_t>(l) * B * k;
    const int s0 = per_segment ? l : (segment >= 0 ? segment : 0);
    const int s1 = per_segment ? l + 1 : (segment >= 0 ? segment + 1 : e->nseg);
    int* cnt = counts + static_cast<size_t>(l) * B;
    int rc;
    if (exchange) {
      rc = exchange_search_locked(e, s0, s1, q_dev, B, k, metric, keys + o, scores + o, gids + o, cnt, st, min_score, use_allow);
    } else if (filtered) {
      // small batches on the filter path consult the bitmap / floor in the fused filter kernel (candidates only);
      // everything else takes the streaming kernel
      rc = WDBX_B200_ERR_OOM;
      if (e->gemm_mode != 1 && use_gemm(e, s0, s1, B, k) && filter_fused_tail(B)) {
        rc = filter_segments(e, s0, s1, q_dev, B, k, metric, keys + o, scores + o, gids + o, cnt, st, 0u, min_score, use_allow);
        if (rc == WDBX_B200_ERR_OOM) {
          cudaGetLastError();
          e->shadow_min_bytes = -1;
        }
      }
      if (rc == WDBX_B200_ERR_OOM)
        rc = scan_any_k(e, s0, s1, q_dev, B, k, metric, keys + o, scores + o, gids + o, cnt, st, min_score, use_allow);
    } else {
      rc = search_segments(e, s0, s1, q_dev, B, k, metric, keys + o, scores + o, gids + o, cnt, st);
    }
    if (rc != WDBX_B200_OK) return rc;
  }
  return WDBX_B200_OK;
}

// One engine, host buffers in and out: pinned H2D -> launches -> ONE D2H of the packed result -> synchronise.
int host_search(wdbx_b200_engine* e, int segment, bool exchange, const float* q_host, int B, int k, int metric,
                float min_score, const uint32_t* const* allow_bitmaps, float* scores_host, int64_t* gids_host,
                uint64_t* keys_host, int32_t* counts_host) {
  int rc = check_search_args(e, B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (!q_host) return fail(WDBX_B200_ERR_ARG, "q_host is NULL");
  if (segment < WDBX_B200_EACH_SEGMENT || segment >= e->nseg)
    return fail(WDBX_B200_ERR_ARG, "segment %d outside [-2, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> hlk(e->host_mu);
  const int lists = segment == WDBX_B200_EACH_SEGMENT ? e->nseg : 1;
  const size_t nq = static_cast<size_t>(B) * e->dim;
  const ResLayout R(lists, B, k);
  rc = ensure_host_buffers(e, nq, R.bytes, true);
  if (rc != WDBX_B200_OK) return rc;
  memcpy(e->hq_pinned, q_host, nq * 4);
  cudaStream_t st = e->hstream;
  CU_TRY(cudaMemcpyAsync(e->dq, e->hq_pinned, nq * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaEventRecord(e->ev0, st));
  rc = launch_host_lists(e, segment, exchange, e->dq, B, k, metric, min_score, allow_bitmaps, R, st);
  if (rc != WDBX_B200_OK) return rc;
  CU_TRY(cudaEventRecord(e->ev1, st));
  CU_TRY(cudaMemcpyAsync(e->hres_pinned, e->dres, R.bytes, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0.0f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_search_ms = ms;
  if (keys_host) memcpy(keys_host, e->hres_pinned + R.off_keys, R.nres * 8);
  if (gids_host) memcpy(gids_host, e->hres_pinned + R.off_gids, R.nres * 8);
  if (scores_host) memcpy(scores_host, e->hres_pinned + R.off_scores, R.nres * 4);
  if (counts_host) memcpy(counts_host, e->hres_pinned + R.off_counts, static_cast<size_t>(lists) * B * 4);
  e->searches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

}  // namespace

extern "C" {

int wdbx_b200_version(void) { return WDBX_B200_ABI_VERSION; }

const char* wdbx_b200_last_error(void) { return g_err; }

int wdbx_b200_device_count(void) {
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(WDBX_B200_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(err));
  }
  return n;
}

int wdbx_b200_create(int device, int dim, int dtype, int num_segments, wdbx_b200_engine** out) {
  if (!out) return fail(WDBX_B200_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (dim <= 0) return fail(WDBX_B200_ERR_ARG, "dim must be positive (got %d)", dim);
  if (dtype != WDBX_B200_F32 && dtype != WDBX_B200_BF16) return fail(WDBX_B200_ERR_ARG, "unknown dtype %d", dtype);
  if (num_segments < 1 || num_segments > WDBX_B200_MAX_SEGMENTS)
    return fail(WDBX_B200_ERR_LIMIT, "num_segments=%d outside [1, %d]", num_segments, WDBX_B200_MAX_SEGMENTS);
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(WDBX_B200_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                err != cudaSuccess ? cudaGetErrorString(err) : "0 devices");
  }
  if (device < 0 || device >= ndev) return fail(WDBX_B200_ERR_ARG, "device %d outside [0, %d)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(WDBX_B200_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(WDBX_B200_ERR_CUDA, "device %d is sm_%d%d; libwdbx_b200 is built for sm_100a only", device, prop.major,
                prop.minor);
  wdbx_b200_engine* e = new (std::nothrow) wdbx_b200_engine();
  if (!e) return fail(WDBX_B200_ERR_OOM, "host allocation failed");
  e->device = device;
  e->dim = dim;
  e->dtype = dtype;
  e->elem_bytes = dtype == WDBX_B200_BF16 ? 2 : 4;
  const int epc = 16 / e->elem_bytes;
  e->dpad = (dim + epc - 1) / epc * epc;
  e->nseg = num_segments;
  e->sm_count = prop.multiProcessorCount;
  e->total_mem = prop.totalGlobalMem;
  e->tune.warps = env_int("WDBX_B200_WARPS", 0);
  e->tune.stages = env_int("WDBX_B200_STAGES", 0);
  e->tune.rows_unroll = env_int("WDBX_B200_UNROLL", 0);
  e->tune.grid = env_int("WDBX_B200_GRID", 0);
  e->tune.evict_first = env_int("WDBX_B200_EVICT_FIRST", -1);
  e->tune.queries_per_pass = env_int("WDBX_B200_QUERIES_PER_PASS", 0);
  e->gemm_min_batch = env_int("WDBX_B200_GEMM_MIN_BATCH", e->gemm_min_batch);
  {
    const int mb = env_int("WDBX_B200_SHADOW_MIN_MB", -2);  // -1 = never; >= 0: threshold in MiB of stored rows
    if (mb >= -1) e->shadow_min_bytes = mb < 0 ? -1 : static_cast<long long>(mb) << 20;
  }
  e->gemm_mode = env_int("WDBX_B200_GEMM_MODE", 0);
  e->pdl = env_int("WDBX_B200_PDL", 1);
  e->overlap = env_int("WDBX_B200_OVERLAP", 0);
  e->filter_i8 = env_int("WDBX_B200_FILTER_I8", 1);
  ScanPlan plan;
  if (scan_plan(dim, e->dpad, e->elem_bytes, 10, 1, e->sm_count, e->tune, &plan) != 0) {
    delete e;
    return fail(WDBX_B200_ERR_LIMIT, "dim=%d is too large for the scan kernel (row must fit a shared-memory stage)", dim);
  }
  cudaError_t e1 = cudaStreamCreateWithFlags(&e->mstream, cudaStreamNonBlocking);
  cudaError_t e2 = cudaStreamCreateWithFlags(&e->hstream, cudaStreamNonBlocking);
  cudaError_t e3 = cudaEventCreate(&e->ev0);
  cudaError_t e4 = cudaEventCreate(&e->ev1);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
    wdbx_b200_destroy(e);
    return fail(WDBX_B200_ERR_CUDA, "stream/event creation failed");
  }
  *out = e;
  return WDBX_B200_OK;
}

void wdbx_b200_destroy(wdbx_b200_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  for (int s = 0; s < kMaxSeg; ++s) free_segment(e->seg[s]);
  for (auto& w : e->ws) {
    cudaFree(w.cand);
    cudaFree(w.counters);
    cudaFree(w.qsplit);
    cudaFree(w.fws);
    cudaFree(w.fcand);
    cudaFree(w.fzero);
    cudaFree(w.fpart);
    cudaFree(w.sel_all);
    cudaFree(w.sel_keys);
    cudaFree(w.sel_ws);
    cudaFree(w.done);
  }
  for (void* ptr : e->retired) cudaFree(ptr);
  for (int r = 0; r < kMaxPeers; ++r)
    if (e->xopened[r]) cudaIpcCloseMemHandle(e->xpeer[r]);
  cudaFree(e->xbuf);
  cudaFree(e->xkeys);
  cudaFree(e->stage_rows);
  cudaFree(e->stage_gids);
  cudaFree(e->dq);
  cudaFree(e->dres);
  cudaFreeHost(e->hq_pinned);
  cudaFreeHost(e->hres_pinned);
  if (e->kev0) cudaEventDestroy(e->kev0);
  if (e->kev1) cudaEventDestroy(e->kev1);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->mstream) cudaStreamDestroy(e->mstream);
  if (e->hstream) cudaStreamDestroy(e->hstream);
  cudaGetLastError();
  delete e;
}

int wdbx_b200_set_tuning(wdbx_b200_engine* e, int warps, int stages, int rows_unroll, int grid, int evict_first) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  std::lock_guard<std::mutex> lk(e->mu);
  ScanTuning t{warps, stages, rows_unroll, grid, evict_first, e->tune.queries_per_pass};
  ScanPlan plan;
  if (scan_plan(e->dim, e->dpad, e->elem_bytes, 10, 1, e->sm_count, t, &plan) != 0)
    return fail(WDBX_B200_ERR_ARG, "tuning does not fit shared memory");
  e->tune = t;
  return WDBX_B200_OK;
}

int wdbx_b200_reserve(wdbx_b200_engine* e, int segment, int64_t rows) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  if (segment < 0 || segment >= e->nseg) return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  if (rows < 0 || rows > 0xFFFFFFF0ll) return fail(WDBX_B200_ERR_LIMIT, "rows=%lld outside [0, 2^32)", (long long)rows);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  return ensure_capacity(e, e->seg[segment], rows);
}

int wdbx_b200_append(wdbx_b200_engine* e, int segment, const float* rows, int64_t n, int src_is_device,
                     const uint32_t* gids, int64_t* first_row_out) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  if (segment < 0 || segment >= e->nseg) return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  if (n < 0) return fail(WDBX_B200_ERR_ARG, "n must be >= 0");
  if (n > 0 && !rows) return fail(WDBX_B200_ERR_ARG, "rows is NULL");
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  Segment& s = e->seg[segment];
  if (first_row_out) *first_row_out = s.n_rows;
  if (n == 0) return WDBX_B200_OK;
  if (s.n_rows + n > 0xFFFFFFF0ll) return fail(WDBX_B200_ERR_LIMIT, "segment would exceed 2^32 rows");
  int rc = ensure_capacity(e, s, s.n_rows + n);
  if (rc != WDBX_B200_OK) return rc;
  const size_t rb = row_bytes(e);
  const bool bf16 = e->dtype == WDBX_B200_BF16;
  // host sources go through a 64 MB device staging buffer; device sources are ingested in one launch
  const int64_t chunk_rows =
      src_is_device ? n : std::max<int64_t>(1, (64ll << 20) / (static_cast<int64_t>(e->dim) * 4));
  for (int64_t done = 0; done < n; done += chunk_rows) {
    const int64_t m = std::min(chunk_rows, n - done);
    const float* src = rows + done * e->dim;
    if (!src_is_device) {
      const size_t need = static_cast<size_t>(m) * e->dim * 4;
      if (e->stage_rows_bytes < need) {
        cudaFree(e->stage_rows);
        e->stage_rows = nullptr;
        e->stage_rows_bytes = 0;
        CU_TRY(cudaMalloc(&e->stage_rows, need));
        e->stage_rows_bytes = need;
      }
      CU_TRY(cudaMemcpyAsync(e->stage_rows, src, need, cudaMemcpyHostToDevice, e->mstream));
      src = e->stage_rows;
    }
    const uint32_t* gsrc = nullptr;
    if (gids) {
      if (e->stage_gids_n < static_cast<size_t>(m)) {
        cudaFree(e->stage_gids);
        e->stage_gids = nullptr;
        e->stage_gids_n = 0;
        CU_TRY(cudaMalloc(&e->stage_gids, static_cast<size_t>(m) * 4));
        e->stage_gids_n = static_cast<size_t>(m);
      }
      CU_TRY(cudaMemcpyAsync(e->stage_gids, gids + done, static_cast<size_t>(m) * 4, cudaMemcpyHostToDevice, e->mstream));
      gsrc = e->stage_gids;
    }
    const int64_t r0 = s.n_rows + done;
    CU_TRY(launch_append_rows(src, m, e->dim, e->dpad, bf16, s.rows + static_cast<size_t>(r0) * rb, s.inv_norm + r0,
                              s.sqnorm + r0, s.gids + r0, gsrc, e->next_gid + static_cast<uint32_t>(done), e->mstream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    if (s.shadow && s.shadow_rows == r0 && s.shadow_cap >= r0 + m) {
      // eager bf16 shadow (fp32 stores): built right behind the rows, so no search ever allocates or builds it
      const int ld16 = filter_ld16(e->dim);
      CU_TRY(launch_shadow_rows(reinterpret_cast<const float*>(s.rows + static_cast<size_t>(r0) * rb), m, e->dpad, ld16,
                                static_cast<unsigned char*>(s.shadow) + static_cast<size_t>(r0) * ld16 * 2, s.rres + r0,
                                e->mstream));
      e->launches.fetch_add(1, std::memory_order_relaxed);
      s.shadow_rows = r0 + m;
    }
    if (s.shadow8 && s.shadow8_rows == r0 && s.shadow8_cap >= r0 + m) {
      const int brc = build_shadow8(e, s, r0, m, e->mstream);
      if (brc != WDBX_B200_OK) return brc;
      s.shadow8_rows = r0 + m;
    }
    // the staging buffers are reused by the next chunk
    CU_TRY(cudaStreamSynchronize(e->mstream));
  }
  s.n_rows += n;
  e->next_gid += static_cast<uint32_t>(n);
  return WDBX_B200_OK;
}

int wdbx_b200_overwrite(wdbx_b200_engine* e, int segment, int64_t row, const float* v_host) {
  if (!e || !v_host) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= e->nseg) return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  Segment& s = e->seg[segment];
  if (row < 0 || row >= s.n_rows) return fail(WDBX_B200_ERR_ARG, "row %lld outside [0, %lld)", (long long)row, (long long)s.n_rows);
  const size_t need = static_cast<size_t>(e->dim) * 4 + 4;
  if (e->stage_rows_bytes < need) {
    cudaFree(e->stage_rows);
    e->stage_rows = nullptr;
    e->stage_rows_bytes = 0;
    CU_TRY(cudaMalloc(&e->stage_rows, need));
    e->stage_rows_bytes = need;
  }
  if (e->stage_gids_n < 1) {
    CU_TRY(cudaMalloc(&e->stage_gids, 4 * 32));
    e->stage_gids_n = 32;
  }
  CU_TRY(cudaMemcpyAsync(e->stage_rows, v_host, static_cast<size_t>(e->dim) * 4, cudaMemcpyHostToDevice, e->mstream));
  // keep the row's gid
  CU_TRY(cudaMemcpyAsync(e->stage_gids, s.gids + row, 4, cudaMemcpyDeviceToDevice, e->mstream));
  CU_TRY(launch_append_rows(e->stage_rows, 1, e->dim, e->dpad, e->dtype == WDBX_B200_BF16,
                            s.rows + static_cast<size_t>(row) * row_bytes(e), s.inv_norm + row, s.sqnorm + row,
                            s.gids + row, e->stage_gids, 0, e->mstream));
  e->launches.fetch_add(1, std::memory_order_relaxed);
  if (s.shadow && row < s.shadow_rows) {
    const int ld16 = filter_ld16(e->dim);
    CU_TRY(launch_shadow_rows(reinterpret_cast<const float*>(s.rows + static_cast<size_t>(row) * row_bytes(e)), 1, e->dpad,
                              ld16, static_cast<unsigned char*>(s.shadow) + static_cast<size_t>(row) * ld16 * 2, s.rres + row,
                              e->mstream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (s.shadow8 && row < s.shadow8_rows) {
    const int brc = build_shadow8(e, s, row, 1, e->mstream);
    if (brc != WDBX_B200_OK) return brc;
  }
  if (s.tomb && ((s.tomb_host[row >> 5] >> (row & 31)) & 1u)) {
    s.tomb_host[row >> 5] &= ~(1u << (row & 31));
    s.n_dead -= 1;
    CU_TRY(cudaMemcpyAsync(s.tomb + (row >> 5), &s.tomb_host[row >> 5], 4, cudaMemcpyHostToDevice, e->mstream));
  }
  CU_TRY(cudaStreamSynchronize(e->mstream));
  return WDBX_B200_OK;
}

int wdbx_b200_tombstone(wdbx_b200_engine* e, int segment, int64_t row, int dead) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  if (segment < 0 || segment >= e->nseg) return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  Segment& s = e->seg[segment];
  if (row < 0 || row >= s.n_rows) return fail(WDBX_B200_ERR_ARG, "row %lld outside [0, %lld)", (long long)row, (long long)s.n_rows);
  if (!s.tomb) {
    if (!dead) return WDBX_B200_OK;
    const size_t words = static_cast<size_t>(s.cap_rows / 32);
    const size_t max_words = static_cast<size_t>(std::max<int64_t>(segment_max_rows(e), s.cap_rows) / 32 + 1);
    const cudaError_t ge = growbuf_grow(s.b_tomb, e->device, words * 4, 0, max_words * 4, e->mstream);
    refresh_pointers(s);
    if (ge != cudaSuccess) {
      cudaGetLastError();
      return fail(WDBX_B200_ERR_OOM, "cannot allocate the tombstone bitmap: %s", cudaGetErrorString(ge));
    }
    CU_TRY(cudaMemsetAsync(s.tomb, 0, words * 4, e->mstream));
    s.tomb_host.assign(words, 0u);
  }
  uint32_t& w = s.tomb_host[row >> 5];
  const uint32_t bit = 1u << (row & 31);
  const bool was = (w & bit) != 0;
  if (dead && !was) { w |= bit; s.n_dead += 1; }
  else if (!dead && was) { w &= ~bit; s.n_dead -= 1; }
  else return WDBX_B200_OK;
  CU_TRY(cudaMemcpyAsync(s.tomb + (row >> 5), &w, 4, cudaMemcpyHostToDevice, e->mstream));
  CU_TRY(cudaStreamSynchronize(e->mstream));
  return WDBX_B200_OK;
}

int wdbx_b200_clear(wdbx_b200_engine* e, int segment) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "engine is NULL");
  if (segment != WDBX_B200_ALL_SEGMENTS && (segment < 0 || segment >= e->nseg))
    return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  CU_TRY(cudaDeviceSynchronize());
  const int s0 = segment == WDBX_B200_ALL_SEGMENTS ? 0 : segment;
  const int s1 = segment == WDBX_B200_ALL_SEGMENTS ? e->nseg : segment + 1;
  for (int i = s0; i < s1; ++i) {
    Segment& s = e->seg[i];
    s.n_rows = 0;
    s.n_dead = 0;
    s.shadow_rows = 0;
    s.shadow8_rows = 0;
    if (s.tomb) {
      std::fill(s.tomb_host.begin(), s.tomb_host.end(), 0u);
      CU_TRY(cudaMemset(s.tomb, 0, s.tomb_host.size() * 4));
    }
  }
  return WDBX_B200_OK;
}

int wdbx_b200_read_rows(wdbx_b200_engine* e, int segment, int64_t row0, int64_t n, float* out_host) {
  if (!e || !out_host) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= e->nseg) return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  Segment& s = e->seg[segment];
  if (row0 < 0 || n < 0 || row0 + n > s.n_rows)
    return fail(WDBX_B200_ERR_ARG, "rows [%lld, %lld) outside [0, %lld)", (long long)row0, (long long)(row0 + n), (long long)s.n_rows);
  const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / (static_cast<int64_t>(e->dim) * 4));
  for (int64_t done = 0; done < n; done += chunk_rows) {
    const int64_t m = std::min(chunk_rows, n - done);
    const size_t need = static_cast<size_t>(m) * e->dim * 4;
    if (e->stage_rows_bytes < need) {
      cudaFree(e->stage_rows);
      e->stage_rows = nullptr;
      e->stage_rows_bytes = 0;
      CU_TRY(cudaMalloc(&e->stage_rows, need));
      e->stage_rows_bytes = need;
    }
    CU_TRY(launch_export_rows(s.rows + static_cast<size_t>(row0 + done) * row_bytes(e), m, e->dim,
                              static_cast<int>(row_bytes(e)), e->dtype == WDBX_B200_BF16, e->stage_rows, e->mstream));
    e->launches.fetch_add(1, std::memory_order_relaxed);
    CU_TRY(cudaMemcpyAsync(out_host + done * e->dim, e->stage_rows, need, cudaMemcpyDeviceToHost, e->mstream));
    CU_TRY(cudaStreamSynchronize(e->mstream));
  }
  return WDBX_B200_OK;
}

int wdbx_b200_read_row(wdbx_b200_engine* e, int segment, int64_t row, float* out_host) {
  return wdbx_b200_read_rows(e, segment, row, 1, out_host);
}

int wdbx_b200_search(wdbx_b200_engine* e, int segment, const float* q_dev, int B, int k, int metric,
                     uint64_t* keys_out, float* scores_out, int64_t* gids_out, int32_t* counts_out,
                     void* cuda_stream) {
  int rc = check_search_args(e, B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (!q_dev) return fail(WDBX_B200_ERR_ARG, "q_dev is NULL");
  if (segment != WDBX_B200_ALL_SEGMENTS && (segment < 0 || segment >= e->nseg))
    return fail(WDBX_B200_ERR_ARG, "segment %d outside [0, %d)", segment, e->nseg);
  DeviceGuard guard(e->device);
  CaptureRelax relax(static_cast<cudaStream_t>(cuda_stream));
  std::lock_guard<std::mutex> lk(e->mu);
  const int s0 = segment == WDBX_B200_ALL_SEGMENTS ? 0 : segment;
  const int s1 = segment == WDBX_B200_ALL_SEGMENTS ? e->nseg : segment + 1;
  rc = search_segments(e, s0, s1, q_dev, B, k, metric, keys_out, scores_out, reinterpret_cast<long long*>(gids_out),
                     counts_out, static_cast<cudaStream_t>(cuda_stream));
  if (rc == WDBX_B200_OK) e->searches.fetch_add(1, std::memory_order_relaxed);
  return rc;
}

int wdbx_b200_search_host(wdbx_b200_engine* e, int segment, const float* q_host, int B, int k, int metric,
                          float* scores_host, int64_t* gids_host, uint64_t* keys_host, int32_t* counts_host) {
  return host_search(e, segment, false, q_host, B, k, metric, -INFINITY, nullptr, scores_host, gids_host, keys_host, counts_host);
}

int wdbx_b200_search_filtered_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric, float min_score,
                                   const uint32_t* const* allow_bitmaps, float* scores_host, int64_t* gids_host,
                                   int32_t* counts_host) {
  return host_search(e, WDBX_B200_ALL_SEGMENTS, false, q_host, B, k, metric, min_score, allow_bitmaps, scores_host, gids_host,
                     nullptr, counts_host);
}

int wdbx_b200_merge(wdbx_b200_engine* e, const uint64_t* keys_dev, int G, int B, int k, uint64_t* keys_out,
                    float* scores_out, int64_t* gids_out, int32_t* counts_out, void* cuda_stream) {
  int rc = check_search_args(e, B, k, 0);
  if (rc != WDBX_B200_OK) return rc;
  if (!keys_dev) return fail(WDBX_B200_ERR_ARG, "keys_dev is NULL");
  if (G <= 0) return fail(WDBX_B200_ERR_ARG, "G must be positive");
  DeviceGuard guard(e->device);
  CU_TRY(launch_merge_topk(keys_dev, G, B, k, keys_out, scores_out, reinterpret_cast<long long*>(gids_out), counts_out,
                           static_cast<cudaStream_t>(cuda_stream)));
  e->launches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

int wdbx_b200_exchange_init(wdbx_b200_engine* e, int rank, int world, void* ipc_handle_out) {
  if (!e || !ipc_handle_out) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(WDBX_B200_ERR_ARG, "exchange needs 2 <= world <= %d and 0 <= rank < world", kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == WDBX_B200_IPC_HANDLE_BYTES, "ipc handle size");
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->xbuf) {
    CU_TRY(cudaMalloc(&e->xbuf, kXchgBytes));
    CU_TRY(cudaMemset(e->xbuf, 0, kXchgBytes));
  }
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, e->xbuf));
  memcpy(ipc_handle_out, &h, sizeof(h));
  e->xrank = rank;
  e->xworld = 0;  // armed by attach
  e->xseq = 0;
  (void)world;
  return WDBX_B200_OK;
}

int wdbx_b200_exchange_attach(wdbx_b200_engine* e, int world, const void* ipc_handles) {
  if (!e || !ipc_handles) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  if (!e->xbuf) return fail(WDBX_B200_ERR_ARG, "call wdbx_b200_exchange_init first");
  if (world < 2 || world > kMaxPeers) return fail(WDBX_B200_ERR_ARG, "world outside [2, %d]", kMaxPeers);
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  const unsigned char* hs = static_cast<const unsigned char*>(ipc_handles);
  for (int r = 0; r < world; ++r) {
    if (r == e->xrank) {
      e->xpeer[r] = e->xbuf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    e->xpeer[r] = static_cast<uint64_t*>(ptr);
    e->xopened[r] = true;
  }
  e->xworld = world;
  return WDBX_B200_OK;
}


int wdbx_b200_search_exchange(wdbx_b200_engine* e, const float* q_dev, int B, int k, int metric, uint64_t* keys_out,
                              float* scores_out, int64_t* gids_out, int32_t* counts_out, void* cuda_stream) {
  int rc = check_search_args(e, B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (!q_dev) return fail(WDBX_B200_ERR_ARG, "q_dev is NULL");
  DeviceGuard guard(e->device);
  CaptureRelax relax(static_cast<cudaStream_t>(cuda_stream));
  std::lock_guard<std::mutex> lk(e->mu);
  rc = exchange_search_locked(e, 0, e->nseg, q_dev, B, k, metric, keys_out, scores_out,
                              reinterpret_cast<long long*>(gids_out), counts_out, static_cast<cudaStream_t>(cuda_stream));
  if (rc == WDBX_B200_OK) e->searches.fetch_add(1, std::memory_order_relaxed);
  return rc;
}

int wdbx_b200_search_exchange_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric,
                                   float* scores_host, int64_t* gids_host, uint64_t* keys_host, int32_t* counts_host) {
  return host_search(e, WDBX_B200_ALL_SEGMENTS, true, q_host, B, k, metric, -INFINITY, nullptr, scores_host, gids_host,
                     keys_host, counts_host);
}

int wdbx_b200_search_exchange_filtered_host(wdbx_b200_engine* e, const float* q_host, int B, int k, int metric,
                                            float min_score, const uint32_t* const* allow_bitmaps, float* scores_host,
                                            int64_t* gids_host, int32_t* counts_host) {
  return host_search(e, WDBX_B200_ALL_SEGMENTS, true, q_host, B, k, metric, min_score, allow_bitmaps, scores_host, gids_host,
                     nullptr, counts_host);
}

int wdbx_b200_set_option(wdbx_b200_engine* e, const char* name, long long value) {
  if (!e || !name) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  const std::string n(name);
  if (n == "shadow_min_mb") e->shadow_min_bytes = value < 0 ? -1 : value << 20;
  else if (n == "gemm_min_batch") e->gemm_min_batch = static_cast<int>(value);
  else if (n == "gemm_mode") e->gemm_mode = static_cast<int>(value);
  else if (n == "pdl") e->pdl = value != 0;
  else if (n == "overlap") e->overlap = value != 0;
  else if (n == "filter_i8") e->filter_i8 = value != 0;
  else if (n == "queries_per_pass") e->tune.queries_per_pass = static_cast<int>(value);
  else return fail(WDBX_B200_ERR_ARG, "unknown option '%s'", name);
  return WDBX_B200_OK;
}

int wdbx_b200_set_kernel_timing(wdbx_b200_engine* e, int enable) {
  if (!e) return fail(WDBX_B200_ERR_ARG, "NULL engine");
  DeviceGuard guard(e->device);
  std::lock_guard<std::mutex> lk(e->mu);
  if (enable && !e->kev0) {
    CU_TRY(cudaEventCreate(&e->kev0));
    CU_TRY(cudaEventCreate(&e->kev1));
  }
  e->ktiming = enable != 0;
  e->kpending = false;
  return WDBX_B200_OK;
}

int wdbx_b200_get_stats(wdbx_b200_engine* e, wdbx_b200_stats* out) {
  if (!e || !out) return fail(WDBX_B200_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(e->mu);
  memset(out, 0, sizeof(*out));
  if (e->kpending && e->kev1) {
    DeviceGuard guard(e->device);
    float ms = 0.0f;
    if (cudaEventSynchronize(e->kev1) == cudaSuccess && cudaEventElapsedTime(&ms, e->kev0, e->kev1) == cudaSuccess) {
      out->last_kernel = e->last_kernel;
      out->last_kernel_ms = ms;
      if (e->last_kernel >= 2 && e->last_fcount && e->last_fregions > 0) {
        // measurement hook only: rows the filter passed on to the exact refine (summed over queries)
        std::vector<unsigned int> h(e->last_fregions);
        if (cudaMemcpy(h.data(), e->last_fcount, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
          long long sum = 0;
          for (unsigned int v : h) sum += v;
          out->last_candidates = sum;
        } else {
          cudaGetLastError();
        }
      }
    } else {
      cudaGetLastError();
    }
  }
  out->abi_version = WDBX_B200_ABI_VERSION;
  out->device = e->device;
  out->dim = e->dim;
  out->dim_padded = e->dpad;
  out->dtype = e->dtype;
  out->num_segments = e->nseg;
  out->sm_count = e->sm_count;
  const int64_t per_row = static_cast<int64_t>(row_bytes(e)) + 12;
  for (int s = 0; s < e->nseg; ++s) {
    const Segment& sg = e->seg[s];
    out->rows_total += sg.n_rows;
    out->rows_live += sg.n_rows - sg.n_dead;
    out->capacity_rows += sg.cap_rows;
    out->bytes_resident += static_cast<int64_t>(sg.b_rows.bytes + sg.b_inv.bytes + sg.b_sq.bytes + sg.b_gids.bytes +
                                                sg.b_tomb.bytes + sg.b_shadow.bytes + sg.b_rres.bytes + sg.b_shadow8.bytes +
                                                sg.b_sc8.bytes + sg.b_rres8.bytes);
    out->seg_rows[s] = sg.n_rows;
    out->seg_live[s] = sg.n_rows - sg.n_dead;
  }
  out->kernel_launches = e->launches.load();
  out->searches = e->searches.load();
  out->last_search_ms = e->last_search_ms;
  return WDBX_B200_OK;
}

// ---- single process, several GPUs ------------------------------------------------------------------------
// north_star: "the shard manager maps num_shards onto the 8 GPUs of one box" with the REST server / CLI
// (wdbx/api/server.py:141-152, wdbx/cli.py:541) untouched, i.e. from ONE ordinary Python process.  A group ties
// G engines (one per device, rows of every segment striped over them by the host layer) together: peer access is
// enabled directly (no CUDA IPC), every search is launched on all G devices from the calling thread, and the
// devices merge their top-k among themselves with the same NVLink exchange the one-process-per-GPU layout uses
// (small batches) or by peer copies of the packed keys to device 0 + merge kernel K3 (any B, k).

int wdbx_b200_group_create(wdbx_b200_engine* const* engines, int n, wdbx_b200_group** out) {
  if (!out) return fail(WDBX_B200_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!engines || n < 2 || n > kMaxPeers) return fail(WDBX_B200_ERR_ARG, "a group needs 2..%d engines", kMaxPeers);
  for (int i = 0; i < n; ++i) {
    wdbx_b200_engine* e = engines[i];
    if (!e) return fail(WDBX_B200_ERR_ARG, "engine %d is NULL", i);
    if (e->dim != engines[0]->dim || e->dtype != engines[0]->dtype || e->nseg != engines[0]->nseg)
      return fail(WDBX_B200_ERR_ARG, "engines of a group must share dim, dtype and num_segments");
    if (e->xworld != 0 || e->xbuf) return fail(WDBX_B200_ERR_ARG, "engine %d already takes part in an exchange", i);
    for (int j = 0; j < i; ++j)
      if (engines[j]->device == e->device) return fail(WDBX_B200_ERR_ARG, "engines %d and %d share device %d", j, i, e->device);
  }
  for (int i = 0; i < n; ++i) {
    DeviceGuard guard(engines[i]->device);
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      int can = 0;
      CU_TRY(cudaDeviceCanAccessPeer(&can, engines[i]->device, engines[j]->device));
      if (!can) return fail(WDBX_B200_ERR_CUDA, "device %d cannot access device %d (no NVLink / P2P)", engines[i]->device,
                            engines[j]->device);
      const cudaError_t pe = cudaDeviceEnablePeerAccess(engines[j]->device, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return fail(WDBX_B200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", engines[i]->device, engines[j]->device,
                    cudaGetErrorString(pe));
      }
      cudaGetLastError();
    }
    CU_TRY(cudaMalloc(&engines[i]->xbuf, kXchgBytes));
    CU_TRY(cudaMemset(engines[i]->xbuf, 0, kXchgBytes));
  }
  wdbx_b200_group* g = new (std::nothrow) wdbx_b200_group();
  if (!g) return fail(WDBX_B200_ERR_OOM, "host allocation failed");
  g->eng.assign(engines, engines + n);
  g->ev.resize(n, nullptr);
  for (int i = 0; i < n; ++i) {
    wdbx_b200_engine* e = engines[i];
    std::lock_guard<std::mutex> lk(e->mu);
    for (int r = 0; r < n; ++r) e->xpeer[r] = engines[r]->xbuf;   // peer access: plain device pointers
    e->xrank = i;
    e->xworld = n;
    e->xseq = 0;
    DeviceGuard guard(e->device);
    if (cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      wdbx_b200_group_destroy(g);
      return fail(WDBX_B200_ERR_CUDA, "event creation failed");
    }
  }
  g->rcs.assign(n, WDBX_B200_OK);
  g->errs.assign(n, std::string());
  if (env_int("WDBX_B200_GROUP_THREADS", 1) != 0)
    for (int i = 1; i < n; ++i) g->workers.emplace_back(group_worker, g, i);
  *out = g;
  return WDBX_B200_OK;
}

void wdbx_b200_group_destroy(wdbx_b200_group* g) {
  if (!g) return;
  g->stop.store(true, std::memory_order_release);
  {
    std::lock_guard<std::mutex> lk(g->wmu);
    g->wcv.notify_all();
  }
  for (auto& t : g->workers) t.join();
  for (size_t i = 0; i < g->eng.size(); ++i) {
    wdbx_b200_engine* e = g->eng[i];
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    {
      std::lock_guard<std::mutex> lk(e->mu);
      e->xworld = 0;   // the engines live on (their owner destroys them); they just stop exchanging
    }
    if (g->ev[i]) cudaEventDestroy(g->ev[i]);
    if (i == 0) {
      cudaFree(g->gkeys);
      if (g->ev_in) cudaEventDestroy(g->ev_in);
      if (g->ev_out) cudaEventDestroy(g->ev_out);
    }
  }
  cudaFreeHost(g->hq);
  cudaFreeHost(g->hres);
  cudaGetLastError();
  delete g;
}

}  // extern "C"

namespace {

// One engine's share of a group search: queries in (from the group's pinned buffer, or from device 0 after
// `ev_in`), launches, and -- batches beyond the exchange's limits -- its keys to device 0.
int group_launch_one(wdbx_b200_group* g, int i, const GroupJob& j) {
  const int G = static_cast<int>(g->eng.size());
  wdbx_b200_engine* e = g->eng[i];
  wdbx_b200_engine* e0 = g->eng[0];
  const int nseg = e0->nseg;
  const int lists = j.segment == WDBX_B200_EACH_SEGMENT ? nseg : 1;
  const size_t nq = static_cast<size_t>(j.B) * e0->dim;
  const ResLayout R(lists, j.B, j.k);
  DeviceGuard guard(e->device);
  int rc = ensure_host_buffers(e, nq, R.bytes, false);
  if (rc != WDBX_B200_OK) return rc;
  cudaStream_t st = e->hstream;
  if (j.q_dev0 == nullptr) {
    CU_TRY(cudaMemcpyAsync(e->dq, g->hq, nq * 4, cudaMemcpyHostToDevice, st));
  } else {
    CU_TRY(cudaStreamWaitEvent(st, j.ev_in, 0));
    CU_TRY(cudaMemcpyPeerAsync(e->dq, e->device, j.q_dev0, e0->device, nq * 4, st));
  }
  if (i == 0) CU_TRY(cudaEventRecord(e->ev0, st));
  const uint32_t* const* al = j.allow ? j.allow + static_cast<size_t>(i) * nseg : nullptr;
  const bool direct = j.exchange && i == 0;   // engine 0's exchange kernel may write the caller's buffers itself
  rc = launch_host_lists(e, j.segment, j.exchange, e->dq, j.B, j.k, j.metric, j.min_score, al, R, st,
                         direct ? j.out0_keys : nullptr, direct ? j.out0_scores : nullptr, direct ? j.out0_gids : nullptr,
                         direct ? j.out0_counts : nullptr);
  if (rc != WDBX_B200_OK) return rc;
  if (!j.exchange) {
    // this device's [lists][B][k] keys -> device 0, laid out [lists][G][B][k] for the merge kernel
    const size_t per = static_cast<size_t>(j.B) * j.k;
    for (int l = 0; l < lists; ++l)
      CU_TRY(cudaMemcpyPeerAsync(g->gkeys + (static_cast<size_t>(l) * G + i) * per, e0->device,
                                 reinterpret_cast<uint64_t*>(e->dres + R.off_keys) + static_cast<size_t>(l) * per, e->device,
                                 per * 8, st));
    CU_TRY(cudaEventRecord(g->ev[i], st));
  }
  return WDBX_B200_OK;
}

// Launcher threads: engine i > 0 is driven by its own thread so that the G devices start within a few
// microseconds of each other (from one thread the last of 8 devices started ~140 us after the first, and the
// on-device exchange makes everybody wait for it).  Workers spin briefly after a job, then sleep.
void group_worker(wdbx_b200_group* g, int i) {
  cudaSetDevice(g->eng[i]->device);
  uint64_t seen = 0;
  for (;;) {
    int spins = 0;
    while (g->gen.load(std::memory_order_acquire) == seen && !g->stop.load(std::memory_order_acquire)) {
      if (++spins < 4000) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        continue;
      }
      std::unique_lock<std::mutex> lk(g->wmu);
      g->sleepers.fetch_add(1);
      g->wcv.wait(lk, [&] { return g->gen.load(std::memory_order_acquire) != seen || g->stop.load(); });
      g->sleepers.fetch_sub(1);
    }
    if (g->stop.load(std::memory_order_acquire)) return;
    seen = g->gen.load(std::memory_order_acquire);
    g_err[0] = 0;
    const int rc = group_launch_one(g, i, g->job);
    g->rcs[i] = rc;
    if (rc != WDBX_B200_OK) g->errs[i] = g_err;
    g->done.fetch_add(1, std::memory_order_release);
  }
}

// Launch one search on every engine of the group.  The queries are either in the group's pinned host buffer
// (q_dev0 == NULL) or on device 0 (q_dev0, ordered after `ev_in`).  The merged result ends up on device 0: in
// engine 0's packed result block, or in the caller's device pointers (out0_*).  Caller holds g->mu.
int group_launch(wdbx_b200_group* g, int segment, const float* q_dev0, cudaEvent_t ev_in, int B, int k, int metric,
                 float min_score, const uint32_t* const* allow, const ResLayout& R, uint64_t* out0_keys, float* out0_scores,
                 long long* out0_gids, int* out0_counts) {
  const int G = static_cast<int>(g->eng.size());
  wdbx_b200_engine* e0 = g->eng[0];
  const int lists = segment == WDBX_B200_EACH_SEGMENT ? e0->nseg : 1;
  const bool exchange = B <= kXchgMaxB && k <= kXchgMaxK;
  if (!exchange) {
    const size_t need = static_cast<size_t>(lists) * G * B * k;
    if (g->gkeys_n < need) {
      DeviceGuard guard(e0->device);
      CU_TRY(cudaStreamSynchronize(e0->hstream));
      cudaFree(g->gkeys);
      g->gkeys = nullptr;
      g->gkeys_n = 0;
      CU_TRY(cudaMalloc(&g->gkeys, need * 8));
      g->gkeys_n = need;
    }
  }
  GroupJob& j = g->job;
  j.segment = segment; j.q_dev0 = q_dev0; j.ev_in = ev_in; j.B = B; j.k = k; j.metric = metric;
  j.min_score = min_score; j.allow = allow; j.exchange = exchange;
  j.out0_keys = out0_keys; j.out0_scores = out0_scores; j.out0_gids = out0_gids; j.out0_counts = out0_counts;
  int rc = WDBX_B200_OK;
  if (g->workers.empty()) {
    for (int i = 0; i < G && rc == WDBX_B200_OK; ++i) rc = group_launch_one(g, i, j);
  } else {
    g->done.store(0, std::memory_order_relaxed);
    g->gen.fetch_add(1, std::memory_order_release);
    if (g->sleepers.load() > 0) {
      std::lock_guard<std::mutex> lk(g->wmu);
      g->wcv.notify_all();
    }
    rc = group_launch_one(g, 0, j);
    for (int spins = 0; g->done.load(std::memory_order_acquire) < G - 1; ++spins)
      if (spins > 20000) std::this_thread::yield();
    for (int i = 1; i < G; ++i)
      if (rc == WDBX_B200_OK && g->rcs[i] != WDBX_B200_OK) rc = fail(g->rcs[i], "device %d: %s", g->eng[i]->device, g->errs[i].c_str());
  }
  if (rc != WDBX_B200_OK) return rc;
  if (!exchange) {
    DeviceGuard guard(e0->device);
    cudaStream_t st = e0->hstream;
    for (int i = 1; i < G; ++i) CU_TRY(cudaStreamWaitEvent(st, g->ev[i], 0));
    const size_t per = static_cast<size_t>(B) * k;
    uint64_t* keys = out0_keys ? out0_keys : reinterpret_cast<uint64_t*>(e0->dres + R.off_keys);
    float* scores = out0_scores ? out0_scores : reinterpret_cast<float*>(e0->dres + R.off_scores);
    long long* gids = out0_gids ? out0_gids : reinterpret_cast<long long*>(e0->dres + R.off_gids);
    int* counts = out0_counts ? out0_counts : reinterpret_cast<int*>(e0->dres + R.off_counts);
    for (int l = 0; l < lists; ++l) {
      CU_TRY(launch_merge_topk(g->gkeys + static_cast<size_t>(l) * G * per, G, B, k, keys + l * per, scores + l * per,
                               gids + l * per, counts + static_cast<size_t>(l) * B, st));
      e0->launches.fetch_add(1, std::memory_order_relaxed);
    }
  }
  return WDBX_B200_OK;
}

int check_group_args(wdbx_b200_group* g, int segment, int B, int k, int metric) {
  if (!g) return fail(WDBX_B200_ERR_ARG, "group is NULL");
  const int rc = check_search_args(g->eng[0], B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (segment < WDBX_B200_EACH_SEGMENT || segment >= g->eng[0]->nseg)
    return fail(WDBX_B200_ERR_ARG, "segment %d outside [-2, %d)", segment, g->eng[0]->nseg);
  return WDBX_B200_OK;
}

}  // namespace

extern "C" {

int wdbx_b200_group_search_host(wdbx_b200_group* g, int segment, const float* q_host, int B, int k, int metric,
                                float min_score, const uint32_t* const* allow_bitmaps, float* scores_host,
                                int64_t* gids_host, uint64_t* keys_host, int32_t* counts_host) {
  int rc = check_group_args(g, segment, B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (!q_host) return fail(WDBX_B200_ERR_ARG, "q_host is NULL");
  std::lock_guard<std::mutex> glk(g->mu);
  wdbx_b200_engine* e0 = g->eng[0];
  const int lists = segment == WDBX_B200_EACH_SEGMENT ? e0->nseg : 1;
  const size_t nq = static_cast<size_t>(B) * e0->dim;
  const ResLayout R(lists, B, k);
  {
    DeviceGuard guard(e0->device);
    if (g->hq_floats < nq) {
      cudaFreeHost(g->hq); g->hq = nullptr; g->hq_floats = 0;
      CU_TRY(cudaHostAlloc(&g->hq, nq * 4, cudaHostAllocPortable));
      g->hq_floats = nq;
    }
    if (g->hres_bytes < R.bytes) {
      cudaFreeHost(g->hres); g->hres = nullptr; g->hres_bytes = 0;
      CU_TRY(cudaHostAlloc(&g->hres, R.bytes, cudaHostAllocPortable));
      g->hres_bytes = R.bytes;
    }
  }
  memcpy(g->hq, q_host, nq * 4);
  rc = group_launch(g, segment, nullptr, nullptr, B, k, metric, min_score, allow_bitmaps, R, nullptr, nullptr, nullptr, nullptr);
  // every launched kernel must finish before the buffers are reused, error or not (peers wait for each other)
  int first_err = rc;
  {
    DeviceGuard guard(e0->device);
    if (rc == WDBX_B200_OK) {
      cudaError_t ce = cudaEventRecord(e0->ev1, e0->hstream);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(g->hres, e0->dres, R.bytes, cudaMemcpyDeviceToHost, e0->hstream);
      if (ce != cudaSuccess) first_err = fail(WDBX_B200_ERR_CUDA, "result copy: %s", cudaGetErrorString(ce));
    }
  }
  // device 0's stream ends after the merge, i.e. after every peer's keys arrived (exchange) or were copied over
  // (gather) -- hence after every device's H2D of the shared pinned query buffer.  The peers' streams need no host
  // synchronisation of their own on the good path (their buffers are reused in stream order).
  for (size_t i = 0; i < g->eng.size(); ++i) {
    if (i > 0 && first_err == WDBX_B200_OK) break;
    wdbx_b200_engine* e = g->eng[i];
    DeviceGuard guard(e->device);
    const cudaError_t ce = cudaStreamSynchronize(e->hstream);
    if (ce != cudaSuccess && first_err == WDBX_B200_OK)
      first_err = fail(WDBX_B200_ERR_CUDA, "device %d: %s", e->device, cudaGetErrorString(ce));
  }
  if (first_err != WDBX_B200_OK) return first_err;
  float ms = 0.0f;
  {
    DeviceGuard guard(e0->device);
    if (cudaEventElapsedTime(&ms, e0->ev0, e0->ev1) == cudaSuccess) e0->last_search_ms = ms;
    else cudaGetLastError();
  }
  if (keys_host) memcpy(keys_host, g->hres + R.off_keys, R.nres * 8);
  if (gids_host) memcpy(gids_host, g->hres + R.off_gids, R.nres * 8);
  if (scores_host) memcpy(scores_host, g->hres + R.off_scores, R.nres * 4);
  if (counts_host) memcpy(counts_host, g->hres + R.off_counts, static_cast<size_t>(lists) * B * 4);
  for (wdbx_b200_engine* e : g->eng) e->searches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

int wdbx_b200_group_search(wdbx_b200_group* g, const float* q_dev0, int B, int k, int metric, uint64_t* keys_out,
                           float* scores_out, int64_t* gids_out, int32_t* counts_out, void* cuda_stream) {
  int rc = check_group_args(g, WDBX_B200_ALL_SEGMENTS, B, k, metric);
  if (rc != WDBX_B200_OK) return rc;
  if (!q_dev0) return fail(WDBX_B200_ERR_ARG, "q_dev0 is NULL");
  std::lock_guard<std::mutex> glk(g->mu);
  wdbx_b200_engine* e0 = g->eng[0];
  cudaStream_t user = static_cast<cudaStream_t>(cuda_stream);
  const ResLayout R(1, B, k);
  {
    DeviceGuard guard(e0->device);
    if (!g->ev_in) {
      CU_TRY(cudaEventCreateWithFlags(&g->ev_in, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&g->ev_out, cudaEventDisableTiming));
    }
    CU_TRY(cudaEventRecord(g->ev_in, user));   // the queries are ready once the caller's stream gets here
  }
  rc = group_launch(g, WDBX_B200_ALL_SEGMENTS, q_dev0, g->ev_in, B, k, metric, -INFINITY, nullptr, R, keys_out, scores_out,
                    reinterpret_cast<long long*>(gids_out), counts_out);
  if (rc != WDBX_B200_OK) return rc;
  {
    DeviceGuard guard(e0->device);
    CU_TRY(cudaEventRecord(g->ev_out, e0->hstream));
    CU_TRY(cudaStreamWaitEvent(user, g->ev_out, 0));   // the caller's stream continues after device 0's merge
  }
  for (wdbx_b200_engine* e : g->eng) e->searches.fetch_add(1, std::memory_order_relaxed);
  return WDBX_B200_OK;
}

}  // extern "C"
