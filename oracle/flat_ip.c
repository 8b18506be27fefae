/* flat_ip.c -- plain C restatement of the reference's exact search arithmetic.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never linked into the product library.
 *
 * Restates, for one query at a time:
 *   - FaissIndex._normalize_vector            wdbx/core/indexing.py:851-856   (oracle_normalize_rows)
 *   - faiss.IndexFlatIP.search as called at   wdbx/core/indexing.py:1013      (oracle_flat_search)
 *     [third-party faiss-cpu>=1.7.0, requirements.txt:20, absent from the image: published
 *      algorithm = fp32 inner product of the query with every stored row, k best kept in a
 *      binary heap, results reordered best-first]
 *   - the cross-shard merge of VectorStore.search vector_store.py:324-330 is the per-thread heap
 *     merge at the end (row blocks play the role of shards).
 * Extensions: metric 1 = raw inner product, metric 2 = negative squared L2 distance.
 * Tie rule: higher score first, then lower row.  NaN scores rank as -inf.
 *
 * Pinned against tests/golden/reference_golden.json (outputs of the reference's own code) by
 * tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  float s;
  int64_t r;
} hit_t;

/* a "worse" than b?  (min-heap on quality: root = worst kept hit) */
static inline int worse(hit_t a, hit_t b) { return a.s < b.s || (a.s == b.s && a.r > b.r); }

static void heap_sift_down(hit_t* h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && worse(h[l], h[m])) m = l;
    if (r < n && worse(h[r], h[m])) m = r;
    if (m == i) return;
    hit_t t = h[i];
    h[i] = h[m];
    h[m] = t;
    i = m;
  }
}

static void heap_push(hit_t* h, int* n, int k, hit_t x) {
  if (*n < k) {
    int i = (*n)++;
    h[i] = x;
    while (i > 0) {
      int p = (i - 1) / 2;
      if (!worse(h[i], h[p])) break;
      hit_t t = h[i];
      h[i] = h[p];
      h[p] = t;
      i = p;
    }
  } else if (worse(h[0], x)) {
    h[0] = x;
    heap_sift_down(h, k, 0);
  }
}

static int cmp_best_first(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a;
  const hit_t* y = (const hit_t*)b;
  if (worse(*y, *x)) return -1;
  if (worse(*x, *y)) return 1;
  return 0;
}

/* x <- x / ||x||_2 per row, zero rows untouched (indexing.py:851-856). */
void oracle_normalize_rows(float* X, int64_t n, int d) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float* x = X + i * (int64_t)d;
    float ss = 0.0f;
    for (int j = 0; j < d; ++j) ss += x[j] * x[j];
    float nrm = sqrtf(ss);
    if (nrm > 0.0f)
      for (int j = 0; j < d; ++j) x[j] = x[j] / nrm;
  }
}

static inline float score_row(const float* x, const float* q, int d, int metric) {
  float acc = 0.0f;
  if (metric == 2) {
#pragma omp simd reduction(+ : acc)
    for (int j = 0; j < d; ++j) {
      float t = x[j] - q[j];
      acc += t * t;
    }
    return -acc;
  }
#pragma omp simd reduction(+ : acc)
  for (int j = 0; j < d; ++j) acc += x[j] * q[j];
  return acc;
}

/* Exact top-k of one query over X [n, d].  metric 0/1: inner product (for cosine pass rows
 * normalised by oracle_normalize_rows and a normalised query), metric 2: -||x-q||^2.
 * dead: optional byte mask of excluded rows.  Returns the number of hits written (<= k),
 * best-first.  nthreads <= 0: all OpenMP threads. */
int oracle_flat_search(const float* X, int64_t n, int d, const float* q, int metric, int k, const uint8_t* dead,
                       int64_t* out_rows, float* out_scores, int nthreads) {
  if (k <= 0 || n <= 0) return 0;
  int nt = 1;
#ifdef _OPENMP
  nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
  if ((int64_t)nt > n) nt = (int)n;
  hit_t* heaps = (hit_t*)malloc(sizeof(hit_t) * (size_t)nt * (size_t)k);
  int* cnt = (int*)calloc((size_t)nt, sizeof(int));
  if (!heaps || !cnt) {
    free(heaps);
    free(cnt);
    return -1;
  }
#pragma omp parallel num_threads(nt)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    hit_t* h = heaps + (size_t)t * k;
    int m = 0;
    int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (int64_t i = lo; i < hi; ++i) {
      if (dead && dead[i]) continue;
      float s = score_row(X + i * (int64_t)d, q, d, metric);
      if (s != s) s = -INFINITY;
      hit_t x = {s, i};
      heap_push(h, &m, k, x);
    }
    cnt[t] = m;
  }
  /* merge the per-thread heaps (the "cross-shard" merge) */
  int total = 0;
  for (int t = 0; t < nt; ++t) total += cnt[t];
  hit_t* all = (hit_t*)malloc(sizeof(hit_t) * (size_t)(total > 0 ? total : 1));
  int w = 0;
  for (int t = 0; t < nt; ++t) {
    memcpy(all + w, heaps + (size_t)t * k, sizeof(hit_t) * (size_t)cnt[t]);
    w += cnt[t];
  }
  qsort(all, (size_t)total, sizeof(hit_t), cmp_best_first);
  int out = total < k ? total : k;
  for (int i = 0; i < out; ++i) {
    out_rows[i] = all[i].r;
    out_scores[i] = all[i].s;
  }
  free(all);
  free(heaps);
  free(cnt);
  return out;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
