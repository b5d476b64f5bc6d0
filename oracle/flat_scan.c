/*
 * flat_scan.c -- C restatement of the reference's CPU search path.  TEST / BASELINE
 * INFRASTRUCTURE ONLY: linked by tests, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs, never by the product.
 *
 * What it restates.  The reference calls faiss-cpu `IndexFlatIP::search` /
 * `IndexFlatL2::search` (utils/vector_store.py:191; faiss-cpu>=1.7.0, requirements.txt:5, not
 * vendored, not installable here).  FAISS' published algorithm for small batches (nq < 20) is:
 * one SIMD inner product (`fvec_inner_product`) or squared distance (`fvec_L2sqr`) per stored
 * row, results collected in a k-heap per query, OpenMP parallelism over QUERIES only -- so a
 * single query runs on ONE thread.  `oracle_flat_search(..., nthreads=1)` is that path.
 * `nthreads > 1` additionally splits the ROWS of each query across threads (per-thread heaps,
 * merged at the end): FAISS does not do this for nq=1, it is offered so the CPU baseline can use
 * every host core ("all the host threads it can use").
 *
 * Pinning: parity unpinned at the arithmetic level (no FAISS numeric fixture exists in the
 * reference); this file is cross-checked against oracle/flat_ip.py, which is
 * pinned on the reference's format / tie / normalisation fixtures (tests/test_oracle_golden.py).
 * Order of equal scores: lower id first (FAISS leaves it implementation-defined).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float s;   /* larger is better */
    int64_t id;
} hit_t;

/* a is worse than b */
static inline int worse(hit_t a, hit_t b) { return a.s < b.s || (a.s == b.s && a.id > b.id); }

/* min-heap on "goodness": root = worst kept hit */
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
static void heap_push(hit_t* h, int* n, int k, hit_t v) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = v;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (!worse(h[i], h[p])) break;
            hit_t t = h[i];
            h[i] = h[p];
            h[p] = t;
            i = p;
        }
    } else if (worse(h[0], v)) {
        h[0] = v;
        heap_sift_down(h, k, 0);
    }
}
static int cmp_best_first(const void* a, const void* b) {
    const hit_t *x = (const hit_t*)a, *y = (const hit_t*)b;
    if (worse(*y, *x)) return -1;
    if (worse(*x, *y)) return 1;
    return 0;
}

static inline float dot_f32(const float* a, const float* b, int d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; ++i) s += a[i] * b[i];
    return s;
}
static inline float l2sqr_f32(const float* a, const float* b, int d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; ++i) {
        float t = a[i] - b[i];
        s += t * t;
    }
    return s;
}

/* packed EXIF word vs filter: same conjunction as core/searcher.py:1884-1950, layout include/psx.h */
typedef struct {
    uint32_t flags, season, period, year, month, reserved;
    uint64_t start, end;
} oracle_filter;

static inline int attr_pass(uint64_t a, const oracle_filter* f) {
    uint32_t fl = f->flags;
    if (fl & 0x0Fu) {
        if (!(a >> 63)) return 0;
        if ((fl & 0x01u) && ((a >> 60) & 7u) != f->season) return 0;
        if ((fl & 0x02u) && ((a >> 57) & 7u) != f->period) return 0;
        if ((fl & 0x04u) && ((a >> 43) & 0x3fffu) != f->year) return 0;
        if ((fl & 0x08u) && ((a >> 39) & 0xfu) != f->month) return 0;
    }
    if (fl & 0x10u) {
        uint64_t dt = a & ((1ull << 39) - 1);
        if (!dt) return 0;
        if ((fl & 0x20u) && dt < f->start) return 0;
        if ((fl & 0x40u) && dt > f->end) return 0;
    }
    return 1;
}

/*
 * X [n][d] fp32 row major, Q [nq][d]; metric 0 = inner product, 1 = squared L2.
 * attrs/filter may be NULL.  D [nq][k], I [nq][k]; unfilled slots (-inf|+inf, -1).
 * Returns 0, or -1 on allocation failure.
 */
int oracle_flat_search(const float* X, int64_t n, int d, const float* Q, int nq, int k, int metric,
                       const uint64_t* attrs, const oracle_filter* filter, float* D, int64_t* I, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#ifndef _OPENMP
    nthreads = 1;
#endif
    const int use_filter = attrs && filter && filter->flags;
    for (int qi = 0; qi < nq; ++qi) {
        const float* q = Q + (size_t)qi * d;
        hit_t* heaps = (hit_t*)malloc((size_t)nthreads * k * sizeof(hit_t));
        int* counts = (int*)calloc(nthreads, sizeof(int));
        if (!heaps || !counts) {
            free(heaps);
            free(counts);
            return -1;
        }
#pragma omp parallel num_threads(nthreads)
        {
#ifdef _OPENMP
            const int t = omp_get_thread_num();
#else
            const int t = 0;
#endif
            hit_t* h = heaps + (size_t)t * k;
            int cnt = 0;
            const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            for (int64_t j = lo; j < hi; ++j) {
                if (use_filter && !attr_pass(attrs[j], filter)) continue;
                const float* x = X + (size_t)j * d;
                hit_t v;
                v.s = metric == 0 ? dot_f32(x, q, d) : -l2sqr_f32(x, q, d);
                v.id = j;
                heap_push(h, &cnt, k, v);
            }
            counts[t] = cnt;
        }
        /* merge the per-thread heaps */
        int total = 0;
        for (int t = 0; t < nthreads; ++t) total += counts[t];
        hit_t* all = (hit_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof(hit_t));
        if (!all) {
            free(heaps);
            free(counts);
            return -1;
        }
        int m = 0;
        for (int t = 0; t < nthreads; ++t) {
            memcpy(all + m, heaps + (size_t)t * k, (size_t)counts[t] * sizeof(hit_t));
            m += counts[t];
        }
        qsort(all, (size_t)m, sizeof(hit_t), cmp_best_first);
        for (int i = 0; i < k; ++i) {
            if (i < m) {
                D[(size_t)qi * k + i] = metric == 0 ? all[i].s : -all[i].s;
                I[(size_t)qi * k + i] = all[i].id;
            } else {
                D[(size_t)qi * k + i] = metric == 0 ? -INFINITY : INFINITY;
                I[(size_t)qi * k + i] = -1;
            }
        }
        free(all);
        free(heaps);
        free(counts);
    }
    return 0;
}

/* Deterministic filler for the CPU baseline: unit-norm pseudo-random rows, parallel over rows. */
void oracle_fill_unit_rows(float* X, int64_t n, int d, uint64_t seed) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
        uint64_t s = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(j + 1));
        float* x = X + (size_t)j * d;
        double nrm = 0.0;
        for (int i = 0; i < d; ++i) {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            /* sum of two uniforms, centred: cheap bell-ish distribution, good enough for timing */
            float v = (float)((double)(s & 0xffffffu) / 16777216.0 + (double)((s >> 24) & 0xffffffu) / 16777216.0 - 1.0);
            x[i] = v;
            nrm += (double)v * v;
        }
        const float inv = nrm > 0 ? (float)(1.0 / sqrt(nrm)) : 0.f;
        for (int i = 0; i < d; ++i) x[i] *= inv;
    }
}

/* Pin the OpenMP pool regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1, which would
 * silently turn the all-core CPU arm into a one-thread arm). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
