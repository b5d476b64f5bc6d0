// psx_fuse.cuh -- K5: the numeric core of Searcher._hybrid_search (core/searcher.py:893-986) and
// Searcher._distance_to_score (core/searcher.py:605-625) on the device, over the merged vector
// candidates and the (sparse) keyword hits of each query.  One CTA per query.
//
// The reference computes in Python floats (IEEE double) and rounds with round(x, 6); both are
// reproduced bit for bit: every double operation uses an explicit round-to-nearest intrinsic (no
// FMA contraction) and round6() implements CPython's correctly-rounded decimal rounding
// (round-half-even on the EXACT binary value) with an error-free product.
#pragma once
#include "psx_common.cuh"

namespace psx {

constexpr int FUSE_MAX_ENTRIES = 2048;  // vector hits + keyword hits per query (11 index bits in the key)

// CPython float.__round__(x, 6) for 0 <= x < 2^21 * 1e-6: returns (n / 1e6, n), n = nearest integer to
// the exact value of x * 10^6, ties to even.
__device__ __forceinline__ double round6(double x, long long* n_out) {
    const double p = __dmul_rn(x, 1e6);
    const double e = __fma_rn(x, 1e6, -p);          // exact: x*1e6 == p + e
    const double f = floor(p);
    const double t = __dadd_rn(__dadd_rn(p, -f), -0.5);  // exact (p - f in [0,1))
    long long n = (long long)f;
    // exact fractional part is (p - f) + e; compare it with 1/2 without rounding
    if (t > -e || (t == -e && (n & 1))) n += 1;
    if (n_out) *n_out = n;
    return __ddiv_rn((double)n, 1e6);
}

// Searcher._distance_to_score
__device__ __forceinline__ double distance_to_score(float distance, int metric) {
    if (metric == PSX_METRIC_IP) {
        double sim = (double)distance;
        sim = sim > 1.0 ? 1.0 : sim;      // min(1.0, distance)
        sim = sim < -1.0 ? -1.0 : sim;    // max(-1.0, .)
        double score = __ddiv_rn(__dadd_rn(sim, 1.0), 2.0);
        if (score > 0.7)
            score = __dadd_rn(0.7, __dmul_rn(__dadd_rn(score, -0.7), 1.3));
        else if (score < 0.3)
            score = __dmul_rn(score, 0.8);
        score = score > 1.0 ? 1.0 : score;
        score = score < 0.0 ? 0.0 : score;
        return round6(score, nullptr);
    }
    double dist = (double)distance;
    if (dist < 0) dist = 0;
    return round6(exp(__dmul_rn(-0.5, dist)), nullptr);
}

struct FuseParams {
    const float* vec_dist;      // [nq][kv]   raw distances as VectorStore.search reports them
    const long long* vec_ids;   // [nq][kv]   -1 = empty
    const double* vec_boost;    // [nq][kv]   metadata boost per hit, or nullptr (1.0)
    const long long* kw_ids;    // [nq][kw]   -1 = empty
    const double* kw_scores;    // [nq][kw]
    const double* kw_boost;     // [nq][kw]   boost for keyword-only hits, or nullptr
    long long* out_ids;         // [nq][kv+kw]
    double* out_fused;          // [nq][kv+kw]
    double* out_vscore;         // [nq][kv+kw]
    double* out_kscore;         // [nq][kv+kw]
    int* out_count;             // [nq]
    int kv, kw;
    double wv, wk;
    int metric, allow_keyword_only, keyword_filtered;
};

__global__ void __launch_bounds__(256) hybrid_fuse_kernel(const FuseParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int E = p.kv + p.kw;
    int np = 64;
    while (np < E) np <<= 1;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);      // [np]
    double* e_vs = reinterpret_cast<double*>(keys + np);          // [E] vector score per entry (or 0)
    double* e_ks = e_vs + E;                                      // [E] keyword score per entry (or 0)
    long long* e_id = reinterpret_cast<long long*>(e_ks + E);     // [E]
    int* v_kw = reinterpret_cast<int*>(e_id + E);                 // [kv] index of the matching keyword hit, or -1
    const size_t q = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int i = tid; i < p.kv; i += nt) {
        e_id[i] = p.vec_ids[q * p.kv + i];
        v_kw[i] = -1;
    }
    for (int j = tid; j < p.kw; j += nt) e_id[p.kv + j] = p.kw_ids[q * p.kw + j];
    for (int i = tid; i < np; i += nt) keys[i] = 0ull;
    __syncthreads();
    // join: keyword hit j <-> vector hit with the same id
    for (int j = tid; j < p.kw; j += nt) {
        const long long id = e_id[p.kv + j];
        if (id < 0) continue;
        for (int i = 0; i < p.kv; ++i) {
            if (e_id[i] == id) {
                v_kw[i] = j;
                e_id[p.kv + j] = -2;  // consumed by the vector entry
                break;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < E; e += nt) {
        const long long id = e_id[e];
        if (id < 0) continue;
        const bool is_vec = e < p.kv;
        bool has_v = is_vec, has_k = false;
        double vs = 0.0, ks = 0.0, boost = 1.0;
        if (is_vec) {
            vs = distance_to_score(p.vec_dist[q * p.kv + e], p.metric);
            if (v_kw[e] >= 0) {
                has_k = true;
                ks = p.kw_scores[q * p.kw + v_kw[e]];
            }
            if (p.vec_boost) boost = p.vec_boost[q * p.kv + e];
        } else {
            if (!p.allow_keyword_only) continue;
            has_k = true;
            ks = p.kw_scores[q * p.kw + (e - p.kv)];
            if (p.kw_boost) boost = p.kw_boost[q * p.kw + (e - p.kv)];
        }
        const double avail = __dadd_rn(has_v ? p.wv : 0.0, has_k ? p.wk : 0.0);
        if (!(avail > 0.0)) continue;
        const double weighted = __dadd_rn(has_v ? __dmul_rn(p.wv, vs) : 0.0, has_k ? __dmul_rn(p.wk, ks) : 0.0);
        double score = __dmul_rn(__ddiv_rn(weighted, avail), boost);
        if (has_k && !has_v) {
            score = __dmul_rn(score, 0.65);
            if (!p.keyword_filtered && ks < 0.45) continue;
        }
        long long n = 0;
        round6(score, &n);
        if (n < 0) n = 0;
        if (n > (1ll << 21) - 1) n = (1ll << 21) - 1;
        e_vs[e] = vs;
        e_ks[e] = ks;
        // fused score (21 bits) | ~id (32 bits) | entry index (11 bits): fused desc, then id asc
        keys[e] = ((uint64_t)n << 43) | ((uint64_t)(uint32_t)(~(uint32_t)id) << 11) | (uint64_t)e;
    }
    __syncthreads();
    block_bitonic_sort_desc(keys, np);
    for (int i = tid; i < E; i += nt) {
        const uint64_t key = keys[i];
        const size_t o = q * E + i;
        if (key) {
            const int e = (int)(key & 0x7ffu);
            p.out_ids[o] = e_id[e];
            p.out_fused[o] = __ddiv_rn((double)(key >> 43), 1e6);
            p.out_vscore[o] = round6(e_vs[e], nullptr);
            p.out_kscore[o] = round6(e_ks[e], nullptr);
        } else {
            p.out_ids[o] = -1;
            p.out_fused[o] = 0.0;
            p.out_vscore[o] = 0.0;
            p.out_kscore[o] = 0.0;
        }
    }
    if (tid == 0) {
        // keys are sorted descending: the non-zero ones form a prefix
        int lo = 0, hi = E;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid]) lo = mid + 1; else hi = mid;
        }
        p.out_count[q] = lo;
    }
}

// ---- K5b: the numeric part of Searcher._finalize_results on the device --------------------------------------------------
// Per query, over its fused scores in candidate order (best first, as hybrid_fuse_kernel leaves them):
//   dynamic threshold   Searcher._calculate_dynamic_threshold (core/searcher.py:627-674): np.percentile(25 / 75) with
//                       numpy's linear interpolation (a + (b-a) t, or b - (b-a)(1-t) for t >= 1/2), np.median, the
//                       coefficient-of-variation rule, the top_k guard, round(., 6);
//   strict / broad      core/searcher.py:1497-1509 from the round's score floors (_get_round_score_floors, host);
//   bucket per hit      the score-only part of _assign_confidence_bucket (:828-840): 3 reliable, 2 generalised, 1 rest.
// Python floats are doubles: every operation is an explicit round-to-nearest double intrinsic, round(x, 6) is round6().
struct FinalizeParams {
    const double* scores;   // [nq][m]  fused scores, descending
    const int* counts;      // [nq]     valid entries per query
    int m, top_k;
    double strict_floor, broad_floor, threshold_floor;
    double* out_strict;     // [nq]
    double* out_broad;      // [nq]
    int* out_bucket;        // [nq][m]  (0 beyond count)
    int* out_counts;        // [nq][2]  reliable, generalised
};

// numpy's _lerp between the order statistics around virtual index (n-1) q of an ASCENDING view of the scores
__device__ __forceinline__ double np_percentile_desc(const double* desc, int n, double q) {
    const double vi = __dmul_rn((double)(n - 1), q);
    const double pf = floor(vi);
    const int prev = (int)pf;
    const int next = prev + 1 < n ? prev + 1 : n - 1;
    const double gamma = __dadd_rn(vi, -pf);
    const double a = desc[n - 1 - prev], b = desc[n - 1 - next];
    const double diff = __dadd_rn(b, -a);
    if (gamma >= 0.5) return __dadd_rn(b, -__dmul_rn(diff, __dadd_rn(1.0, -gamma)));
    return __dadd_rn(a, __dmul_rn(diff, gamma));
}
// The reference mixes Python floats and numpy float64 scalars (np.percentile / np.median return the latter, and
// arithmetic keeps the numpy type), and `round(x, 6)` means different algorithms for the two: CPython rounds the EXACT
// binary value correctly (round6), numpy computes rint(x * 1e6) / 1e6.  They differ when x * 1e6 rounds onto a tie --
// frequent here, where x is 0.85 x (a score on the 6-digit grid).  So every value carries its type, and max() / min()
// return the OBJECT Python would return (the first argument unless the second is strictly greater / smaller).
struct PyNum {
    double v;
    bool is_np;
};
__device__ __forceinline__ PyNum py_float(double v) { return PyNum{v, false}; }
__device__ __forceinline__ PyNum np_float(double v) { return PyNum{v, true}; }
__device__ __forceinline__ PyNum py_max(PyNum a, PyNum b) { return b.v > a.v ? b : a; }
__device__ __forceinline__ PyNum py_min(PyNum a, PyNum b) { return b.v < a.v ? b : a; }
__device__ __forceinline__ PyNum py_mul(PyNum a, double c) { return PyNum{__dmul_rn(a.v, c), a.is_np}; }
__device__ __forceinline__ PyNum py_add(PyNum a, double c) { return PyNum{__dadd_rn(a.v, c), a.is_np}; }
__device__ __forceinline__ PyNum py_round6(PyNum x) {
    if (x.is_np) return PyNum{__ddiv_rn(rint(__dmul_rn(x.v, 1e6)), 1e6), true};  // np.float64.__round__
    return PyNum{round6(x.v, nullptr), false};                                     // float.__round__
}

__global__ void __launch_bounds__(128) finalize_kernel(const FinalizeParams p) {
    __shared__ double s_strict, s_broad;
    __shared__ int s_cnt[2];
    const int qi = blockIdx.x;
    const double* sc = p.scores + (size_t)qi * p.m;
    int n = p.counts[qi];
    n = n < 0 ? 0 : (n > p.m ? p.m : n);
    if (threadIdx.x == 0) {
        double strict_v = p.strict_floor, broad_v = p.broad_floor;
        if (n > 0) {
            PyNum dyn;
            if (n <= p.top_k * 2) {
                dyn = py_max(py_float(__dmul_rn(sc[n - 1], 0.9)), py_float(p.threshold_floor));  // (not rounded in the reference either)
            } else {
                const PyNum q25 = np_float(np_percentile_desc(sc, n, 0.25)), q75 = np_float(np_percentile_desc(sc, n, 0.75));
                // np.median: mean of the middle element(s)
                const PyNum median = np_float((n & 1) ? sc[n - 1 - n / 2]
                                                      : __ddiv_rn(__dadd_rn(sc[n - 1 - (n / 2 - 1)], sc[n - 1 - n / 2]), 2.0));
                const double cv = median.v > 0 ? __ddiv_rn(__dadd_rn(q75.v, -q25.v), median.v) : 1.0;
                PyNum thr;
                if (cv < 0.2)
                    thr = py_max(py_mul(median, 0.85), py_mul(q25, 0.9));
                else if (cv < 0.5)
                    thr = q25;
                else
                    thr = py_max(py_mul(q25, 0.7), py_mul(median, 0.7));
                if (n >= p.top_k) thr = py_max(thr, py_float(__dmul_rn(sc[p.top_k - 1], 0.8)));
                dyn = py_round6(py_max(thr, py_float(p.threshold_floor)));
            }
            const PyNum strict = py_max(dyn, py_float(p.strict_floor));
            PyNum broad = py_min(py_add(strict, -0.05), py_max(py_float(p.broad_floor), py_mul(strict, 0.84)));
            broad = py_round6(py_max(py_float(p.broad_floor), broad));
            strict_v = strict.v;
            broad_v = broad.v;
        }
        s_strict = strict_v;
        s_broad = broad_v;
        s_cnt[0] = s_cnt[1] = 0;
        p.out_strict[qi] = strict_v;
        p.out_broad[qi] = broad_v;
    }
    __syncthreads();
    const double strict = s_strict, broad = s_broad;
    int rel = 0, gen = 0;
    for (int i = threadIdx.x; i < p.m; i += blockDim.x) {
        int b = 0;
        if (i < n) {
            const double s = sc[i];
            b = s >= strict ? 3 : s >= broad ? 2 : 1;
            rel += b == 3;
            gen += b == 2;
        }
        p.out_bucket[(size_t)qi * p.m + i] = b;
    }
    if (rel) atomicAdd(&s_cnt[0], rel);
    if (gen) atomicAdd(&s_cnt[1], gen);
    __syncthreads();
    if (threadIdx.x == 0) {
        p.out_counts[2 * qi] = s_cnt[0];
        p.out_counts[2 * qi + 1] = s_cnt[1];
    }
}

}  // namespace psx
