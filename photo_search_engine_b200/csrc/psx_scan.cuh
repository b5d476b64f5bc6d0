// psx_scan.cuh -- K1 (streaming exact scan + per-CTA top-k) and K2 (cross-CTA merge), fused in
// one launch.  Replaces FAISS IndexFlat::search for small nq (utils/vector_store.py:191).
//
// Data movement.  Every consumer warp owns a private ring of `stages` shared-memory slots of
// PSX_SLOT_BYTES.  Lane 0 arms the slot's mbarrier with the byte count and issues one
// cp.async.bulk (TMA 1-D) per slot -- or, when the EXIF predicate rejects some rows of a
// multi-row slot, one bulk copy per surviving row -- so rejected rows never leave HBM.  The
// warp then waits on the mbarrier, reads the slot with conflict-free 128-bit LDS, and refills
// the slot it just drained.  With W warps x S stages x 4 KB per SM, ~24 MB is in flight
// chip-wide, several times bandwidth x latency.
//
// Arithmetic.  fp32 FMA; every row uses the same reduction tree (4 lane-local accumulators
// combined as (a0+a1)+(a2+a3), then a 5-step xor butterfly), so equal rows give bit-equal
// scores regardless of which warp/CTA/GPU scans them -- required by the reference's exact-tie
// test (tests/test_vector_store.py:35-51).
//
// Selection.  Scores become sortable 64-bit keys (psx_common.cuh).  Keys above the CTA's
// running threshold tau are appended to a shared candidate buffer by a warp-aggregated
// atomic; when the buffer passes its high-water mark the CTA bitonic-sorts it, keeps k and
// raises tau to the k-th key.  After the stream each CTA publishes its k best (sorted) and the
// last CTA to finish merges all lists (bitonic merge tree) and emits scores/ids/keys.
#pragma once
#include "psx_common.cuh"

namespace psx {

struct ScanParams {
    const unsigned char* x;   // [n][row_bytes]
    const uint64_t* attrs;    // [n] or nullptr
    const float* q;           // [d] fp32
    const uint64_t* ceil_ptr; // nullptr, or key that every accepted key must be strictly below
    uint64_t* lists;          // [grid][kpad] scratch
    unsigned int* counter;    // ticket for "last CTA merges"
    float* out_scores;        // [k]
    long long* out_ids;       // [k]
    uint64_t* out_keys;       // [kpad] or nullptr
    long long n;
    int d;          // logical dimension
    int ld;         // elements per stored row (zero padded, row_bytes % 16 == 0)
    int row_bytes;
    int k, kpad;
    int rpi;        // rows per item   (item = what one warp consumes between two CTA barriers)
    int cpr;        // chunks (slots) per row, > 1 only when rpi == 1
    int stages;
    int cand_cap, high_water;
    int sync_every; // CTA-wide overflow check every this many iterations (>= 1)
    int metric;
    int has_filter;
    uint32_t id_base;
    psx_filter f;
};

__device__ __forceinline__ bool attr_pass(uint64_t a, const psx_filter& f) {
    const uint32_t fl = f.flags;
    if (fl & (PSX_F_SEASON | PSX_F_PERIOD | PSX_F_YEAR | PSX_F_MONTH)) {
        if (!(a >> 63)) return false;
        if ((fl & PSX_F_SEASON) && (uint32_t)((a >> 60) & 7u) != f.season) return false;
        if ((fl & PSX_F_PERIOD) && (uint32_t)((a >> 57) & 7u) != f.period) return false;
        if ((fl & PSX_F_YEAR) && (uint32_t)((a >> 43) & 0x3fffu) != f.year) return false;
        if ((fl & PSX_F_MONTH) && (uint32_t)((a >> 39) & 0xfu) != f.month) return false;
    }
    if (fl & PSX_F_NEED_DT) {
        const uint64_t dt = a & ((1ull << 39) - 1);
        if (!dt) return false;
        if ((fl & PSX_F_START) && dt < f.start) return false;
        if ((fl & PSX_F_END) && dt > f.end) return false;
    }
    return true;
}

// 16 bytes of a stored row against the matching query elements, accumulated into a[0..3].
template <typename T, int METRIC>
__device__ __forceinline__ void piece_fma(const uint4& raw, const float4* __restrict__ q4, int piece, float (&a)[4]);

template <>
__device__ __forceinline__ void piece_fma<float, PSX_METRIC_IP>(const uint4& raw, const float4* __restrict__ q4, int piece,
                                                                 float (&a)[4]) {
    const float4 q = q4[piece];
    a[0] = fmaf(__uint_as_float(raw.x), q.x, a[0]);
    a[1] = fmaf(__uint_as_float(raw.y), q.y, a[1]);
    a[2] = fmaf(__uint_as_float(raw.z), q.z, a[2]);
    a[3] = fmaf(__uint_as_float(raw.w), q.w, a[3]);
}
template <>
__device__ __forceinline__ void piece_fma<float, PSX_METRIC_L2>(const uint4& raw, const float4* __restrict__ q4, int piece,
                                                                 float (&a)[4]) {
    const float4 q = q4[piece];
    const float d0 = __uint_as_float(raw.x) - q.x, d1 = __uint_as_float(raw.y) - q.y;
    const float d2 = __uint_as_float(raw.z) - q.z, d3 = __uint_as_float(raw.w) - q.w;
    a[0] = fmaf(d0, d0, a[0]);
    a[1] = fmaf(d1, d1, a[1]);
    a[2] = fmaf(d2, d2, a[2]);
    a[3] = fmaf(d3, d3, a[3]);
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <>
__device__ __forceinline__ void piece_fma<__nv_bfloat16, PSX_METRIC_IP>(const uint4& raw, const float4* __restrict__ q4,
                                                                         int piece, float (&a)[4]) {
    const float4 q0 = q4[2 * piece], q1 = q4[2 * piece + 1];
    a[0] = fmaf(bf16lo(raw.x), q0.x, a[0]);
    a[1] = fmaf(bf16hi(raw.x), q0.y, a[1]);
    a[2] = fmaf(bf16lo(raw.y), q0.z, a[2]);
    a[3] = fmaf(bf16hi(raw.y), q0.w, a[3]);
    a[0] = fmaf(bf16lo(raw.z), q1.x, a[0]);
    a[1] = fmaf(bf16hi(raw.z), q1.y, a[1]);
    a[2] = fmaf(bf16lo(raw.w), q1.z, a[2]);
    a[3] = fmaf(bf16hi(raw.w), q1.w, a[3]);
}
template <>
__device__ __forceinline__ void piece_fma<__nv_bfloat16, PSX_METRIC_L2>(const uint4& raw, const float4* __restrict__ q4,
                                                                         int piece, float (&a)[4]) {
    const float4 q0 = q4[2 * piece], q1 = q4[2 * piece + 1];
    float d;
    d = bf16lo(raw.x) - q0.x; a[0] = fmaf(d, d, a[0]);
    d = bf16hi(raw.x) - q0.y; a[1] = fmaf(d, d, a[1]);
    d = bf16lo(raw.y) - q0.z; a[2] = fmaf(d, d, a[2]);
    d = bf16hi(raw.y) - q0.w; a[3] = fmaf(d, d, a[3]);
    d = bf16lo(raw.z) - q1.x; a[0] = fmaf(d, d, a[0]);
    d = bf16hi(raw.z) - q1.y; a[1] = fmaf(d, d, a[1]);
    d = bf16lo(raw.w) - q1.z; a[2] = fmaf(d, d, a[2]);
    d = bf16hi(raw.w) - q1.w; a[3] = fmaf(d, d, a[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sort the candidate buffer, keep the k best, raise tau.  Block-wide.
__device__ __forceinline__ void compact_candidates(uint64_t* cand, int* s_count, uint64_t* s_tau, int k) {
    const int n = *s_count;
    int np = 2;
    while (np < n) np <<= 1;
    __syncthreads();  // everyone has read n
    for (int i = n + threadIdx.x; i < np; i += blockDim.x) cand[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(cand, np);
    if (threadIdx.x == 0) {
        if (n >= k) {
            *s_tau = cand[k - 1];
            *s_count = k;
        }
    }
    __syncthreads();
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(PSX_MAX_THREADS, 1) scan_topk_kernel(const ScanParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;

    unsigned char* ring = smem_raw;
    float* sq = reinterpret_cast<float*>(ring + (size_t)W * S * PSX_SLOT_BYTES);
    // q is kept padded to a multiple of 8 floats so bf16 pieces never read past it
    const int qpad = (p.ld + 7) & ~7;
    uint64_t* cand = reinterpret_cast<uint64_t*>(sq + qpad);
    uint64_t* bars = cand + p.cand_cap;
    uint64_t* s_tau = bars + W * S;
    uint32_t* masks = reinterpret_cast<uint32_t*>(s_tau + 1);
    int* s_count = reinterpret_cast<int*>(masks + W * S);
    int* s_flag = s_count + 1;

    // ---- prologue ------------------------------------------------------------------------
    for (int i = threadIdx.x; i < qpad; i += blockDim.x) sq[i] = i < p.d ? p.q[i] : 0.0f;
    if (threadIdx.x == 0) {
        *s_tau = 0ull;
        *s_count = 0;
        *s_flag = 0;
    }
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(smem_u32(bars + warp * S + s), 1);
    }
    mbar_fence_init();
    __syncthreads();

    const uint64_t ceil_key = p.ceil_ptr ? *p.ceil_ptr : ~0ull;
    const int rpi = p.rpi, cpr = p.cpr, row_bytes = p.row_bytes;
    const long long num_items = (p.n + rpi - 1) / rpi;
    const long long Wt = (long long)gridDim.x * W;
    const long long gw = (long long)blockIdx.x * W + warp;
    const int iters = (int)((num_items + Wt - 1) / Wt);
    const int my_items = num_items > gw ? (int)((num_items - gw + Wt - 1) / Wt) : 0;
    const int total_loads = my_items * cpr;

    const uint32_t ring_base = smem_u32(ring) + (uint32_t)(warp * S) * PSX_SLOT_BYTES;
    const uint32_t bar_base = smem_u32(bars + warp * S);
    uint32_t* my_masks = masks + warp * S;

    // producer state: attribute word of the next item to be issued (prefetched one item ahead)
    uint64_t attr_pf = 0;
    uint32_t issue_mask = 0;
    auto prefetch_attr = [&](int item_no) {
        attr_pf = 0;
        if (p.has_filter && item_no < my_items) {
            const long long row = ((long long)item_no * Wt + gw) * rpi + lane;
            if (lane < rpi && row < p.n) attr_pf = __ldg(p.attrs + row);
        }
    };
    int p_item = 0, p_chunk = 0;  // next load to issue = chunk p_chunk of this warp's item p_item
    auto issue = [&](int slot) {
        const int item_no = p_item, c = p_chunk;
        if (++p_chunk == cpr) {
            p_chunk = 0;
            ++p_item;
        }
        const long long row0 = ((long long)item_no * Wt + gw) * rpi;
        const long long left = p.n - row0;
        const int rows = left < rpi ? (int)left : rpi;
        const uint32_t full = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);
        if (c == 0) {
            if (p.has_filter) {
                const bool ok = lane < rows && attr_pass(attr_pf, p.f);
                issue_mask = __ballot_sync(0xffffffffu, ok);
                prefetch_attr(item_no + 1);
            } else {
                issue_mask = full;
            }
        }
        const uint32_t mask = issue_mask;
        const uint32_t bar = bar_base + slot * 8;
        const uint32_t dst = ring_base + (uint32_t)slot * PSX_SLOT_BYTES;
        const unsigned char* src = p.x + (size_t)row0 * row_bytes;
        if (lane == 0) my_masks[slot] = mask;
        if (mask == 0) {
            if (lane == 0) mbar_arrive(bar);
        } else if (cpr > 1) {
            const int off = c * PSX_SLOT_BYTES;
            const int bytes = row_bytes - off < PSX_SLOT_BYTES ? row_bytes - off : PSX_SLOT_BYTES;
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(dst, src + off, bytes, bar);
            }
        } else if (mask == full) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, rows * row_bytes);
                bulk_g2s(dst, src, rows * row_bytes, bar);
            }
        } else {
            if (lane == 0) mbar_arrive_expect_tx(bar, __popc(mask) * row_bytes);
            __syncwarp();
            if ((mask >> lane) & 1u) bulk_g2s(dst + lane * row_bytes, src + (size_t)lane * row_bytes, row_bytes, bar);
        }
        __syncwarp();
    };

    prefetch_attr(0);
    int issued = total_loads < S ? total_loads : S;
    for (int L = 0; L < issued; ++L) issue(L);
    // consumer state: slot and phase parity of the next load to consume
    int c_slot = 0;
    uint32_t c_phase = 0;
    auto advance = [&]() {  // slot just drained -> refill it with the next load, step the ring
        __syncwarp();
        if (issued < total_loads) {
            issue(c_slot);
            ++issued;
        }
        if (++c_slot == S) {
            c_slot = 0;
            c_phase ^= 1u;
        }
    };

    const float4* q4 = reinterpret_cast<const float4*>(sq);
    const int pieces_per_row = row_bytes >> 4;
    constexpr int QP = sizeof(T) == 4 ? 1 : 2;  // float4 of q per 16-byte piece (documented only)
    (void)QP;

    // ---- main stream -----------------------------------------------------------------------
    uint64_t tau = 0ull;
    int over = 0, since_sync = 0;
    for (int it = 0; it < iters; ++it) {
        float myscore = 0.0f;
        uint32_t item_mask = 0;
        long long row0 = 0;
        if (it < my_items) {
            row0 = ((long long)it * Wt + gw) * rpi;
            if (cpr == 1) {
                const int slot = c_slot;
                mbar_wait(bar_base + slot * 8, c_phase);
                item_mask = my_masks[slot];
                const uint4* xs = reinterpret_cast<const uint4*>(ring + ((size_t)(warp * S + slot)) * PSX_SLOT_BYTES);
                uint32_t m = item_mask;
                while (m) {
                    const int r = __ffs(m) - 1;
                    m &= m - 1;
                    const uint4* xr = xs + (size_t)r * pieces_per_row;
                    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                    for (int pc = lane; pc < pieces_per_row; pc += 32) piece_fma<T, METRIC>(xr[pc], q4, pc, a);
                    const float s = warp_sum((a[0] + a[1]) + (a[2] + a[3]));
                    if (lane == r) myscore = s;
                }
                advance();
            } else {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                for (int c = 0; c < cpr; ++c) {
                    const int slot = c_slot;
                    mbar_wait(bar_base + slot * 8, c_phase);
                    item_mask = my_masks[slot];
                    if (item_mask) {
                        const uint4* xs =
                            reinterpret_cast<const uint4*>(ring + ((size_t)(warp * S + slot)) * PSX_SLOT_BYTES);
                        const int piece0 = c * (PSX_SLOT_BYTES >> 4);
                        int np = pieces_per_row - piece0;
                        if (np > (PSX_SLOT_BYTES >> 4)) np = PSX_SLOT_BYTES >> 4;
#pragma unroll 4
                        for (int pc = lane; pc < np; pc += 32) piece_fma<T, METRIC>(xs[pc], q4, piece0 + pc, a);
                    }
                    advance();
                }
                const float s = warp_sum((a[0] + a[1]) + (a[2] + a[3]));
                if (lane == 0) myscore = s;
            }
        }
        // ---- push survivors ------------------------------------------------------------
        bool want = false;
        uint64_t key = 0;
        if ((item_mask >> lane) & 1u) {
            const float s = METRIC == PSX_METRIC_L2 ? -myscore : myscore;
            key = make_key(s, p.id_base + (uint32_t)(row0 + lane));
            want = key > tau && key < ceil_key;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, want);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(s_count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (want) cand[base + __popc(bal & ((1u << lane) - 1u))] = key;
            over |= base + __popc(bal) > p.high_water;
        }
        // Warps run free between checks (they overlap each other's latencies); the buffer has
        // room for sync_every iterations of appends above the high-water mark.
        if (++since_sync == p.sync_every || it + 1 == iters) {
            since_sync = 0;
            if (__syncthreads_or(over)) compact_candidates(cand, s_count, s_tau, p.k);
            over = 0;
            tau = *s_tau;
        }
    }

    // ---- publish this CTA's k best -------------------------------------------------------------
    compact_candidates(cand, s_count, s_tau, p.k);
    {
        const int cnt = *s_count < p.k ? *s_count : p.k;
        uint64_t* mine = p.lists + (size_t)blockIdx.x * p.kpad;
        for (int i = threadIdx.x; i < p.kpad; i += blockDim.x) mine[i] = i < cnt ? cand[i] : 0ull;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(p.counter, 1u);
        *s_flag = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!*s_flag) return;

    // ---- last CTA: merge all lists (K2) ---------------------------------------------------------
    __threadfence();
    uint64_t* buf = reinterpret_cast<uint64_t*>(ring);
    const int cap_lists = (int)(((size_t)W * S * PSX_SLOT_BYTES / 8) / p.kpad);
    block_merge_lists(p.lists, gridDim.x, p.kpad, buf, cap_lists);
    block_emit_results(buf, p.k, p.kpad, p.metric, p.out_scores, p.out_ids, p.out_keys);
    if (threadIdx.x == 0) *p.counter = 0u;
}

// Standalone merge (K4 final merge of all-gathered shard lists): one CTA per query.
__global__ void __launch_bounds__(256, 1)
merge_keys_kernel(const uint64_t* __restrict__ keys, int nlists, int k, int kpad, int cap_lists, int metric,
                  float* out_scores, long long* out_ids) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
    const size_t qi = blockIdx.x;
    block_merge_lists(keys + qi * (size_t)nlists * kpad, nlists, kpad, buf, cap_lists);
    block_emit_results(buf, k, kpad, metric, out_scores + qi * k, out_ids + qi * k, nullptr);
}

}  // namespace psx
