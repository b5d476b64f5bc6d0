// psx_scan.cuh -- K1 (streaming exact scan + per-CTA top-k) and K2 (cross-CTA merge), fused in
// one launch.  Replaces FAISS IndexFlat::search for small nq (utils/vector_store.py:191).
//
// Data movement.  Rows are dealt to warps in groups of up to 32 consecutive rows.  Every warp owns
// a private ring of `stages` shared-memory slots of PSX_SLOT_BYTES; a slot receives one window
// (as many consecutive rows as fit, or one 4 KB chunk of a longer row).  For each group the 32
// lanes evaluate the EXIF predicate on the packed attribute words (one coalesced load, prefetched
// two groups ahead) and ballot; windows with no passing row are skipped without touching the
// ring, partially passing windows get one bulk copy per surviving row -- rejected rows never
// leave HBM.  Lane 0 arms the slot's mbarrier with the byte count and issues cp.async.bulk
// (TMA 1-D); the warp waits on the mbarrier, reads the slot with conflict-free 128-bit LDS and
// refills the slot it just drained.  With W warps x S stages x 4 KB per SM, ~19 MB is in flight
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

constexpr int PSX_FUSE_TICKET = 0, PSX_FUSE_DONE = 32, PSX_FUSE_ARRIVED = 64, PSX_FUSE_WORDS = 96;

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
    int rps;        // rows per slot (window), <= 32
    int gsize;      // rows per group = rps * (32 / rps): one predicate ballot per group
    int cpr;        // chunks (slots) per row, > 1 only when rps == 1
    int stages;
    int cand_cap, high_water;
    int sync_every; // CTA-wide overflow check every this many slots per warp (>= 1)
    int metric;
    int has_filter;
    uint32_t id_base;
    psx_filter f;
    // fused cross-GPU exchange (row shards): the last CTA stores this shard's k keys straight into
    // every peer's receive buffer over NVLink and raises a per-source flag there.  world 0 = off.
    const int* cond_flag;  // nullptr, or: run only if *cond_flag != 0 (conditional fallback scan)
    unsigned long long* trace;  // nullptr, or [grid][8] phase timestamps in ns (psx_set_trace_device)
    // PSX_SCAN_DEAL
    const uint32_t* rowlist;    // nullptr = the rows of the arena in order; else the ids of the rows to scan
    unsigned int* list_count;   // number of entries of rowlist (device value, reset by the last CTA)
    unsigned int* work;         // ticket counter of the dynamically dealt tail (0 at launch, reset by the last CTA)
    int static_batch;           // units per statically dealt batch (and cap of a dynamic one)
    int dyn_tail;               // 0 = deal everything statically
    // Row list compacted by THIS launch (no filter_list_kernel before it): phase 1 evaluates the predicate over `attrs`
    // in 256-row chunks dealt by ticket and appends to rowlist / list_count, a counter of finished chunks is the barrier,
    // phase 2 scans the list.  fuse[PSX_FUSE_TICKET] = tickets handed out, [PSX_FUSE_DONE] = a 64-bit word: list entries
    // reserved (low half) and tickets finished (high half), [PSX_FUSE_ARRIVED] = CTAs arrived (all 0 at launch, reset by
    // the last CTA; 32 words apart: one cache line each).
    // nullptr = the list was written by the kernel before this one.
    unsigned int* fuse;
    int fuse_sub;               // blocks of (warps x 256) rows per ticket, 1..4
    int xchg_world, xchg_rank;
    int xchg_targets;      // receive buffers this shard publishes to: xchg_recv[0 .. xchg_targets) (all ranks when one
                           // process per GPU; only the merging device when one process drives every GPU)
    uint32_t xchg_seq;
    uint64_t* xchg_recv[PSX_XCHG_MAX_WORLD];   // peer p: [2 parities][world][PSX_K_PASS_MAX] keys
    uint32_t* xchg_flag[PSX_XCHG_MAX_WORLD];   // peer p: [2 parities][PSX_XCHG_MAX_WORLD] sequence numbers
    // The merging side fused into this launch: after publishing, the last CTA waits (bounded) until every rank's list
    // for this query has landed in THIS rank's receive buffer and selects the global top-k itself -- no second kernel,
    // and the next query's scan (programmatic dependent launch) streams on the other SMs meanwhile.  nullptr = a
    // separate merge_wait_kernel does it (needed when several ranks share one stream: emulated ranks).
    const uint64_t* xchg_my_recv;
    const uint32_t* xchg_my_flag;
    float* xchg_out_scores;
    long long* xchg_out_ids;
    int* xchg_status;                          // host-mapped: 1 + the first rank that never published (0 = fine)
    unsigned long long xchg_spin_limit;
};

__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// diagnostics: phase `i` of this CTA reached (entry, prologue done, stream done, list published, merged, emitted)
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int i) {
    if (trace && threadIdx.x == 0) trace[(size_t)blockIdx.x * 8 + i] = globaltimer_ns();
}

// second block of stamps (phase 1 of a launch that compacts its own row list): ticket known, words evaluated, list space
// reserved, entries written and fenced, every ticket finished
__device__ __forceinline__ void trace_stamp2(unsigned long long* trace, int i) {
    if (trace && threadIdx.x == 0) trace[(size_t)(gridDim.x + blockIdx.x) * 8 + i] = globaltimer_ns();
}

// Programmatic dependent launch.  A scan launched with programmaticStreamSerialization may become resident
// while the scan before it is still finishing.  Everything up to the end of the row stream only reads the
// corpus / the query and uses this launch's own ticket word (two alternate), so it may overlap the previous
// scan's final sorts; pdl_wait() -- "the previous grid has completed and its writes are visible" -- comes
// before the first write to the scratch both launches share (candidate lists, merge tickets, outputs), and
// only then does the CTA let the NEXT launch go, which bounds the overlap to two scans in flight.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// The same predicate without data-dependent branches, for the compaction phase (16 words per lane and trip): the
// equality constraints as one masked compare, the datetime window as one unsigned range test.  Built once per thread
// from the (launch-uniform) filter.
struct AttrTest {
    uint64_t mask, want;  // ((a ^ want) & mask) == 0
    uint64_t lo, span;    // (dt - lo) <= span  (unsigned)
    bool none;            // the constraints contradict the encoding: nothing passes (folded into the two tests)
};
__device__ __forceinline__ AttrTest make_attr_test(const psx_filter& f) {
    constexpr uint64_t DT = (1ull << 39) - 1;
    AttrTest t{0ull, 0ull, 0ull, DT, false};
    const uint32_t fl = f.flags;
    if (fl & (PSX_F_SEASON | PSX_F_PERIOD | PSX_F_YEAR | PSX_F_MONTH)) {
        t.mask |= 1ull << 63;
        t.want |= 1ull << 63;
        if (fl & PSX_F_SEASON) t.mask |= 7ull << 60, t.want |= (uint64_t)(f.season & 7u) << 60, t.none |= f.season > 7u;
        if (fl & PSX_F_PERIOD) t.mask |= 7ull << 57, t.want |= (uint64_t)(f.period & 7u) << 57, t.none |= f.period > 7u;
        if (fl & PSX_F_YEAR) t.mask |= 0x3fffull << 43, t.want |= (uint64_t)(f.year & 0x3fffu) << 43, t.none |= f.year > 0x3fffu;
        if (fl & PSX_F_MONTH) t.mask |= 0xfull << 39, t.want |= (uint64_t)(f.month & 0xfu) << 39, t.none |= f.month > 0xfu;
    }
    if (fl & PSX_F_NEED_DT) {
        uint64_t lo = 1ull, hi = DT;
        if ((fl & PSX_F_START) && f.start > lo) lo = f.start;
        if ((fl & PSX_F_END) && f.end < hi) hi = f.end;
        if (lo > hi) t.none = true;
        t.lo = lo;
        t.span = hi - lo;
    }
    if (t.none) {  // an unsatisfiable pair of tests: the word would need dt = 2^39 - 1 and dt = 1 at once
        t.mask = ~0ull;
        t.want = ~0ull >> 1;
        t.lo = 1ull;
        t.span = 0ull;
    }
    return t;
}
__device__ __forceinline__ bool attr_pass_fast(uint64_t a, const AttrTest& t) {
    return (((a ^ t.want) & t.mask) == 0ull) & (((a & ((1ull << 39) - 1)) - t.lo) <= t.span);
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

// Sums of R rows' lane partials at once.  The xor butterfly of warp_sum leaves every row's total in all
// 32 lanes -- 5 shuffles per row.  Here the first log2(R) steps halve the number of rows a lane carries
// instead (a lane keeps the rows selected by its high lane bits and sends the others to its partner), so
// R rows cost R - 1 + (5 - log2 R) shuffles in total.  Every row still goes through exactly the additions
// of the butterfly (partner pairs at distance 16, 8, 4, 2, 1, in that order), hence bit-identical sums.
// Returns the total of row (lane >> (5 - log2 R)).
template <int N>
__device__ __forceinline__ void reduce_rows_step(float (&s)[N < 1 ? 1 : N], int lane, int o) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float keep = upper ? s[i + N / 2] : s[i];
        const float send = upper ? s[i] : s[i + N / 2];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
}
template <int R>
__device__ __forceinline__ float reduce_rows(float (&s)[R], int lane) {
    static_assert(R == 1 || R == 2 || R == 4 || R == 8, "rows per slot");
    if constexpr (R == 8) {
        reduce_rows_step<8>(s, lane, 16);
        float t4[4] = {s[0], s[1], s[2], s[3]};
        reduce_rows_step<4>(t4, lane, 8);
        float t2[2] = {t4[0], t4[1]};
        reduce_rows_step<2>(t2, lane, 4);
        float v = t2[0];
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
    } else if constexpr (R == 4) {
        reduce_rows_step<4>(s, lane, 16);
        float t2[2] = {s[0], s[1]};
        reduce_rows_step<2>(t2, lane, 8);
        float v = t2[0];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
    } else if constexpr (R == 2) {
        reduce_rows_step<2>(s, lane, 16);
        float v = s[0];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
    } else {
        return warp_sum(s[0]);
    }
}

// Sort the candidate buffer, keep the k best, raise tau.  Block-wide.
__device__ __forceinline__ void compact_candidates(uint64_t* cand, int* s_count, uint64_t* s_tau, int* s_over, int k) {
    const int n = *s_count;
    int np = 2;
    while (np < n) np <<= 1;
    __syncthreads();  // everyone has read n (and s_over)
    for (int i = n + threadIdx.x; i < np; i += blockDim.x) cand[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(cand, np);
    if (threadIdx.x == 0) {
        if (n >= k) {
            *s_tau = cand[k - 1];
            *s_count = k;
        }
        *s_over = 0;
    }
    __syncthreads();
}

// One stored row (or one chunk of a long row) against the query.
//   PPL > 0 : the row/chunk is exactly PPL*32 16-byte pieces, fully unrolled;
//             QREG: the query lives in registers (rows that fit one slot), else in shared memory.
//   PPL == 0: generic strided loop, query in shared memory.
template <typename T, int METRIC, int PPL, bool QREG>
struct RowDot {
    static constexpr int QF4 = (sizeof(T) == 4 ? 1 : 2) * (PPL > 0 ? PPL : 1);
    float4 qreg[QREG ? QF4 : 1];

    __device__ __forceinline__ void load_query(const float4* __restrict__ q4, int lane) {
        if constexpr (QREG) {
#pragma unroll
            for (int t = 0; t < PPL; ++t) {
                if constexpr (sizeof(T) == 4) {
                    qreg[t] = q4[lane + 32 * t];
                } else {
                    qreg[2 * t] = q4[2 * (lane + 32 * t)];
                    qreg[2 * t + 1] = q4[2 * (lane + 32 * t) + 1];
                }
            }
        }
    }
    // xs: first piece of the row/chunk in shared memory; piece0: index of that piece in the row;
    // npieces: pieces in this row/chunk (ignored when PPL > 0)
    __device__ __forceinline__ void accumulate(const uint4* __restrict__ xs, const float4* __restrict__ q4, int piece0,
                                               int npieces, int lane, float (&a)[4]) const {
        if constexpr (PPL > 0) {
            uint4 v[PPL];
#pragma unroll
            for (int t = 0; t < PPL; ++t) v[t] = xs[lane + 32 * t];
#pragma unroll
            for (int t = 0; t < PPL; ++t) {
                if constexpr (QREG) {
                    piece_fma<T, METRIC>(v[t], qreg, t, a);
                } else {
                    piece_fma<T, METRIC>(v[t], q4, piece0 + lane + 32 * t, a);
                }
            }
        } else {
#pragma unroll 4
            for (int pc = lane; pc < npieces; pc += 32) piece_fma<T, METRIC>(xs[pc], q4, piece0 + pc, a);
        }
    }
};

// How the rows of a launch reach the warps.
//   PSX_SCAN_DEAL   : the launch is a range of `units` (one unit = one window = the rows of one ring
//                     slot, or one long row).  Most units are dealt round-robin in small static batches;
//                     the tail of the range is dealt dynamically from a global ticket counter in batches
//                     that shrink towards the end (guided self-scheduling), so that every warp drains at
//                     the same moment whatever the per-SM bandwidth was.  A unit is either `rps`
//                     consecutive rows of the arena, or `rps` consecutive entries of a row-id list (the
//                     rows that pass the EXIF predicate, written by filter_list_kernel).
//   PSX_SCAN_GROUPS : the predicate is evaluated inside the scan, one ballot per group of <= 32 rows
//                     dealt round-robin (attribute words prefetched two groups ahead).
#define PSX_SCAN_DEAL 0
#define PSX_SCAN_GROUPS 1

// Phase 1 of a launch that compacts its own row list (ScanParams::fuse).  Kept out of line: the streaming loop of
// scan_topk_kernel is compiled exactly as it is without this phase.  Block-wide; returns the length of the list, and the
// order in which this CTA arrived (high half of the result).  `scratch` = a few words of shared memory that are idle until the stream starts.
struct CompactArgs {  // the few launch parameters the phase reads (by value: the kernel's parameter block stays in the constant bank)
    const uint64_t* attrs;
    long long n;
    psx_filter f;
    const uint32_t* rowlist;
    unsigned int* fuse;
    int fuse_sub;
    unsigned long long* trace;
};
static __device__ __noinline__ unsigned long long compact_own_list(const CompactArgs p, uint32_t* scratch, uint32_t fuse_ticket,
                                                                   uint32_t fuse_arrival) {
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- phase 1: the predicate as a stream compaction, by the scan's own CTAs -------------------------
    // Tickets of W x 256 rows go to whichever CTA asks next, so the barrier below waits for RUNNING CTAs only: one
    // that becomes resident late (its SM still merges the previous query) finds no ticket left and holds nobody up.
    // One atomic per ticket for the ticket, one for the list space (every counter on a cache line of its own).
    uint32_t* list = const_cast<uint32_t*>(p.rowlist);
    uint32_t* s_f = scratch;  // [0] ticket, [1] list base, [2 + w] offset of warp w (cand is idle until the stream starts)
    // ONE 64-bit word carries both counters: entries reserved in the list (low half) and tickets finished (high
    // half) -- the poll that sees the last ticket finished has the length of the list in the same load
    unsigned long long* fuse_word = reinterpret_cast<unsigned long long*>(p.fuse + PSX_FUSE_DONE);
    // a ticket = `sub` (1..4) consecutive blocks of W x 256 rows, sized by the host so that a corpus of a few
    // million rows is about one ticket per CTA (one round of the latency chain load -> count -> reserve -> write)
    const uint32_t sub = (uint32_t)p.fuse_sub;
    const uint32_t blk = (uint32_t)W << 8;
    const uint32_t per = blk * sub;
    const uint32_t tickets = (uint32_t)((p.n + per - 1) / per);
    if (threadIdx.x == 0) {
        s_f[0] = fuse_ticket;
        s_f[1] = fuse_arrival;
    }
    __syncthreads();
    const AttrTest at = make_attr_test(p.f);
    uint32_t t = s_f[0];
    const uint32_t cta_arrival = s_f[1];
    uint32_t finished = 0;
    __syncthreads();
    trace_stamp2(p.trace, 0);
    while (t < tickets) {
        uint32_t next = 0;
        if (threadIdx.x == 0) next = atomicAdd(p.fuse + PSX_FUSE_TICKET, 1u);
        // lane l of warp w, block b of the ticket: row pairs t*per + b*blk + w*256 + 64 h + 2 l, h = 0..3
        const long long t0 = (long long)t * per + ((long long)warp << 8) + 2 * lane;
        uint32_t bits = 0;  // byte b = the 8 rows of block b
#pragma unroll
        for (int b2 = 0; b2 < 4; b2 += 2) {  // two blocks (8 loads of 16 bytes per lane) in flight at a time
            uint64_t a[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = b2 + (j >> 2), h = j & 3;
                const long long r = t0 + (long long)b * blk + h * 64;
                a[2 * j] = a[2 * j + 1] = 0ull;
                if ((uint32_t)b < sub) {
                    if (r + 1 < p.n) {
                        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p.attrs + r));
                        a[2 * j] = v.x;
                        a[2 * j + 1] = v.y;
                    } else if (r < p.n) {
                        a[2 * j] = __ldg(p.attrs + r);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = b2 + (j >> 2), h = j & 3;
                const long long r = t0 + (long long)b * blk + h * 64;
                if ((uint32_t)b < sub) {
                    // (words past the end of the corpus were loaded as 0 = "no EXIF"; an unconstrained filter never gets here)
                    bits |= (uint32_t)(r < p.n && attr_pass_fast(a[2 * j], at)) << (8 * b + 2 * h);
                    bits |= (uint32_t)(r + 1 < p.n && attr_pass_fast(a[2 * j + 1], at)) << (8 * b + 2 * h + 1);
                }
            }
        }
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_f[2 + warp] = incl;
        __syncthreads();
        if (finished == 0) trace_stamp2(p.trace, 1);
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < W; ++w) {
                const uint32_t c = s_f[2 + w];
                s_f[2 + w] = tot;
                tot += c;
            }
            s_f[1] = tot ? (uint32_t)atomicAdd(fuse_word, (unsigned long long)tot) : 0u;
            s_f[0] = next;
        }
        __syncthreads();
        if (finished == 0) trace_stamp2(p.trace, 2);
        uint32_t pos = s_f[1] + s_f[2 + warp] + incl - mine;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t r = (uint32_t)(t0 + (long long)(j >> 2) * blk + (j & 3) * 64);
            if (bits & (1u << (2 * j))) list[pos++] = r;
            if (bits & (2u << (2 * j))) list[pos++] = r + 1;
        }
        ++finished;
        t = s_f[0];
        __syncthreads();  // s_f is rewritten by the next trip
    }
    __threadfence();  // this thread's list entries, before the CTA reports its tickets
    __syncthreads();
    trace_stamp2(p.trace, 3);
    if (threadIdx.x == 0) {
        if (finished) atomicAdd(fuse_word, (unsigned long long)finished << 32);
        unsigned long long v;
        while ((uint32_t)((v = ld_acquire_gpu_u64(fuse_word)) >> 32) < tickets) __nanosleep(20);
        s_f[0] = (uint32_t)v;
    }
    __syncthreads();
    const uint32_t self_rows = s_f[0];
    __syncthreads();  // (cand is handed to the stream phase)
    trace_stamp2(p.trace, 4);
    return ((unsigned long long)cta_arrival << 32) | self_rows;
}

template <typename T, int METRIC, int PPL, bool QREG, int MODE>
__global__ void __launch_bounds__(PSX_MAX_THREADS, 1) scan_topk_kernel(const ScanParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (p.cond_flag && *p.cond_flag == 0) return;  // uniform: the whole grid skips
    trace_stamp(p.trace, 0);
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    // a launch that compacts its own row list: this CTA's arrival number and first ticket, asked for now so that their
    // round trips run under the prologue
    uint32_t fuse_arrival = 0, fuse_ticket = 0;
    if constexpr (MODE == PSX_SCAN_DEAL) {
        if (p.fuse && threadIdx.x == 0) {
            fuse_arrival = atomicAdd(p.fuse + PSX_FUSE_ARRIVED, 1u);
            fuse_ticket = atomicAdd(p.fuse + PSX_FUSE_TICKET, 1u);
        }
    }

    unsigned char* ring = smem_raw;
    float* sq = reinterpret_cast<float*>(ring + (size_t)W * S * PSX_SLOT_BYTES);
    // q is kept padded to a multiple of 8 floats so bf16 pieces never read past it
    const int qpad = (p.ld + 7) & ~7;
    uint64_t* cand = reinterpret_cast<uint64_t*>(sq + qpad);
    uint64_t* bars = cand + p.cand_cap;
    uint64_t* s_tau = bars + W * S;
    uint32_t* masks = reinterpret_cast<uint32_t*>(s_tau + 1);
    uint32_t* row0s = masks + W * S;
    int* s_count = reinterpret_cast<int*>(row0s + W * S);
    int* s_flag = s_count + 1;
    int* s_over = s_count + 2;
    uint32_t* rowids = reinterpret_cast<uint32_t*>(s_count + 4);  // [W*S][32], list launches only

    // ---- prologue ------------------------------------------------------------------------
    for (int i = threadIdx.x; i < qpad; i += blockDim.x) sq[i] = i < p.d ? p.q[i] : 0.0f;
    if (threadIdx.x == 0) {
        *s_tau = 0ull;
        *s_count = 0;
        *s_flag = 0;
        *s_over = 0;
    }
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(smem_u32(bars + warp * S + s), 1);
    }
    mbar_fence_init();
    __syncthreads();
    trace_stamp(p.trace, 1);

    const uint64_t ceil_key = p.ceil_ptr ? *p.ceil_ptr : ~0ull;
    const int R = p.rps, cpr = p.cpr, row_bytes = p.row_bytes;
    uint32_t Wt = gridDim.x * (uint32_t)W;             // warps the work is dealt to ...
    uint32_t gw = blockIdx.x * (uint32_t)W + warp;     // ... and this warp's number among them

    const uint32_t ring_base = smem_u32(ring) + (uint32_t)(warp * S) * PSX_SLOT_BYTES;
    const uint32_t bar_base = smem_u32(bars + warp * S);
    uint32_t* my_masks = masks + warp * S;
    uint32_t* my_row0s = row0s + warp * S;
    uint32_t* my_rowids = rowids + (size_t)warp * S * 32;
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    RowDot<T, METRIC, PPL, QREG> dot;

    const bool listed = MODE == PSX_SCAN_DEAL && p.rowlist != nullptr;
    const bool self_listed = listed && p.fuse != nullptr;
    // a row-list launch follows the kernel that wrote the list: everything above overlapped it, the list is read below
    if (listed && !self_listed) pdl_wait();
    bool late_warp = false;
    uint32_t self_rows = 0, cta_arrival = 0;  // length of the list this launch compacted itself; this CTA's arrival number
    if constexpr (MODE == PSX_SCAN_DEAL) {
        if (self_listed) {
            const CompactArgs ca{p.attrs, p.n, p.f, p.rowlist, p.fuse, p.fuse_sub, p.trace};
            const unsigned long long packed = compact_own_list(ca, reinterpret_cast<uint32_t*>(cand), fuse_ticket, fuse_arrival);
            self_rows = (uint32_t)packed;
            cta_arrival = (uint32_t)(packed >> 32);
            // ---- phase 2 deals the list to the CTAs in order of ARRIVAL, the last one excepted: that CTA (a late one, or
            // simply the last to get going) takes no fixed share and only helps with the dynamically dealt tail
            // (a list long enough for the dynamically dealt tail absorbs a late CTA by itself: everybody gets a share)
            const uint32_t units = (self_rows + (uint32_t)R - 1u) / (uint32_t)R;
            if (gridDim.x > 1 && !(p.dyn_tail && units >= 64u * Wt)) Wt -= (uint32_t)W;
            gw = cta_arrival * (uint32_t)W + warp;
            late_warp = gw >= Wt;
        }
    }
    dot.load_query(q4, lane);
    bool p_exhausted = false;
    int p_chunk = 0;         // long rows: next chunk of the current row
    int in_flight = 0;
    uint32_t cur_mask = 0;   // window being streamed (bit r = r-th row of the window)
    uint32_t cur_row0 = 0;   // its first row (arena launches) / its row (long rows of a list launch)

    // ---- PSX_SCAN_DEAL producer state --------------------------------------------------------
    uint32_t n_rows = 0, n_units = 0, n_static = 0, bs = 1;  // rows, units, statically dealt units, static batch
    uint32_t s_next = 0;              // next static batch of this warp
    uint32_t u_cur = 0, u_end = 0;    // open batch of units
    uint32_t u_first = 0;             // list launches: unit whose entries sit in lane 0.. of rid_batch
    uint32_t rid_batch = 0;           // list launches: lane l holds list[u_first * R + l]
    uint32_t grab_base = 0, grab_cnt = 0;  // dynamic batch requested ahead of time (base valid in lane 0)
    bool grab_pending = false;
    // list launches, statically dealt part: the ids of this warp's next batches are fetched 32 entries at a time (one
    // independent load per lane) instead of one dependent load per batch -- a short list gives a warp a handful of
    // single-row batches, and their id loads in series were a third of such a query
    uint32_t pf_ids = 0;              // lane e holds entry e of the prefetched run
    uint32_t pf_first = 0;            // static batch index (multiple of Wt past gw) the run starts at
    uint32_t pf_batches = 0, pf_used = 0;  // batches in the run / already opened
    // ---- PSX_SCAN_GROUPS producer state ------------------------------------------------------
    const int G = p.gsize;
    const uint32_t rmask = R >= 32 ? 0xffffffffu : ((1u << R) - 1u);
    const long long num_groups = (p.n + G - 1) / G;
    long long g_next = gw;   // next group to open
    long long g_row0 = 0;    // first row of the open group
    uint32_t g_mask = 0;     // rows of the open group still to be streamed (bit i = row g_row0 + i)
    uint64_t attr_a = 0, attr_b = 0;  // attribute words of the next two groups (prefetched)

    if constexpr (MODE == PSX_SCAN_DEAL) {
        n_rows = self_listed ? self_rows : listed ? __ldcg(p.list_count) : (uint32_t)p.n;
        n_units = (n_rows + (uint32_t)R - 1u) / (uint32_t)R;
        // the dynamically dealt tail: an eighth of the launch, at least 16 and at most 48 units per warp;
        // launches too small for that are dealt statically (their warps finish within one unit of each other)
        uint32_t dyn = 0;
        if (p.dyn_tail && n_units >= 64u * Wt) {
            dyn = n_units >> 3;
            if (dyn < 16u * Wt) dyn = 16u * Wt;
            if (dyn > 48u * Wt) dyn = 48u * Wt;
        }
        n_static = n_units - dyn;
        bs = (uint32_t)p.static_batch;
        while (bs > 1 && n_static / (Wt * bs) < 8) bs >>= 1;
        s_next = gw;
        if (late_warp) s_next = n_static;  // (bs >= 1: beyond every statically dealt batch)
    }
    auto request_grab = [&](uint32_t remaining) {  // ask for the next dynamic batch; the reply is read later
        uint32_t c = remaining / Wt;
        c = c < 1u ? 1u : (c > (uint32_t)p.static_batch ? (uint32_t)p.static_batch : c);
        grab_cnt = c;
        if (lane == 0) grab_base = atomicAdd(p.work, c);
        grab_pending = true;
    };
    auto open_batch = [&]() -> bool {
        const bool was_static = (uint64_t)s_next * bs < n_static;
        if (was_static) {
            u_cur = s_next * bs;
            u_end = u_cur + bs < n_static ? u_cur + bs : n_static;
            s_next += Wt;
            if (n_static < n_units && (uint64_t)s_next * bs >= n_static) request_grab(n_units - n_static);
        } else {
            if (n_static == n_units) return false;
            if (!grab_pending) request_grab(n_units - n_static);  // this warp had no static batch at all
            const uint32_t base = n_static + __shfl_sync(0xffffffffu, grab_base, 0);
            grab_pending = false;
            if (base >= n_units || base < n_static) return false;
            u_cur = base;
            u_end = n_units - base < grab_cnt ? n_units : base + grab_cnt;
            request_grab(n_units - u_end);
        }
        if (listed) {  // the row ids of the whole batch: (u_end - u_cur) * R <= 32 entries
            u_first = u_cur;
            const uint32_t per = bs * (uint32_t)R;  // list entries per static batch (<= 32)
            if (was_static && per <= 16u) {
                if (pf_used == pf_batches) {  // fetch the ids of the next 32 / per static batches of this warp
                    pf_first = u_cur / bs;    // (= the batch index just opened)
                    pf_batches = 32u / per;
                    pf_used = 0;
                    const uint32_t j = (uint32_t)lane / per, o = (uint32_t)lane % per;
                    const uint64_t b = (uint64_t)pf_first + (uint64_t)j * Wt;
                    const uint64_t idx = b * per + o;
                    pf_ids = (j < pf_batches && b * bs < n_static && idx < n_rows && (b * bs + o / (uint32_t)R) < n_static)
                                 ? __ldcg(p.rowlist + idx) : 0u;
                }
                rid_batch = __shfl_sync(0xffffffffu, pf_ids, (pf_used * per + (uint32_t)lane) & 31u);
                ++pf_used;
            } else {
                const uint32_t idx = u_cur * (uint32_t)R + lane;
                rid_batch = (lane < (u_end - u_cur) * (uint32_t)R && idx < n_rows) ? __ldcg(p.rowlist + idx) : 0u;
            }
        }
        return true;
    };

    auto load_attr = [&](long long g) -> uint64_t {
        const long long row = g * G + lane;
        return (p.has_filter && g < num_groups && lane < G && row < p.n) ? __ldg(p.attrs + row) : 0ull;
    };
    if constexpr (MODE == PSX_SCAN_GROUPS) {
        attr_a = load_attr(g_next);
        attr_b = load_attr(g_next + Wt);
    }
    auto open_group = [&]() -> bool {
        if (g_next >= num_groups) return false;
        g_row0 = g_next * G;
        const long long left = p.n - g_row0;
        const int rows = left < G ? (int)left : G;
        if (p.has_filter) {
            g_mask = __ballot_sync(0xffffffffu, lane < rows && attr_pass(attr_a, p.f));
            attr_a = attr_b;
            attr_b = load_attr(g_next + 2 * (long long)Wt);
        } else {
            g_mask = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);
        }
        g_next += Wt;
        return true;
    };

    // Fill `slot` with the next window (or the next chunk of a long row).  false = stream exhausted.
    auto produce = [&](int slot) -> bool {
        uint32_t rid = 0;  // list launches: the row this lane copies
        if (p_chunk == 0) {
            if constexpr (MODE == PSX_SCAN_DEAL) {
                if (u_cur == u_end && !open_batch()) {
                    p_exhausted = true;
                    return false;
                }
                const uint32_t u = u_cur++;
                const uint32_t first = u * (uint32_t)R;  // first row / first list entry of the unit
                const uint32_t cnt = n_rows - first < (uint32_t)R ? n_rows - first : (uint32_t)R;
                cur_mask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
                if (listed) {
                    rid = __shfl_sync(0xffffffffu, rid_batch, ((u - u_first) * (uint32_t)R + lane) & 31);
                    cur_row0 = __shfl_sync(0xffffffffu, rid, 0);
                } else {
                    cur_row0 = first;
                }
            } else {
                while (g_mask == 0) {
                    if (!open_group()) {
                        p_exhausted = true;
                        return false;
                    }
                }
                const int b = __ffs(g_mask) - 1;
                const int w = R == 1 ? b : b / R;
                cur_mask = (g_mask >> (w * R)) & rmask;
                g_mask &= ~(rmask << (w * R));
                cur_row0 = (uint32_t)(g_row0 + (long long)w * R);
            }
        }
        const uint32_t mask = cur_mask;
        const uint32_t bar = bar_base + slot * 8;
        const uint32_t dst = ring_base + (uint32_t)slot * PSX_SLOT_BYTES;
        const unsigned char* src = p.x + (size_t)cur_row0 * row_bytes;
        if (lane == 0) {
            my_masks[slot] = mask;
            my_row0s[slot] = cur_row0;
        }
        if (cpr > 1) {
            const int off = p_chunk * PSX_SLOT_BYTES;
            const int bytes = row_bytes - off < PSX_SLOT_BYTES ? row_bytes - off : PSX_SLOT_BYTES;
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(dst, src + off, bytes, bar);
            }
            if (++p_chunk == cpr) p_chunk = 0;
        } else if (listed) {  // one bulk copy per listed row
            if (lane == 0) mbar_arrive_expect_tx(bar, __popc(mask) * row_bytes);
            if ((mask >> lane) & 1u) my_rowids[slot * 32 + lane] = rid;
            __syncwarp();
            // (evict-first: the listed rows pass through L2 once, the attribute words of the next query's predicate stay)
            if ((mask >> lane) & 1u) bulk_g2s_hint(dst + lane * row_bytes, p.x + (size_t)rid * row_bytes, row_bytes, bar, l2_policy_evict_first());
        } else {
            const int hi = 32 - __clz(mask);  // rows [0, hi) of the window, all passing <=> one copy
            if (mask == (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u))) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, hi * row_bytes);
                    bulk_g2s(dst, src, hi * row_bytes, bar);
                }
            } else {
                if (lane == 0) mbar_arrive_expect_tx(bar, __popc(mask) * row_bytes);
                __syncwarp();
                if ((mask >> lane) & 1u) bulk_g2s(dst + lane * row_bytes, src + (size_t)lane * row_bytes, row_bytes, bar);
            }
        }
        __syncwarp();
        return true;
    };
    for (int s = 0; s < S; ++s) {
        if (!produce(s)) break;
        ++in_flight;
    }

    // ---- consumer ------------------------------------------------------------------------------
    int c_slot = 0;
    uint32_t c_phase = 0;
    auto advance = [&]() {  // slot drained -> refill it with the next window, step the ring
        __syncwarp();
        --in_flight;
        if (!p_exhausted && produce(c_slot)) ++in_flight;
        if (++c_slot == S) {
            c_slot = 0;
            c_phase ^= 1u;
        }
    };
    const int pieces_per_row = row_bytes >> 4;
    constexpr int SLOT_PIECES = PSX_SLOT_BYTES >> 4;
    uint64_t tau = 0ull;
    int over = 0;
    float a[4] = {0.f, 0.f, 0.f, 0.f};  // long rows: carried across the chunks of a row
    int c_chunk = 0;
    int n_compact = 0;
    for (;;) {
        for (int m = 0; m < p.sync_every && in_flight > 0; ++m) {
            const int slot = c_slot;
            mbar_wait(bar_base + slot * 8, c_phase);
            const uint32_t mask = my_masks[slot];
            // global row of this lane's score: consecutive rows of the arena, or the listed row
            const uint32_t myrow = (listed && cpr == 1) ? my_rowids[slot * 32 + lane] : my_row0s[slot] + (uint32_t)lane;
            const uint4* xs = reinterpret_cast<const uint4*>(ring + ((size_t)(warp * S + slot)) * PSX_SLOT_BYTES);
            float myscore = 0.0f;
            bool row_done = true;
            bool fast = false;
            if constexpr (PPL > 0 && QREG) {
                // rows of PPL*32 pieces: a full window is a compile-time number of rows -- no mask walking,
                // all dot products in flight together, one transposed reduction for the whole window
                constexpr int RC = SLOT_PIECES / (PPL * 32);
                constexpr int SH = RC == 8 ? 2 : RC == 4 ? 3 : RC == 2 ? 4 : 5;
                if (mask == (RC >= 32 ? 0xffffffffu : (1u << RC) - 1u)) {
                    float s[RC];
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        float b[4] = {0.f, 0.f, 0.f, 0.f};
                        dot.accumulate(xs + r * (PPL * 32), q4, 0, PPL * 32, lane, b);
                        s[r] = (b[0] + b[1]) + (b[2] + b[3]);
                    }
                    const float v = reduce_rows<RC>(s, lane);
                    myscore = RC == 1 ? v : __shfl_sync(0xffffffffu, v, (lane << SH) & 31);  // lane r takes row r
                    fast = true;
                }
            }
            if (fast) {
            } else if (cpr == 1) {
                uint32_t mm = mask;
                while (mm) {
                    // two rows per trip: their FMA chains and butterflies are independent, which
                    // halves the latency-bound part for short rows (several rows per slot)
                    const int r0 = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const int r1 = mm ? __ffs(mm) - 1 : -1;
                    if (r1 >= 0) mm &= mm - 1;
                    float b[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f};
                    dot.accumulate(xs + (size_t)r0 * pieces_per_row, q4, 0, pieces_per_row, lane, b);
                    if (r1 >= 0) dot.accumulate(xs + (size_t)r1 * pieces_per_row, q4, 0, pieces_per_row, lane, c);
                    float s0 = (b[0] + b[1]) + (b[2] + b[3]), s1 = (c[0] + c[1]) + (c[2] + c[3]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    }
                    if (lane == r0) myscore = s0;
                    if (lane == r1) myscore = s1;
                }
            } else {
                const int piece0 = c_chunk * SLOT_PIECES;
                int np = pieces_per_row - piece0;
                if (np > SLOT_PIECES) np = SLOT_PIECES;
                dot.accumulate(xs, q4, piece0, np, lane, a);
                if (++c_chunk == cpr) {
                    c_chunk = 0;
                    myscore = warp_sum((a[0] + a[1]) + (a[2] + a[3]));
                    a[0] = a[1] = a[2] = a[3] = 0.f;
                } else {
                    row_done = false;
                }
            }
            advance();
            if (!row_done) continue;
            // ---- push survivors --------------------------------------------------------------
            bool want = false;
            uint64_t key = 0;
            if ((mask >> lane) & 1u) {
                const float s = METRIC == PSX_METRIC_L2 ? -myscore : myscore;
                key = make_key(s, p.id_base + myrow);
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
        }
        // Warps run free between checks (they overlap each other's latencies); the buffer has
        // room for sync_every slots of appends per warp above the high-water mark.
        if (over) *s_over = 1;
        over = 0;
        const int alive = __syncthreads_count(in_flight > 0);
        if (*s_over) {
            compact_candidates(cand, s_count, s_tau, s_over, p.k);
            ++n_compact;
        }
        tau = *s_tau;
        if (!alive) break;
    }

    pdl_wait();
    pdl_launch_dependents();
    trace_stamp(p.trace, 2);
    if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)n_compact;
    // ---- publish this CTA's k best -------------------------------------------------------------
    compact_candidates(cand, s_count, s_tau, s_over, p.k);
    {
        const int cnt = *s_count < p.k ? *s_count : p.k;
        uint64_t* mine = p.lists + (size_t)blockIdx.x * p.kpad;
        for (int i = threadIdx.x; i < p.kpad; i += blockDim.x) mine[i] = i < cnt ? cand[i] : 0ull;
    }
    __threadfence();
    __syncthreads();
    trace_stamp(p.trace, 3);
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(p.counter, 1u);
        *s_flag = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!*s_flag) return;

    // ---- last CTA: merge all lists (K2) ---------------------------------------------------------
    __threadfence();
    uint64_t* buf = reinterpret_cast<uint64_t*>(ring);
    block_select_from_lists(p.lists, gridDim.x, p.k, p.kpad, buf, (int)((size_t)W * S * PSX_SLOT_BYTES / 8));
    trace_stamp(p.trace, 4);
    block_emit_results(buf, p.k, p.kpad, p.metric, p.out_scores, p.out_ids, p.out_keys);
    if (threadIdx.x == 0) {
        // every warp of the grid has read the reply to its last ticket request before its CTA took a
        // merge ticket, so both counters are idle now: leave them ready for the next launch
        *p.counter = 0u;
        if (MODE == PSX_SCAN_DEAL) {
            *p.work = 0u;
            if (listed) *p.list_count = 0u;
            if (p.fuse) p.fuse[PSX_FUSE_TICKET] = p.fuse[PSX_FUSE_DONE] = p.fuse[PSX_FUSE_DONE + 1] = p.fuse[PSX_FUSE_ARRIVED] = 0u;
        }
    }
    if (p.xchg_world > 0) {
        // K4 fused: publish this shard's list to every rank (NVLink P2P stores), then the flags
        const int slot = (int)(p.xchg_seq & 1u);
        for (int peer = 0; peer < p.xchg_targets; ++peer) {
            uint64_t* dst = p.xchg_recv[peer] + ((size_t)slot * p.xchg_world + p.xchg_rank) * PSX_K_PASS_MAX;
            for (int i = threadIdx.x; i < p.kpad; i += blockDim.x) st_relaxed_sys_u64(dst + i, i < p.k ? buf[i] : 0ull);
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < p.xchg_targets)
            st_release_sys_u32(p.xchg_flag[threadIdx.x] + slot * PSX_XCHG_MAX_WORLD + p.xchg_rank, p.xchg_seq);
        if (p.xchg_my_recv) {
            // ---- K4 receiving side, inline: wait for `world` lists, select the global top-k ----------------------
            __syncthreads();
            if (threadIdx.x == 0) *s_flag = 0;
            __syncthreads();
            if ((int)threadIdx.x < p.xchg_world) {
                const uint32_t* f = p.xchg_my_flag + slot * PSX_XCHG_MAX_WORLD + threadIdx.x;
                unsigned long long spins = 0;
                while (ld_acquire_sys_u32(f) != p.xchg_seq) {
                    __nanosleep(64);
                    if (++spins > p.xchg_spin_limit) {
                        *s_flag = 1 + (int)threadIdx.x;
                        break;
                    }
                }
            }
            __syncthreads();
            if (*s_flag) {  // a rank never published: report it (the host re-runs the query over the collective path)
                if (threadIdx.x == 0 && p.xchg_status) {
                    *p.xchg_status = *s_flag;
                    __threadfence_system();
                }
            } else {
                const uint64_t* lists = p.xchg_my_recv + (size_t)slot * p.xchg_world * PSX_K_PASS_MAX;
                int np = p.kpad;
                while (np < p.xchg_world * p.kpad) np <<= 1;
                for (int idx = threadIdx.x; idx < np; idx += blockDim.x) {
                    uint64_t v = 0ull;
                    if (idx < p.xchg_world * p.kpad) v = ld_cg_u64(lists + (size_t)(idx / p.kpad) * PSX_K_PASS_MAX + (idx % p.kpad));
                    buf[idx] = v;
                }
                __syncthreads();
                block_bitonic_sort_desc(buf, np);
                block_emit_results(buf, p.k, p.kpad, p.metric, p.xchg_out_scores, p.xchg_out_ids, nullptr);
            }
        }
    }
    __syncthreads();
    trace_stamp(p.trace, 5);
}

}  // namespace psx
