// psx_gemm.cuh -- K3: batched queries as a dense contraction on the 5th-generation tensor cores.
//
// The reference issues one FAISS search per query (utils/vector_store.py:190); its multi-round
// drivers (core/searcher.py:1352-1458) and many concurrent users make batches.  For nq >= a
// handful, S = Q . X^T is a GEMM: 2*N*d*nq FLOP over the same N*d*4 bytes, so the corpus is
// streamed ONCE for the whole batch.
//
//   tcgen05.mma kind::tf32, cta_group::1, M = 128 queries per accumulator, N = BN corpus rows,
//   K = 8 per instruction; operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle,
//   K-major for both) through a STAGES-deep mbarrier ring; fp32 accumulators in TMEM, double
//   buffered when they fit, read back with tcgen05.ld by four epilogue warps.
//
// The score matrix never reaches HBM: the epilogue compares every score with a per-query
// threshold theta_q and appends only the survivors' row ids to a small per-query candidate list.
// TF32 keeps 10 mantissa bits, so these scores are approximate; exactness is restored by
//   (1) rescore_select_kernel: exact fp32 dot (same reduction tree as the streaming scan, hence
//       bit-identical scores) of every candidate, exact top-k by integer sort, and
//   (2) a proof obligation per query: the k-th exact score must be >= theta_q + eps, eps = the
//       TF32 error bound.  Then no row outside the list can belong to the top-k.  Queries that
//       fail it (or overflow their list) are flagged and re-run by the streaming scan.
// theta_q comes from a strided sample pass of the same GEMM kernel (MODE_SAMPLE): every (CTA, query)
// keeps its 8 best sample scores in registers, theta_kernel takes the ~16th best of the whole sample.
#pragma once
#include <cuda.h>

#include "psx_common.cuh"

namespace psx {

// One k-block = one 128-byte swizzle row per operand row: 32 fp32 elements (kind::tf32, K = 8 per MMA) or
// 64 bf16 elements (kind::f16, K = 16 per MMA) -- four MMAs per k-block either way, and the same bytes per
// pipeline stage.  BF selects the bf16 form (PSX_STORE_BF16_MASTER: the GEMM streams the bf16 rows, half
// the HBM bytes at twice the tensor rate; exactness still comes from the fp32 re-score).
constexpr int GEMM_KB_BYTES = 128;
constexpr int GEMM_MMAS_PER_KB = 4;
constexpr int GEMM_BK = 32;        // fp32 elements per k-block
template <bool BF>
__host__ __device__ constexpr int gemm_bk() { return BF ? 64 : 32; }
constexpr int GEMM_M = 128;        // queries per accumulator tile
constexpr int GEMM_THREADS = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-7 epilogue
constexpr int GEMM_MODE_SAMPLE = 0, GEMM_MODE_FILTER = 1;

struct GemmParams {
    long long n;             // corpus rows
    int d;                   // logical dimension
    int nq;                  // queries in this batch (<= MT*128)
    int num_tiles;           // ceil(n / BN)
    int tile_step;           // MODE_SAMPLE: visit tiles 0, tile_step, 2*tile_step, ...
    int mode;
    const float* theta;      // [nq] thresholds (MODE_FILTER)
    uint64_t* cand;          // [nq][cand_cap] survivors: orderable(approximate score) << 32 | row
    int* cand_count;         // [nq]
    int cand_cap;
    float* sample_scores;    // [nq][sample_ld] (MODE_SAMPLE): SAMPLE_KEEP scores per sample CTA (pair)
    int sample_ld;
    int sample_cols;         // MODE_SAMPLE: only the first sample_cols rows of a visited tile are sampled (multiple of 32)
    // MODE_FILTER with the threshold sample FOLDED IN (gemm_filter_pair_kernel): every CTA pair's first tile doubles as its
    // sample tile.  Its accumulator stays in TMEM while the grid agrees on the thresholds -- sample scores to memory, grid
    // barrier, each CTA selects theta for its share of the queries, grid barrier -- and is then read a second time, as a
    // filter tile; the MMA warp computes the second tile meanwhile.  No sample kernel, no theta kernel, no recomputation.
    int fold;
    int theta_rank;          // fold: theta = the theta_rank-th largest sample score of a query
    float* theta_out;        // fold: [nq] thresholds, written here
    unsigned int* gbar;      // fold: grid barrier counter (monotonic; the launch owns [gbar_base, gbar_base + 2 * gridDim.x))
    unsigned int gbar_base;
    const uint64_t* attrs;   // EXIF words [n], or nullptr: rows failing `f` are neither sampled nor kept
    psx_filter f;
};

// predicate on packed EXIF words (same conjunction as attr_pass in psx_scan.cuh; this header is included first)
__device__ __forceinline__ bool gemm_attr_pass(uint64_t a, const psx_filter& f) {
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

// Programmatic dependent launch inside one batch: sample pass -> theta_kernel -> filter pass -> rescore_select_kernel are
// launched back to back, each (but the first) allowed to become resident before its predecessor has drained.  Every
// thread of a kernel executes griddepcontrol.wait ("the predecessor has completed, its writes are visible") BEFORE the
// kernel signals its own dependents -- the order the scan kernel uses -- so "this grid has completed" keeps implying
// "every earlier grid has":
//   sample pass   signals at once (it is launched in plain stream order); theta_kernel's CTAs sit in their wait meanwhile;
//   theta_kernel  waits, signals, then sorts: the filter pass launches while the thresholds are being computed;
//   filter pass   its TMA and MMA warps run two tiles ahead before they wait, its epilogue warps wait before they read
//                 the thresholds -- the tensor pipe is busy while theta_kernel still runs; it signals when it is done;
//   rescore       loads its query, waits, re-scores.
__device__ __forceinline__ void gemm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gemm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 (or bf16) inputs, fp32 accumulate
template <bool BF>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (BF) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B tf32 (format 2 of kind::tf32) or bf16 (format 1 of kind::f16), both K-major, M x N
template <bool BF>
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | ((BF ? 1u : 2u) << 7) | ((BF ? 1u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// 32 consecutive fp32 columns of this warp's 32 TMEM lanes -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// r[j] for a run-time j.  Indexing a register array with a run-time value would move the whole array to local memory;
// a switch compiles to a branch table with one move per case.  Called once per survivor (~2 % of the lanes per block).
__device__ __forceinline__ uint32_t pick32(const uint32_t (&r)[32], int j) {
    switch (j) {
#define PSX_PICK(i) case i: return r[i];
        PSX_PICK(0) PSX_PICK(1) PSX_PICK(2) PSX_PICK(3) PSX_PICK(4) PSX_PICK(5) PSX_PICK(6) PSX_PICK(7)
        PSX_PICK(8) PSX_PICK(9) PSX_PICK(10) PSX_PICK(11) PSX_PICK(12) PSX_PICK(13) PSX_PICK(14) PSX_PICK(15)
        PSX_PICK(16) PSX_PICK(17) PSX_PICK(18) PSX_PICK(19) PSX_PICK(20) PSX_PICK(21) PSX_PICK(22) PSX_PICK(23)
        PSX_PICK(24) PSX_PICK(25) PSX_PICK(26) PSX_PICK(27) PSX_PICK(28) PSX_PICK(29) PSX_PICK(30)
#undef PSX_PICK
        default: return r[31];
    }
}
// a survivor of the threshold test: its approximate score travels with the row id (the exact re-score only visits
// the rows whose approximate score can still reach the top-k)
__device__ __forceinline__ uint64_t cand_entry(uint32_t score_bits, uint32_t row) {
    return ((uint64_t)f32_to_ord(__uint_as_float(score_bits)) << 32) | row;
}
__device__ __forceinline__ float cand_score(uint64_t e) { return ord_to_f32((uint32_t)(e >> 32)); }
__device__ __forceinline__ uint32_t cand_row(uint64_t e) { return (uint32_t)e; }

// Survivors are staged per epilogue thread (= per query) in shared memory and appended to the query's global list in runs
// of GEMM_STAGE_N: ONE atomicAdd -- whose reply the thread has to wait for -- per run instead of per survivor.  With ~1400
// listed rows per query (bf16 tier) the per-survivor atomics paced the epilogue, and through it the tensor pipe
// (78 % active; 89 % with a third of the survivors).
constexpr int GEMM_STAGE_N = 8;
constexpr int GEMM_STAGE_BYTES_PER_ACC = 128 * GEMM_STAGE_N * 8;  // 128 epilogue threads per accumulator tile
struct SurvivorStage {
    uint64_t* slot;  // this thread's GEMM_STAGE_N entries in shared memory
    int n;
    __device__ __forceinline__ void flush(uint64_t* list, int* count, int cap) {
        if (n == 0) return;
        const int pos = atomicAdd(count, n);
        for (int i = 0; i < n; ++i)
            if (pos + i < cap) list[pos + i] = slot[i];
        n = 0;
    }
    __device__ __forceinline__ void push(uint64_t entry, uint64_t* list, int* count, int cap) {
        slot[n++] = entry;
        if (n == GEMM_STAGE_N) flush(list, count, cap);
    }
};

// Sample pass: the 8 largest scores one (CTA, query) has seen, kept sorted in registers.  Only these
// reach memory -- the thresholds need the ~16th largest score of the whole sample, and no CTA holds more
// than 8 of the top 16 (probability < 3e-4 even with 9 CTAs; the error only lowers theta, i.e. admits more
// candidates).
constexpr int SAMPLE_KEEP = 8;
struct SampleTop {
    float t[SAMPLE_KEEP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < SAMPLE_KEEP; ++i) t[i] = -INFINITY;
    }
    __device__ __forceinline__ void offer(float v) {
        if (v > t[SAMPLE_KEEP - 1]) {
            t[SAMPLE_KEEP - 1] = v;
#pragma unroll
            for (int i = SAMPLE_KEEP - 1; i > 0; --i) {
                const float hi = fmaxf(t[i - 1], t[i]), lo = fminf(t[i - 1], t[i]);
                t[i - 1] = hi;
                t[i] = lo;
            }
        }
    }
    __device__ __forceinline__ void store(float* dst) const {
#pragma unroll
        for (int i = 0; i < SAMPLE_KEEP; ++i) dst[i] = t[i];
    }
};

// ---- the GEMM + fused selection kernel ---------------------------------------------------------------
// MT accumulator tiles of 128 queries each, BN corpus rows per tile.  TMEM columns: MT*BN per
// accumulator stage, 512 in total.
template <int MT, int BN, int STAGES, bool BF>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const GemmParams p) {
    constexpr int ACC_COLS = MT * BN;
    constexpr int ACC_STAGES = 512 / ACC_COLS >= 2 ? 2 : 1;
    constexpr int A_BYTES = MT * GEMM_M * GEMM_KB_BYTES;
    constexpr int B_BYTES = BN * GEMM_KB_BYTES;
    constexpr int BK = gemm_bk<BF>();
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(ACC_COLS <= 512 && BN % 32 == 0 && BN <= 256, "bad tile");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;  // [STAGES][A | B], every operand tile 1024-byte aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    uint64_t* stage_mem = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES + 256);  // [MT][128][GEMM_STAGE_N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.d + BK - 1) / BK;
    if (p.mode == GEMM_MODE_SAMPLE) gemm_pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);
            mbar_init(smem_u32(empty_bar + s), 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(smem_u32(acc_full + a), 1);
            mbar_init(smem_u32(acc_empty + a), 4);  // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int first = blockIdx.x * p.tile_step;
    const int stride = gridDim.x * p.tile_step;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            bool waited = false;
            for (int t = first; t < p.num_tiles; t += stride) {
                const int row0 = t * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + s), ph ^ 1u);
                    const uint32_t bar = smem_u32(full_bar + s);
                    const uint32_t a_dst = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    mbar_arrive_expect_tx(bar, STAGE_BYTES);
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tma_load_2d(a_dst + m * (GEMM_M * GEMM_KB_BYTES), &map_q, bar, kb * BK, m * GEMM_M);
                    tma_load_2d(a_dst + A_BYTES, &map_x, bar, kb * BK, row0);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (!waited) {  // one tile ahead of the predecessor (it only reads the corpus and the staged queries)
                    gemm_pdl_wait();
                    waited = true;
                }
            }
            if (!waited) gemm_pdl_wait();
        } else {
            gemm_pdl_wait();
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc<BF>(GEMM_M, BN);
            int s = 0, a = 0;
            uint32_t ph = 0, aph = 0;
            bool waited = false;
            for (int t = first; t < p.num_tiles; t += stride) {
                if (a == 1 && !waited) {  // the first accumulator is on its way: now wait for the predecessor
                    gemm_pdl_wait();
                    waited = true;
                }
                mbar_wait(smem_u32(acc_empty + a), aph ^ 1u);  // epilogue drained this accumulator
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t b_desc = umma_smem_desc(a_addr + A_BYTES);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const uint64_t a_desc = umma_smem_desc(a_addr + m * (GEMM_M * GEMM_KB_BYTES));
                        const uint32_t d_addr = tmem_base + (uint32_t)(a * ACC_COLS + m * BN);
#pragma unroll
                        for (int k = 0; k < GEMM_MMAS_PER_KB; ++k) {
                            // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in the
                            // (address >> 4) field of both descriptors
                            umma_ss<BF>(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(smem_u32(empty_bar + s));  // smem slot reusable once these MMAs retire
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit(smem_u32(acc_full + a));  // accumulator complete
                if (++a == ACC_STAGES) {
                    a = 0;
                    aph ^= 1u;
                }
                if (ACC_STAGES == 1 && !waited) {
                    gemm_pdl_wait();
                    waited = true;
                }
            }
            if (!waited) gemm_pdl_wait();
        } else {
            gemm_pdl_wait();
        }
    } else if (warp < 4) {
        gemm_pdl_wait();  // idle warps (2: TMEM allocation, 3: spare)
    } else {
        // ===== epilogue: TMEM -> registers -> threshold test -> candidate lists =====
        const int ew = warp - 4;            // TMEM lanes [32*ew, 32*ew + 32)
        const int qlane = ew * 32 + lane;   // query row inside an accumulator tile
        gemm_pdl_wait();                    // thresholds / zeroed list counters come from the kernel before
        float theta[MT];
        SampleTop top[MT];
        SurvivorStage stage[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int qi = m * GEMM_M + qlane;
            theta[m] = (p.mode == GEMM_MODE_FILTER && qi < p.nq) ? p.theta[qi] : INFINITY;
            top[m].init();
            stage[m].slot = stage_mem + ((size_t)m * GEMM_M + qlane) * GEMM_STAGE_N;
            stage[m].n = 0;
        }
        int a = 0, tile_no = 0;
        uint32_t aph = 0;
        for (int t = first; t < p.num_tiles; t += stride, ++tile_no) {
            mbar_wait(smem_u32(acc_full + a), aph);
            tc_fence_after();
            const long long row0 = (long long)t * BN;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int qi = m * GEMM_M + qlane;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * ACC_COLS + m * BN);
                const int nblocks = p.mode == GEMM_MODE_FILTER ? BN / 32 : p.sample_cols / 32;
#pragma unroll 1
                for (int c = 0; c < nblocks; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    if (p.mode == GEMM_MODE_FILTER) {
                        float mx = __uint_as_float(r[0]);
#pragma unroll
                        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                        if (mx >= theta[m]) {
                            // survivors as a bit mask (straight-line code), then one trip per set bit: a lane has a
                            // survivor in ~2 % of its blocks, but some lane of the warp has one in ~half of them, so the
                            // block must stay cheap for the lanes that only tag along
                            uint32_t hit = 0;
#pragma unroll
                            for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(r[j]) >= theta[m] ? 1u : 0u) << j;
                            while (hit) {
                                const int j = __ffs(hit) - 1;
                                hit &= hit - 1;
                                const long long row = row0 + c * 32 + j;
                                if (row < p.n && (!p.attrs || gemm_attr_pass(__ldg(p.attrs + row), p.f)))
                                    stage[m].push(cand_entry(pick32(r, j), (uint32_t)row), p.cand + (size_t)qi * p.cand_cap, p.cand_count + qi,
                                                  p.cand_cap);
                            }
                        }
                    } else if (qi < p.nq) {
                        // sample pass: only this query's running top-8 of the visited tiles is kept
                        float mx = __uint_as_float(r[0]);
#pragma unroll
                        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                        if (mx > top[m].t[SAMPLE_KEEP - 1]) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const long long row = row0 + c * 32 + j;
                                if (__uint_as_float(r[j]) > top[m].t[SAMPLE_KEEP - 1] && row < p.n &&
                                    (!p.attrs || gemm_attr_pass(__ldg(p.attrs + row), p.f)))
                                    top[m].offer(__uint_as_float(r[j]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1u;
            }
        }
        if (p.mode == GEMM_MODE_SAMPLE) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int qi = m * GEMM_M + qlane;
                if (qi < p.nq) top[m].store(p.sample_scores + (size_t)qi * p.sample_ld + (size_t)blockIdx.x * SAMPLE_KEEP);
            }
        } else {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int qi = m * GEMM_M + qlane;
                stage[m].flush(p.cand + (size_t)qi * p.cand_cap, p.cand_count + qi, p.cand_cap);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.mode == GEMM_MODE_FILTER) gemm_pdl_launch_dependents();  // every thread has waited for the predecessor by now
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- CTA-pair variant (cta_group::2) for 129..256 queries ------------------------------------------------
// Two CTAs of a cluster (= two SMs of a TPC) cooperate on one 256 x 256 tile per step: CTA r holds queries
// [128r, 128r+128) (A rows) and corpus rows [256t + 128r, +128) (half of B); the leader issues ONE
// tcgen05.mma.cta_group::2 (M = 256) that reads both halves and writes each CTA's own 128 TMEM lanes.  Per
// CTA a pipeline stage is 16 KB (A) + 16 KB (B) instead of 32 + 16, so the ring is 6 deep instead of 4 and
// the query block is re-read from L2 once per 256 corpus rows instead of once per 128.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (as CUTLASS' ClusterBarrier::arrive): an explicit .release.cluster here costs ~1500 cycles per
    // call -- it waits for the thread's outstanding TMA issues to become visible cluster-wide
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
template <bool BF>
__device__ __forceinline__ void umma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (BF) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// barrier among the 128 epilogue threads of a CTA (named barrier 1; __syncthreads would involve the TMA / MMA warps)
__device__ __forceinline__ void epilogue_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// One thread per CTA: arrive at the grid barrier and wait until `target` arrivals are visible.  All CTAs of the grid are
// resident (one per SM, grid <= SM count), so the wait ends; it is bounded anyway -- false = gave up.
__device__ __forceinline__ bool grid_barrier(unsigned int* gbar, unsigned int target) {
    __threadfence();
    atomicAdd(gbar, 1u);
    for (unsigned int spin = 0; (int)(ld_acquire_gpu_u32(gbar) - target) < 0; ++spin) {
        __nanosleep(32);
        if (spin > (1u << 24)) return false;
    }
    return true;
}

template <int STAGES, bool BF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_filter_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const GemmParams p) {
    constexpr int BN = 256;              // corpus rows per pair tile (128 per CTA)
    constexpr int ACC_STAGES = 2;        // 2 x 256 TMEM columns
    constexpr int A_BYTES = GEMM_M * GEMM_KB_BYTES;        // this CTA's 128 queries
    constexpr int B_BYTES = (BN / 2) * GEMM_KB_BYTES;      // this CTA's 128 corpus rows
    constexpr int BK = gemm_bk<BF>();
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    uint64_t* stage_mem = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES + 256);  // [128][GEMM_STAGE_N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();
    const bool leader = cta == 0;
    const int kblocks = (p.d + BK - 1) / BK;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    if (p.mode == GEMM_MODE_SAMPLE) gemm_pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(full_bar + s), 2);   // leader: own arrive.expect_tx + the peer's arrive
            mbar_init(smem_u32(empty_bar + s), 1);  // multicast commit of the leader's MMA thread
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(smem_u32(acc_full + a), 1);   // multicast commit
            mbar_init(smem_u32(acc_empty + a), 8);  // four epilogue warps in each of the two CTAs
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int first = pair * p.tile_step;
    const int stride = npairs * p.tile_step;
    // This pair's tiles: first, first + stride, ... -- n_mine of them, visited from position `rot` on (wrapping).  Folded
    // launches rotate every pair's sequence by a different amount so that the FIRST tiles of the pairs (= the threshold
    // sample) are spread evenly over the corpus instead of being its first rows.
    const int n_mine = first < p.num_tiles ? (p.num_tiles - first + stride - 1) / stride : 0;
    const int rot = p.fold && n_mine > 0 ? (int)((long long)pair * n_mine / npairs) : 0;
    auto tile_at = [&](int j) { const int jj = j + rot; return first + stride * (jj >= n_mine ? jj - n_mine : jj); };

    if (warp == 0) {
        // ===== TMA producer (both CTAs; bytes land on the leader's full barrier) =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            bool waited = false;
            for (int j = 0; j < n_mine; ++j) {
                const int row0 = tile_at(j) * BN + (int)cta * (BN / 2);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + s), ph ^ 1u);
                    const uint32_t bar = smem_u32(full_bar + s);
                    const uint32_t a_dst = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    if (leader) mbar_arrive_expect_tx(bar, 2 * STAGE_BYTES);
                    tma_load_2d_2sm(a_dst, &map_q, bar, kb * BK, (int)cta * GEMM_M);
                    tma_load_2d_2sm(a_dst + A_BYTES, &map_x, bar, kb * BK, row0);
                    if (!leader) mbar_arrive_cluster(mapa_shared(bar, 0));
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (!waited) {  // one tile ahead of the predecessor (see gemm_filter_kernel)
                    gemm_pdl_wait();
                    waited = true;
                }
            }
            if (!waited) gemm_pdl_wait();
        } else {
            gemm_pdl_wait();
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc<BF>(2 * GEMM_M, BN);
            int s = 0, a = 0;
            uint32_t ph = 0, aph = 0;
            bool waited = false;
            for (int j = 0; j < n_mine; ++j) {
                if (a == 1 && !waited) {  // the first accumulator is on its way: now wait for the predecessor
                    gemm_pdl_wait();
                    waited = true;
                }
                mbar_wait(smem_u32(acc_empty + a), aph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + s), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    const uint64_t a_desc = umma_smem_desc(a_addr);
                    const uint64_t b_desc = umma_smem_desc(a_addr + A_BYTES);
                    const uint32_t d_addr = tmem_base + (uint32_t)(a * BN);
#pragma unroll
                    for (int k = 0; k < GEMM_MMAS_PER_KB; ++k) umma_ss_2sm<BF>(d_addr, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2sm(smem_u32(empty_bar + s), 3);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit_2sm(smem_u32(acc_full + a), 3);
                if (++a == ACC_STAGES) {
                    a = 0;
                    aph ^= 1u;
                }
            }
            if (!waited) gemm_pdl_wait();
        } else {
            gemm_pdl_wait();
        }
    } else if (warp < 4) {
        gemm_pdl_wait();  // idle warps
    } else {
        // ===== epilogue (each CTA drains its own 128 TMEM lanes = its 128 queries) =====
        const int ew = warp - 4;
        const int etid = ew * 32 + lane;
        const int qi = (int)cta * GEMM_M + etid;
        gemm_pdl_wait();  // thresholds / zeroed list counters come from the kernel before
        float theta = (p.mode == GEMM_MODE_FILTER && !p.fold && qi < p.nq) ? p.theta[qi] : INFINITY;
        SampleTop top;
        top.init();
        SurvivorStage stage;
        stage.slot = stage_mem + (size_t)etid * GEMM_STAGE_N;
        stage.n = 0;
        const uint32_t leader_acc_empty = mapa_shared(smem_u32(acc_empty), 0);
        // one accumulator tile against the threshold (survivors -> the query's list) ...
        auto filter_tile = [&](int t, int a) {
            const long long row0 = (long long)t * BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                float mx = __uint_as_float(r[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                if (mx >= theta) {
                    uint32_t hit = 0;  // see gemm_filter_kernel
#pragma unroll
                    for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(r[j]) >= theta ? 1u : 0u) << j;
                    while (hit) {
                        const int j = __ffs(hit) - 1;
                        hit &= hit - 1;
                        const long long row = row0 + c * 32 + j;
                        if (row < p.n && (!p.attrs || gemm_attr_pass(__ldg(p.attrs + row), p.f)))
                            stage.push(cand_entry(pick32(r, j), (uint32_t)row), p.cand + (size_t)qi * p.cand_cap, p.cand_count + qi, p.cand_cap);
                    }
                }
            }
        };
        // ... or into the query's running top-8 (threshold sample)
        auto sample_tile = [&](int t, int a, int nblocks) {
            const long long row0 = (long long)t * BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * BN);
#pragma unroll 1
            for (int c = 0; c < nblocks; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                if (qi < p.nq) {
                    float mx = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                    if (mx > top.t[SAMPLE_KEEP - 1]) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const long long row = row0 + c * 32 + j;
                            if (__uint_as_float(r[j]) > top.t[SAMPLE_KEEP - 1] && row < p.n &&
                                (!p.attrs || gemm_attr_pass(__ldg(p.attrs + row), p.f)))
                                top.offer(__uint_as_float(r[j]));
                        }
                    }
                }
            }
        };
        int a = 0;
        uint32_t aph = 0;
        int j0 = 0;
        if (p.fold) {
            // ---- the pair's first tile is its share of the threshold sample; its accumulator is kept -----------------
            if (n_mine > 0) {
                mbar_wait(smem_u32(acc_full + 0), 0);
                tc_fence_after();
                sample_tile(tile_at(0), 0, BN / 32);
            }
            if (qi < p.nq) top.store(p.sample_scores + (size_t)qi * p.sample_ld + (size_t)pair * SAMPLE_KEEP);
            __shared__ int s_ok;
            epilogue_bar();
            if (etid == 0) s_ok = grid_barrier(p.gbar, p.gbar_base + gridDim.x) ? 1 : 0;
            epilogue_bar();
            // this CTA selects theta for queries blockIdx.x, blockIdx.x + gridDim.x, ...: the theta_rank-th largest of the
            // npairs * 8 kept sample scores, by bisection on the orderable 32-bit image of the scores (32 rounds of "how
            // many keys are >= mid", each a register count + warp reduction + 4 partial sums: ~1.5 us per query, while the
            // MMA warp is computing the second tile)
            __shared__ int s_cnt[2][4];
            const int ns = npairs * SAMPLE_KEEP;
            constexpr int PER = 8;  // keys per epilogue thread: ns <= 128 * PER (at most 128 pairs)
            for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
                uint32_t key[PER];
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int idx = etid + 128 * i;
                    const float v = idx < ns ? ld_cg_f32(p.sample_scores + (size_t)q * p.sample_ld + idx) : -INFINITY;
                    key[i] = (idx < ns && v > -INFINITY) ? f32_to_ord(v) : 0u;  // 0 = "no sample" (below every real score)
                }
                uint32_t lo = 0u, hi = 0xffffffffu;  // invariant: count(key >= lo) >= rank (lo = 0 counts everything)
                for (int it = 0; it < 32 && lo < hi; ++it) {
                    const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;  // upper middle, so that lo = mid makes progress
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < PER; ++i) c += key[i] >= mid ? 1 : 0;
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == 0) s_cnt[it & 1][ew] = c;
                    epilogue_bar();
                    const int total = s_cnt[it & 1][0] + s_cnt[it & 1][1] + s_cnt[it & 1][2] + s_cnt[it & 1][3];
                    if (total >= p.theta_rank)
                        lo = mid;
                    else
                        hi = mid - 1u;
                }
                if (etid == 0) {
                    // lo = the theta_rank-th largest key; 0 when there are fewer real samples than the rank: take everything
                    p.theta_out[q] = lo ? ord_to_f32(lo) : -INFINITY;
                    // a failed barrier leaves the sample incomplete: overflow the list on purpose -> the query is re-run by the scan
                    p.cand_count[q] = s_ok ? 0 : p.cand_cap + 1;
                }
                epilogue_bar();  // s_cnt is reused by the next query
            }
            if (etid == 0) s_ok = (grid_barrier(p.gbar, p.gbar_base + 2u * gridDim.x) && s_ok) ? 1 : 0;
            epilogue_bar();
            theta = (qi < p.nq && s_ok) ? ld_cg_f32(p.theta_out + qi) : INFINITY;
            // ---- the kept accumulator again, as a filter tile ----------------------------------------------------------
            if (n_mine > 0) {
                filter_tile(tile_at(0), 0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_acc_empty + 0 * 8);
                a = 1;
                j0 = 1;
            }
        }
        for (int j = j0; j < n_mine; ++j) {
            mbar_wait(smem_u32(acc_full + a), aph);
            tc_fence_after();
            if (p.mode == GEMM_MODE_FILTER)
                filter_tile(tile_at(j), a);
            else
                sample_tile(tile_at(j), a, p.sample_cols / 32);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_acc_empty + a * 8);
            if (++a == ACC_STAGES) {
                a = 0;
                aph ^= 1u;
            }
        }
        if (p.mode == GEMM_MODE_SAMPLE && qi < p.nq) top.store(p.sample_scores + (size_t)qi * p.sample_ld + (size_t)pair * SAMPLE_KEEP);
        if (p.mode == GEMM_MODE_FILTER) stage.flush(p.cand + (size_t)qi * p.cand_cap, p.cand_count + qi, p.cand_cap);
    }
    tc_fence_before();
    cluster_sync_all();
    if (p.mode == GEMM_MODE_FILTER) gemm_pdl_launch_dependents();  // every thread has waited for the predecessor by now
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

// ---- theta: per query, the rank-th largest of its sample ------------------------------------------------------
// One CTA per query over the ns = (sample CTAs) x 8 kept scores (ns <= 1184 <= THETA_SORT): all of them are
// sorted and the element at the requested rank is taken.  Approximate by design: the threshold only has to land
// between the k-th and roughly the (cand_cap)-th score -- exactness comes from the proof obligation checked after
// the exact re-score.
constexpr int THETA_SORT = 2048;
__global__ void __launch_bounds__(512) theta_kernel(const float* __restrict__ sample, int sample_ld, int ns, int rank,
                                                    float* __restrict__ theta, int* __restrict__ cand_count) {
    __shared__ uint64_t keys[THETA_SORT];
    gemm_pdl_wait();  // the sample pass has completed
    gemm_pdl_launch_dependents();
    const int qi = blockIdx.x;
    const float* s = sample + (size_t)qi * sample_ld;
    int np = 64;
    while (np < ns) np <<= 1;
    for (int i = threadIdx.x; i < np; i += blockDim.x) keys[i] = i < ns ? make_key(s[i], (uint32_t)i) : 0ull;
    __syncthreads();
    block_bitonic_sort_desc(keys, np);
    if (threadIdx.x == 0) {
        int r = rank < 1 ? 1 : rank;
        if (r > ns) r = ns;
        float v = (r >= 1 && keys[r - 1]) ? key_score(keys[r - 1]) : -INFINITY;
        if (!(v > -INFINITY)) v = -INFINITY;  // fewer samples than the rank: take everything
        theta[qi] = v;
        cand_count[qi] = 0;
    }
}

// ---- exact re-score + selection ------------------------------------------------------------------------------
// One CTA per query.  The list holds every row whose APPROXIMATE score s~ (TF32 / bf16 GEMM, or the bf16 scan) is
// >= theta, together with that score; |s~ - s| <= eps for every row (eps = eps_coef |q| max|x|, computed here).
//   1. sort the list by s~;  tau = the k-th largest s~;
//   2. only the BAND { s~ >= tau - 2 eps } is re-scored exactly: a row below the band has s <= s~ + eps < tau - eps,
//      while the k rows with s~ >= tau have s >= tau - eps -- it cannot be among the k best.  (At k = 100 over 1M
//      1024-d rows the band holds ~170 of the ~460 listed rows: the random 4 KB row reads, which dominated this
//      kernel, shrink by that factor.)
//   3. exact fp32 dot of the band rows with exactly the reduction tree of scan_topk_kernel (per-lane pieces lane,
//      lane+32, ... into four accumulators, (a0+a1)+(a2+a3), xor butterfly), so the scores are bit-identical to the
//      single-query path; integer sort of the exact keys; the first k are emitted.
// flags[qi] != 0  <=>  the result is NOT proven exact and the caller re-runs the query on the scan:
//   1  the list overflowed (survivors were dropped);
//   2  fewer than k rows above a finite threshold;
//   3  rows OUTSIDE the list could matter: neither  tau - 2 eps >= theta  (the band is complete inside the list)
//      nor  k-th exact score >= theta + eps  (every unlisted row has s < theta + eps) holds.
__global__ void __launch_bounds__(512, 2)
rescore_select_kernel(const float* __restrict__ x, int ld, int d, long long n, const float* __restrict__ q, int k, int kpad,
                      const uint64_t* __restrict__ cand, const int* __restrict__ cand_count, int cand_cap,
                      const float* __restrict__ theta, float eps_coef, const float* __restrict__ max_sumsq,
                      const float* __restrict__ eps_dev, uint32_t id_base,
                      float* out_scores, long long* out_ids, uint64_t* out_keys, long long keys_stride, int* flags) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ float s_part[16];
    __shared__ float s_eps;
    __shared__ int s_band;
    const int qi = blockIdx.x;
    const int qpad = (ld + 3) & ~3;
    float* sq = reinterpret_cast<float*>(smem_raw);                       // [qpad]
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + (((size_t)qpad * 4 + 15) & ~(size_t)15));  // [np]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // the query, and its squared norm for the rounding bound (inputs of the whole batch: nothing to wait for)
    float qq = 0.f;
    for (int i = threadIdx.x; i < qpad; i += blockDim.x) {
        const float v = i < d ? q[(size_t)qi * d + i] : 0.f;
        sq[i] = v;
        qq = fmaf(v, v, qq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    if (lane == 0) s_part[warp] = qq;
    gemm_pdl_wait();  // the lists (filter pass / bf16 scan) are complete
    gemm_pdl_launch_dependents();
    const int raw_count = cand_count[qi];
    const int count = raw_count < cand_cap ? raw_count : cand_cap;
    int np = kpad;
    while (np < count) np <<= 1;
    const uint64_t* mine = cand + (size_t)qi * cand_cap;
    for (int i = threadIdx.x; i < np; i += blockDim.x) keys[i] = i < count ? mine[i] : 0ull;
    if (threadIdx.x == 0) s_band = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < nwarps; ++w) tot += s_part[w];
        s_eps = eps_dev ? eps_dev[qi] : eps_coef * sqrtf(tot) * sqrtf(fmaxf(*max_sumsq, 0.f)) * 1.001f;
    }
    // 1. by approximate score
    block_bitonic_sort_desc(keys, np);
    const long long need = n < k ? n : (long long)k;
    const float eps = s_eps;
    const float th = theta[qi];
    float beta = -INFINITY;  // lower edge of the band
    if (need > 0 && count >= need) beta = cand_score(keys[need - 1]) - 2.f * eps;
    // 2. the band is a prefix of the sorted list
    for (int i = threadIdx.x; i < count; i += blockDim.x)
        if (cand_score(keys[i]) >= beta && (i + 1 == count || !(cand_score(keys[i + 1]) >= beta))) s_band = i + 1;
    __syncthreads();
    const int m = s_band;
    // 3. exact scores of the band, two rows per warp and trip (16 warps per query, two queries per SM)
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    const int pieces = ld >> 2;
    for (int c = 2 * warp; c < m; c += 2 * nwarps) {
        const bool two = c + 1 < m;
        const uint32_t r0 = cand_row(keys[c]), r1 = two ? cand_row(keys[c + 1]) : r0;
        const float4* x0 = reinterpret_cast<const float4*>(x + (size_t)r0 * ld);
        const float4* x1 = reinterpret_cast<const float4*>(x + (size_t)r1 * ld);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll 2
        for (int pc = lane; pc < pieces; pc += 32) {
            const float4 v = __ldg(x0 + pc);
            const float4 u = __ldg(x1 + pc);
            const float4 w = q4[pc];
            a0 = fmaf(v.x, w.x, a0);
            a1 = fmaf(v.y, w.y, a1);
            a2 = fmaf(v.z, w.z, a2);
            a3 = fmaf(v.w, w.w, a3);
            b0 = fmaf(u.x, w.x, b0);
            b1 = fmaf(u.y, w.y, b1);
            b2 = fmaf(u.z, w.z, b2);
            b3 = fmaf(u.w, w.w, b3);
        }
        float s0 = (a0 + a1) + (a2 + a3), s1 = (b0 + b1) + (b2 + b3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        __syncwarp();
        if (lane == 0) {
            keys[c] = make_key(s0, id_base + r0);
            if (two) keys[c + 1] = make_key(s1, id_base + r1);
        }
    }
    int np2 = kpad;
    while (np2 < m) np2 <<= 1;
    for (int i = m + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(keys, np2);
    // keys_stride: distance between two queries' key lists (kpad, or world * kpad when the lists of several shards
    // interleave in one buffer -- possibly peer memory on the merging GPU)
    block_emit_results(keys, k, kpad, PSX_METRIC_IP, out_scores ? out_scores + (size_t)qi * k : nullptr,
                       out_ids ? out_ids + (size_t)qi * k : nullptr, out_keys ? out_keys + (size_t)qi * keys_stride : nullptr);
    if (threadIdx.x == 0) {
        int bad = 0;
        const bool everything_listed = !(th > -INFINITY);  // no finite threshold: every eligible row is in the list
        if (raw_count > cand_cap) {
            bad = 1;
        } else if (count < need) {
            bad = everything_listed ? 0 : 2;
        } else if (need > 0 && n > count && !everything_listed) {
            const float kth = key_score(keys[need - 1]);
            if (!(beta >= th) && !(kth >= th + eps)) bad = 3;
        }
        flags[qi] = bad;
    }
}

// bf16-prefilter tier: the k' sorted keys of the bf16 scan -> the candidate list rescore_select_kernel
// consumes (query 0), with theta = the k'-th bf16 score (or -inf when fewer than k' rows exist /
// pass the predicate, in which case every eligible row is already a candidate).
// Also computes the rounding bound for this query: storing x as bf16 (8 significant bits, round to
// nearest) perturbs <q,x> by at most 2^-8 * sum|q_i x_i| <= 2^-8 |q| |x|; 4.1e-3 adds 5 % slack.
__global__ void keys_to_cands_kernel(const uint64_t* __restrict__ keys, int kprime, uint32_t id_base, uint64_t* cand,
                                     int* cand_count, float* theta, const float* __restrict__ q, int d,
                                     const float* __restrict__ max_sumsq, float* eps_out) {
    __shared__ int cnt;
    __shared__ float qsq[32];
    if (threadIdx.x == 0) cnt = 0;
    float acc = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) acc = fmaf(q[i], q[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) qsq[threadIdx.x >> 5] = acc;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < kprime; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key) {
            // keys are sorted: the non-empty ones form a prefix.  Entry = (orderable bf16-scan score, local row)
            cand[i] = (key & 0xffffffff00000000ull) | (uint64_t)(key_id(key) - id_base);
            ++mine;
        }
    }
    atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        cand_count[0] = cnt;
        theta[0] = cnt == kprime ? key_score(keys[kprime - 1]) : -INFINITY;
        float qq = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) qq += qsq[w];
        eps_out[0] = 4.1e-3f * sqrtf(qq) * sqrtf(fmaxf(*max_sumsq, 0.f)) * 1.001f;
    }
}

}  // namespace psx
