// psx_api.cu -- host side of the C ABI declared in include/psx.h: HBM arena, staged appends,
// launch configuration, paging for k > PSX_K_PASS_MAX.  No CPU compute path exists here: every
// score is produced by the kernels in psx_scan.cuh.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "psx_aux.cuh"
#include "psx_fuse.cuh"
#include "psx_gemm.cuh"
#include "psx_merge.cuh"
#include "psx_scan_launch.cuh"

using namespace psx;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? PSX_ERR_OOM : PSX_ERR_CUDA, "%s: %s",    \
                        #call, cudaGetErrorString(e_));                                            \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------
// index object
// ------------------------------------------------------------------------------------------
constexpr int BATCH_MAX_Q_ROWS = 256;
struct psx_index {
    int d = 0, ld = 0, metric = 0, dtype = 0, device = 0;
    size_t esize = 4, row_bytes = 0;
    unsigned char* x = nullptr;  // [cap][row_bytes]  the rows the scan streams (fp32 or bf16)
    // PSX_STORE_BF16_MASTER: an fp32 master copy next to the bf16 rows.  The scan streams the bf16 rows
    // (half the bytes) for k' > k candidates, the master re-scores them exactly.
    unsigned char* xm = nullptr;  // [cap][mrow_bytes] fp32
    int ldm = 0;
    size_t mrow_bytes = 0;
    uint64_t* attrs = nullptr;   // [cap]
    bool attrs_set = false;
    std::atomic<long long> n{0};  // rows in HBM (written under `mu`; psx_ntotal reads it without)
    long long cap = 0;
    // rows staged by psx_add, not yet in HBM.  Guarded by `pmu` alone, so that appends (the index build thread,
    // core/indexer.py:858) never queue behind a search that holds `mu` for a whole GPU round trip.  Lock order: mu, then pmu.
    std::vector<float> pending;
    std::atomic<long long> pending_n{0};
    std::mutex pmu;

    cudaStream_t stream = nullptr;
    cudaEvent_t last_ev = nullptr;   // completion of the last search (scratch reuse across streams)
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    int sm_count = 148;
    int warps = 16, stages = 2, ctas_per_sm = 1;
    bool stages_auto = true;

    uint64_t* lists = nullptr;  // [grid][kpad]
    size_t lists_cap = 0;
    unsigned int* counter = nullptr;  // [0] merge tickets, [1] / [3] dynamic-tail tickets of even / odd launches, [2] / [4] entries of the even / odd rowlist,
                                      // [32 + slot * PSX_FUSE_WORDS ..] tickets, finished tickets, arrived CTAs of a scan that compacts the even / odd rowlist itself
    unsigned long long list_seq = 0;  // row lists written so far (alternates the two lists)
    unsigned long long scan_seq = 0;  // launches so far (alternates the dynamic-tail ticket word)
    // Programmatic dependent launch of the scans: 0 = never; 1 = the 2nd, 3rd ... scan of ONE API call overlaps its
    // predecessor (default: the call's first launch is in plain stream order, so whatever produced the queries is
    // complete and visible before any of the call's kernels starts); 2 = also the first scan of a call -- the caller
    // guarantees that the query of a call is never written by the kernel enqueued right before the call on that stream
    // (queries resident on the device, or delivered by a memcpy).
    int pdl = 1;
    bool call_first = true;  // no scan launched yet in the current API call
    uint32_t* rowlist = nullptr;      // [2][cap] ids of the rows that pass the current / the next query's predicate
    long long rowlist_cap = 0;
    bool deal = true;        // unfiltered scans: dealt units with a dynamic tail (false: static predicate groups)
    bool dyn_tail = true;
    int static_batch = 8;
    int filter_mode = 0;     // 0 = auto, 1 = predicate evaluated group by group inside the scan, 2 = compacted into a row list by a kernel
                             // before the scan, 3 = compacted into a row list by the scan's own first phase (auto: 3 up to 4M rows, else 2)
    // device + pinned staging for the host-buffer API
    float* dq = nullptr;
    size_t dq_cap = 0;
    float* dscores = nullptr;
    long long* dids = nullptr;
    uint64_t* dkeys = nullptr;
    size_t dout_cap = 0;
    float* hq = nullptr;
    float* hscores = nullptr;
    long long* hids = nullptr;
    size_t hq_cap = 0, hout_cap = 0;
    // batched (GEMM) path scratch
    float* dmax_sumsq = nullptr;  // device: max ||stored row||^2, maintained by the pack kernel
    float max_norm = 0.f;         // host copy of its square root
    int batch_min = 4;        // smallest nq routed to the tensor-core path
    bool batch_pair = true;   // 129..256 queries: CTA-pair (cta_group::2) kernel instead of two accumulators per CTA
    bool batch_fold = true;   // the threshold sample folded into the filter kernel where the shape allows (tunable "batch_fold")
    unsigned int* gbar = nullptr;     // device: grid-barrier counter of the folded kernel (monotonic)
    unsigned int gbar_epoch = 0;      // host: arrivals consumed by the launches so far
    bool batch_pdl = true;    // the kernels of one batch chain by programmatic dependent launch (tunable "batch_pdl")
    bool batch_bf16 = true;   // PSX_STORE_BF16_MASTER: the batched GEMM reads the bf16 rows (kind::f16) instead of the fp32 master (kind::tf32)
    float* bq = nullptr;      // [256][ld] zero-padded query block
    int bq_dirty_rows = BATCH_MAX_Q_ROWS;  // leading rows of bq that may hold non-zero data (rows beyond nq must read as zero)
    size_t bq_row_bytes = 0;               // row pitch those rows were written with
    // tensor maps of the batched GEMM, re-encoded only when what they describe changes
    struct MapKey {
        const void* base = nullptr;
        long long rows = 0;
        int cols = 0, ld = 0, box_rows = 0, bf16 = -1;
        bool operator==(const MapKey& o) const {
            return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && bf16 == o.bf16;
        }
    };
    MapKey mq_key, mx_key;
    CUtensorMap mq_map, mx_map;
    float* btheta = nullptr;  // [256]
    int* bcount = nullptr;    // [256]
    int* bflags = nullptr;    // [256]
    uint64_t* bcand = nullptr;  // [256][bcand_cap]  (orderable approximate score << 32 | row)
    int bcand_cap = 0;          // entries per query the lists currently have room for
    uint64_t* mkeys = nullptr;  // [PSX_K_PASS_MAX] bf16-prefilter keys (PSX_STORE_BF16_MASTER)
    float* meps = nullptr;      // [1] its rounding bound
    float* bsample = nullptr;
    size_t bsample_cap = 0;
    int* hflags = nullptr;    // pinned [256]
    long long batch_fallbacks = 0, batch_queries = 0, mixed_queries = 0;
    // bulk ingest (load(), add_batch): two pinned staging buffers + two device bounce buffers, so that the host-side copy
    // of chunk i+1 overlaps the H2D transfer and the pack kernel of chunk i.  Allocated by the first large upload.
    float* up_pin[2] = {nullptr, nullptr};
    float* up_dev[2] = {nullptr, nullptr};
    cudaEvent_t up_ev[2] = {nullptr, nullptr};
    size_t up_bytes = 0;
    double up_last_gbps = 0.0;            // throughput of the last large upload (host bytes / wall time)
    unsigned long long* trace = nullptr;  // diagnostics: phase timestamps of the next scans (caller-owned)
    // fused exchange: status word of merge_wait_kernel in host-mapped memory (0 = fine, 1 + r = rank r never published)
    int* xstatus_host = nullptr;
    int* xstatus_dev = nullptr;
    // ---- multi-device group (psx_create_sharded): this handle owns one child index per device entry and no rows itself.
    // Rows are split in contiguous id ranges: child s holds [s*quota, (s+1)*quota) of the rows in HBM, the last child also
    // everything beyond (appends land there until the layout is rebalanced).  Child 0 is the "home": it merges.
    std::vector<psx_index*> shards;
    long long quota = 0;
    long long shard_min_rows = 8192;     // never split a corpus into shards smaller than this (tunable "shard_min_rows")
    unsigned char* gx = nullptr;         // home device: receive buffer of the fused exchange (psx_exchange_bytes())
    uint32_t gseq = 0;
    uint32_t g_synced_seq = 0;           // fused queries up to this sequence number are known complete (host synchronised)
    uint64_t* gkeys = nullptr;           // home device: [nq][active shards][kpad] key lists (batches, paging, mixed tier)
    size_t gkeys_cap = 0;
    cudaEvent_t g_merged[2] = {nullptr, nullptr};  // home stream: the merge of query i (slot i & 1) has completed
    std::vector<cudaEvent_t> g_done;     // per child: its part of the current step is complete
    std::vector<cudaEvent_t> g_h2d;      // per child: its copy out of the query staging is complete
    std::vector<char> g_h2d_busy;
    float* g_hq = nullptr;               // pinned (portable) staging of the queries
    size_t g_hq_cap = 0;
    bool xchg_inline = true;             // tunable "xchg_inline": the scan's last CTA also does the cross-rank merge
    int fault_skip_publish = -1;         // test hook (tunable): this shard never publishes to the fused exchange
    long long xchg_timeout_ms = 0;       // tunable: bounded spin of the fused wait (0 = default, PSX_XCHG_TIMEOUT_MS or 20 s)
    long long g_fused = 0, g_keyed = 0, g_timeouts = 0;  // queries served by the fused exchange / the key-list path / fused timeouts
    std::mutex mu;
};
static inline bool is_group(const psx_index* h) { return !h->shards.empty(); }
#define NOT_ON_GROUP(h, what)                                                                                          \
    do {                                                                                                               \
        if (is_group(h)) return fail(PSX_ERR_STATE, what " is a single-device call; a multi-device handle is driven through the host-buffer API"); \
    } while (0)

static void group_destroy_parts(psx_index* g);
static int group_reset(psx_index* g);
static int group_reserve(psx_index* g, long long n);
static int group_add_device(psx_index* g, const float* x_dev, long long n, int normalize, cudaStream_t st);
static int group_set_attrs(psx_index* g, long long row0, const uint64_t* attrs, long long n, bool from_device, cudaStream_t st);
static int group_read_rows(psx_index* g, long long row0, long long n, float* out);
static int group_set_tunable(psx_index* g, const char* key, int value);

// bounded spin of merge_wait_kernel: ~64 ns sleep + one system-scope load per poll.  PSX_XCHG_TIMEOUT_MS overrides (tests).
static unsigned long long xchg_spin_limit(long long override_ms = 0) {
    static const double env_ms = [] {
        const char* e = getenv("PSX_XCHG_TIMEOUT_MS");
        return e && *e ? atof(e) : 20000.0;
    }();
    const double ms = override_ms > 0 ? (double)override_ms : env_ms;
    return (unsigned long long)std::max(1000.0, ms * 1e6 / 600.0);  // ~0.6 us per poll
}

static inline int scan_dtype(const psx_index* h) { return h->dtype == PSX_STORE_F32 ? PSX_STORE_F32 : PSX_STORE_BF16; }
static inline bool has_fp32_rows(const psx_index* h) { return h->dtype != PSX_STORE_BF16; }
static inline const float* fp32_rows(const psx_index* h) { return (const float*)(h->dtype == PSX_STORE_F32 ? h->x : h->xm); }
static inline int fp32_ld(const psx_index* h) { return h->dtype == PSX_STORE_F32 ? h->ld : h->ldm; }

constexpr int BATCH_MAX_Q = 256;
// survivors per query the candidate lists hold: twice the expected number (T), a power of two in [4096, 16384]
constexpr int BATCH_CAND_CAP_MIN = 4096, BATCH_CAND_CAP_MAX = 16384;
static int batch_cand_cap(int T) {
    int cap = BATCH_CAND_CAP_MIN;
    while (cap < 2 * T && cap < BATCH_CAND_CAP_MAX) cap <<= 1;
    return cap;
}
// Tile shapes (measured on B200, 1M x 1024): up to 128 queries use one accumulator and 256-row corpus tiles
// (TMEM double buffered, half the L2 re-reads of the query block); 129..256 queries use two accumulators
// and 128-row tiles -- 256-row tiles would leave no TMEM for double buffering and lose the MMA/epilogue overlap.
template <int MT>
struct BatchCfg {
    static constexpr int BN = MT == 2 ? 128 : 256;
    static constexpr int STAGES = 4;  // 48 KB per stage either way
};
constexpr int PAIR_STAGES = 6;
static int batch_bn(int mt, bool pair) { return mt == 2 ? (pair ? 256 : BatchCfg<2>::BN) : BatchCfg<1>::BN; }

// launch with (optionally) programmatic stream serialization: the kernel may become resident while its predecessor in
// the stream still runs; it synchronises with it through griddepcontrol.wait where it consumes its results
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cta_group::2 variant for 129..256 queries: grid = 2 * pairs, cluster (2,1,1) is a kernel attribute
template <bool BF>
static int launch_gemm_pair(psx_index* h, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& gp, int pairs, cudaStream_t st,
                            bool pdl) {
    constexpr size_t smem = (size_t)PAIR_STAGES * (GEMM_M + 128) * GEMM_KB_BYTES + 256 + GEMM_STAGE_BYTES_PER_ACC;
    static std::atomic<bool> ready[64];
    if (h->device < 64 && !ready[h->device].load()) {
        CU(cudaFuncSetAttribute(gemm_filter_pair_kernel<PAIR_STAGES, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ready[h->device].store(true);
    }
    CU(launch_ex(gemm_filter_pair_kernel<PAIR_STAGES, BF>, 2u * pairs, GEMM_THREADS, smem, st, pdl, mq, mx, gp));
    g_launches++;
    return PSX_OK;
}

static int pow2ceil(long long v) {
    long long p = 1;
    while (p < v) p <<= 1;
    return (int)p;
}

static int ensure_xstatus(psx_index* h, int** dev) {
    if (!h->xstatus_host) {
        CU(cudaHostAlloc(&h->xstatus_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *h->xstatus_host = 0;
        CU(cudaHostGetDevicePointer(&h->xstatus_dev, h->xstatus_host, 0));
    }
    *dev = h->xstatus_dev;
    return PSX_OK;
}

extern "C" int64_t psx_kpad(int64_t k) {
    if (k < 1) k = 1;
    if (k > PSX_K_PASS_MAX) k = PSX_K_PASS_MAX;
    int p = pow2ceil(k);
    return p < 32 ? 32 : p;
}

extern "C" const char* psx_last_error(void) { return g_err.c_str(); }
extern "C" int psx_abi_version(void) { return PSX_ABI_VERSION; }
extern "C" int64_t psx_launch_count(void) { return g_launches.load(); }

template <typename K>
static int set_max_smem(K kernel) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT));
    return PSX_OK;
}

extern "C" int psx_create(int d, int metric, int store_dtype, int device, psx_index** out) {
    if (!out) return fail(PSX_ERR_INVALID, "out is null");
    *out = nullptr;
    if (d <= 0 || d > 32768) return fail(PSX_ERR_INVALID, "dimension %d out of range [1, 32768]", d);
    if (metric != PSX_METRIC_IP && metric != PSX_METRIC_L2) return fail(PSX_ERR_INVALID, "bad metric %d", metric);
    if (store_dtype != PSX_STORE_F32 && store_dtype != PSX_STORE_BF16 && store_dtype != PSX_STORE_BF16_MASTER)
        return fail(PSX_ERR_INVALID, "bad store dtype %d", store_dtype);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(PSX_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(PSX_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PSX_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    DeviceGuard g(device);
    if (!g.ok) return fail(PSX_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    psx_index* h = new (std::nothrow) psx_index();
    if (!h) return fail(PSX_ERR_OOM, "host allocation failed");
    h->d = d;
    h->metric = metric;
    h->dtype = store_dtype;
    h->device = device;
    h->esize = store_dtype == PSX_STORE_F32 ? 4 : 2;
    const int per16 = 16 / (int)h->esize;
    h->ld = (d + per16 - 1) / per16 * per16;
    h->row_bytes = (size_t)h->ld * h->esize;
    h->ldm = (d + 3) / 4 * 4;
    h->mrow_bytes = store_dtype == PSX_STORE_BF16_MASTER ? (size_t)h->ldm * 4 : 0;
    h->sm_count = prop.multiProcessorCount;
    int rc = PSX_OK;
    auto init = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->last_ev, cudaEventDisableTiming));
        CU(cudaMalloc(&h->counter, (32 + 2 * PSX_FUSE_WORDS) * sizeof(unsigned int)));
        CU(cudaMemset(h->counter, 0, (32 + 2 * PSX_FUSE_WORDS) * sizeof(unsigned int)));
        CU(cudaMalloc(&h->dmax_sumsq, sizeof(float)));
        CU(cudaMemset(h->dmax_sumsq, 0, sizeof(float)));
        return set_max_smem(merge_keys_kernel);
    };
    rc = init();
    if (rc != PSX_OK) {
        psx_destroy(h);
        return rc;
    }
    *out = h;
    return PSX_OK;
}

static void free_all(psx_index* h) {
    cudaFree(h->x);
    cudaFree(h->xm);
    cudaFree(h->attrs);
    cudaFree(h->lists);
    cudaFree(h->counter);
    cudaFree(h->rowlist);
    cudaFree(h->dq);
    cudaFree(h->dscores);
    cudaFree(h->dids);
    cudaFree(h->dkeys);
    cudaFreeHost(h->hq);
    cudaFreeHost(h->hscores);
    cudaFreeHost(h->hids);
    cudaFree(h->dmax_sumsq);
    cudaFree(h->bq);
    cudaFree(h->btheta);
    cudaFree(h->bcount);
    cudaFree(h->bflags);
    cudaFree(h->bcand);
    cudaFree(h->mkeys);
    cudaFree(h->meps);
    cudaFree(h->bsample);
    cudaFree(h->gbar);
    cudaFreeHost(h->hflags);
    cudaFreeHost(h->xstatus_host);
    for (int b = 0; b < 2; ++b) {
        cudaFreeHost(h->up_pin[b]);
        cudaFree(h->up_dev[b]);
        if (h->up_ev[b]) cudaEventDestroy(h->up_ev[b]);
    }
    if (h->last_ev) cudaEventDestroy(h->last_ev);
    if (h->stream) cudaStreamDestroy(h->stream);
}

extern "C" int psx_destroy(psx_index* h) {
    if (!h) return PSX_OK;
    if (is_group(h)) {
        {
            std::lock_guard<std::mutex> lk(h->mu);
            group_destroy_parts(h);
        }
        delete h;
        return PSX_OK;
    }
    {
        // wait for a call still running on another thread (the caller must not START new calls on a handle it destroys)
        std::lock_guard<std::mutex> lk(h->mu);
        DeviceGuard g(h->device);
        cudaDeviceSynchronize();
        free_all(h);
    }
    delete h;
    return PSX_OK;
}

extern "C" int64_t psx_ntotal(const psx_index* h) { return h ? h->n + h->pending_n : 0; }
extern "C" int psx_dim(const psx_index* h) { return h ? h->d : 0; }
extern "C" int psx_metric(const psx_index* h) { return h ? h->metric : 0; }

extern "C" int psx_reset(psx_index* h) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (is_group(h)) return group_reset(h);
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->x);
    cudaFree(h->xm);
    cudaFree(h->attrs);
    h->x = nullptr;
    h->xm = nullptr;
    h->attrs = nullptr;
    h->attrs_set = false;
    cudaMemset(h->dmax_sumsq, 0, sizeof(float));
    h->max_norm = 0.f;
    h->n = 0;
    h->cap = 0;
    {
        std::lock_guard<std::mutex> pk(h->pmu);
        h->pending.clear();
        h->pending.shrink_to_fit();
        h->pending_n = 0;
    }
    return PSX_OK;
}

// grow the arena to hold at least `need` rows (amortised doubling, device-to-device move)
static int ensure_capacity(psx_index* h, long long need, bool exact) {
    if (need <= h->cap) return PSX_OK;
    long long ncap = need;
    if (!exact) {
        long long dbl = h->cap * 2;
        if (dbl > ncap) ncap = dbl;
        if (ncap < 1024) ncap = 1024;
    }
    unsigned char* nx = nullptr;
    uint64_t* na = nullptr;
    cudaError_t e = cudaMalloc(&nx, (size_t)ncap * h->row_bytes);
    if (e != cudaSuccess && !exact && ncap > need) {  // retry without the growth slack
        cudaGetLastError();
        ncap = need;
        e = cudaMalloc(&nx, (size_t)ncap * h->row_bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PSX_ERR_OOM, "cudaMalloc of %lld rows x %zu B failed: %s", ncap, h->row_bytes, cudaGetErrorString(e));
    }
    e = cudaMalloc(&na, (size_t)ncap * sizeof(uint64_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(nx);
        return fail(PSX_ERR_OOM, "cudaMalloc of attribute words failed: %s", cudaGetErrorString(e));
    }
    unsigned char* nm = nullptr;
    if (h->mrow_bytes) {
        e = cudaMalloc(&nm, (size_t)ncap * h->mrow_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(nx);
            cudaFree(na);
            return fail(PSX_ERR_OOM, "cudaMalloc of the fp32 master (%lld rows) failed: %s", ncap, cudaGetErrorString(e));
        }
    }
    CU(cudaMemsetAsync(na, 0, (size_t)ncap * sizeof(uint64_t), h->stream));
    if (h->n > 0) {
        CU(cudaMemcpyAsync(nx, h->x, (size_t)h->n * h->row_bytes, cudaMemcpyDeviceToDevice, h->stream));
        CU(cudaMemcpyAsync(na, h->attrs, (size_t)h->n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, h->stream));
        if (nm) CU(cudaMemcpyAsync(nm, h->xm, (size_t)h->n * h->mrow_bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    // a search enqueued earlier on a caller's stream may still be reading the old arena
    if (h->has_last) CU(cudaEventSynchronize(h->last_ev));
    cudaFree(h->x);
    cudaFree(h->xm);
    cudaFree(h->attrs);
    h->xm = nm;
    h->x = nx;
    h->attrs = na;
    h->cap = ncap;
    return PSX_OK;
}

static int launch_pack(psx_index* h, const float* src_dev, long long row0, long long n, int normalize, cudaStream_t st) {
    if (n <= 0) return PSX_OK;
    long long blocks = (n + 7) / 8;  // 8 warps per block
    if (blocks > (long long)h->sm_count * 16) blocks = (long long)h->sm_count * 16;
    unsigned char* dst = h->x + (size_t)row0 * h->row_bytes;
    if (h->dtype == PSX_STORE_F32)
        pack_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(src_dev, (float*)dst, n, h->d, h->ld, normalize, h->dmax_sumsq);
    else
        pack_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(src_dev, (__nv_bfloat16*)dst, n, h->d, h->ld, normalize,
                                                                        h->dmax_sumsq);
    g_launches++;
    if (h->xm) {
        pack_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(src_dev, (float*)(h->xm + (size_t)row0 * h->mrow_bytes), n, h->d, h->ldm,
                                                                 normalize, h->dmax_sumsq);
        g_launches++;
    }
    h->max_norm = 0.f;  // re-read lazily
    CU(cudaGetLastError());
    return PSX_OK;
}

// host -> pinned staging with several threads (one thread tops out near 10 GB/s; the PCIe link takes ~50)
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min(8u, hw / 2 ? hw / 2 : 1u), bytes >> 22));
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t off = (size_t)t * per;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        pool.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    for (std::thread& t : pool) t.join();
}

constexpr size_t UPLOAD_CHUNK_BYTES = 64ull << 20;

// upload host rows into the arena behind row h->n; the caller holds h->mu (or owns h as the child of a group) and the
// device is current.  Small uploads (add_item trickles) go through one bounded bounce buffer; large ones are pipelined:
// chunk i+1 is copied into pinned staging by several host threads while chunk i crosses PCIe and is packed.
static int upload_host_rows(psx_index* h, const float* rows, long long rows_n) {
    if (rows_n <= 0) return PSX_OK;
    int rc = ensure_capacity(h, h->n + rows_n, false);
    if (rc) return rc;
    const size_t row_bytes = (size_t)h->d * sizeof(float);
    const size_t total = (size_t)rows_n * row_bytes;
    if (total < (8ull << 20)) {
        float* bounce = nullptr;
        if (cudaMalloc(&bounce, total) != cudaSuccess) {
            cudaGetLastError();
            return fail(PSX_ERR_OOM, "cudaMalloc of the upload bounce buffer failed");
        }
        cudaError_t e = cudaMemcpyAsync(bounce, rows, total, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) {
            rc = launch_pack(h, bounce, h->n, rows_n, 0, h->stream);
            if (rc == PSX_OK) e = cudaStreamSynchronize(h->stream);
        }
        cudaFree(bounce);
        if (e != cudaSuccess || rc != PSX_OK) return rc ? rc : fail(PSX_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
        return PSX_OK;
    }
    if (!h->up_bytes) {
        for (int b = 0; b < 2; ++b) {
            cudaError_t e = cudaHostAlloc(&h->up_pin[b], UPLOAD_CHUNK_BYTES, cudaHostAllocPortable);
            if (e == cudaSuccess) e = cudaMalloc(&h->up_dev[b], UPLOAD_CHUNK_BYTES);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->up_ev[b], cudaEventDisableTiming);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(PSX_ERR_OOM, "allocating the upload staging failed: %s", cudaGetErrorString(e));
            }
        }
        h->up_bytes = UPLOAD_CHUNK_BYTES;
    }
    const long long chunk_rows = std::max<long long>(1, (long long)(h->up_bytes / row_bytes));
    const auto t0 = std::chrono::steady_clock::now();
    long long done = 0;
    for (int i = 0; done < rows_n; ++i) {
        const int b = i & 1;
        const long long m = std::min(chunk_rows, rows_n - done);
        if (i >= 2) CU(cudaEventSynchronize(h->up_ev[b]));  // the transfer that last used this staging pair is over
        parallel_memcpy(h->up_pin[b], rows + (size_t)done * h->d, (size_t)m * row_bytes);
        CU(cudaMemcpyAsync(h->up_dev[b], h->up_pin[b], (size_t)m * row_bytes, cudaMemcpyHostToDevice, h->stream));
        if ((rc = launch_pack(h, h->up_dev[b], h->n + done, m, 0, h->stream))) return rc;
        CU(cudaEventRecord(h->up_ev[b], h->stream));
        done += m;
    }
    CU(cudaStreamSynchronize(h->stream));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sec > 0) h->up_last_gbps = (double)total / sec / 1e9;
    return PSX_OK;
}

static int group_place_host_rows(psx_index* g, const float* rows, long long rows_n);

// move everything psx_add staged into HBM (single device: behind the arena's last row; group: to the tail shard(s))
static int flush_pending(psx_index* h) {
    if (h->pending_n == 0) return PSX_OK;
    // take the staged rows; appends arriving from now on start a fresh staging vector
    std::vector<float> rows;
    long long rows_n = 0;
    {
        std::lock_guard<std::mutex> pk(h->pmu);
        rows.swap(h->pending);
        rows_n = h->pending_n;
    }
    const int rc = is_group(h) ? group_place_host_rows(h, rows.data(), rows_n) : upload_host_rows(h, rows.data(), rows_n);
    std::lock_guard<std::mutex> pk(h->pmu);
    if (rc) {  // failure: the rows stay staged (in order, in front of whatever arrived meanwhile)
        rows.insert(rows.end(), h->pending.begin(), h->pending.end());
        h->pending.swap(rows);
        return rc;
    }
    // n and pending_n move together under pmu: psx_ntotal never sees the rows twice or not at all
    h->n += rows_n;  // (a group's n is the total over its children)
    h->pending_n -= rows_n;
    return PSX_OK;
}

extern "C" int psx_add(psx_index* h, const float* x, int64_t n) {
    if (!h || (!x && n > 0) || n < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_add");
    // large blocks (load(), add_batch) go straight to HBM through the pipelined upload -- no copy into the staging vector
    if ((size_t)n * h->d * sizeof(float) >= (32ull << 20)) {
        std::lock_guard<std::mutex> lk(h->mu);
        DeviceGuard g(h->device);
        int rc = flush_pending(h);
        if (rc) return rc;
        if ((unsigned long long)(h->n + n) >= 0xffffffffull) return fail(PSX_ERR_RANGE, "more than 2^32-1 rows");
        rc = is_group(h) ? group_place_host_rows(h, x, n) : upload_host_rows(h, x, n);
        if (rc) return rc;
        std::lock_guard<std::mutex> pk(h->pmu);
        h->n += n;
        return PSX_OK;
    }
    bool big = false;
    {
        std::lock_guard<std::mutex> pk(h->pmu);
        if ((unsigned long long)(h->n + h->pending_n + n) >= 0xffffffffull) return fail(PSX_ERR_RANGE, "more than 2^32-1 rows");
        try {
            h->pending.insert(h->pending.end(), x, x + (size_t)n * h->d);
        } catch (const std::bad_alloc&) {
            return fail(PSX_ERR_OOM, "host staging allocation failed");
        }
        h->pending_n += n;
        big = (size_t)h->pending_n * h->d * sizeof(float) >= (512ull << 20);
    }
    // keep the host staging bounded: big batches go to HBM right away
    if (big) {
        std::lock_guard<std::mutex> lk(h->mu);
        DeviceGuard g(h->device);
        return flush_pending(h);
    }
    return PSX_OK;
}

extern "C" int psx_sync(psx_index* h) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    return flush_pending(h);
}

extern "C" int psx_reserve(psx_index* h, int64_t n) {
    if (!h || n < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_reserve");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    if (is_group(h)) {
        int rc = flush_pending(h);
        return rc ? rc : group_reserve(h, n);
    }
    return ensure_capacity(h, n, true);
}

extern "C" int psx_add_device(psx_index* h, const float* x_dev, int64_t n, int normalize, void* stream) {
    if (!h || (!x_dev && n > 0) || n < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_add_device");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (is_group(h)) return group_add_device(h, x_dev, n, normalize, (cudaStream_t)stream);
    if ((unsigned long long)(h->n + n) >= 0xffffffffull) return fail(PSX_ERR_RANGE, "more than 2^32-1 rows");
    rc = ensure_capacity(h, h->n + n, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_pack(h, x_dev, h->n, n, normalize, st);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    h->n += n;
    return PSX_OK;
}

extern "C" int psx_set_attrs(psx_index* h, int64_t row0, const uint64_t* attrs, int64_t n) {
    if (!h || !attrs || n < 0 || row0 < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_set_attrs");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (is_group(h)) return group_set_attrs(h, row0, attrs, n, false, nullptr);
    if (row0 + n > h->n) return fail(PSX_ERR_RANGE, "attribute rows [%lld,%lld) exceed ntotal %lld", (long long)row0,
                                     (long long)(row0 + n), h->n.load());
    if (n == 0) return PSX_OK;
    CU(cudaMemcpyAsync(h->attrs + row0, attrs, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->attrs_set = true;
    return PSX_OK;
}

extern "C" int psx_set_attrs_device(psx_index* h, int64_t row0, const uint64_t* attrs_dev, int64_t n, void* stream) {
    if (!h || !attrs_dev || n < 0 || row0 < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_set_attrs_device");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (is_group(h)) return group_set_attrs(h, row0, attrs_dev, n, true, (cudaStream_t)stream);
    if (row0 + n > h->n) return fail(PSX_ERR_RANGE, "attribute rows exceed ntotal");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(h->attrs + row0, attrs_dev, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    h->attrs_set = true;
    return PSX_OK;
}

// ------------------------------------------------------------------------------------------
// scan launch
// ------------------------------------------------------------------------------------------
struct ScanPlan {
    ScanParams p;
    int grid, block, mode;
    bool pdl;
    size_t smem;
};

static int plan_scan_with(psx_index* h, bool master, int k, int W, int mode, bool listed, ScanPlan& plan);

// Very long rows (d up to 32768: a 128 KB query block in shared memory) do not leave room for 16 warp
// rings: retry with fewer warps before giving up.
static int plan_scan(psx_index* h, bool master, int k, int mode, bool listed, ScanPlan& plan) {
    int rc = PSX_ERR_INVALID;
    for (int W = h->warps; W >= 2; W >>= 1) {
        rc = plan_scan_with(h, master, k, W, mode, listed, plan);
        if (rc == PSX_OK) return rc;
    }
    return rc;
}

static int plan_scan_with(psx_index* h, bool master, int k, int W, int mode, bool listed, ScanPlan& plan) {
    ScanParams& p = plan.p;
    memset(&p, 0, sizeof p);
    plan.mode = mode;
    // which arena is streamed: the scan rows, or the fp32 master of a PSX_STORE_BF16_MASTER index
    struct { int ld; size_t row_bytes; } a = {master ? h->ldm : h->ld, master ? h->mrow_bytes : h->row_bytes};
    p.n = h->n;
    p.d = h->d;
    p.ld = a.ld;
    p.row_bytes = (int)a.row_bytes;
    p.k = k;
    p.kpad = (int)psx_kpad(k);
    p.metric = h->metric;
    if (a.row_bytes <= PSX_SLOT_BYTES) {
        p.rps = (int)std::min<size_t>(32, PSX_SLOT_BYTES / a.row_bytes);
        p.cpr = 1;
    } else {
        p.rps = 1;
        p.cpr = (int)((a.row_bytes + PSX_SLOT_BYTES - 1) / PSX_SLOT_BYTES);
    }
    // One predicate ballot covers a group of up to 32 rows, and groups are dealt round-robin to all
    // warps of the grid.  A lone warp streams slowly (latency bound), so the kernel's tail is one
    // group: keep groups small enough that every warp gets >= 256 of them (imbalance < 0.4 %), but
    // not below 4 windows unless the corpus is so small that warps would otherwise stay idle.
    int wpg = 32 / p.rps;  // windows per group
    const long long all_warps = (long long)h->sm_count * h->ctas_per_sm * W;
    auto groups_for = [&](int w) { return (h->n + (long long)p.rps * w - 1) / ((long long)p.rps * w); };
    while (wpg > 4 && groups_for(wpg) < 256 * all_warps) wpg >>= 1;
    while (wpg > 1 && groups_for(wpg) < all_warps) wpg >>= 1;
    p.gsize = p.rps * wpg;
    // CTA-wide overflow checks are spaced so that a few hundred appends fit between two of them
    p.sync_every = std::max(1, std::min(8, 256 / (W * p.rps)));
    const int burst = W * p.rps * p.sync_every;  // most keys a CTA can append between two checks
    int cap = pow2ceil((long long)k + burst);
    if (cap - burst - k < std::max(burst, 64)) cap <<= 1;
    if (cap < 1024) cap = 1024;
    p.cand_cap = cap;
    p.high_water = cap - burst;
    const int qpad = (a.ld + 7) & ~7;
    auto smem_for = [&](int S) {
        return (size_t)W * S * PSX_SLOT_BYTES + (size_t)qpad * 4 + (size_t)cap * 8 + (size_t)W * S * 8 + 8 +
               (size_t)W * S * 8 + 16 + (listed ? (size_t)W * S * 32 * 4 : 0);
    };
    // co-resident CTAs share the SM's 228 KB (1 KB per CTA is reserved by the driver)
    const size_t limit = h->ctas_per_sm <= 1 ? (size_t)PSX_SMEM_LIMIT : (size_t)(228 * 1024) / h->ctas_per_sm - 1024;
    int S = h->stages;
    // short rows leave part of every 4 KB slot unused: keep the bytes in flight up with a third stage
    if (h->stages_auto && p.cpr == 1 && (size_t)p.rps * a.row_bytes * 5 < (size_t)PSX_SLOT_BYTES * 4) S = 3;
    // row lists are short (that is the point of a predicate): a warp's share is a handful of rows, fetched one bulk copy
    // each -- the number of round trips, not the bytes, sets the time, so give the ring its third slot
    if (h->stages_auto && listed && p.cpr == 1) S = 3;
    while (S > 2 && smem_for(S) > limit) --S;
    // the merge reuses the ring: it must hold at least two lists
    while ((size_t)W * S * PSX_SLOT_BYTES / 8 < (size_t)2 * p.kpad && smem_for(S + 1) <= PSX_SMEM_LIMIT) ++S;
    if (smem_for(S) > PSX_SMEM_LIMIT || (size_t)W * S * PSX_SLOT_BYTES / 8 < (size_t)2 * p.kpad)
        return fail(PSX_ERR_INVALID, "d=%d k=%d does not fit the shared-memory plan", h->d, k);
    p.stages = S;
    plan.smem = smem_for(S);
    // one warp per group (predicate groups) / per unit (dealt windows) until the machine is full; the
    // length of a row list is only known on the device, so a list launch always takes the full grid
    const long long num_groups = mode == PSX_SCAN_GROUPS ? (h->n + p.gsize - 1) / p.gsize : (h->n + p.rps - 1) / p.rps;
    long long grid = (num_groups + W - 1) / W;
    const long long maxgrid = (long long)h->sm_count * h->ctas_per_sm;
    if (grid > maxgrid || listed) grid = maxgrid;
    // a row-list batch is the 32 list entries one lane-wide load fetches (its latency is paid once per batch)
    p.static_batch = listed ? std::max(1, 32 / p.rps) : std::max(1, std::min(h->static_batch, 32));
    p.dyn_tail = h->dyn_tail ? 1 : 0;
    if (grid < 1) grid = 1;
    plan.grid = (int)grid;
    plan.block = W * 32;
    return PSX_OK;
}

static int launch_scan_variant(int device, int dtype, int metric, int ppl, bool qreg, const ScanPlan& plan, cudaStream_t st) {
    const ScanLaunch l{plan.grid, plan.block, plan.smem, plan.pdl};
    cudaError_t e;
    if (dtype == PSX_STORE_F32)
        e = metric == PSX_METRIC_IP ? launch_scan_shape<float, PSX_METRIC_IP>(device, ppl, qreg, plan.mode, plan.p, l, st)
                                    : launch_scan_shape<float, PSX_METRIC_L2>(device, ppl, qreg, plan.mode, plan.p, l, st);
    else
        e = metric == PSX_METRIC_IP ? launch_scan_shape<__nv_bfloat16, PSX_METRIC_IP>(device, ppl, qreg, plan.mode, plan.p, l, st)
                                    : launch_scan_shape<__nv_bfloat16, PSX_METRIC_L2>(device, ppl, qreg, plan.mode, plan.p, l, st);
    if (e != cudaSuccess) return fail(PSX_ERR_CUDA, "scan_topk_kernel launch: %s", cudaGetErrorString(e));
    return PSX_OK;
}

struct XchgArgs {
    int world, rank;
    uint32_t seq;
    const uint64_t* bases;  // host array [targets]: base address of every receive buffer this shard publishes to
    int targets;            // 0 = world (one process per GPU: every rank merges); 1 = only bases[0] merges (one process, G devices)
    // merge fused into the scan's last CTA (see ScanParams::xchg_my_recv): this rank's own receive buffer + outputs
    uint64_t my_base = 0;
    float* out_scores = nullptr;
    long long* out_ids = nullptr;
    int* status = nullptr;
    unsigned long long spin_limit = 0;
};
static size_t xchg_flag_offset() { return (size_t)2 * PSX_XCHG_MAX_WORLD * PSX_K_PASS_MAX * sizeof(uint64_t); }

static int launch_scan(psx_index* h, const float* q_dev, int k, const psx_filter* f, uint32_t id_base,
                       const uint64_t* ceil_ptr, float* out_scores, long long* out_ids, uint64_t* out_keys, cudaStream_t st,
                       const XchgArgs* xa = nullptr, bool master = false, const int* cond_flag = nullptr) {
    const bool filtered = f && f->flags;
    // the predicate: compacted into a row list first (default), or evaluated group by group inside the scan
    const bool listed = filtered && h->filter_mode != 1;
    const int mode = listed || (!filtered && h->deal) ? PSX_SCAN_DEAL : PSX_SCAN_GROUPS;
    ScanPlan plan;
    int rc = plan_scan(h, master, k, mode, listed, plan);
    if (rc) return rc;
    ScanParams& p = plan.p;
    // may this launch start while the kernel before it in the stream still runs?  (back-to-back queries, also behind the
    // merge kernel of a sharded query.)  Launches that consume a condition flag or a page ceiling written right before
    // them, or that are traced, keep full stream order.
    const bool overlap_ok = !cond_flag && !ceil_ptr && !h->trace;
    const bool overlap_prev = (h->pdl >= 2 || (h->pdl == 1 && !h->call_first)) && overlap_ok;
    p.list_count = h->counter + 2;
    // the list is compacted by the scan's own first phase (one launch per query, and the whole of it may overlap the
    // merge of the query before), or by a kernel of its own ahead of the scan
    // (auto: by the scan itself up to 4M rows -- 5-13 % faster at 1M rows, level at 10M, where the words are 80 MB and the
    // compaction kernel's overlap with the previous query's tail is worth as much as the saved launch)
    const bool self_listed = listed && (h->filter_mode == 3 || (h->filter_mode != 2 && h->n <= (1ll << 22)));
    if (listed) {
        if (h->rowlist_cap < h->cap) {
            CU(cudaStreamSynchronize(st));
            cudaFree(h->rowlist);
            h->rowlist = nullptr;
            h->rowlist_cap = 0;
            CU(cudaMalloc(&h->rowlist, (size_t)2 * h->cap * sizeof(uint32_t)));
            h->rowlist_cap = h->cap;
        }
        // two lists alternate, so that the list of query i+1 can be compacted while the scan of query i sorts and merges
        const int slot = (int)(h->list_seq++ & 1ull);
        uint32_t* list = h->rowlist + (size_t)slot * h->rowlist_cap;
        p.list_count = h->counter + (slot ? 4 : 2);
        if (self_listed) {
            p.fuse = h->counter + 32 + slot * PSX_FUSE_WORDS;
            // about one ticket per CTA (one CTA may be late: see the kernel) while a ticket stays within 4 blocks
            const long long blk = (long long)(plan.block / 32) * 256;
            const long long share = (h->n + std::max(plan.grid - 1, 1) - 1) / std::max(plan.grid - 1, 1);
            p.fuse_sub = (int)std::max<long long>(1, std::min<long long>(4, (share + blk - 1) / blk));
        } else {
            const long long chunks = (h->n + 2047) / 2048;  // 256 threads x 8 rows per trip
            const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(chunks, (long long)h->sm_count * 8));
            CU(launch_ex(filter_list_kernel, blocks, 256u, 0, st, overlap_prev, (const uint64_t*)h->attrs, (long long)h->n, *f, list,
                         p.list_count, cond_flag));
            g_launches++;
        }
        p.rowlist = list;
    }
    p.work = h->counter + ((h->scan_seq++ & 1ull) ? 3 : 1);
    // a row-list scan always overlaps the kernel that writes its list (it waits for it before reading the list)
    plan.pdl = listed && !self_listed ? (h->pdl >= 1 && overlap_ok) : overlap_prev;
    h->call_first = false;
    const size_t need_lists = (size_t)plan.grid * p.kpad;
    if (need_lists > h->lists_cap) {
        CU(cudaStreamSynchronize(st));
        cudaFree(h->lists);
        h->lists = nullptr;
        h->lists_cap = 0;
        const size_t want = std::max(need_lists, (size_t)h->sm_count * h->ctas_per_sm * 128);
        CU(cudaMalloc(&h->lists, want * sizeof(uint64_t)));
        h->lists_cap = want;
    }
    p.x = master ? h->xm : h->x;
    p.cond_flag = cond_flag;
    p.trace = h->trace;
    p.q = q_dev;
    p.ceil_ptr = ceil_ptr;
    p.lists = h->lists;
    p.counter = h->counter;
    p.out_scores = out_scores;
    p.out_ids = out_ids;
    p.out_keys = out_keys;
    p.id_base = id_base;
    p.has_filter = 0;
    p.attrs = nullptr;
    if (filtered && !listed) {
        p.has_filter = 1;
        p.attrs = h->attrs;
        p.f = *f;
    } else if (self_listed) {  // phase 1 of the launch reads the words; the streaming phase only sees the list
        p.attrs = h->attrs;
        p.f = *f;
    }
    if (xa) {
        p.xchg_world = xa->world;
        p.xchg_rank = xa->rank;
        p.xchg_seq = xa->seq;
        p.xchg_targets = xa->targets > 0 ? xa->targets : xa->targets < 0 ? 0 : xa->world;
        for (int r = 0; r < p.xchg_targets; ++r) {
            p.xchg_recv[r] = (uint64_t*)(uintptr_t)xa->bases[r];
            p.xchg_flag[r] = (uint32_t*)(uintptr_t)(xa->bases[r] + xchg_flag_offset());
        }
        if (xa->my_base) {
            p.xchg_my_recv = (const uint64_t*)(uintptr_t)xa->my_base;
            p.xchg_my_flag = (const uint32_t*)(uintptr_t)(xa->my_base + xchg_flag_offset());
            p.xchg_out_scores = xa->out_scores;
            p.xchg_out_ids = xa->out_ids;
            p.xchg_status = xa->status;
            p.xchg_spin_limit = xa->spin_limit;
        }
    }
    const size_t arena_row_bytes = master ? h->mrow_bytes : h->row_bytes;
    const int ppr = (int)(arena_row_bytes >> 4);  // 16-byte pieces per row
    int ppl = 0;
    bool qreg = false;
    if (p.cpr == 1 && ppr % 32 == 0) {
        ppl = ppr / 32;
        qreg = true;
        if (ppl == 5 || ppl == 7) ppl = 0, qreg = false;
    } else if (p.cpr > 1 && arena_row_bytes % PSX_SLOT_BYTES == 0) {
        ppl = 8;
    }
    rc = launch_scan_variant(h->device, master ? PSX_STORE_F32 : scan_dtype(h), h->metric, ppl, qreg, plan, st);
    if (rc) return rc;
    g_launches++;
    CU(cudaGetLastError());
    return PSX_OK;
}

// ------------------------------------------------------------------------------------------
// batched path (K3): TF32 tensor-core GEMM with fused threshold selection + exact re-score
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// 2-D fp32 (or bf16) tensor [rows][cols] with row stride ld elements, box = one 128-byte swizzle row x box_rows
static int make_map(CUtensorMap* map, const void* base, bool bf16, long long rows, int cols, int ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(PSX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const size_t esz = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)(GEMM_KB_BYTES / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PSX_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return PSX_OK;
}

// survivors per query the threshold aims at.  (bf16 operands: the rounding bound 8.2e-3 |q||x| is ~0.26 sigma of the
// score distribution of 1024-d unit vectors, i.e. ~2.6-3 x k rows lie within eps of the k-th score and ~7 x k within the
// 2 eps band the re-score visits: 12k+192 keeps them above theta.  Listing a row costs 8 bytes; only band rows are re-read.)
// bf16 operands are used for k <= 512 only: beyond, the 2-eps band alone outgrows the lists (the TF32 form over the
// master rows serves those k)
static bool batch_uses_bf16(const psx_index* h, int k) { return h->dtype == PSX_STORE_BF16_MASTER && h->batch_bf16 && k <= 512; }
static int batch_T(const psx_index* h, int k) {
    if (batch_uses_bf16(h, k)) return 12 * k + 192;
    // TF32: 4k+64 rows keep the eps band above theta; listing a row is cheap (8 bytes, only the band is re-read), so the
    // same number again (at most 2048, and never more than the lists can absorb at k = 2048) buys a threshold sample of
    // half the size for the same noise
    return std::min(std::min(8 * k + 128, 4 * k + 64 + 2048), 9000);
}
// shape test of the tensor-core path: an fp32 inner-product index, and a corpus large enough that the threshold's
// survivors are a small fraction of it (T <= n / 24: otherwise a sampled threshold is meaningless and the lists approach
// the corpus itself -- such small corpora are latency-bound on the scan anyway)
static bool batch_shape_ok(const psx_index* h, int64_t k) {
    return h->metric == PSX_METRIC_IP && has_fp32_rows(h) && k >= 1 && k <= PSX_K_PASS_MAX && h->n >= 65536 && h->d >= 32 &&
           (long long)batch_T(h, (int)k) * 24 <= h->n;
}
// cached cuTensorMapEncodeTiled
static int cached_map(psx_index::MapKey& key, CUtensorMap& map, const void* base, bool bf16, long long rows, int cols, int ld, int box_rows) {
    psx_index::MapKey want;
    want.base = base;
    want.rows = rows;
    want.cols = cols;
    want.ld = ld;
    want.box_rows = box_rows;
    want.bf16 = bf16 ? 1 : 0;
    if (key == want) return PSX_OK;
    int rc = make_map(&map, base, bf16, rows, cols, ld, box_rows);
    if (rc) return rc;
    key = want;
    return PSX_OK;
}

static bool batch_eligible(const psx_index* h, int64_t nq, int64_t k, const psx_filter* f) {
    (void)f;  // the predicate is applied in the epilogue (candidates) and in the sample (thresholds)
    return h->batch_min > 0 && nq >= h->batch_min && batch_shape_ok(h, k);
}

static int ensure_batch_scratch(psx_index* h, size_t sample_floats, int cand_cap = BATCH_CAND_CAP_MIN) {
    if (cand_cap > h->bcand_cap) {
        if (h->bcand) {
            CU(cudaDeviceSynchronize());  // a batch enqueued earlier may still be filling the old lists
            cudaFree(h->bcand);
            h->bcand = nullptr;
            h->bcand_cap = 0;
        }
        CU(cudaMalloc(&h->bcand, (size_t)BATCH_MAX_Q * cand_cap * sizeof(uint64_t)));
        h->bcand_cap = cand_cap;
    }
    if (!h->bq) {
        CU(cudaMalloc(&h->bq, (size_t)BATCH_MAX_Q * (h->ldm + 8) * sizeof(float)));
        CU(cudaMalloc(&h->btheta, BATCH_MAX_Q * sizeof(float)));
        CU(cudaMalloc(&h->bcount, BATCH_MAX_Q * sizeof(int)));
        CU(cudaMalloc(&h->bflags, BATCH_MAX_Q * sizeof(int)));
        CU(cudaMalloc(&h->mkeys, (size_t)PSX_K_PASS_MAX * sizeof(uint64_t)));
        CU(cudaMalloc(&h->meps, sizeof(float)));
        CU(cudaMallocHost(&h->hflags, BATCH_MAX_Q * sizeof(int)));
    }
    if (sample_floats > h->bsample_cap) {
        cudaFree(h->bsample);
        h->bsample = nullptr;
        h->bsample_cap = 0;
        CU(cudaMalloc(&h->bsample, sample_floats * sizeof(float)));
        h->bsample_cap = sample_floats;
    }
    return PSX_OK;
}

// PSX_DEBUG_SYNC=1: synchronise after every launch of the batched path and name the kernel that faulted
static bool debug_sync() {
    static const bool on = [] { const char* e = getenv("PSX_DEBUG_SYNC"); return e && *e == '1'; }();
    return on;
}
#define DBG_SYNC(st, what)                                                                                     \
    do {                                                                                                       \
        if (debug_sync()) {                                                                                    \
            cudaError_t e_ = cudaStreamSynchronize(st);                                                        \
            if (e_ != cudaSuccess) return fail(PSX_ERR_CUDA, "%s faulted: %s", what, cudaGetErrorString(e_)); \
        }                                                                                                      \
    } while (0)

// PSX_BATCH_TIMING=1: CUDA events between the phases of a batch, printed to stderr (synchronises the stream)
struct BatchTimer {
    static bool enabled() {
        static const bool on = [] { const char* e = getenv("PSX_BATCH_TIMING"); return e && *e == '1'; }();
        return on;
    }
    cudaStream_t st;
    cudaEvent_t ev[8];
    const char* names[8];
    int n = 0;
    explicit BatchTimer(cudaStream_t s) : st(s) {
        if (enabled()) mark("start");
    }
    void mark(const char* name) {
        if (!enabled() || n >= 8) return;
        cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        names[n++] = name;
    }
    void report() {
        if (!enabled()) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "psx batch phases:");
        for (int i = 1; i < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, " %s %.1f us;", names[i], ms * 1e3f);
        }
        fprintf(stderr, "\n");
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        n = 0;
    }
};

template <int MT, bool BF>
static int launch_gemm(psx_index* h, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& gp, int grid, cudaStream_t st, bool pdl) {
    constexpr int STAGES = BatchCfg<MT>::STAGES;
    constexpr int BATCH_BN = BatchCfg<MT>::BN;
    constexpr int STAGE_BYTES = (MT * GEMM_M + BATCH_BN) * GEMM_KB_BYTES;
    constexpr size_t smem = (size_t)STAGES * STAGE_BYTES + 256 + (size_t)MT * GEMM_STAGE_BYTES_PER_ACC;
    static std::atomic<bool> ready[64];
    if (h->device < 64 && !ready[h->device].load()) {
        CU(cudaFuncSetAttribute(gemm_filter_kernel<MT, BATCH_BN, STAGES, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ready[h->device].store(true);
    }
    CU(launch_ex(gemm_filter_kernel<MT, BATCH_BN, STAGES, BF>, (unsigned)grid, GEMM_THREADS, smem, st, pdl, mq, mx, gp));
    g_launches++;
    return PSX_OK;
}

// One batch of nq <= 256 queries (device pointers).  flags_dev[qi] != 0 marks results that are not
// proven exact; the caller re-runs those queries on the streaming scan.
static int launch_batch(psx_index* h, const float* q_dev, int nq, int k, const psx_filter* f, uint32_t id_base, float qnorm_max,
                        float* out_scores, long long* out_ids, uint64_t* out_keys, int* flags_dev, cudaStream_t st,
                        long long keys_stride = 0) {
    const int MT = nq > GEMM_M ? 2 : 1;
    const bool pair = MT == 2 && h->batch_pair;
    const int BATCH_BN = batch_bn(MT, pair);
    const int num_tiles = (int)((h->n + BATCH_BN - 1) / BATCH_BN);
    // theta from a strided sample: aim at ~SAMPLE_RANK sample scores above the threshold that ~T rows pass.  The
    // threshold has to land between the k-th score and roughly the (cand_cap)-th; the certificate decides.  The
    // number of rows above the sample's r-th score is ~ T * Gamma(r)/r; it must not undercut the ~1.3-1.7 k rows of
    // the eps band (history: 8 sample scores with T = 3k+48 failed 2.7 % of the queries of a 10M-row corpus, each of
    // which costs a full scan; 16 with T = 4k+64 none in 10 240; now 12 with T = 8k+128, i.e. twice the head room).
    const int T = batch_T(h, k);
    // The sample: every tile_step-th tile, and of a visited tile only the first sample_cols rows, sized so that about
    // SAMPLE_RANK sample rows score above the threshold that T rows of the corpus pass.  A sample CTA keeps its 8 best
    // scores per query, so a CTA must not see more than a few rows above theta: narrow the columns when T/n is large
    // (small corpora with a large k) and never give one CTA more than its share.
    constexpr double SAMPLE_RANK = 12.0;
    const double frac = (double)T / (double)h->n;  // fraction of the rows above theta
    int sample_cols = 256;
    while (sample_cols > 32 && frac * sample_cols > 1.5) sample_cols >>= 1;
    if (sample_cols > BATCH_BN) sample_cols = BATCH_BN;
    const double want_rows = SAMPLE_RANK / frac;
    int sample_tiles = (int)std::min<double>(num_tiles, std::max(1.0, want_rows / sample_cols + 0.5));
    int tile_step = num_tiles / sample_tiles;
    if (tile_step < 1) tile_step = 1;
    sample_tiles = (num_tiles + tile_step - 1) / tile_step;
    const int grid_f = pair ? std::min(h->sm_count / 2, num_tiles) : std::min(h->sm_count, num_tiles);
    // Fold the sample into the filter kernel (CTA-pair kernel): every pair's first tile -- a full 256-row tile -- is its
    // sample.  Possible when one tile per pair is sample enough (want_rows <= grid_f tiles) and a tile holds no more than
    // ~1.5 rows above theta (else a pair's 8 kept scores saturate).  At k = 100 that is 160k .. 1.46M rows per GPU.
    const bool fold = pair && h->batch_fold && !BatchTimer::enabled() && !debug_sync() && frac * 256.0 <= 1.5 &&
                      want_rows <= (double)grid_f * 256.0 && grid_f >= 8;
    const int grid_s = fold ? grid_f : pair ? std::min(h->sm_count / 2, sample_tiles) : std::min(h->sm_count, sample_tiles);
    const int sample_ld = grid_s * SAMPLE_KEEP;  // every sample CTA (pair) leaves its 8 best scores per query
    const int cand_cap = batch_cand_cap(T);
    int rc = ensure_batch_scratch(h, (size_t)MT * GEMM_M * sample_ld, cand_cap);
    if (rc) return rc;
    // queries -> zero-padded [MT*128][ld] block (rows beyond nq and columns beyond d are zero)
    const int fld = fp32_ld(h);
    // bf16 + fp32 master: the GEMM streams the bf16 rows (half the HBM bytes, kind::f16 at twice the TF32 rate);
    // the survivors are re-scored on the master exactly as in the TF32 form
    const bool bf = batch_uses_bf16(h, k);
    // Query staging: the block the TMA reads is [MT*128][ld] with rows beyond nq and columns beyond d zero.  Only the rows
    // an earlier, larger batch left behind are cleared (none in steady state); the padding columns of a row are written
    // by the copy / pack itself.
    const size_t q_row_bytes = bf ? (size_t)h->ld * sizeof(__nv_bfloat16) : (size_t)fld * sizeof(float);
    if (h->bq_row_bytes != q_row_bytes) {  // first use, or the operand type changed: everything is stale
        h->bq_dirty_rows = BATCH_MAX_Q_ROWS;
        h->bq_row_bytes = q_row_bytes;
    }
    if (h->bq_dirty_rows > nq)
        CU(cudaMemsetAsync((unsigned char*)h->bq + (size_t)nq * q_row_bytes, 0, (size_t)(h->bq_dirty_rows - nq) * q_row_bytes, st));
    h->bq_dirty_rows = nq;
    if (bf) {
        pack_rows_kernel<__nv_bfloat16><<<(nq + 7) / 8, 256, 0, st>>>(q_dev, (__nv_bfloat16*)h->bq, nq, h->d, h->ld, 0, nullptr);
        g_launches++;
        CU(cudaGetLastError());
        if ((rc = cached_map(h->mq_key, h->mq_map, h->bq, true, (long long)MT * GEMM_M, h->d, h->ld, GEMM_M))) return rc;
        if ((rc = cached_map(h->mx_key, h->mx_map, h->x, true, h->n, h->d, h->ld, pair ? 128 : BATCH_BN))) return rc;
    } else {
        if (fld == h->d) {
            CU(cudaMemcpyAsync(h->bq, q_dev, (size_t)nq * h->d * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {  // d not a multiple of 4: rows are padded to ld floats (pack = copy + zero padding)
            pack_rows_kernel<float><<<(nq + 7) / 8, 256, 0, st>>>(q_dev, h->bq, nq, h->d, fld, 0, nullptr);
            g_launches++;
            CU(cudaGetLastError());
        }
        if ((rc = cached_map(h->mq_key, h->mq_map, h->bq, false, (long long)MT * GEMM_M, h->d, fld, GEMM_M))) return rc;
        if ((rc = cached_map(h->mx_key, h->mx_map, fp32_rows(h), false, h->n, h->d, fld, pair ? 128 : BATCH_BN))) return rc;
    }
    const CUtensorMap& mq = h->mq_map;
    const CUtensorMap& mx = h->mx_map;
    GemmParams gp;
    memset(&gp, 0, sizeof gp);
    gp.n = h->n;
    gp.d = h->d;
    gp.nq = nq;
    gp.num_tiles = num_tiles;
    gp.theta = h->btheta;
    gp.cand = h->bcand;
    gp.cand_count = h->bcount;
    gp.cand_cap = cand_cap;
    gp.sample_scores = h->bsample;
    gp.sample_ld = sample_ld;
    gp.sample_cols = sample_cols;
    if (f && f->flags) {
        gp.attrs = h->attrs;
        gp.f = *f;
    }
    // the kernels of one batch chain by programmatic dependent launch (see psx_gemm.cuh) unless a phase is being timed
    const bool chain = h->batch_pdl && !BatchTimer::enabled() && !debug_sync();
    auto run_gemm = [&](int grid, bool pdl) -> int {
        if (bf)
            return pair ? launch_gemm_pair<true>(h, mq, mx, gp, grid, st, pdl)
                        : MT == 2 ? launch_gemm<2, true>(h, mq, mx, gp, grid, st, pdl) : launch_gemm<1, true>(h, mq, mx, gp, grid, st, pdl);
        return pair ? launch_gemm_pair<false>(h, mq, mx, gp, grid, st, pdl)
                    : MT == 2 ? launch_gemm<2, false>(h, mq, mx, gp, grid, st, pdl) : launch_gemm<1, false>(h, mq, mx, gp, grid, st, pdl);
    };
    BatchTimer bt(st);
    if (fold) {
        // ONE kernel: sample (first tile of every pair, kept in TMEM) -> grid barrier -> thresholds -> grid barrier -> filter
        if (!h->gbar) {
            CU(cudaMalloc(&h->gbar, sizeof(unsigned int)));
            CU(cudaMemsetAsync(h->gbar, 0, sizeof(unsigned int), st));
            h->gbar_epoch = 0;
        }
        long long sample_rows = 0;  // the first tile of pair p under the kernel's rotation of the pair's tile sequence
        for (int pr = 0; pr < grid_f; ++pr) {
            const int n_mine = pr < num_tiles ? (num_tiles - pr + grid_f - 1) / grid_f : 0;
            if (n_mine == 0) continue;
            const int t0 = pr + grid_f * (int)((long long)pr * n_mine / grid_f);
            sample_rows += std::max<long long>(0, std::min<long long>(BATCH_BN, h->n - (long long)t0 * BATCH_BN));
        }
        int rank = (int)((double)T * (double)sample_rows / (double)h->n + 0.5);
        if (rank < 2) rank = 2;
        gp.mode = GEMM_MODE_FILTER;
        gp.tile_step = 1;
        gp.fold = 1;
        gp.theta_rank = rank;
        gp.theta_out = h->btheta;
        gp.gbar = h->gbar;
        gp.gbar_base = h->gbar_epoch;
        h->gbar_epoch += 2u * 2u * (unsigned)grid_f;  // two barriers x (2 CTAs per pair)
        DBG_SYNC(st, "query staging");
        rc = run_gemm(grid_f, false);  // behind the staging copies in plain stream order
        if (rc) return rc;
    } else {
        // pass 1: sample
        gp.mode = GEMM_MODE_SAMPLE;
        gp.tile_step = tile_step;
        DBG_SYNC(st, "query staging");
        rc = run_gemm(grid_s, false);  // behind the staging copies in plain stream order
        if (rc) return rc;
        DBG_SYNC(st, "gemm_filter_kernel(sample)");
        bt.mark("sample pass");
        // rows actually sampled: sample_cols of every visited tile (the last tile may be short)
        long long sample_rows = 0;
        for (int t = 0; t < num_tiles; t += tile_step)
            sample_rows += std::max<long long>(0, std::min<long long>(sample_cols, h->n - (long long)t * BATCH_BN));
        int rank = (int)((double)T * (double)sample_rows / (double)h->n + 0.5);
        if (rank < 2) rank = 2;
        CU(launch_ex(theta_kernel, (unsigned)nq, 512u, 0, st, chain, (const float*)h->bsample, sample_ld, sample_ld, rank, h->btheta, h->bcount));
        g_launches++;
        DBG_SYNC(st, "theta_kernel");
        bt.mark("theta");
        // pass 2: every tile, threshold test fused into the epilogue
        gp.mode = GEMM_MODE_FILTER;
        gp.tile_step = 1;
        rc = run_gemm(grid_f, chain);
        if (rc) return rc;
    }
    DBG_SYNC(st, "gemm_filter_kernel(filter)");
    bt.mark("filter pass");
    // exact re-score of the survivors + top-k + proof obligation
    const int kpad = (int)psx_kpad(k);
    const size_t smem = (size_t)std::max(cand_cap, kpad) * 8 + (size_t)(fld + 8) * 4;
    static std::atomic<bool> ready[64];
    if (h->device < 64 && !ready[h->device].load()) {
        CU(cudaFuncSetAttribute(rescore_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT - 1024));
        ready[h->device].store(true);
    }
    // TF32 drops 13 mantissa bits of each operand (truncation: relative error < 2^-10 each):
    // |score_tf32 - score| <= 2 * 2^-10 * sum|q_i x_i| <= 2^-9 |q| |x| (Cauchy-Schwarz); 2.2e-3 leaves 12 % for the accumulation.
    // bf16 operands are rounded to nearest with 8 significant bits (unit roundoff 2^-8 EACH, q and x both rounded):
    // |score_bf16 - score| <= (2 * 2^-8 + 2^-16) |q| |x| = 7.83e-3 |q| |x|; 8.2e-3 leaves 5 % for the accumulation.
    // The kernel multiplies the coefficient by the query's own norm and the largest stored row norm (both on the device).
    (void)qnorm_max;
    const float eps_coef = bf ? 8.2e-3f : 2.2e-3f;
    CU(launch_ex(rescore_select_kernel, (unsigned)nq, 512u, smem, st, chain, fp32_rows(h), fld, h->d, (long long)h->n, q_dev, k, kpad,
                 (const uint64_t*)h->bcand, (const int*)h->bcount, cand_cap, (const float*)h->btheta, eps_coef, (const float*)h->dmax_sumsq,
                 (const float*)nullptr, id_base, out_scores, out_ids, out_keys, (long long)(keys_stride > 0 ? keys_stride : kpad), flags_dev));
    g_launches++;
    DBG_SYNC(st, "rescore_select_kernel");
    bt.mark("rescore");
    bt.report();
    return PSX_OK;
}

// ------------------------------------------------------------------------------------------
// one exact query, whatever the storage tier
// ------------------------------------------------------------------------------------------
// fp32 scan over the rows that hold the exact values (the arena itself, or the master copy)
static int launch_exact_scan(psx_index* h, const float* q_dev, int k, const psx_filter* f, uint32_t id_base, const uint64_t* ceil_ptr,
                             float* out_scores, long long* out_ids, uint64_t* out_keys, cudaStream_t st, const int* cond_flag = nullptr) {
    return launch_scan(h, q_dev, k, f, id_base, ceil_ptr, out_scores, out_ids, out_keys, st, nullptr, h->dtype == PSX_STORE_BF16_MASTER,
                       cond_flag);
}

// PSX_STORE_BF16_MASTER: stream the bf16 rows for k' candidates, re-score them on the fp32 master,
// certify, and fall back to the master scan (a conditional launch: the grid exits at once when the
// certificate holds) otherwise.  No host synchronisation.
static int launch_mixed(psx_index* h, const float* q_dev, int k, const psx_filter* f, uint32_t id_base, float* out_scores,
                        long long* out_ids, uint64_t* out_keys, cudaStream_t st) {
    int rc = ensure_batch_scratch(h, 0);
    if (rc) return rc;
    int kprime = 4 * k + 64;
    if (kprime > PSX_K_PASS_MAX) kprime = PSX_K_PASS_MAX;
    if ((rc = launch_scan(h, q_dev, kprime, f, id_base, nullptr, nullptr, nullptr, h->mkeys, st))) return rc;
    keys_to_cands_kernel<<<1, 256, 0, st>>>(h->mkeys, kprime, id_base, h->bcand, h->bcount, h->btheta, q_dev, h->d, h->dmax_sumsq, h->meps);
    g_launches++;
    CU(cudaGetLastError());
    const int kpad = (int)psx_kpad(k);
    const size_t smem = (size_t)h->bcand_cap * 8 + (size_t)(h->ldm + 8) * 4;
    static std::atomic<bool> ready[64];
    if (h->device < 64 && !ready[h->device].load()) {
        CU(cudaFuncSetAttribute(rescore_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT - 1024));
        ready[h->device].store(true);
    }
    rescore_select_kernel<<<1, 512, smem, st>>>((const float*)h->xm, h->ldm, h->d, h->n, q_dev, k, kpad, h->bcand, h->bcount, h->bcand_cap,
                                                h->btheta, 0.f, h->dmax_sumsq, h->meps, id_base, out_scores, out_ids, out_keys, kpad, h->bflags);
    g_launches++;
    CU(cudaGetLastError());
    h->mixed_queries++;
    return launch_exact_scan(h, q_dev, k, f, id_base, nullptr, out_scores, out_ids, out_keys, st, h->bflags);
}

// the single-query entry every API funnels through
static int launch_query(psx_index* h, const float* q_dev, int k, const psx_filter* f, uint32_t id_base, const uint64_t* ceil_ptr,
                        float* out_scores, long long* out_ids, uint64_t* out_keys, cudaStream_t st) {
    if (h->dtype == PSX_STORE_BF16_MASTER) {
        // IP only: the rounding bound of the L2 form also involves |x|^2; paged continuations (ceil_ptr) go straight to the master
        if (h->metric == PSX_METRIC_IP && !ceil_ptr) return launch_mixed(h, q_dev, k, f, id_base, out_scores, out_ids, out_keys, st);
        return launch_exact_scan(h, q_dev, k, f, id_base, ceil_ptr, out_scores, out_ids, out_keys, st);
    }
    return launch_scan(h, q_dev, k, f, id_base, ceil_ptr, out_scores, out_ids, out_keys, st);
}

// scratch is per index: order this search after the previous one if it ran on another stream
static int enter_stream(psx_index* h, cudaStream_t st) {
    h->call_first = true;
    if (h->has_last && h->last_stream != st) CU(cudaStreamWaitEvent(st, h->last_ev, 0));
    return PSX_OK;
}
static int leave_stream(psx_index* h, cudaStream_t st) {
    CU(cudaEventRecord(h->last_ev, st));
    h->last_stream = st;
    h->has_last = true;
    return PSX_OK;
}

extern "C" int psx_search_device(psx_index* h, const float* q_dev, int64_t nq, int64_t k, const psx_filter* filter,
                                 uint32_t id_base, float* out_scores_dev, int64_t* out_ids_dev, uint64_t* out_keys_dev,
                                 void* stream) {
    if (!h || !q_dev || nq < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_search_device");
    if (k < 1 || k > PSX_K_PASS_MAX) return fail(PSX_ERR_INVALID, "k=%lld not in [1,%d]", (long long)k, PSX_K_PASS_MAX);
    NOT_ON_GROUP(h, "psx_search_device");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = enter_stream(h, st))) return rc;
    const int64_t kpad = psx_kpad(k);
    for (int64_t qi = 0; qi < nq; ++qi) {
        rc = launch_query(h, q_dev + qi * h->d, (int)k, filter, id_base, nullptr,
                          out_scores_dev ? out_scores_dev + qi * k : nullptr,
                         out_ids_dev ? (long long*)out_ids_dev + qi * k : nullptr,
                         out_keys_dev ? out_keys_dev + qi * kpad : nullptr, st);
        if (rc) return rc;
    }
    return leave_stream(h, st);
}

extern "C" int64_t psx_exchange_bytes(void) {
    return (int64_t)(xchg_flag_offset() + 2 * PSX_XCHG_MAX_WORLD * sizeof(uint32_t) + 192);
}

extern "C" int psx_search_exchange_device(psx_index* h, const float* q_dev, int64_t k, const psx_filter* filter, uint32_t id_base,
                                          int rank, int world, const uint64_t* peer_bases, uint32_t seq, int phases,
                                          float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
    if (!h || !peer_bases || (!(phases & 1) && !(phases & 2)) || ((phases & 1) && !q_dev) ||
        ((phases & 2) && (!out_scores_dev || !out_ids_dev)))
        return fail(PSX_ERR_INVALID, "bad arguments to psx_search_exchange_device");
    if (world < 1 || world > PSX_XCHG_MAX_WORLD || rank < 0 || rank >= world || seq == 0)
        return fail(PSX_ERR_INVALID, "exchange needs 1 <= world <= %d, 0 <= rank < world, seq >= 1", PSX_XCHG_MAX_WORLD);
    if (k < 1 || k > PSX_K_PASS_MAX) return fail(PSX_ERR_INVALID, "k=%lld not in [1,%d]", (long long)k, PSX_K_PASS_MAX);
    NOT_ON_GROUP(h, "psx_search_exchange_device");
    if (h->dtype == PSX_STORE_BF16_MASTER)
        return fail(PSX_ERR_STATE, "the fused exchange publishes the scan's own keys; a bf16+master index exchanges through psx_search_device keys");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = enter_stream(h, st))) return rc;
    XchgArgs xa{world, rank, seq, peer_bases, 0};
    int* status_dev = nullptr;
    if ((rc = ensure_xstatus(h, &status_dev))) return rc;
    // the normal call (both phases): the scan's last CTA also waits for the peers' lists and merges -- one kernel per query
    const bool inline_merge = phases == 3 && h->xchg_inline;
    if (inline_merge) {
        // the ring the merge sorts in must hold world * kpad keys: always true for the default geometry (128 KB)
        xa.my_base = peer_bases[rank];
        xa.out_scores = out_scores_dev;
        xa.out_ids = (long long*)out_ids_dev;
        xa.status = status_dev;
        xa.spin_limit = xchg_spin_limit(h->xchg_timeout_ms);
    }
    if ((phases & 1) && (rc = launch_scan(h, q_dev, (int)k, filter, id_base, nullptr, nullptr, nullptr, nullptr, st, &xa))) return rc;
    if (!(phases & 2) || inline_merge) return leave_stream(h, st);
    const int kpad = (int)psx_kpad(k);
    int np = kpad;
    while (np < world * kpad) np <<= 1;
    const size_t smem = (size_t)np * 8;
    static std::atomic<bool> ready[64];
    if (h->device < 64 && !ready[h->device].load()) {
        CU(cudaFuncSetAttribute(merge_wait_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT - 1024));
        ready[h->device].store(true);
    }
    const uint64_t mine = peer_bases[rank];
    merge_wait_kernel<<<1, 256, smem, st>>>((const uint64_t*)(uintptr_t)mine, (const uint32_t*)(uintptr_t)(mine + xchg_flag_offset()), world,
                                           seq, (int)k, kpad, np, h->metric, out_scores_dev, (long long*)out_ids_dev, nullptr, status_dev,
                                           xchg_spin_limit(h->xchg_timeout_ms));
    g_launches++;
    CU(cudaGetLastError());
    return leave_stream(h, st);
}

extern "C" int psx_search_batch_device(psx_index* h, const float* q_dev, int64_t nq, int64_t k, const psx_filter* filter,
                                       float qnorm_max, uint32_t id_base, float* out_scores_dev, int64_t* out_ids_dev,
                                       uint64_t* out_keys_dev, int* flags_dev, void* stream) {
    if (!h || !q_dev || !out_scores_dev || !out_ids_dev || !flags_dev || nq < 1)
        return fail(PSX_ERR_INVALID, "bad arguments to psx_search_batch_device");
    NOT_ON_GROUP(h, "psx_search_batch_device");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (!batch_shape_ok(h, k))
        return fail(PSX_ERR_STATE, "tensor-core batch path needs an fp32 inner-product index with >= max(65536, 24 * (4k+64)) rows, d >= 32, k <= %d",
                    PSX_K_PASS_MAX);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = enter_stream(h, st))) return rc;
    const int64_t kpad = psx_kpad(k);
    for (int64_t q0 = 0; q0 < nq; q0 += BATCH_MAX_Q) {
        const int gq = (int)std::min<int64_t>(BATCH_MAX_Q, nq - q0);
        rc = launch_batch(h, q_dev + q0 * h->d, gq, (int)k, filter, id_base, qnorm_max, out_scores_dev + q0 * k,
                          (long long*)out_ids_dev + q0 * k, out_keys_dev ? out_keys_dev + q0 * kpad : nullptr, flags_dev + q0, st);
        if (rc) return rc;
    }
    h->batch_queries += nq;
    return leave_stream(h, st);
}

extern "C" int psx_merge_keys_device(int device, const uint64_t* keys_dev, int64_t nq, int64_t nlists, int64_t k, int metric,
                                     float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
    if (!keys_dev || nq < 0 || nlists < 1 || k < 1 || k > PSX_K_PASS_MAX)
        return fail(PSX_ERR_INVALID, "bad arguments to psx_merge_keys_device");
    if (nq == 0) return PSX_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(PSX_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    static std::atomic<bool> attr_done[64];
    if (device >= 0 && device < 64 && !attr_done[device].load()) {
        CU(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT));
        attr_done[device].store(true);
    }
    const int kpad = (int)psx_kpad(k);
    int cap_lists = (int)std::min<int64_t>(nlists + 1, (200 * 1024 / 8) / kpad);
    if (cap_lists < 2) cap_lists = 2;
    const size_t smem = (size_t)cap_lists * kpad * 8;
    merge_keys_kernel<<<(unsigned)nq, 256, smem, (cudaStream_t)stream>>>(keys_dev, (int)nlists, (int)k, kpad, cap_lists, metric,
                                                                       out_scores_dev, (long long*)out_ids_dev, nullptr);
    g_launches++;
    CU(cudaGetLastError());
    return PSX_OK;
}

static int ensure_io(psx_index* h, size_t qfloats, size_t outs) {
    if (qfloats > h->dq_cap) {
        cudaFree(h->dq);
        h->dq = nullptr;
        h->dq_cap = 0;
        CU(cudaMalloc(&h->dq, qfloats * sizeof(float)));
        h->dq_cap = qfloats;
    }
    if (qfloats > h->hq_cap) {
        cudaFreeHost(h->hq);
        h->hq = nullptr;
        h->hq_cap = 0;
        CU(cudaMallocHost(&h->hq, qfloats * sizeof(float)));
        h->hq_cap = qfloats;
    }
    if (outs > h->dout_cap) {
        cudaFree(h->dscores);
        cudaFree(h->dids);
        cudaFree(h->dkeys);
        h->dscores = nullptr;
        h->dids = nullptr;
        h->dkeys = nullptr;
        h->dout_cap = 0;
        CU(cudaMalloc(&h->dscores, outs * sizeof(float)));
        CU(cudaMalloc(&h->dids, outs * sizeof(long long)));
        CU(cudaMalloc(&h->dkeys, outs * sizeof(uint64_t)));
        h->dout_cap = outs;
    }
    if (outs > h->hout_cap) {
        cudaFreeHost(h->hscores);
        cudaFreeHost(h->hids);
        h->hscores = nullptr;
        h->hids = nullptr;
        h->hout_cap = 0;
        CU(cudaMallocHost(&h->hscores, outs * sizeof(float)));
        CU(cudaMallocHost(&h->hids, outs * sizeof(long long)));
        h->hout_cap = outs;
    }
    return PSX_OK;
}

// Host-buffer batch search through the tensor-core path; unproven queries are re-run on the scan.
static int search_batched_host(psx_index* h, const float* q, int64_t nq, int64_t kk, int64_t k, const psx_filter* filter,
                               uint32_t id_base, float* out_scores, int64_t* out_ids) {
    int rc;
    cudaStream_t st = h->stream;
    if ((rc = ensure_batch_scratch(h, 0))) return rc;  // h->bflags / h->hflags must exist before they are passed on
    if ((rc = enter_stream(h, st))) return rc;
    for (int64_t q0 = 0; q0 < nq; q0 += BATCH_MAX_Q) {
        const int gq = (int)std::min<int64_t>(BATCH_MAX_Q, nq - q0);
        if ((rc = ensure_io(h, (size_t)gq * h->d, (size_t)gq * kk))) return rc;
        memcpy(h->hq, q + q0 * h->d, (size_t)gq * h->d * sizeof(float));
        float qn2 = 0.f;
        for (int qi = 0; qi < gq; ++qi) {
            double acc = 0.0;
            const float* v = h->hq + (size_t)qi * h->d;
            for (int i = 0; i < h->d; ++i) acc += (double)v[i] * v[i];
            qn2 = std::max(qn2, (float)acc);
        }
        CU(cudaMemcpyAsync(h->dq, h->hq, (size_t)gq * h->d * sizeof(float), cudaMemcpyHostToDevice, st));
        if ((rc = launch_batch(h, h->dq, gq, (int)kk, filter, id_base, sqrtf(qn2) * 1.0001f, h->dscores, h->dids, nullptr, h->bflags, st)))
            return rc;
        // the certificates travel with the results: ONE synchronisation per batch unless a query has to be re-run
        CU(cudaMemcpyAsync(h->hflags, h->bflags, gq * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->hscores, h->dscores, (size_t)gq * kk * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->hids, h->dids, (size_t)gq * kk * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        h->batch_queries += gq;
        bool rerun = false;
        for (int qi = 0; qi < gq; ++qi) {
            if (!h->hflags[qi]) continue;
            h->batch_fallbacks++;
            rerun = true;
            rc = launch_exact_scan(h, h->dq + (size_t)qi * h->d, (int)kk, filter, id_base, nullptr, h->dscores + (size_t)qi * kk,
                                   h->dids + (size_t)qi * kk, nullptr, st);
            if (rc) return rc;
        }
        if (rerun) {
            CU(cudaMemcpyAsync(h->hscores, h->dscores, (size_t)gq * kk * sizeof(float), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h->hids, h->dids, (size_t)gq * kk * sizeof(long long), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        for (int qi = 0; qi < gq; ++qi) {
            float* os = out_scores + (q0 + qi) * k;
            int64_t* oi = out_ids + (q0 + qi) * k;
            memcpy(os, h->hscores + (size_t)qi * kk, (size_t)kk * sizeof(float));
            memcpy(oi, h->hids + (size_t)qi * kk, (size_t)kk * sizeof(long long));
            for (int64_t i = kk; i < k; ++i) {
                os[i] = -INFINITY;
                oi[i] = -1;
            }
        }
    }
    return leave_stream(h, st);
}

static int group_search(psx_index* g, const float* q, int64_t nq, int64_t k, const psx_filter* filter, float* out_scores,
                        int64_t* out_ids);

// Host-buffer search of ONE device's rows (ids reported as id_base + row).  The caller holds h->mu (or owns h as
// the child of a group), the device is current and nothing is staged.
static int search_single(psx_index* h, const float* q, int64_t nq, int64_t k, const psx_filter* filter, uint32_t id_base,
                         float* out_scores, int64_t* out_ids) {
    int rc;
    const float empty = h->metric == PSX_METRIC_L2 ? INFINITY : -INFINITY;
    if (h->n == 0) {
        for (int64_t i = 0; i < nq * k; ++i) {
            out_scores[i] = empty;
            out_ids[i] = -1;
        }
        return PSX_OK;
    }
    // results beyond ntotal can never be filled: scan for min(k, n) and pad on the host
    const int64_t kk = std::min<int64_t>(k, h->n);
    if (batch_eligible(h, nq, kk, filter)) return search_batched_host(h, q, nq, kk, k, filter, id_base, out_scores, out_ids);
    const int64_t pages = (kk + PSX_K_PASS_MAX - 1) / PSX_K_PASS_MAX;
    // per-query stride of the device outputs
    const int64_t kslot = pages == 1 ? psx_kpad(kk) : pages * PSX_K_PASS_MAX;
    // bound the device/pinned staging: process queries in groups
    const int64_t group = std::max<int64_t>(1, std::min<int64_t>(nq, (8ll << 20) / kslot));
    if ((rc = ensure_io(h, (size_t)group * h->d, (size_t)group * kslot))) return rc;
    cudaStream_t st = h->stream;
    if ((rc = enter_stream(h, st))) return rc;
    for (int64_t q0 = 0; q0 < nq; q0 += group) {
        const int64_t gq = std::min(group, nq - q0);
        memcpy(h->hq, q + q0 * h->d, (size_t)gq * h->d * sizeof(float));
        CU(cudaMemcpyAsync(h->dq, h->hq, (size_t)gq * h->d * sizeof(float), cudaMemcpyHostToDevice, st));
        for (int64_t qi = 0; qi < gq; ++qi) {
            for (int64_t pg = 0; pg < pages; ++pg) {
                const int kp = (int)std::min<int64_t>(PSX_K_PASS_MAX, kk - pg * PSX_K_PASS_MAX);
                const size_t off = (size_t)qi * kslot + (size_t)pg * PSX_K_PASS_MAX;
                // page pg continues strictly below the last key of page pg-1 (a full page)
                const uint64_t* ceil_ptr = pg ? h->dkeys + off - 1 : nullptr;
                rc = launch_query(h, h->dq + qi * h->d, kp, filter, id_base, ceil_ptr, h->dscores + off, h->dids + off,
                                  h->dkeys + off, st);
                if (rc) return rc;
            }
        }
        CU(cudaMemcpyAsync(h->hscores, h->dscores, (size_t)gq * kslot * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->hids, h->dids, (size_t)gq * kslot * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int64_t qi = 0; qi < gq; ++qi) {
            float* os = out_scores + (q0 + qi) * k;
            int64_t* oi = out_ids + (q0 + qi) * k;
            memcpy(os, h->hscores + qi * kslot, (size_t)kk * sizeof(float));
            memcpy(oi, h->hids + qi * kslot, (size_t)kk * sizeof(long long));
            for (int64_t i = kk; i < k; ++i) {
                os[i] = empty;
                oi[i] = -1;
            }
        }
    }
    return leave_stream(h, st);
}

extern "C" int psx_search(psx_index* h, const float* q, int64_t nq, int64_t k, const psx_filter* filter, float* out_scores,
                          int64_t* out_ids) {
    if (!h || !q || !out_scores || !out_ids || nq < 0) return fail(PSX_ERR_INVALID, "bad arguments to psx_search");
    if (k < 1) return fail(PSX_ERR_INVALID, "k=%lld must be >= 1", (long long)k);
    if (nq == 0) return PSX_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (is_group(h)) return group_search(h, q, nq, k, filter, out_scores, out_ids);
    return search_single(h, q, nq, k, filter, 0, out_scores, out_ids);
}

// ------------------------------------------------------------------------------------------
// reading rows back
// ------------------------------------------------------------------------------------------
static int read_rows_locked(psx_index* h, long long row0, long long n, float* out) {
    if (n == 0) return PSX_OK;
    const long long chunk = std::max<long long>(1, (128ll << 20) / ((long long)h->d * 4));
    float* tmp = nullptr;
    const long long trows = std::min(chunk, n);
    CU(cudaMalloc(&tmp, (size_t)trows * h->d * sizeof(float)));
    long long done = 0;
    int rc = PSX_OK;
    while (done < n && rc == PSX_OK) {
        const long long m = std::min(trows, n - done);
        const long long total = m * h->d;
        unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 32);
        const unsigned char* src = h->x + (size_t)(row0 + done) * h->row_bytes;
        if (has_fp32_rows(h))
            unpack_rows_kernel<float><<<blocks, 256, 0, h->stream>>>(fp32_rows(h) + (size_t)(row0 + done) * fp32_ld(h), tmp, m, h->d,
                                                                    fp32_ld(h));
        else
            unpack_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, h->stream>>>((const __nv_bfloat16*)src, tmp, m, h->d, h->ld);
        g_launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(out + (size_t)done * h->d, tmp, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(PSX_ERR_CUDA, "row read-back failed: %s", cudaGetErrorString(e));
        done += m;
    }
    cudaFree(tmp);
    return rc;
}

extern "C" int psx_read_rows(psx_index* h, int64_t row0, int64_t n, float* out) {
    if (!h || row0 < 0 || n < 0 || (!out && n > 0)) return fail(PSX_ERR_INVALID, "bad arguments to psx_read_rows");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (row0 + n > h->n) return fail(PSX_ERR_RANGE, "rows [%lld,%lld) exceed ntotal %lld", (long long)row0, (long long)(row0 + n), h->n.load());
    if (is_group(h)) return group_read_rows(h, row0, n, out);
    return read_rows_locked(h, row0, n, out);
}

extern "C" int psx_reconstruct(psx_index* h, int64_t id, float* out) {
    if (!h || !out) return fail(PSX_ERR_INVALID, "bad arguments to psx_reconstruct");
    std::lock_guard<std::mutex> lk(h->mu);
    if (id < 0 || id >= h->n + h->pending_n) return fail(PSX_ERR_RANGE, "id %lld not in [0,%lld)", (long long)id, (long long)(h->n + h->pending_n));
    if (id >= h->n && has_fp32_rows(h)) {  // still staged on the host, stored precision == fp32
        std::lock_guard<std::mutex> pk(h->pmu);
        const long long off = id - h->n;
        if (off >= 0 && off < h->pending_n) {
            memcpy(out, h->pending.data() + (size_t)off * h->d, (size_t)h->d * sizeof(float));
            return PSX_OK;
        }
    }
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (is_group(h)) return group_read_rows(h, id, 1, out);
    return read_rows_locked(h, id, 1, out);
}

extern "C" int psx_storage_device(psx_index* h, const void** rows_dev, int64_t* ld_elems, int* store_dtype) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    NOT_ON_GROUP(h, "psx_storage_device");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = flush_pending(h);
    if (rc) return rc;
    if (rows_dev) *rows_dev = h->x;
    if (ld_elems) *ld_elems = h->ld;
    if (store_dtype) *store_dtype = scan_dtype(h);
    return PSX_OK;
}

extern "C" int psx_hybrid_fuse_device(int device, int64_t nq, int64_t kv, const float* vec_dist_dev, const int64_t* vec_ids_dev,
                                      const double* vec_boost_dev, int64_t kw, const int64_t* kw_ids_dev,
                                      const double* kw_scores_dev, const double* kw_boost_dev, double vector_weight,
                                      double keyword_weight, int metric, int allow_keyword_only, int keyword_filtered,
                                      int64_t* out_ids_dev, double* out_fused_dev, double* out_vscore_dev,
                                      double* out_kscore_dev, int* out_count_dev, void* stream) {
    if (nq < 0 || kv < 0 || kw < 0 || kv + kw < 1 || kv + kw > FUSE_MAX_ENTRIES)
        return fail(PSX_ERR_INVALID, "hybrid fusion needs 1 <= kv + kw <= %d", FUSE_MAX_ENTRIES);
    if ((kv && (!vec_dist_dev || !vec_ids_dev)) || (kw && (!kw_ids_dev || !kw_scores_dev)) || !out_ids_dev || !out_fused_dev ||
        !out_vscore_dev || !out_kscore_dev || !out_count_dev)
        return fail(PSX_ERR_INVALID, "null pointer passed to psx_hybrid_fuse_device");
    if (nq == 0) return PSX_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(PSX_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    FuseParams p;
    p.vec_dist = vec_dist_dev;
    p.vec_ids = (const long long*)vec_ids_dev;
    p.vec_boost = vec_boost_dev;
    p.kw_ids = (const long long*)kw_ids_dev;
    p.kw_scores = kw_scores_dev;
    p.kw_boost = kw_boost_dev;
    p.out_ids = (long long*)out_ids_dev;
    p.out_fused = out_fused_dev;
    p.out_vscore = out_vscore_dev;
    p.out_kscore = out_kscore_dev;
    p.out_count = out_count_dev;
    p.kv = (int)kv;
    p.kw = (int)kw;
    p.wv = vector_weight;
    p.wk = keyword_weight;
    p.metric = metric;
    p.allow_keyword_only = allow_keyword_only;
    p.keyword_filtered = keyword_filtered;
    const int E = (int)(kv + kw);
    int np = 64;
    while (np < E) np <<= 1;
    const size_t smem = (size_t)np * 8 + (size_t)E * 24 + (size_t)kv * 4 + 16;
    static std::atomic<bool> ready[64];
    if (device >= 0 && device < 64 && !ready[device].load()) {
        CU(cudaFuncSetAttribute(hybrid_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSX_SMEM_LIMIT));
        ready[device].store(true);
    }
    hybrid_fuse_kernel<<<(unsigned)nq, 256, smem, (cudaStream_t)stream>>>(p);
    g_launches++;
    CU(cudaGetLastError());
    return PSX_OK;
}

extern "C" int psx_finalize_device(int device, int64_t nq, int64_t m, const double* scores_dev, const int* counts_dev, int top_k,
                                   double strict_floor, double broad_floor, double threshold_floor, double* out_strict_dev,
                                   double* out_broad_dev, int* out_bucket_dev, int* out_counts_dev, void* stream) {
    if (nq < 0 || m < 1 || top_k < 1 || !scores_dev || !counts_dev || !out_strict_dev || !out_broad_dev || !out_bucket_dev || !out_counts_dev)
        return fail(PSX_ERR_INVALID, "bad arguments to psx_finalize_device");
    if (nq == 0) return PSX_OK;
    DeviceGuard g(device);
    if (!g.ok) return fail(PSX_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    FinalizeParams p;
    p.scores = scores_dev;
    p.counts = counts_dev;
    p.m = (int)m;
    p.top_k = top_k;
    p.strict_floor = strict_floor;
    p.broad_floor = broad_floor;
    p.threshold_floor = threshold_floor;
    p.out_strict = out_strict_dev;
    p.out_broad = out_broad_dev;
    p.out_bucket = out_bucket_dev;
    p.out_counts = out_counts_dev;
    finalize_kernel<<<(unsigned)nq, 128, 0, (cudaStream_t)stream>>>(p);
    g_launches++;
    CU(cudaGetLastError());
    return PSX_OK;
}

extern "C" double psx_upload_gbps(psx_index* h) {
    if (!h) return 0.0;
    double best = h->up_last_gbps;
    for (psx_index* c : h->shards) best = std::max(best, c->up_last_gbps);
    return best;
}

extern "C" int psx_batch_supported(psx_index* h, int64_t k) {
    if (!h || is_group(h)) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    if (flush_pending(h) != PSX_OK) return 0;
    return batch_shape_ok(h, k) ? 1 : 0;
}

extern "C" int psx_exchange_status(psx_index* h, int* status) {
    if (!h || !status) return fail(PSX_ERR_INVALID, "bad arguments to psx_exchange_status");
    std::lock_guard<std::mutex> lk(h->mu);
    *status = h->xstatus_host ? *h->xstatus_host : 0;
    if (h->xstatus_host) *h->xstatus_host = 0;
    return PSX_OK;
}

extern "C" int psx_batch_stats(psx_index* h, int64_t* queries, int64_t* fallbacks) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    long long bq = h->batch_queries, bf = h->batch_fallbacks;
    // a group counts a query once (every child serves every query of a batch) and a fallback per child re-run
    if (is_group(h)) {
        for (psx_index* c : h->shards) {
            bq = std::max(bq, (long long)c->batch_queries);
            bf += c->batch_fallbacks;
        }
    }
    if (queries) *queries = bq;
    if (fallbacks) *fallbacks = bf;
    return PSX_OK;
}

extern "C" int psx_set_trace_device(psx_index* h, uint64_t* trace_dev) {
    if (!h) return fail(PSX_ERR_INVALID, "null handle");
    NOT_ON_GROUP(h, "psx_set_trace_device");
    std::lock_guard<std::mutex> lk(h->mu);
    h->trace = (unsigned long long*)trace_dev;
    return PSX_OK;
}

extern "C" int psx_set_tunable(psx_index* h, const char* key, int value) {
    if (!h || !key) return fail(PSX_ERR_INVALID, "bad arguments to psx_set_tunable");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!strcmp(key, "warps")) {
        h->warps = value <= 0 ? 16 : std::min(value, PSX_MAX_WARPS);
    } else if (!strcmp(key, "stages")) {
        h->stages = value <= 0 ? 2 : std::max(2, std::min(value, 12));
        h->stages_auto = value <= 0;
    } else if (!strcmp(key, "ctas_per_sm")) {
        h->ctas_per_sm = value <= 0 ? 1 : std::min(value, 8);
    } else if (!strcmp(key, "pdl")) {  // 0 never, 1 within one call (default), 2 also across calls (see psx_index::pdl)
        h->pdl = value < 0 ? 1 : std::min(value, 2);
    } else if (!strcmp(key, "deal")) {  // unfiltered scans: 1 = dealt units (default), 0 = static groups
        h->deal = value != 0;
    } else if (!strcmp(key, "dyn_tail")) {  // dealt units: 1 = dynamic tail (default), 0 = all static
        h->dyn_tail = value != 0;
    } else if (!strcmp(key, "static_batch")) {
        h->static_batch = value <= 0 ? 8 : std::min(value, 32);
    } else if (!strcmp(key, "filter_mode")) {  // 0 auto (3 up to 4M rows, else 2), 1 predicate inside the scan, 2 row list by a kernel, 3 row list by the scan
        h->filter_mode = value < 0 || value > 3 ? 0 : value;
    } else if (!strcmp(key, "batch_pair")) {
        h->batch_pair = value > 0;
    } else if (!strcmp(key, "batch_fold")) {  // 1 = threshold sample folded into the filter kernel where possible (default), 0 = separate passes
        h->batch_fold = value != 0;
    } else if (!strcmp(key, "batch_pdl")) {  // 1 = the kernels of a batch overlap by programmatic dependent launch (default), 0 = plain stream order
        h->batch_pdl = value != 0;
    } else if (!strcmp(key, "batch_bf16")) {  // bf16+master indexes: 1 = bf16 GEMM over the bf16 rows (default), 0 = TF32 GEMM over the master
        h->batch_bf16 = value > 0;
    } else if (!strcmp(key, "batch_min")) {  // smallest nq sent to the tensor-core path; 0 disables it
        h->batch_min = value < 0 ? 4 : value;
    } else if (!strcmp(key, "xchg_inline")) {  // 1 = cross-rank merge fused into the scan's last CTA (default), 0 = separate merge kernel
        h->xchg_inline = value != 0;
    } else if (!strcmp(key, "xchg_timeout_ms")) {  // bounded wait of the fused exchange's merge kernel
        h->xchg_timeout_ms = value > 0 ? value : 0;
    } else if (!strcmp(key, "fault_skip_publish")) {  // test hook: shard `value` of a multi-device handle stays silent
        h->fault_skip_publish = value;
        return PSX_OK;
    } else if (!strcmp(key, "shard_min_rows") && is_group(h)) {
        // multi-device handles: a corpus is never split into shards smaller than this (default 8192)
    } else if (!strcmp(key, "shard_min_rows")) {
        return PSX_OK;  // meaningless on one device; accepted so that callers need not care
    } else {
        return fail(PSX_ERR_INVALID, "unknown tunable '%s'", key);
    }
    if (is_group(h)) return group_set_tunable(h, key, value);
    return PSX_OK;
}

// ------------------------------------------------------------------------------------------
// one handle over several devices (psx_create_sharded)
// ------------------------------------------------------------------------------------------
#include "psx_group.cuh"
