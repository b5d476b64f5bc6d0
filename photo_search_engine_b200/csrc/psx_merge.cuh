// psx_merge.cuh -- the small non-template kernels around the scan: K1a (EXIF predicate -> row list),
// K4 (merge of the per-shard key lists: after an all-gather, or fused with the NVLink exchange).
// Included by psx_api.cu only (the scan kernel template is instantiated in several translation units).
#pragma once
#include "psx_scan.cuh"

namespace psx {

// K1a: the EXIF predicate as a stream compaction.  Evaluates the packed attribute word of every row
// and appends the ids of the passing rows to `list` (order: by block-sized chunks, chunks in ticket
// order); *count must be 0 at launch (the scan that consumes the list resets it).  HBM-bound: 8 bytes
// read per row, 4 bytes written per passing row.
__global__ void __launch_bounds__(256) filter_list_kernel(const uint64_t* __restrict__ attrs, long long n, psx_filter f,
                                                          uint32_t* __restrict__ list, unsigned int* count, const int* cond_flag) {
    __shared__ unsigned int s_warp[8];
    __shared__ unsigned int s_base;
    if (cond_flag && *cond_flag == 0) return;  // the conditional scan this list is for will not run either
    // Programmatic dependent launch: the scan that consumes this list may become resident now (it waits for this grid
    // before it reads the list).  This kernel itself may have started while the scan BEFORE it still sorts and merges: it
    // only reads the attribute words and writes the list / counter slot that scan does not use (two alternate), and it
    // waits for that scan at its very end, so "this grid has completed" keeps implying "every earlier grid has".
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PER = 8;  // rows per thread and trip: four 16-byte loads in flight per thread
    const long long chunk = (long long)blockDim.x * PER;
    for (long long c0 = (long long)blockIdx.x * chunk; c0 < n; c0 += (long long)gridDim.x * chunk) {
        // thread t covers the row pairs c0 + h*512 + 2t, +1 for h = 0..3
        uint64_t a[PER];
#pragma unroll
        for (int h = 0; h < PER / 2; ++h) {
            const long long r = c0 + (long long)h * 2 * blockDim.x + 2 * threadIdx.x;
            a[2 * h] = a[2 * h + 1] = 0ull;
            if (r + 1 < n) {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(attrs + r));
                a[2 * h] = v.x;
                a[2 * h + 1] = v.y;
            } else if (r < n) {
                a[2 * h] = __ldg(attrs + r);
            }
        }
        uint32_t bits = 0;
#pragma unroll
        for (int h = 0; h < PER / 2; ++h) {
            const long long r = c0 + (long long)h * 2 * blockDim.x + 2 * threadIdx.x;
            if (r < n && attr_pass(a[2 * h], f)) bits |= 1u << (2 * h);
            if (r + 1 < n && attr_pass(a[2 * h + 1], f)) bits |= 2u << (2 * h);
        }
        const unsigned int mine = __popc(bits);
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
            for (int w = 0; w < 8; ++w) {
                const unsigned int c = s_warp[w];
                s_warp[w] = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(count, tot) : 0u;
        }
        __syncthreads();
        unsigned int pos = s_base + s_warp[warp] + incl - mine;
#pragma unroll
        for (int h = 0; h < PER / 2; ++h) {
            const uint32_t r = (uint32_t)(c0 + (long long)h * 2 * blockDim.x + 2 * threadIdx.x);
            if (bits & (1u << (2 * h))) list[pos++] = r;
            if (bits & (2u << (2 * h))) list[pos++] = r + 1;
        }
        __syncthreads();  // s_warp / s_base are reused by the next trip
    }
    pdl_wait();
}

// K4 fused, receiving side: wait until every rank's list for query `seq` has landed in this GPU's
// receive buffer, then select the global top-k.  One CTA.  The spin is bounded: a peer that never
// publishes (dead process, faulted GPU) does not hang this GPU.  With a `status` word (host-mapped
// memory) the kernel then reports the timeout there and returns without a result -- the context stays
// healthy and the host re-runs the query over the collective path; without one it traps.
__global__ void __launch_bounds__(256, 1)
merge_wait_kernel(const uint64_t* __restrict__ recv, const uint32_t* flags, int world, uint32_t seq, int k, int kpad, int cap_keys,
                  int metric, float* out_scores, long long* out_ids, uint64_t* out_keys, int* status, unsigned long long spin_limit) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ int s_timeout;
    // the next query's scan may start streaming now: it touches nothing this merge reads, and it publishes only
    // after its own pdl_wait(), i.e. after this kernel has completed
    pdl_launch_dependents();
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    const int slot = (int)(seq & 1u);
    if ((int)threadIdx.x < world) {
        const uint32_t* f = flags + slot * PSX_XCHG_MAX_WORLD + threadIdx.x;
        unsigned long long spins = 0;
        while (ld_acquire_sys_u32(f) != seq) {
            __nanosleep(64);
            if (++spins > spin_limit) {
                if (!status) __trap();
                s_timeout = 1 + (int)threadIdx.x;
                break;
            }
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (threadIdx.x == 0) {
            *status = s_timeout;  // 1 + a rank that did not publish
            __threadfence_system();
        }
        return;
    }
    const uint64_t* lists = recv + (size_t)slot * world * PSX_K_PASS_MAX;
    // lists are PSX_K_PASS_MAX apart; compact them to a kpad stride view by reading through an index map
    // (block_select_from_lists expects stride kpad): gather the heads into shared memory first
    const int tid = threadIdx.x, nt = blockDim.x;
    int np = kpad;
    while (np < world * kpad) np <<= 1;
    for (int idx = tid; idx < np; idx += nt) {
        uint64_t v = 0ull;
        if (idx < world * kpad) v = ld_cg_u64(lists + (size_t)(idx / kpad) * PSX_K_PASS_MAX + (idx % kpad));
        buf[idx] = v;
    }
    __syncthreads();
    (void)cap_keys;
    block_bitonic_sort_desc(buf, np);
    block_emit_results(buf, k, kpad, metric, out_scores, out_ids, out_keys);
}

// Standalone merge (K4 final merge of all-gathered shard lists): one CTA per query.
__global__ void __launch_bounds__(256, 1)
merge_keys_kernel(const uint64_t* __restrict__ keys, int nlists, int k, int kpad, int cap_lists, int metric,
                  float* out_scores, long long* out_ids, uint64_t* out_keys) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
    pdl_launch_dependents();  // see merge_wait_kernel
    const size_t qi = blockIdx.x;
    block_select_from_lists(keys + qi * (size_t)nlists * kpad, nlists, k, kpad, buf, cap_lists * kpad);
    block_emit_results(buf, k, kpad, metric, out_scores + qi * k, out_ids + qi * k, out_keys ? out_keys + qi * kpad : nullptr);
}

}  // namespace psx
