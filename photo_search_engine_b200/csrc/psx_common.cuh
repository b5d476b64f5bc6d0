// psx_common.cuh -- shared device helpers: sortable (score,id) keys, mbarrier / bulk-copy PTX,
// block-wide bitonic primitives.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/psx.h"

#define PSX_SLOT_BYTES 4096      // one ring slot = one bulk copy of at most this many bytes
#define PSX_MAX_WARPS 16
#define PSX_MAX_THREADS (PSX_MAX_WARPS * 32)
#define PSX_SMEM_LIMIT (227 * 1024)
#define PSX_XCHG_MAX_WORLD 8

namespace psx {

// ---------------------------------------------------------------------------------------
// sortable keys.  A hit is (score fp32, row id u32); "better" = higher score, then lower id.
// key = orderable(score) << 32 | ~id  makes "better" == "larger unsigned 64-bit integer", so the
// per-CTA selection, the cross-CTA merge and the cross-GPU merge are all plain integer
// compare-exchange networks and ties resolve identically everywhere.  key 0 = empty slot
// (the smallest key of a real hit is 0x007FFFFF'xxxxxxxx, the image of -inf).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_to_ord(float s) {
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
    if (score != score) score = -INFINITY;  // NaN rows sort last
    score += 0.0f;                          // -0.0 -> +0.0 so that equal floats give equal keys
    return ((uint64_t)f32_to_ord(score) << 32) | (uint32_t)(~id);
}
__device__ __forceinline__ float key_score(uint64_t key) { return ord_to_f32((uint32_t)(key >> 32)); }
__device__ __forceinline__ uint32_t key_id(uint64_t key) { return ~(uint32_t)key; }

// ---------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D, SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on the phase with the given parity.  try_wait suspends in hardware between polls; the
// poll bound turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
// global -> shared bulk copy, completion signalled on `bar` (bytes: multiple of 16, both
// addresses 16-byte aligned).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// the same with an L2 eviction-priority hint (createpolicy): rows streamed once should leave L2 before anything else
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------
// block-wide sorting networks over 64-bit keys in shared memory (descending)
// ---------------------------------------------------------------------------------------
// Compare-exchange strides 32..1 of the bitonic network, for the phases size_from..size_to, on
// 64-element chunks held two keys per lane: no shared-memory round trips and no block barriers.
__device__ __forceinline__ void warp_bitonic_strides_le32(uint64_t* buf, int n, int size_from, int size_to) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int base = warp * 64; base < n; base += nwarps * 64) {
        const int i0 = base + lane, i1 = i0 + 32;
        uint64_t e0 = buf[i0], e1 = buf[i1];
        for (int sz = size_from; sz <= size_to; sz <<= 1) {
            const bool d0 = (i0 & sz) == 0, d1 = (i1 & sz) == 0;  // descending blocks
            if (sz >= 64) {  // stride 32: partner is the other key of this lane
                const uint64_t hi = e0 > e1 ? e0 : e1, lo = e0 > e1 ? e1 : e0;
                e0 = d0 ? hi : lo;
                e1 = d0 ? lo : hi;
            }
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                if (st < sz) {
                    const bool lower = (lane & st) == 0;
                    const uint64_t p0 = __shfl_xor_sync(0xffffffffu, e0, st);
                    const uint64_t p1 = __shfl_xor_sync(0xffffffffu, e1, st);
                    e0 = ((e0 > p0) == (lower == d0)) ? e0 : p0;
                    e1 = ((e1 > p1) == (lower == d1)) ? e1 : p1;
                }
            }
        }
        buf[i0] = e0;
        buf[i1] = e1;
    }
}

// Full bitonic sort of buf[0..n) in shared memory, descending, n a power of two.  All threads of
// the block must call.  Strides >= 64 go through shared memory with a block barrier each; the
// six innermost strides of every phase run in registers with warp shuffles, so a 1024-key sort
// needs 15 barriers instead of 55.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* buf, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n < 64) {  // tiny: plain network
        for (int size = 2; size <= n; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (n >> 1); t += nt) {
                    const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                    const int j = i | stride;
                    const bool desc = (i & size) == 0;
                    const uint64_t a = buf[i], b = buf[j];
                    if ((a < b) == desc) {
                        buf[i] = b;
                        buf[j] = a;
                    }
                }
                __syncthreads();
            }
        }
        return;
    }
    warp_bitonic_strides_le32(buf, n, 2, 64);
    __syncthreads();
    for (int size = 128; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride >= 64; stride >>= 1) {
#pragma unroll 4
            for (int t = tid; t < (n >> 1); t += nt) {
                const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int j = i | stride;
                const bool desc = (i & size) == 0;
                const uint64_t a = buf[i], b = buf[j];
                if ((a < b) == desc) {
                    buf[i] = b;
                    buf[j] = a;
                }
            }
            __syncthreads();
        }
        warp_bitonic_strides_le32(buf, n, size, size);
        __syncthreads();
    }
}

// Merge `L` descending lists of `kp` keys each (kp a power of two >= 2, padded with 0) living in
// global memory at src[l*kp + i] into the `kp` best keys, descending, left in buf[0..kp).
// `buf` is shared memory with room for cap_lists*kp keys (cap_lists >= 2).  Uses only
// compare-exchange of whole keys, so the result is the exact (score desc, id asc) order.
__device__ __forceinline__ void block_merge_lists(const uint64_t* __restrict__ src, int L, int kp, uint64_t* buf,
                                                  int cap_lists) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int half = kp >> 1;
    const int lg = 31 - __clz(kp);  // kp is a power of two
    bool have_acc = false;
    int pos = 0;
    while (pos < L) {
        const int base = have_acc ? 1 : 0;
        int nb = cap_lists - base;
        if (nb > L - pos) nb = L - pos;
        for (int idx = tid; idx < nb * kp; idx += nt) buf[base * kp + idx] = ld_cg_u64(src + (size_t)pos * kp + idx);
        __syncthreads();
        const int m = base + nb;
        for (int stride = 1; stride < m; stride <<= 1) {
            // pairs (a = p*2*stride, b = a + stride) with b < m
            const int npairs = (m - stride + 2 * stride - 1) / (2 * stride);
            // half-cleaner across the pair: the kp largest of A u B, as a bitonic sequence in A
#pragma unroll 4
            for (int w = tid; w < npairs * kp; w += nt) {
                const int p = w >> lg, i = w & (kp - 1);
                uint64_t* A = buf + (size_t)(p * 2 * stride) * kp;
                const uint64_t* B = A + (size_t)stride * kp;
                const uint64_t x = A[i], y = B[kp - 1 - i];
                A[i] = x > y ? x : y;
            }
            __syncthreads();
            for (int j = half; j >= 1; j >>= 1) {
#pragma unroll 4
                for (int w = tid; w < npairs * half; w += nt) {
                    const int p = w >> (lg - 1), t = w & (half - 1);
                    uint64_t* A = buf + (size_t)(p * 2 * stride) * kp;
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const uint64_t a = A[i], b = A[i | j];
                    if (a < b) {
                        A[i] = b;
                        A[i | j] = a;
                    }
                }
                __syncthreads();
            }
        }
        have_acc = true;
        pos += nb;
    }
}

// Exact top-k of `L` descending lists WITHOUT merging them all: gather only the first m keys of
// every list, sort those, and accept the result once no list's last gathered key still beats the
// k-th gathered key (every deeper key of a list is smaller than its last gathered one, so it can
// not enter the top-k).  m starts near 2k/L (at least 4, then widened to fill the power of two the
// sort pads to anyway) and doubles on failure; data dealt evenly to the CTAs passes on the first try
// (k = 100 over 148 lists: m = 6, a 1024-key sort, retry probability ~1 %), so the cost is ~one small
// sort whatever k is.  Falls back to the full merge tree when the prefixes outgrow `cap_keys`.
// Result in buf[0..kp).
__device__ __forceinline__ void block_select_from_lists(const uint64_t* __restrict__ src, int L, int k, int kp, uint64_t* buf,
                                                        int cap_keys) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int m = (2 * k + L - 1) / L;
    if (m < 4) m = 4;
    if (m > kp) m = kp;
    for (;;) {
        long long total = (long long)L * m;
        if (total > cap_keys) break;
        int np = kp;
        while (np < total) np <<= 1;
        if (np > cap_keys) break;
        if (np / L > m) {  // the sort costs the same up to np keys: take deeper prefixes for free
            m = np / L < kp ? np / L : kp;
            total = (long long)L * m;
        }
        // the last gathered key of every list stays in a register of the thread that fetched it
        uint64_t tail_key = 0ull;
        for (int idx = tid; idx < np; idx += nt) {
            uint64_t v = 0ull;
            if (idx < total) {
                const int l = idx / m, j = idx - l * m;
                v = ld_cg_u64(src + (size_t)l * kp + j);
                if (j == m - 1 && v > tail_key) tail_key = v;
            }
            buf[idx] = v;
        }
        __syncthreads();
        block_bitonic_sort_desc(buf, np);
        const uint64_t kth = total >= k ? buf[k - 1] : 0ull;
        const int more = m < kp && tail_key > kth;
        if (!__syncthreads_or(more)) return;  // buf[0..kp) holds the answer (np >= kp)
        m = m * 2 > kp ? kp : m * 2;
    }
    block_merge_lists(src, L, kp, buf, cap_keys / kp);
}

// Decode the first k keys of a sorted list into FAISS-shaped outputs.
__device__ __forceinline__ void block_emit_results(const uint64_t* sorted, int k, int kp, int metric, float* out_scores,
                                                   long long* out_ids, uint64_t* out_keys) {
    for (int i = threadIdx.x; i < kp; i += blockDim.x) {
        const uint64_t key = i < k ? sorted[i] : 0ull;
        if (out_keys) out_keys[i] = key;
        if (i < k) {
            float s = key ? key_score(key) : -INFINITY;
            if (metric == PSX_METRIC_L2) s = -s;
            if (out_scores) out_scores[i] = s;
            if (out_ids) out_ids[i] = key ? (long long)key_id(key) : -1ll;
        }
    }
}

}  // namespace psx
