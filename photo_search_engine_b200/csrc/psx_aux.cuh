// psx_aux.cuh -- K0: row packing (fp32 -> stored dtype, zero padding to the row stride, optional
// L2 normalisation as utils/vector_store.py:83-90) and its inverse used by reconstruct / save.
// Pure streaming kernels: one warp per row, 128-bit accesses where the layout allows.
#pragma once
#include "psx_common.cuh"

namespace psx {

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// src [n][d] fp32 (dense)  ->  dst [n][ld] T (zero padded).  normalize: x / ||x||_2 in fp32, rows
// of zero norm are stored unchanged.
// max_sumsq (device, may be null) receives max over rows of ||stored row||^2 (atomicMax on the bits
// of a non-negative float): the batched path scales its TF32 error bound with it.
template <typename T>
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n,
                                                        int d, int ld, int normalize, float* max_sumsq) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; row < n; row += warps) {
        const float* s = src + row * d;
        T* o = dst + row * ld;
        float inv_is_div = 1.0f;
        bool scale = false;
        float acc = 0.f;
        for (int i = lane; i < d; i += 32) acc = fmaf(s[i], s[i], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (normalize) {
            const float norm = sqrtf(acc);
            if (norm != 0.f) {
                inv_is_div = norm;
                scale = true;
                acc = 1.0f;
            }
        }
        if (max_sumsq && lane == 0 && acc == acc) atomicMax(reinterpret_cast<int*>(max_sumsq), __float_as_int(acc));
        for (int i = lane; i < ld; i += 32) {
            float v = i < d ? s[i] : 0.f;
            if (scale) v = __fdiv_rn(v, inv_is_div);
            o[i] = from_f32<T>(v);
        }
    }
}

// src [n][ld] T  ->  dst [n][d] fp32
template <typename T>
__global__ void __launch_bounds__(256) unpack_rows_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n,
                                                          int d, int ld) {
    const long long total = n * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / d;
        const int c = (int)(i - row * d);
        dst[i] = to_f32(src[row * ld + c]);
    }
}

}  // namespace psx
