// Shared helpers for all avsr_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avsr_b200.h"

// Implemented in capi.cu
void avsr_set_error(const char* fmt, ...);

#define AVSR_CHECK_CUDA(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            avsr_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,               \
                           cudaGetErrorString(_e));                                             \
            return AVSR_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define AVSR_REQUIRE(cond, ...)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            avsr_set_error(__VA_ARGS__);                                                        \
            return AVSR_ERR_ARG;                                                                \
        }                                                                                       \
    } while (0)

#define AVSR_LAUNCH_CHECK()                                                                     \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            avsr_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,               \
                           cudaGetErrorString(_e));                                             \
            return AVSR_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): the kernels of one decode position form a chain in which every kernel is
// launched while its predecessor is still running; it may do work that does not depend on the predecessor (barrier /
// TMEM set-up, weight prefetch) and must call pdl_wait() before it touches anything an earlier kernel of the chain wrote
// or still reads.  Data written before the chain started (weights, cross-attention K/V, posteriors) needs no wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t avsr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reductions; `red` is a shared array of >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// erf-GELU for GEMM epilogues that are bound by instruction issue (erff costs ~60 instructions; the epilogue of a
// 128x256 tile has ~45 per element before it outlasts the MMAs): Abramowitz-Stegun 7.1.26,
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1 / (1 + p z),  z >= 0,   |error| <= 1.5e-7 (+ MUFU rounding),
// i.e. fp32-rounding level for a result that is stored as bf16.  2 MUFU + ~14 FMA-pipe instructions.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));      // one MUFU each (no IEEE fix-up code)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erf_abs = fmaf(-p * t, e, 1.f);              // erf(|x| / sqrt 2)
    const float hx = 0.5f * x;
    return fmaf(0.5f * ax, erf_abs, hx);                     // 0.5 x (1 + sign(x) erf_abs)
}

// fp32 -> three bf16 terms (x = a1 + a2 + a3 to ~2^-24 relative) laid out for the "bf16x3" tensor-core GEMM that keeps
// fp32-level accuracy on the decode side: the row [a1 | a1 | a2 | a1 | a2 | a3] (6 blocks of K) is multiplied with the
// weight row [w1 | w2 | w1 | w3 | w2 | w1], i.e. the six largest cross terms of (a1+a2+a3)(w1+w2+w3).
__device__ __forceinline__ void avsr_split3_store(__nv_bfloat16* row_base, int K, int c, float v) {
    const __nv_bfloat16 a1 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(a1);
    const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
    row_base[c] = a1; row_base[K + c] = a1; row_base[3 * K + c] = a1;
    row_base[2 * K + c] = a2; row_base[4 * K + c] = a2;
    row_base[5 * K + c] = a3;
}

// Four consecutive columns at once (c % 4 == 0, K % 4 == 0): six 8-byte stores instead of 24 two-byte ones.
__device__ __forceinline__ void avsr_split3_store4(__nv_bfloat16* row_base, int K, int c, float4 v) {
    const float in[4] = {v.x, v.y, v.z, v.w};
    __align__(8) __nv_bfloat16 t1[4], t2[4], t3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        t1[i] = __float2bfloat16_rn(in[i]);
        const float r1 = in[i] - __bfloat162float(t1[i]);
        t2[i] = __float2bfloat16_rn(r1);
        t3[i] = __float2bfloat16_rn(r1 - __bfloat162float(t2[i]));
    }
    const uint2 u1 = *reinterpret_cast<const uint2*>(t1), u2 = *reinterpret_cast<const uint2*>(t2), u3 = *reinterpret_cast<const uint2*>(t3);
    *reinterpret_cast<uint2*>(row_base + c) = u1;
    *reinterpret_cast<uint2*>(row_base + K + c) = u1;
    *reinterpret_cast<uint2*>(row_base + 3 * K + c) = u1;
    *reinterpret_cast<uint2*>(row_base + 2 * K + c) = u2;
    *reinterpret_cast<uint2*>(row_base + 4 * K + c) = u2;
    *reinterpret_cast<uint2*>(row_base + 5 * K + c) = u3;
}

// Compact form for the decoder step's own GEMM (csrc/gemm_x3.cu): row = [a1 | a2 | a3] (3 blocks of K); the six cross
// terms are formed by six MMAs per k step instead of by repeating the terms along K.
__device__ __forceinline__ void avsr_split3c_store(__nv_bfloat16* row_base, int K, int c, float v) {
    const __nv_bfloat16 a1 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(a1);
    const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
    row_base[c] = a1; row_base[K + c] = a2; row_base[2 * K + c] = a3;
}
__device__ __forceinline__ void avsr_split3c_store4(__nv_bfloat16* row_base, int K, int c, float4 v) {
    const float in[4] = {v.x, v.y, v.z, v.w};
    __align__(8) __nv_bfloat16 t1[4], t2[4], t3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        t1[i] = __float2bfloat16_rn(in[i]);
        const float r1 = in[i] - __bfloat162float(t1[i]);
        t2[i] = __float2bfloat16_rn(r1);
        t3[i] = __float2bfloat16_rn(r1 - __bfloat162float(t2[i]));
    }
    *reinterpret_cast<uint2*>(row_base + c) = *reinterpret_cast<const uint2*>(t1);
    *reinterpret_cast<uint2*>(row_base + K + c) = *reinterpret_cast<const uint2*>(t2);
    *reinterpret_cast<uint2*>(row_base + 2 * K + c) = *reinterpret_cast<const uint2*>(t3);
}

__device__ __forceinline__ float avsr_apply_act(float v, int act, float slope) {
    if (act == AVSR_ACT_GELU) return gelu_erf_fast(v);
    if (act == AVSR_ACT_RELU) return fmaxf(v, 0.f);
    if (act == AVSR_ACT_PRELU) return v >= 0.f ? v : v * slope;
    return v;
}
