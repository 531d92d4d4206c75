// Shared device/host helpers for libtu_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tu_b200.h"

namespace tu {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
void count_launch();   // every kernel launch of the library is counted (tu_launch_count)

#define TU_CHECK_ARG(cond, msg)                         \
    do {                                                \
        if (!(cond)) {                                  \
            ::tu::set_error(std::string("tu: ") + msg); \
            return TU_ERR_ARG;                          \
        }                                               \
    } while (0)

#define TU_CHECK_LAUNCH(what)                                   \
    do {                                                        \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) return ::tu::cuda_fail(_e, what); \
        ::tu::count_launch();                                   \
    } while (0)

// ---- scalar load/store with dtype conversion --------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
// uint8 frames: ToTensor semantics on the way in (x / 255 in fp32, inference.py:65-70, app_overlay.py:298) and
// (out * 255).clamp(0, 255).to(uint8) on the way out (truncation, app_overlay.py:383)
// exact uint8 -> float without the conversion pipe: 2^23 + v has v in its low mantissa bits
__device__ __forceinline__ float u8_to_float(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.f; }
__device__ __forceinline__ float to_f(uint8_t v) { return u8_to_float(v) / 255.f; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
// trunc(clamp(v * 255, 0, 255)): adding 2^23 with round-down leaves floor(t) in the low mantissa bits (t >= 0)
template <> __device__ __forceinline__ uint8_t from_f<uint8_t>(float v) {
    return (uint8_t)(__float_as_uint(__fadd_rd(fminf(fmaxf(v * 255.f, 0.f), 255.f), 8388608.f)) & 0xFFu);
}

// 4 consecutive elements -> float4 (pointer must be aligned to 4 elements)
__device__ __forceinline__ float4 load4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 load4(const bf16 *p) {
    uint2 u = *reinterpret_cast<const uint2 *>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&u.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162 *>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void store4(bf16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&a);
    u.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = u;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// Every kernel of a forward is launched with programmatic stream serialization: it may become resident (and run its prologue:
// barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel of the stream is draining.  pdl_wait() blocks
// until that kernel has completed and its writes are visible, so it must precede EVERY access to global memory; kernels call
// pdl_trigger() first thing so that their own successor can be scheduled early.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_use_pdl;      // api.cu (debug key "pdl")
extern int g_bicubic_pair; // resample.cu (debug key "bicubic_pair")
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline size_t dtype_size(int dtype) { return dtype == TU_BF16 ? 2 : dtype == TU_U8 ? 1 : 4; }

}  // namespace tu
