// Shared device/host helpers for libvarsep_sm100a (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/varsep.h"

namespace vs {

// ---------------------------------------------------------------- errors / launch accounting
extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int fail(const char* fmt, ...);
// after every launch: count it and surface launch-configuration errors
int launched(const char* what);

#define VS_REQUIRE(cond, ...) \
    do {                      \
        if (!(cond)) return ::vs::fail(__VA_ARGS__); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();                 // of the CURRENT device
int current_device();
// One flag per CUDA device: function attributes (cudaFuncSetAttribute) and occupancy facts belong to a device's
// context, so a process that drives several GPUs must establish them once PER DEVICE, not once per process.
struct DeviceOnce {
    bool done[64] = {};
    bool& flag() { const int d = current_device(); return done[(d >= 0 && d < 64) ? d : 0]; }
};

// ---------------------------------------------------------------- element access
template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void st(T* p, float v);
template <> __device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (address must be 4-element aligned)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------- activations (networks/utils.py:50-72)
// none / ReLU / LeakyReLU are evaluated branch-free as  max(z, s*z)  with s = 1 / 0 / 0.2 (loop-invariant): a switch over all
// six activations inside an unrolled epilogue compiles to one indirect branch PER ELEMENT (BRX + the exp / tanh code paths
// replicated per call site), which cost the fused-activation epilogues several times their memory time.
__device__ __forceinline__ float act_pl_slope(int act) { return act == VS_ACT_NONE ? 1.f : act == VS_ACT_RELU ? 0.f : 0.2f; }
// (the transcendental activations sit behind a real call: kept out of line, the common path stays straight-line code)
static __device__ __noinline__ float act_fwd_slow(float z, int act) {
    if (act == VS_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
    if (act == VS_ACT_TANH) return tanhf(z);
    return z > 0.f ? z : expm1f(z);                                     // VS_ACT_ELU
}
__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act <= VS_ACT_LEAKY) return fmaxf(z, act_pl_slope(act) * z);
    return act_fwd_slow(z, act);
}
static __device__ __noinline__ float act_grad_from_out_slow(float a, int act) {
    if (act == VS_ACT_SIGMOID) return a * (1.f - a);
    if (act == VS_ACT_TANH) return 1.f - a * a;
    return a > 0.f ? 1.f : a + 1.f;                                     // VS_ACT_ELU
}
static __device__ __noinline__ float act_grad_from_in_slow(float z, int act) {
    if (act == VS_ACT_SIGMOID) { float s = 1.f / (1.f + expf(-z)); return s * (1.f - s); }
    if (act == VS_ACT_TANH) { float t = tanhf(z); return 1.f - t * t; }
    return z > 0.f ? 1.f : expf(z);                                     // VS_ACT_ELU
}
// derivative expressed with the activation OUTPUT a = act(z) (all six are invertible enough for that;
// ReLU/LeakyReLU use the sign, which the in-place reference ops also do)
__device__ __forceinline__ float act_grad_from_out(float a, int act) {
    if (act <= VS_ACT_LEAKY) return a > 0.f ? 1.f : act_pl_slope(act);
    return act_grad_from_out_slow(a, act);
}
// derivative expressed with the pre-activation z
__device__ __forceinline__ float act_grad_from_in(float z, int act) {
    if (act <= VS_ACT_LEAKY) return z > 0.f ? 1.f : act_pl_slope(act);
    return act_grad_from_in_slow(z, act);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dtype dispatch on the host side
#define VS_DISPATCH_DTYPE(dtype, T, ...)                                     \
    do {                                                                     \
        if ((dtype) == VS_F32) { using T = float; __VA_ARGS__; }             \
        else if ((dtype) == VS_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else return ::vs::fail("unsupported dtype %d", (int)(dtype));        \
    } while (0)

}  // namespace vs
