// Shared helpers for libb200dsp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/b200dsp.h"

struct b200_ctx {
    int device = 0;
    int sms = 148;
    int max_smem_optin = 0;
};

namespace b200 {

// Thread-local error text (b200_last_error()).
std::string& last_error();
int fail(const char* fmt, ...);

#define B200_CUDA_CHECK(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            return ::b200::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                                __FILE__, __LINE__);                                           \
        }                                                                                      \
    } while (0)

#define B200_REQUIRE(cond, ...)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            return ::b200::fail(__VA_ARGS__);                                                  \
        }                                                                                      \
    } while (0)

#define B200_LAUNCH_CHECK() B200_CUDA_CHECK(cudaGetLastError())

// Activates the context's device for the calling thread for the duration of a scope.
struct DeviceGuard {
    int previous = -1;
    bool switched = false;
    explicit DeviceGuard(const b200_ctx* ctx) {
        if (cudaGetDevice(&previous) == cudaSuccess && previous != ctx->device) {
            switched = cudaSetDevice(ctx->device) == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched) {
            cudaSetDevice(previous);
        }
    }
};

inline cudaStream_t as_stream(b200_stream s) { return static_cast<cudaStream_t>(s); }

inline bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
inline int ilog2(uint64_t n) {
    int k = 0;
    while ((1ull << (k + 1)) <= n) ++k;
    return k;
}

// ---- device-side complex helpers -----------------------------------------------------------

// Packed FP32x2 forms (Blackwell FFMA2/FADD2/FMUL2): one instruction per complex add, two per
// complex multiply. ptxas folds the lane swap / per-lane negation / scalar broadcast below into the
// operand modifiers (.LO_HI, .NP/.PN, .F32), so none of them costs a MOV.
#ifndef B200_SCALAR_FP32
#define B200_PACKED 1
#else
#define B200_PACKED 0
#endif

__device__ __forceinline__ float2 cmul(const float2 a, const float2 b) {
    // (a.x + i a.y)(b.x + i b.y); FMA contraction allowed (FFT interior, normwise tolerance).
#if B200_PACKED
    return __ffma2_rn(a, make_float2(b.x, b.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y)));
#else
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#endif
}
__device__ __forceinline__ float2 cmul_conj(const float2 a, const float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// Same rounding sequence as std::complex<float> operator* on a non-FMA x86-64 build
// (src/domains/core/multiply/module_impl_native_cpu.cc:94-100): four products, one sub, one add.
__device__ __forceinline__ float2 cmul_exact(const float2 a, const float2 b) {
    const float ac = __fmul_rn(a.x, b.x);
    const float bd = __fmul_rn(a.y, b.y);
    const float ad = __fmul_rn(a.x, b.y);
    const float bc = __fmul_rn(a.y, b.x);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}
__device__ __forceinline__ float2 cadd(const float2 a, const float2 b) {
#if B200_PACKED
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 csub(const float2 a, const float2 b) {
#if B200_PACKED
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a - i*b and a + i*b
__device__ __forceinline__ float2 csub_i(const float2 a, const float2 b) {
#if B200_PACKED
    return __fadd2_rn(a, make_float2(b.y, -b.x));
#else
    return make_float2(a.x + b.y, a.y - b.x);
#endif
}
__device__ __forceinline__ float2 cadd_i(const float2 a, const float2 b) {
#if B200_PACKED
    return __fadd2_rn(a, make_float2(-b.y, b.x));
#else
    return make_float2(a.x - b.y, a.y + b.x);
#endif
}
// s * a for a real scalar s
__device__ __forceinline__ float2 cscale(const float2 a, const float s) {
#if B200_PACKED
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}

// Streaming (read-once / write-once) global accesses: keep L1 for the window/twiddle tables.
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_f2(float2* p, const float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream_f1(float* p, const float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- amplitude / range scalar math (shared by the standalone modules and the fused epilogue) --

// Backend::ApproxLog10 (include/jetstream/backend/devices/cpu/helpers.hh:61-74), same operation
// order, every step individually rounded (the reference CPU build has no FMA).
__device__ __forceinline__ float approx_log10_exact(const float x) {
    int e;
    const float f = frexpf(fabsf(x), &e);
    float y = 1.23149591368684f;
    y = __fmul_rn(y, f);
    y = __fadd_rn(y, -4.11852516267426f);
    y = __fmul_rn(y, f);
    y = __fadd_rn(y, 6.02197014179219f);
    y = __fmul_rn(y, f);
    y = __fadd_rn(y, -3.13396450166353f);
    y = __fadd_rn(y, static_cast<float>(e));
    return __fmul_rn(y, 0.3010299956639812f);
}

__device__ __forceinline__ float amplitude_exact(const float magnitude, const float coeff) {
    return magnitude == 0.0f ? -INFINITY
                             : __fadd_rn(__fmul_rn(20.0f, approx_log10_exact(magnitude)), coeff);
}

__device__ __forceinline__ float range_exact(const float in, const float scale, const float offset) {
    const float normalized = __fadd_rn(__fmul_rn(in, scale), offset);
    return __fadd_rn(0.5f, __fmul_rn(0.5f, tanhf(__fmul_rn(4.0f, __fsub_rn(normalized, 0.5f)))));
}

}  // namespace b200
