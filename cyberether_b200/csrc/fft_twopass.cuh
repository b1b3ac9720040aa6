// Two-pass plan for power-of-two transforms above the single-CTA limit: n = 16 M, 1024 <= M <= 8192
// (n = 16384 ... 131072). Replaces the six-pass four-step plan (three transposes + twiddle pass + two batched
// sub-transforms) for those lengths; the reference runs pocketfft::c2c / cuFFT here
// (src/domains/dsp/fft/module_impl_native_cpu.cc:129-140, module_impl_native_cuda.cc:321-333,433).
//
// Decimation in frequency with the SMALL factor first, so the strided side is a column access that coalesces by itself:
//   x[n1 M + n2], n1 < 16, n2 < M;  k = k1 + 16 k2
//   pass 1 (this file, fft_col16_kernel):  z[k1][n2] = W_n^(n2 k1) * sum_n1 x[n1 M + n2] W_16^(n1 k1)
//       one thread per column n2: 16 coalesced 8-byte loads (a warp covers 256 contiguous bytes of each of the 16
//       slabs), radix-16 in registers, 15 thread-constant twiddles straight from the F64-evaluated table, 16 coalesced
//       stores into the scratch laid out [row][k1][n2] — no shared memory, no barrier.
//   pass 2 (fft_radix_kernel<log2 M, MODE_C2C_T>): X[k1 + 16 k2] = sum_n2 z[k1][n2] W_M^(n2 k2)
//       the register-radix row kernel on 16 contiguous rows of length M per transform, storing with stride 16.
//
// HBM traffic: the batch is processed in chunks whose scratch (chunk bytes) stays resident in the 126 MB L2 between the
// two launches: pass 1 reads the input from HBM and leaves z dirty in L2, pass 2 reads z from L2 and streams the result
// to HBM, and the next chunk overwrites the same scratch lines before they are evicted. Algorithmic 16 B/sample;
// without residency the plan would move 32 B/sample.
#pragma once

#include "fft_radix.cuh"

namespace b200 {

constexpr int kCol16Threads = 256;

// L2 eviction-priority policies (createpolicy): the scratch is written to stay (evict_last), the input streams
// through (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ float2 ldg_hint_f2(const float2* p, const uint64_t policy) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y)
                 : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void stg_hint_f2(float2* p, const float2 v, const uint64_t policy) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y),
                 "l"(policy)
                 : "memory");
}

struct Col16Params {
    const float2* in;        // [rows][16][M]
    float2* scratch;         // [rows][16][M]
    uint64_t rows;
    uint32_t m;              // M (power of two, multiple of kCol16Threads)
    int inverse;             // swap(re, im) on the way in; pass 2 swaps on the way out
    const float2* twiddle;   // W_n^j, j < n = 16 M
    int hints;               // 1: L2 eviction-priority hints on the loads / stores
};

template <bool HINTS>
__global__ void __launch_bounds__(kCol16Threads, 2) fft_col16_kernel(const Col16Params p) {
    const uint32_t col = blockIdx.x * kCol16Threads + threadIdx.x;     // n2; gridDim.x * 256 == M
    const uint64_t m = p.m;
    const uint64_t n = 16 * m;
    float2 tw[15];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        tw[k - 1] = p.twiddle[static_cast<uint64_t>(col) * k];          // col k < n: no wrap
    }
    uint64_t pol_in = 0, pol_out = 0;
    if constexpr (HINTS) {
        pol_in = l2_policy_evict_first();
        pol_out = l2_policy_evict_last();
    }
    for (uint64_t row = blockIdx.y; row < p.rows; row += gridDim.y) {
        const float2* const src = p.in + row * n + col;
        float2 v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            if constexpr (HINTS) {
                v[a] = ldg_hint_f2(src + a * m, pol_in);
            } else {
                v[a] = ldg_stream_f2(src + a * m);
            }
        }
        if (p.inverse) {
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                v[a] = make_float2(v[a].y, v[a].x);
            }
        }
        dft16(v);
        float2* const dst = p.scratch + row * n + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float2 y = k == 0 ? v[dft16_pos(0)] : cmul(v[dft16_pos(k)], tw[k - 1]);
            if constexpr (HINTS) {
                stg_hint_f2(dst + k * m, y, pol_out);
            } else {
                dst[k * m] = y;
            }
        }
    }
}

}  // namespace b200
