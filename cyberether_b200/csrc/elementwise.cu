// Standalone per-module kernels (API parity with the reference's one-module-one-kernel graph):
// window, invert, multiply (broadcast), multiply_constant, amplitude, range, cast.
// All are HBM-bound streaming kernels: 16-byte vector accesses, grid = k * SM count, grid-stride
// loops, streaming cache hints. The arithmetic follows the reference CPU implementations
// operation by operation (no FMA contraction) so results are bit-identical for finite inputs
// wherever the reference itself is deterministic IEEE arithmetic.
#include "common.cuh"

namespace b200 {

static inline int stream_grid(const b200_ctx* ctx, uint64_t work_items, int threads, int ctas_per_sm) {
    const uint64_t needed = (work_items + threads - 1) / threads;
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * ctas_per_sm;
    return static_cast<int>(needed < cap ? (needed ? needed : 1) : cap);
}

// ---- window -------------------------------------------------------------------------------
// src/domains/dsp/window/module_impl_native_cpu.cc:20-37
__global__ void window_blackman_kernel(float2* __restrict__ out, const uint64_t n) {
    const double kPi = 3.14159265358979323846;  // JST_PI
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        if (n == 1) {
            out[0] = make_float2(1.0f, 0.0f);
            return;
        }
        const double di = static_cast<double>(i);
        const double dm = static_cast<double>(n - 1);
        // 2.0 * JST_PI * i / (N - 1): left-to-right, each step rounded.
        const double a1 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, kPi), di), dm);
        const double a2 = __ddiv_rn(__dmul_rn(__dmul_rn(4.0, kPi), di), dm);
        const double tap = __dadd_rn(__dsub_rn(0.42, __dmul_rn(0.50, cos(a1))), __dmul_rn(0.08, cos(a2)));
        out[i] = make_float2(static_cast<float>(tap), 0.0f);
    }
}

// ---- invert -------------------------------------------------------------------------------
// src/domains/dsp/invert/module_impl_native_cpu.cc:78-103
__global__ void invert_kernel(const float2* __restrict__ in, float2* __restrict__ out, const uint64_t total,
                              const uint64_t n, const uint64_t inner) {
    const double kPi = 3.14159265358979323846;
    const bool even = (n & 1ull) == 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t k = (i / inner) % n;
        const float2 v = in[i];
        if (even) {
            out[i] = (k & 1ull) ? make_float2(-v.x, -v.y) : v;
        } else {
            const double phase = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, kPi), static_cast<double>(n / 2)),
                                                     static_cast<double>(k)),
                                           static_cast<double>(n));
            const float2 w = make_float2(static_cast<float>(cos(phase)), static_cast<float>(sin(phase)));
            out[i] = cmul_exact(v, w);
        }
    }
}

// ---- multiply -----------------------------------------------------------------------------
// src/domains/core/multiply/module_impl_native_cpu.cc:86-100 over the broadcast views of
// src/domains/core/multiply/module_impl.cc:28-83.
struct BroadcastPlan {
    int rank;
    uint64_t shape[8];
    uint64_t stride_a[8];
    uint64_t stride_b[8];
};

template <typename T>
__device__ __forceinline__ T mul_op(const T a, const T b);
template <>
__device__ __forceinline__ float mul_op<float>(const float a, const float b) {
    return __fmul_rn(a, b);
}
template <>
__device__ __forceinline__ float2 mul_op<float2>(const float2 a, const float2 b) {
    return cmul_exact(a, b);
}

template <typename T>
__global__ void multiply_generic_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ c,
                                        const uint64_t total, const BroadcastPlan plan) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t rem = i, oa = 0, ob = 0;
#pragma unroll 1
        for (int d = plan.rank - 1; d >= 0; --d) {
            const uint64_t q = rem / plan.shape[d];
            const uint64_t coord = rem - q * plan.shape[d];
            rem = q;
            oa += coord * plan.stride_a[d];
            ob += coord * plan.stride_b[d];
        }
        c[i] = mul_op<T>(a[oa], b[ob]);
    }
}

// a [rows, n] contiguous, b [n] broadcast over rows (the spectrum_engine multiply), CF32.
// Each thread owns a fixed 16-byte column slot (2 complex) so its b operand stays in registers.
__global__ void multiply_rowbcast_cf32_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                              float4* __restrict__ c, const uint64_t rows,
                                              const uint64_t vec_per_row) {
    for (uint64_t col = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; col < vec_per_row;
         col += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float4 w = b[col];
        const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
        for (uint64_t r = blockIdx.y; r < rows; r += gridDim.y) {
            const float4 v = ldg_stream_f4(a + r * vec_per_row + col);
            const float2 p0 = cmul_exact(make_float2(v.x, v.y), w0);
            const float2 p1 = cmul_exact(make_float2(v.z, v.w), w1);
            stg_stream_f4(c + r * vec_per_row + col, make_float4(p0.x, p0.y, p1.x, p1.y));
        }
    }
}

__global__ void multiply_same_cf32_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                          float4* __restrict__ c, const uint64_t vecs) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < vecs;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float4 x = ldg_stream_f4(a + i), y = ldg_stream_f4(b + i);
        const float2 p0 = cmul_exact(make_float2(x.x, x.y), make_float2(y.x, y.y));
        const float2 p1 = cmul_exact(make_float2(x.z, x.w), make_float2(y.z, y.w));
        stg_stream_f4(c + i, make_float4(p0.x, p0.y, p1.x, p1.y));
    }
}

__global__ void multiply_same_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                         float4* __restrict__ c, const uint64_t vecs) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < vecs;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const float4 x = ldg_stream_f4(a + i), y = ldg_stream_f4(b + i);
        stg_stream_f4(c + i, make_float4(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y), __fmul_rn(x.z, y.z),
                                         __fmul_rn(x.w, y.w)));
    }
}

// ---- multiply_constant ----------------------------------------------------------------------
// src/domains/core/multiply_constant/module_impl_native_cpu.cc:82-100 (CF32 * F32 scalar scales
// both parts; std::complex<float> * float).
__global__ void multiply_constant_kernel(const float* __restrict__ in, float* __restrict__ out,
                                         const uint64_t count, const float constant, const bool vec) {
    const uint64_t vecs = vec ? count / 4 : 0;      // misaligned (offset view) buffers take the scalar loop
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = tid; i < vecs; i += step) {
        const float4 v = ldg_stream_f4(in4 + i);
        stg_stream_f4(out4 + i, make_float4(__fmul_rn(v.x, constant), __fmul_rn(v.y, constant),
                                            __fmul_rn(v.z, constant), __fmul_rn(v.w, constant)));
    }
    for (uint64_t i = vecs * 4 + tid; i < count; i += step) {
        out[i] = __fmul_rn(in[i], constant);
    }
}

// ---- amplitude ----------------------------------------------------------------------------
// src/domains/dsp/amplitude/module_impl_native_cpu.cc:73-99
__global__ void amplitude_cf32_kernel(const float2* __restrict__ in, float* __restrict__ out,
                                      const uint64_t count, const float coeff, const bool vec) {
    const uint64_t quads = vec ? count / 4 : 0;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    auto one = [coeff](const float re, const float im) {
        const float magnitude = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
        return amplitude_exact(magnitude, coeff);
    };
    for (uint64_t i = tid; i < quads; i += step) {
        const float4 a = ldg_stream_f4(in4 + 2 * i), b = ldg_stream_f4(in4 + 2 * i + 1);
        stg_stream_f4(out4 + i, make_float4(one(a.x, a.y), one(a.z, a.w), one(b.x, b.y), one(b.z, b.w)));
    }
    for (uint64_t i = quads * 4 + tid; i < count; i += step) {
        const float2 v = in[i];
        out[i] = one(v.x, v.y);
    }
}

__global__ void amplitude_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                     const uint64_t count, const float coeff, const bool vec) {
    const uint64_t quads = vec ? count / 4 : 0;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = tid; i < quads; i += step) {
        const float4 v = ldg_stream_f4(in4 + i);
        stg_stream_f4(out4 + i, make_float4(amplitude_exact(fabsf(v.x), coeff), amplitude_exact(fabsf(v.y), coeff),
                                            amplitude_exact(fabsf(v.z), coeff), amplitude_exact(fabsf(v.w), coeff)));
    }
    for (uint64_t i = quads * 4 + tid; i < count; i += step) {
        out[i] = amplitude_exact(fabsf(in[i]), coeff);
    }
}

// ---- range --------------------------------------------------------------------------------
// src/domains/core/range/module_impl_native_cpu.cc:67-82
__global__ void range_f32_kernel(const float* __restrict__ in, float* __restrict__ out, const uint64_t count,
                                 const float scale, const float offset, const bool vec) {
    const uint64_t quads = vec ? count / 4 : 0;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const bool flat = scale == 0.0f;
    for (uint64_t i = tid; i < quads; i += step) {
        const float4 v = ldg_stream_f4(in4 + i);
        const float4 r = flat ? make_float4(0.5f, 0.5f, 0.5f, 0.5f)
                              : make_float4(range_exact(v.x, scale, offset), range_exact(v.y, scale, offset),
                                            range_exact(v.z, scale, offset), range_exact(v.w, scale, offset));
        stg_stream_f4(out4 + i, r);
    }
    for (uint64_t i = quads * 4 + tid; i < count; i += step) {
        out[i] = flat ? 0.5f : range_exact(in[i], scale, offset);
    }
}

// ---- cast F32 -> CF32 -----------------------------------------------------------------------
__global__ void cast_f32_cf32_kernel(const float* __restrict__ in, float2* __restrict__ out, const uint64_t count) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        out[i] = make_float2(in[i], 0.0f);
    }
}

// ---- cast integer -> F32 (complex integers are two scalars) ------------------------------------------------------
// src/domains/core/cast/module_impl_native_cpu.cc:163-330: out = static_cast<F32>(in) / scaler. The scaler is a power
// of two, so the IEEE division is exact; cvt.rn is the C++ int->float conversion. 16 bytes of input per thread.
template <typename T>
__global__ void cast_int_f32_kernel(const T* __restrict__ in, float* __restrict__ out, const uint64_t scalars,
                                    const float inv_scaler) {
    // inv_scaler = 1 / scaler with scaler a power of two: x * 2^-k == x / 2^k bit for bit, one instruction instead of
    // the IEEE division sequence. A thread converts FOUR scalars per step — one 4 / 8 / 16-byte load, one float4 store —
    // so that a warp's store instruction covers 512 contiguous bytes (whole 32-byte sectors; round 1 gave every thread
    // 16 input bytes = 64 output bytes and each store instruction filled half of 32 scattered sectors: 49 % of the
    // roofline for CI8 -> CF32), with four such steps in flight per thread.
    struct alignas(4 * sizeof(T)) Quad {
        T v[4];
    };
    const uint64_t quads = (reinterpret_cast<uintptr_t>(in) & (4 * sizeof(T) - 1)) == 0 ? scalars / 4 : 0;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const Quad* const src = reinterpret_cast<const Quad*>(in);
    float4* const dst = reinterpret_cast<float4*>(out);
    auto convert = [inv_scaler](const Quad q) {
        return make_float4(__fmul_rn(static_cast<float>(q.v[0]), inv_scaler), __fmul_rn(static_cast<float>(q.v[1]), inv_scaler),
                           __fmul_rn(static_cast<float>(q.v[2]), inv_scaler), __fmul_rn(static_cast<float>(q.v[3]), inv_scaler));
    };
    uint64_t i = tid;
    for (; i + 3 * step < quads; i += 4 * step) {
        Quad q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            q[u] = src[i + u * step];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            stg_stream_f4(dst + i + u * step, convert(q[u]));
        }
    }
    for (; i < quads; i += step) {
        stg_stream_f4(dst + i, convert(src[i]));
    }
    for (uint64_t k = quads * 4 + tid; k < scalars; k += step) {
        out[k] = __fmul_rn(static_cast<float>(in[k]), inv_scaler);
    }
}

// ---- strided copy (layout gather / scatter) ----------------------------------------------------------
// The role of the reference's fft_layout kernel (src/domains/dsp/fft/module_impl_native_cuda.cc:31-141): bring a
// strided / permuted view into the contiguous [batch, n] layout the fast kernels use, and scatter results back.
struct CopyPlan {
    int rank;
    uint64_t shape[8];
    uint64_t src_stride[8];
    uint64_t dst_stride[8];
};

template <typename T>
__global__ void copy_strided_kernel(const T* __restrict__ src, T* __restrict__ dst, const uint64_t total,
                                    const CopyPlan plan) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint64_t rem = i, os = 0, od = 0;
#pragma unroll 1
        for (int d = plan.rank - 1; d >= 0; --d) {
            const uint64_t q = rem / plan.shape[d];
            const uint64_t coord = rem - q * plan.shape[d];
            rem = q;
            os += coord * plan.src_stride[d];
            od += coord * plan.dst_stride[d];
        }
        dst[od] = src[os];
    }
}

// ---- real-input FFT helpers (R2C / FFTPACK packing around the C2C kernels) --------------------------------
// pocketfft::r2c keeps bins 0..n/2; pocketfft::r2r_fftpack stores [Re X0, Re X1, Im X1, ..., (Re X_{n/2})]
// (src/domains/dsp/fft/module_impl_native_cpu.cc:142-167).
__global__ void fft_r2c_pack_kernel(const float2* __restrict__ full, float2* __restrict__ out, const uint64_t batch,
                                    const uint64_t n) {
    const uint64_t half = n / 2 + 1, total = batch * half;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / half, k = i - row * half;
        out[i] = full[row * n + k];
    }
}
__global__ void fftpack_pack_kernel(const float2* __restrict__ full, float* __restrict__ out, const uint64_t batch,
                                    const uint64_t n) {
    const uint64_t total = batch * n;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / n, j = i - row * n;
        const uint64_t k = (j + 1) / 2;                      // j = 0 -> X0.re; j = 2k-1 -> Xk.re; j = 2k -> Xk.im
        const float2 v = full[row * n + k];
        out[i] = (j == 0 || (j & 1)) ? v.x : v.y;
    }
}
// halfcomplex -> full Hermitian spectrum X[k], X[n-k] = conj(X[k])
__global__ void fftpack_unpack_kernel(const float* __restrict__ in, float2* __restrict__ full, const uint64_t batch,
                                      const uint64_t n) {
    const uint64_t total = batch * n;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / n, k = i - row * n;
        const float* r = in + row * n;
        const uint64_t kk = k <= n / 2 ? k : n - k;          // mirror index
        float re, im;
        if (kk == 0) {
            re = r[0];
            im = 0.0f;
        } else if (2 * kk == n) {
            re = r[n - 1];
            im = 0.0f;
        } else {
            re = r[2 * kk - 1];
            im = r[2 * kk];
        }
        full[i] = make_float2(re, k <= n / 2 ? im : -im);
    }
}
__global__ void real_part_kernel(const float2* __restrict__ in, float* __restrict__ out, const uint64_t count) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        out[i] = in[i].x;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static int multiply_dispatch(b200_ctx* ctx, const T* a, const T* b, T* c, int rank, const uint64_t* shape,
                             const uint64_t* stride_a, const uint64_t* stride_b, b200_stream stream) {
    B200_REQUIRE(ctx && a && b && c && shape && stride_a && stride_b, "b200_multiply: null argument");
    B200_REQUIRE(rank >= 1 && rank <= 8, "b200_multiply: rank %d unsupported (1..8)", rank);
    uint64_t total = 1;
    for (int d = 0; d < rank; ++d) {
        B200_REQUIRE(shape[d] > 0, "b200_multiply: zero-sized dimension");
        total *= shape[d];
    }
    DeviceGuard guard(ctx);
    constexpr bool kComplex = sizeof(T) == 8;
    // Contiguity analysis of the two broadcast views.
    bool a_contig = true, b_contig = true;
    uint64_t expect = 1;
    for (int d = rank - 1; d >= 0; --d) {
        if (shape[d] != 1) {
            a_contig = a_contig && stride_a[d] == expect;
            b_contig = b_contig && stride_b[d] == expect;
        }
        expect *= shape[d];
    }
    const bool vec_ok = aligned16(a) && aligned16(b) && aligned16(c);
    const uint64_t per_vec = 16 / sizeof(T);
    if (a_contig && b_contig && vec_ok && total % per_vec == 0) {
        const uint64_t vecs = total / per_vec;
        const int grid = stream_grid(ctx, vecs, 256, 8);
        if constexpr (kComplex) {
            multiply_same_cf32_kernel<<<grid, 256, 0, as_stream(stream)>>>(
                reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                reinterpret_cast<float4*>(c), vecs);
        } else {
            multiply_same_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(
                reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                reinterpret_cast<float4*>(c), vecs);
        }
        B200_LAUNCH_CHECK();
        return B200_SUCCESS;
    }
    if constexpr (kComplex) {
        // [rows, n] x [1, n] row broadcast (b has stride 0 on every dim but the last).
        bool b_row = stride_b[rank - 1] == 1 || shape[rank - 1] == 1;
        for (int d = 0; d < rank - 1; ++d) {
            b_row = b_row && (stride_b[d] == 0 || shape[d] == 1);
        }
        const uint64_t n = shape[rank - 1];
        if (a_contig && b_row && vec_ok && n % 2 == 0 && n >= 2) {
            const uint64_t rows = total / n, vec_per_row = n / 2;
            const unsigned gx = static_cast<unsigned>((vec_per_row + 255) / 256);
            const uint64_t want_y = (static_cast<uint64_t>(ctx->sms) * 8 + gx - 1) / gx;
            const unsigned gy = static_cast<unsigned>(rows < want_y ? rows : want_y);
            multiply_rowbcast_cf32_kernel<<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(
                reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                reinterpret_cast<float4*>(c), rows, vec_per_row);
            B200_LAUNCH_CHECK();
            return B200_SUCCESS;
        }
    }
    BroadcastPlan plan{};
    plan.rank = rank;
    for (int d = 0; d < rank; ++d) {
        plan.shape[d] = shape[d];
        plan.stride_a[d] = stride_a[d];
        plan.stride_b[d] = stride_b[d];
    }
    const int grid = stream_grid(ctx, total, 256, 8);
    multiply_generic_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(a, b, c, total, plan);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_window_blackman_cf32(b200_ctx* ctx, b200_cf32* out, uint64_t n, b200_stream stream) {
    B200_REQUIRE(ctx && out, "b200_window_blackman_cf32: null argument");
    B200_REQUIRE(n > 0, "b200_window_blackman_cf32: window size cannot be zero");
    DeviceGuard guard(ctx);
    window_blackman_kernel<<<stream_grid(ctx, n, 128, 4), 128, 0, as_stream(stream)>>>(
        reinterpret_cast<float2*>(out), n);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_invert_cf32(b200_ctx* ctx, const b200_cf32* in, b200_cf32* out, uint64_t outer, uint64_t n,
                     uint64_t inner, b200_stream stream) {
    B200_REQUIRE(ctx && in && out, "b200_invert_cf32: null argument");
    const uint64_t total = outer * n * inner;
    if (total == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    invert_kernel<<<stream_grid(ctx, total, 256, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), total, n, inner);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_multiply_cf32(b200_ctx* ctx, const b200_cf32* a, const b200_cf32* b, b200_cf32* c, int rank,
                       const uint64_t* shape, const uint64_t* stride_a, const uint64_t* stride_b,
                       b200_stream stream) {
    return multiply_dispatch<float2>(ctx, reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b),
                                     reinterpret_cast<float2*>(c), rank, shape, stride_a, stride_b, stream);
}

int b200_multiply_f32(b200_ctx* ctx, const float* a, const float* b, float* c, int rank, const uint64_t* shape,
                      const uint64_t* stride_a, const uint64_t* stride_b, b200_stream stream) {
    return multiply_dispatch<float>(ctx, a, b, c, rank, shape, stride_a, stride_b, stream);
}

int b200_multiply_constant_cf32(b200_ctx* ctx, const b200_cf32* in, b200_cf32* out, uint64_t count,
                                float constant, b200_stream stream) {
    return b200_multiply_constant_f32(ctx, reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out),
                                      count * 2, constant, stream);
}

int b200_multiply_constant_f32(b200_ctx* ctx, const float* in, float* out, uint64_t count, float constant,
                               b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_multiply_constant: null argument");
    if (count == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    multiply_constant_kernel<<<stream_grid(ctx, count / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(
        in, out, count, constant, aligned16(in) && aligned16(out));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_amplitude_cf32(b200_ctx* ctx, const b200_cf32* in, float* out, uint64_t count, float coeff,
                        b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_amplitude_cf32: null argument");
    if (count == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    amplitude_cf32_kernel<<<stream_grid(ctx, count / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(in), out, count, coeff, aligned16(in) && aligned16(out));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_amplitude_f32(b200_ctx* ctx, const float* in, float* out, uint64_t count, float coeff,
                       b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_amplitude_f32: null argument");
    if (count == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    amplitude_f32_kernel<<<stream_grid(ctx, count / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(
        in, out, count, coeff, aligned16(in) && aligned16(out));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_range_f32(b200_ctx* ctx, const float* in, float* out, uint64_t count, float scale, float offset,
                   b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_range_f32: null argument");
    if (count == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    range_f32_kernel<<<stream_grid(ctx, count / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(
        in, out, count, scale, offset, aligned16(in) && aligned16(out));
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fft_real_helper(b200_ctx* ctx, int op, const void* in, void* out, uint64_t batch, uint64_t n,
                         b200_stream stream) {
    B200_REQUIRE(ctx && in && out, "b200_fft_real_helper: null argument");
    B200_REQUIRE(op >= 0 && op <= 3, "b200_fft_real_helper: op must be 0..3");
    if (batch * n == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    const int grid = stream_grid(ctx, batch * n, 256, 8);
    const cudaStream_t s = as_stream(stream);
    switch (op) {
        case 0: fft_r2c_pack_kernel<<<grid, 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), batch, n); break;
        case 1: fftpack_pack_kernel<<<grid, 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float*>(out), batch, n); break;
        case 2: fftpack_unpack_kernel<<<grid, 256, 0, s>>>(static_cast<const float*>(in), static_cast<float2*>(out), batch, n); break;
        default: real_part_kernel<<<grid, 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float*>(out), batch * n); break;
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_copy_strided(b200_ctx* ctx, const void* src, void* dst, int elem_bytes, int rank, const uint64_t* shape,
                      const uint64_t* src_stride, const uint64_t* dst_stride, b200_stream stream) {
    B200_REQUIRE(ctx && src && dst && shape && src_stride && dst_stride, "b200_copy_strided: null argument");
    B200_REQUIRE(rank >= 1 && rank <= 8, "b200_copy_strided: rank %d unsupported (1..8)", rank);
    B200_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8,
                 "b200_copy_strided: element size must be 1, 2, 4 or 8 bytes");
    CopyPlan plan{};
    plan.rank = rank;
    uint64_t total = 1;
    for (int d = 0; d < rank; ++d) {
        plan.shape[d] = shape[d];
        plan.src_stride[d] = src_stride[d];
        plan.dst_stride[d] = dst_stride[d];
        total *= shape[d];
    }
    if (total == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    const int grid = stream_grid(ctx, total, 256, 8);
    if (elem_bytes == 8) {
        copy_strided_kernel<float2><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float2*>(src),
                                                                          static_cast<float2*>(dst), total, plan);
    } else if (elem_bytes == 1) {
        copy_strided_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(src),
                                                                           static_cast<uint8_t*>(dst), total, plan);
    } else if (elem_bytes == 2) {
        copy_strided_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint16_t*>(src),
                                                                            static_cast<uint16_t*>(dst), total, plan);
    } else {
        copy_strided_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(src),
                                                                         static_cast<float*>(dst), total, plan);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_cast_f32_cf32(b200_ctx* ctx, const float* in, b200_cf32* out, uint64_t count, b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_cast_f32_cf32: null argument");
    if (count == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    cast_f32_cf32_kernel<<<stream_grid(ctx, count, 256, 8), 256, 0, as_stream(stream)>>>(
        in, reinterpret_cast<float2*>(out), count);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_cast_int(b200_ctx* ctx, const void* in, int in_dtype, void* out, uint64_t count, b200_stream stream) {
    B200_REQUIRE(ctx && (count == 0 || (in && out)), "b200_cast_int: null argument");
    B200_REQUIRE(in_dtype >= B200_DTYPE_I8 && in_dtype <= B200_DTYPE_CU32,
                 "[MODULE_CAST_B200] Unsupported conversion from dtype code %d.", in_dtype);
    if (count == 0) {
        return B200_SUCCESS;
    }
    const bool is_complex = in_dtype >= B200_DTYPE_CI8;
    const int base = is_complex ? in_dtype - (B200_DTYPE_CI8 - B200_DTYPE_I8) : in_dtype;
    const uint64_t scalars = is_complex ? count * 2 : count;
    B200_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b200_cast_int: output must be 16-byte aligned");
    DeviceGuard guard(ctx);
    float* dst = static_cast<float*>(out);
    const cudaStream_t s = as_stream(stream);
    switch (base) {
        case B200_DTYPE_I8:
            cast_int_f32_kernel<int8_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const int8_t*>(in), dst, scalars, 1.0f / 128.0f);
            break;
        case B200_DTYPE_U8:
            cast_int_f32_kernel<uint8_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const uint8_t*>(in), dst, scalars, 1.0f / 128.0f);
            break;
        case B200_DTYPE_I16:
            cast_int_f32_kernel<int16_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const int16_t*>(in), dst, scalars, 1.0f / 32768.0f);
            break;
        case B200_DTYPE_U16:
            cast_int_f32_kernel<uint16_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const uint16_t*>(in), dst, scalars, 1.0f / 32768.0f);
            break;
        case B200_DTYPE_I32:
            cast_int_f32_kernel<int32_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const int32_t*>(in), dst, scalars, 1.0f / 2147483648.0f);
            break;
        default:
            cast_int_f32_kernel<uint32_t><<<stream_grid(ctx, scalars / 16 + 1, 256, 8), 256, 0, s>>>(
                static_cast<const uint32_t*>(in), dst, scalars, 1.0f / 2147483648.0f);
            break;
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // extern "C"
