// Shared pieces of the FFT kernels: parameter block, prologue (window multiply) and epilogue
// (complex store / amplitude dB / amplitude+range) policies, radix butterflies.
#pragma once

#include "common.cuh"

namespace b200 {

// MODE_C2C_T (fft_radix_kernel only): the second pass of the two-pass plan (fft_twopass.cuh) — rows are [transform][k1 < 16]
// slabs of a 16x longer transform; X[k2] of row r goes to out[(r / 16) * 16 n + (r % 16) + 16 k2], no input swap for the
// inverse (the first pass did it), output swap as MODE_C2C.
// MODE_R2C (fft_radix_kernel, 32 <= N <= 8192): the rows are REAL rows of length 2N viewed as N complex samples; the epilogue
// unpacks the N-point spectrum into the 2N-point real-input spectrum (real_out, real_layout) without leaving the CTA.
enum : int { MODE_C2C = 0, MODE_AMP = 1, MODE_AMP_RANGE = 2, MODE_C2C_T = 3, MODE_R2C = 4 };
enum : int { WIN_NONE = 0, WIN_REAL = 1, WIN_COMPLEX = 2 };

struct FftParams {
    const float2* in;       // [rows, n] CF32
    void* out;              // [rows, n] CF32 (MODE_C2C) or F32 (MODE_AMP*)
    uint64_t rows;
    uint32_t n;
    int inverse;            // MODE_C2C only: unnormalised inverse (exp(+j...))
    const float2* twiddle;  // W_n^j = exp(-2*pi*i*j/n), j in [0, n), evaluated in F64 on the host
    const float* win_re;    // [n] real window (WIN_REAL)
    const float2* win_c;    // [n] complex window (WIN_COMPLEX)
    // Epilogue constants (host-computed in F64):
    //   MODE_AMP       : out = Y * amp_scale + amp_coeff, Y = log2-approx(|X|)   (-inf at |X| == 0)
    //   MODE_AMP_RANGE : out = 1 / (1 + 2^(Y * k1 + k0))  ==  0.5 + 0.5*tanh(4*((dB*s + o) - 0.5))
    // range with min == max (scale 0) is the constant 0.5: the host passes k1 = k0 = 0, zero_value = 0.5.
    float amp_scale, amp_coeff, k1, k0;
    float zero_value;       // MODE_AMP_RANGE result where |X|^2 == 0 (0, or 0.5 for a flat range)
    // Fused AGC (spectrum_engine enableAgc: one RMS tile per spectrum, src/domains/dsp/spectrum_engine/block_impl.cc:186-200):
    // every row is scaled by clamp(reference / sqrt(mean |X|^2 + epsilon), min, max) before the amplitude.
    double agc_reference, agc_epsilon, agc_min, agc_max;
    // Fused column sums (the lineplot consumer's batch sum, lineplot/module_impl_native_cpu.cc:93-98): every CTA adds
    // the epilogue results of its rows in registers and writes one [n] partial at the end; colsum_partial is
    // [gridDim.x, n], reduced in CTA order afterwards (deterministic for a given grid).
    float* colsum_partial;
    // MODE_R2C: [rows, N + 1] CF32 (real_layout 0, pocketfft::r2c) or [rows, 2N] F32 FFTPACK half-complex (real_layout 1)
    void* real_out;
    int real_layout;
};

// ---- butterflies (forward sign) ------------------------------------------------------------

__device__ __forceinline__ void bfly2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// 4-point DFT in place: (a,b,c,d) <- (X0,X1,X2,X3), forward (multiplication by -i on the odd leg).
__device__ __forceinline__ void bfly4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
    a = cadd(s0, s2);
    c = csub(s0, s2);
    b = csub_i(s1, s3);
    d = cadd_i(s1, s3);
}

// 16-point DFT in place as 4x4. Result X[k] lands in v[dft16_pos(k)].
__host__ __device__ constexpr int dft16_pos(const int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ void dft16_first_layer(float2 (&v)[16]) {
#pragma unroll
    for (int a0 = 0; a0 < 4; ++a0) {
        bfly4(v[a0], v[a0 + 4], v[a0 + 8], v[a0 + 12]);
    }
}

__device__ __forceinline__ void dft16_rest(float2 (&v)[16]) {
    constexpr float kC = 0.92387953251128674f;  // cos(pi/8)
    constexpr float kS = 0.38268343236508977f;  // sin(pi/8)
    constexpr float kH = 0.70710678118654752f;  // sqrt(1/2)
    // Internal twiddles W16^(a0*q) on v[a0 + 4q], a0,q in 1..3.
    // a0=1: q=1 -> W^1, q=2 -> W^2, q=3 -> W^3
    v[5] = cmul(v[5], make_float2(kC, -kS));
    v[9] = cscale(csub_i(v[9], v[9]), kH);
    v[13] = cmul(v[13], make_float2(kS, -kC));
    // a0=2: q=1 -> W^2, q=2 -> W^4 (= -i), q=3 -> W^6
    v[6] = cscale(csub_i(v[6], v[6]), kH);
    v[10] = make_float2(v[10].y, -v[10].x);
    v[14] = cscale(cadd_i(v[14], v[14]), -kH);
    // a0=3: q=1 -> W^3, q=2 -> W^6, q=3 -> W^9
    v[7] = cmul(v[7], make_float2(kS, -kC));
    v[11] = cscale(cadd_i(v[11], v[11]), -kH);
    v[15] = cmul(v[15], make_float2(-kC, kS));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        bfly4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    dft16_first_layer(v);
    dft16_rest(v);
}

// ---- prologue ------------------------------------------------------------------------------

template <int WIN>
__device__ __forceinline__ float2 apply_window(const float2 x, const float wr, const float2 wc) {
    if constexpr (WIN == WIN_REAL) {
        // (x.re + i x.im) * (w + 0i): the reference's products with the zero imaginary part vanish.
        return cscale(x, wr);
    } else if constexpr (WIN == WIN_COMPLEX) {
        return cmul(x, wc);
    } else {
        return x;
    }
}

// ---- epilogue ------------------------------------------------------------------------------

__device__ __forceinline__ float sqrt_approx(const float x) {
    float y;
    // MUFU.SQRT: max rel. error 2^-23. ftz: |X|^2 below 1.2e-38 (|X| < 1e-19) is treated as zero power.
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(const float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(const float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Two outputs at a time so the polynomial / affine steps issue as packed FFMA2/FADD2.
//   Y = poly(F) + E with |X| = F * 2^E, F in [0.5, 1): Backend::ApproxLog10 / log10(2)
//   (include/jetstream/backend/devices/cpu/helpers.hh:61-74).
//   MODE_AMP       : Y * amp_scale + amp_coeff, -inf where |X|^2 == 0
//   MODE_AMP_RANGE : 1 / (1 + 2^(Y k1 + k0)) == 0.5 + 0.5 tanh(4 ((dB s + o) - 0.5)), 0 where |X|^2 == 0
template <int MODE, bool AGC = false>
__device__ __forceinline__ float2 spectral_epilogue2(const float2 X0, const float2 X1, const FftParams& p,
                                                     const float gain = 1.0f) {
    const float p0 = fmaf(X0.x, X0.x, X0.y * X0.y);
    const float p1 = fmaf(X1.x, X1.x, X1.y * X1.y);
    // AGC: |g X| = g |X| (the reference scales the components in F64 and rounds them to F32 first; the difference is
    // one rounding of the magnitude)
    const int b0 = __float_as_int(AGC ? sqrt_approx(p0) * gain : sqrt_approx(p0));
    const int b1 = __float_as_int(AGC ? sqrt_approx(p1) * gain : sqrt_approx(p1));
    const float2 f = make_float2(__int_as_float((b0 & 0x007fffff) | 0x3f000000),
                                 __int_as_float((b1 & 0x007fffff) | 0x3f000000));
    // (float)(biased exponent field) without I2F (a quarter-rate XU instruction next to the three MUFU per output):
    // (b >> 23) | 0x4B000000 is the F32 2^23 + exponent, exact; the 2^23 and the frexp bias (-126) leave in one exact
    // integer-valued addition below, so the result is bit-identical to the conversion.
    const float2 e = make_float2(__int_as_float((b0 >> 23) | 0x4B000000), __int_as_float((b1 >> 23) | 0x4B000000));
    float2 y = __ffma2_rn(make_float2(1.23149591368684f, 1.23149591368684f), f,
                          make_float2(-4.11852516267426f, -4.11852516267426f));
    y = __ffma2_rn(y, f, make_float2(6.02197014179219f, 6.02197014179219f));
    y = __ffma2_rn(y, f, make_float2(-3.13396450166353f, -3.13396450166353f));
    y = __fadd2_rn(y, __fadd2_rn(e, make_float2(-8388734.0f, -8388734.0f)));      // - (2^23 + 126)
    if constexpr (MODE == MODE_AMP) {
        const float2 r = __ffma2_rn(y, make_float2(p.amp_scale, p.amp_scale), make_float2(p.amp_coeff, p.amp_coeff));
        return make_float2(p0 == 0.0f ? -INFINITY : r.x, p1 == 0.0f ? -INFINITY : r.y);
    } else {
        const float2 a = __ffma2_rn(y, make_float2(p.k1, p.k1), make_float2(p.k0, p.k0));
        const float2 d = __fadd2_rn(make_float2(ex2_approx(a.x), ex2_approx(a.y)), make_float2(1.0f, 1.0f));
        return make_float2(p0 == 0.0f ? p.zero_value : rcp_approx(d.x), p1 == 0.0f ? p.zero_value : rcp_approx(d.y));
    }
}

template <int MODE>
__device__ __forceinline__ float spectral_epilogue(const float2 X, const FftParams& p) {
    return spectral_epilogue2<MODE>(X, X, p).x;
}

}  // namespace b200
