// Single-pass 4096-point CF32 FFT kernel with fused prologue/epilogue — the B200 replacement for
// the spectrum_engine module chain (src/domains/dsp/spectrum_engine/block_impl.cc:120-217) and,
// in MODE_C2C, for the 4096-point `fft` module.
//
// Design:
//   * persistent CTAs (2 per SM, 256 threads), one 4096-point row at a time per CTA;
//   * each row (32 KiB) is staged into shared memory by ONE TMA bulk copy (cp.async.bulk +
//     mbarrier complete_tx), kStages-deep ring so HBM reads run ahead of the math;
//   * 4096 = 16 x 16 x 16: three radix-16 passes held entirely in registers (16 points/thread),
//     two shared-memory exchanges done IN the row's own staging buffer (no extra smem), all
//     exchange accesses bank-conflict-free (XOR swizzle on the second one);
//   * window multiply on the way in, |X| -> dB -> range on the way out, in registers;
//   * HBM traffic = 8 B/sample in + 4 B/sample out (MODE_AMP*), nothing else.
//
// Index algebra (validated against numpy in tests/test_index_algebra.py):
//   n = 256a + t (t = 16b + c),  k = k0 + 16 k1 + 256 k2
//   pass 1  thread t        : Y1[k0] = sum_a x[256a + t] W16^(a k0);  *= W4096^(t k0)  -> smem[256 k0 + t]
//   pass 2  thread 16k0 + c : Y2[k1] = sum_b smem[256 k0 + 16 b + c] W16^(b k1); *= W256^(c k1)
//                                                                      -> smem[272 k1 + 17 k0 + c]
//   pass 3  thread k0+16k1  : X[k0 + 16 k1 + 256 k2] = sum_c smem[272 k1 + 17 k0 + c] W16^(c k2)
#pragma once

#include "fft_common.cuh"

namespace b200 {

constexpr int kFft4096N = 4096;
constexpr int kFft4096Threads = 256;
constexpr int kFft4096RowBytes = kFft4096N * 8;
// A stage holds the 32 KiB row; the second exchange uses a padded [k1=16][k0=16][c=16 (+2 pad)] layout:
// row pitch 18 elements (16-byte aligned rows -> LDS.128 in pass 3, conflict-free), plane pitch 288.
constexpr int kFft4096X2Row = 18;
constexpr int kFft4096X2Plane = 16 * kFft4096X2Row;
constexpr int kFft4096StageBytes = 16 * kFft4096X2Plane * 8;   // 36864
// CTAs per SM -> TMA ring depth (shared memory: 227 KB per SM)
__host__ __device__ constexpr int fft4096_stages(const int ctas) { return ctas >= 3 ? 2 : 3; }
__host__ __device__ constexpr int fft4096_smem_bytes(const int ctas) {
    return fft4096_stages(ctas) * kFft4096StageBytes + 64;
}

// Input sample formats of the fused chain. IN_CF32 is the module-level contract; the complex-integer formats fuse
// the reference's `cast` module (src/domains/core/cast/module_impl_native_cpu.cc:270-330, scaler 128 / 32768) into the
// pass-1 load, so an SDR's native samples are read from HBM once at 2 or 4 bytes instead of being expanded to CF32 first.
enum : int { IN_CF32 = 0, IN_CI8 = 1, IN_CU8 = 2, IN_CI16 = 3, IN_CU16 = 4 };
__host__ __device__ constexpr int fft4096_in_bytes(const int itype) {
    return itype == IN_CF32 ? 8 : (itype <= IN_CU8 ? 2 : 4);
}
// Integer rows land in their own ring (3 x 8 KiB for 8-bit, 2 x 16 KiB for 16-bit samples: 2 CTAs/SM need <= 113 KB
// each) and the exchanges alternate between two buffers.
__host__ __device__ constexpr int fft4096_int_land_stages(const int itype) { return fft4096_in_bytes(itype) == 2 ? 3 : 2; }
__host__ __device__ constexpr int fft4096_int_smem_bytes(const int itype) {
    return 2 * kFft4096StageBytes + fft4096_int_land_stages(itype) * kFft4096N * fft4096_in_bytes(itype) + 64;
}
template <int ITYPE>
__device__ __forceinline__ float2 load_int_sample(const unsigned char* land, const uint32_t index) {
    // (F32)v / 2^k: the power-of-two reciprocal is exact, so the product equals the reference's division bit for bit
    if constexpr (ITYPE == IN_CI8) {
        const char2 c = *reinterpret_cast<const char2*>(land + 2 * index);
        return make_float2(static_cast<float>(c.x) * 0.0078125f, static_cast<float>(c.y) * 0.0078125f);
    } else if constexpr (ITYPE == IN_CU8) {
        const uchar2 c = *reinterpret_cast<const uchar2*>(land + 2 * index);
        return make_float2(static_cast<float>(c.x) * 0.0078125f, static_cast<float>(c.y) * 0.0078125f);
    } else if constexpr (ITYPE == IN_CI16) {
        const short2 c = *reinterpret_cast<const short2*>(land + 4 * index);
        return make_float2(static_cast<float>(c.x) * 3.0517578125e-05f, static_cast<float>(c.y) * 3.0517578125e-05f);
    } else {
        const ushort2 c = *reinterpret_cast<const ushort2*>(land + 4 * index);
        return make_float2(static_cast<float>(c.x) * 3.0517578125e-05f, static_cast<float>(c.y) * 3.0517578125e-05f);
    }
}

// ---- mbarrier / TMA (bulk async copy) PTX wrappers ------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_row(void* smem_dst, const void* gmem_src, const uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Thread-constant twiddle powers w^1, w^2, w^3 (lo) and w^4, w^8, w^12 (hi); w^k = hi[k>>2] * lo[k&3].
struct TwiddleSet {
    float2 lo[3];
    float2 hi[3];
};

__device__ __forceinline__ TwiddleSet load_twiddles(const float2* __restrict__ table, const uint32_t base) {
    // table[j] = W4096^j; powers of w = W4096^base. base * 12 < 4096 is guaranteed by the callers.
    TwiddleSet s;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        s.lo[m] = table[base * (m + 1)];
        s.hi[m] = table[base * 4 * (m + 1)];
    }
    return s;
}

// v[dft16_pos(k)] *= w^k for k = 1..15.
__device__ __forceinline__ void apply_twiddles(float2 (&v)[16], const TwiddleSet& s) {
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        float2 x = v[dft16_pos(k)];
        if ((k & 3) != 0) {
            x = cmul(x, s.lo[(k & 3) - 1]);
        }
        if ((k >> 2) != 0) {
            x = cmul(x, s.hi[(k >> 2) - 1]);
        }
        v[dft16_pos(k)] = x;
    }
}

template <int MODE, int WIN, int CTAS, int ITYPE = IN_CF32, bool AGC = false, bool COLSUM = false>
__global__ void __launch_bounds__(kFft4096Threads, CTAS) fft4096_kernel(const FftParams p) {
    static_assert(!COLSUM || MODE != MODE_C2C, "column sums exist for the amplitude outputs only");
    // AGC: the mean power of the spectrum equals the power of the windowed row (Parseval: sum |X_k|^2 = N sum |x_n w_n|^2),
    // which pass 1 has in registers: per-warp partial sums go to shared memory (double-buffered by row parity; barriers
    // (A) and (C) of the row order them before the epilogue reads them).
    __shared__ float agc_partial[2][kFft4096Threads / 32];
    constexpr bool kInt = ITYPE != IN_CF32;
    constexpr int kFft4096Stages = kInt ? fft4096_int_land_stages(ITYPE) : fft4096_stages(CTAS);   // TMA ring depth
    constexpr int kLandBytes = kFft4096N * fft4096_in_bytes(ITYPE);                        // bytes of one input row
    constexpr int kLandPitch = kInt ? kLandBytes : kFft4096StageBytes;
    constexpr int kLandBase = kInt ? 2 * kFft4096StageBytes : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kLandBase + kFft4096Stages * kLandPitch);
    uint64_t* const reads_done = full + kFft4096Stages;   // split barrier (B): 8 warp arrivals per row

    const uint32_t t = threadIdx.x;
    const uint64_t first = blockIdx.x;
    const uint64_t stride = gridDim.x;
    const uint32_t my_rows =
        first < p.rows ? static_cast<uint32_t>((p.rows - first + stride - 1) / stride) : 0u;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kFft4096Stages; ++s) {
            mbar_init(&full[s], 1);
        }
        mbar_init(reads_done, kFft4096Threads / 32);
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned char* next_src = reinterpret_cast<const unsigned char*>(p.in) + first * kLandBytes;   // thread 0
    const uint64_t src_step = stride * kLandBytes;
    uint32_t issued = 0;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kFft4096Stages; ++s) {
            if (issued < my_rows) {
                mbar_expect_tx(&full[s], kLandBytes);
                tma_load_row(smem_raw + kLandBase + s * kLandPitch, next_src, kLandBytes, &full[s]);
                next_src += src_step;
                ++issued;
            }
        }
    }

    // Thread-constant operands (persistent across rows).
    const uint32_t lo4 = t & 15, hi4 = t >> 4;
    const TwiddleSet tw1 = load_twiddles(p.twiddle, t);          // W4096^(t k0)
    const TwiddleSet tw2 = load_twiddles(p.twiddle, 16 * lo4);   // W256^(c k1) = W4096^(16 c k1)

    float wr[16];
    if constexpr (WIN == WIN_REAL) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            wr[a] = p.win_re[t + 256 * a];
        }
    }

    // Per-thread byte offsets inside a stage buffer (element = 8 bytes).
    //   pass 1 load / exchange-1 store : t + 256 a                       (a, k0 = register index)
    //   exchange-1 load  (t = 16 k0 + c): 256 k0 + c + 16 b
    //   exchange-2 store (t = 16 k0 + c): 18 k0 + c + 288 k1            (row pitch 18 / plane pitch 288:
    //   exchange-2 load  (t = k0 + 16 k1): 18 k0 + 288 k1 + c             conflict-free, immediate offsets)
    const uint32_t off_p1 = t * 8;
    const uint32_t off_x1 = (256 * hi4 + lo4) * 8;
    const uint32_t off_x2s = (kFft4096X2Row * hi4 + lo4) * 8;
    const uint32_t off_x2l = (kFft4096X2Row * lo4 + kFft4096X2Plane * hi4) * 8;

    uint32_t stage = 0, parity = 0, row_parity = 0;
    uint32_t refill_stage = 0;  // stage of the previous row (refilled after barrier (A))
    unsigned char* out_ptr = static_cast<unsigned char*>(p.out) +
                             (first * kFft4096N + t) * (MODE == MODE_C2C ? 8 : 4);
    const uint64_t out_step = stride * kFft4096N * (MODE == MODE_C2C ? 8 : 4);
    // COLSUM: running sums of this CTA's outputs, column t + 256 k in colsum[k / 2].{x, y} (rows in CTA order)
    float2 colsum[COLSUM ? 8 : 1];
#pragma unroll
    for (int k = 0; k < (COLSUM ? 8 : 1); ++k) {
        colsum[k] = make_float2(0.f, 0.f);
    }

    for (uint32_t i = 0; i < my_rows; ++i) {
        // CF32: the row's landing buffer is also its exchange buffer. Integer input: a separate landing ring, the
        // exchanges alternate between two buffers (row i+1 may write its exchange-1 while slow warps still read the
        // exchange-2 of row i).
        unsigned char* const buf = smem_raw + (kInt ? (i & 1u) : stage) * kFft4096StageBytes;
        const unsigned char* const land = smem_raw + kLandBase + stage * kLandPitch;
        mbar_wait(&full[stage], parity);

        // ---- pass 1 -------------------------------------------------------------------------
        float2 v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            float2 x;
            if constexpr (kInt) {
                x = load_int_sample<ITYPE>(land, t + 256 * a);
            } else {
                x = *reinterpret_cast<const float2*>(buf + off_p1 + 2048 * a);
            }
            if constexpr (MODE == MODE_C2C) {
                if (p.inverse) {
                    x = make_float2(x.y, x.x);  // IFFT(x) = swap(FFT(swap(x)))
                }
            }
            if constexpr (WIN == WIN_REAL) {
                x = apply_window<WIN>(x, wr[a], make_float2(0.f, 0.f));
            } else if constexpr (WIN == WIN_COMPLEX) {
                x = apply_window<WIN>(x, 0.f, p.win_c[t + 256 * a]);
            }
            v[a] = x;
        }
        if constexpr (AGC) {
            float power = 0.0f;
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                power = fmaf(v[a].x, v[a].x, fmaf(v[a].y, v[a].y, power));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                power += __shfl_xor_sync(0xffffffffu, power, off);
            }
            if ((t & 31) == 0) {
                agc_partial[i & 1][t >> 5] = power;
            }
        }
        dft16(v);
        apply_twiddles(v, tw1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // this thread's own input slots: no hazard
            *reinterpret_cast<float2*>(buf + off_p1 + 2048 * k) = v[dft16_pos(k)];
        }
        __syncthreads();  // (A)

        // The previous row's buffer is free now (every thread finished its pass-3 reads before
        // arriving at (A)): refill it with the row kStages-1 ahead. Integer input: the landing slot of THIS row is
        // free (every thread has converted its 16 samples), refill it with the row kStages ahead.
        if (t == 0 && (kInt || i >= 1) && issued < my_rows) {
            const uint32_t slot = kInt ? stage : refill_stage;
            fence_proxy_async();
            mbar_expect_tx(&full[slot], kLandBytes);
            tma_load_row(smem_raw + kLandBase + slot * kLandPitch, next_src, kLandBytes, &full[slot]);
            next_src += src_step;
            ++issued;
        }

        // ---- pass 2 -------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            v[b] = *reinterpret_cast<const float2*>(buf + off_x1 + 128 * b);
        }
        dft16_first_layer(v);   // consumes every loaded value: this warp's exchange-1 reads have completed
        // Split barrier (B) "all exchange-1 reads done before any exchange-2 write": arrive now (one lane
        // per warp), wait only right before the stores, so the butterfly math overlaps the other warps.
        __syncwarp();
        if ((t & 31) == 0) {
            mbar_arrive(reads_done);
        }
        dft16_rest(v);
        apply_twiddles(v, tw2);
        mbar_wait(reads_done, row_parity);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            *reinterpret_cast<float2*>(buf + off_x2s + kFft4096X2Plane * 8 * k) = v[dft16_pos(k)];
        }
        __syncthreads();  // (C)

        // ---- pass 3 -------------------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const float4 pair = *reinterpret_cast<const float4*>(buf + off_x2l + 8 * c);
            v[c] = make_float2(pair.x, pair.y);
            v[c + 1] = make_float2(pair.z, pair.w);
        }
        dft16(v);

        // ---- epilogue: X[t + 256 k2] ---------------------------------------------------------
        if constexpr (MODE == MODE_C2C) {
            float2* const out = reinterpret_cast<float2*>(out_ptr);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float2 X = v[dft16_pos(k)];
                if (p.inverse) {
                    X = make_float2(X.y, X.x);
                }
                stg_stream_f2(out + 256 * k, X);
            }
        } else {
            float* const out = reinterpret_cast<float*>(out_ptr);
            float gain = 1.0f;
            if constexpr (AGC) {
                float mean_power = 0.0f;
#pragma unroll
                for (int w = 0; w < kFft4096Threads / 32; ++w) {
                    mean_power += agc_partial[i & 1][w];
                }
                // AgcImplNativeCpu: clamp(reference / sqrt(meanPower + epsilon), minGain, maxGain), F64
                const double g = p.agc_reference / sqrt(static_cast<double>(mean_power) + p.agc_epsilon);
                gain = static_cast<float>(g < p.agc_min ? p.agc_min : (p.agc_max < g ? p.agc_max : g));
            }
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const float2 r = spectral_epilogue2<MODE, AGC>(v[dft16_pos(k)], v[dft16_pos(k + 1)], p, gain);
                stg_stream_f1(out + 256 * k, r.x);
                stg_stream_f1(out + 256 * (k + 1), r.y);
                if constexpr (COLSUM) {
                    colsum[k / 2] = __fadd2_rn(colsum[k / 2], r);
                }
            }
        }
        out_ptr += out_step;

        refill_stage = stage;
        row_parity ^= 1;
        if (++stage == kFft4096Stages) {
            stage = 0;
            parity ^= 1;
        }
    }
    if constexpr (COLSUM) {
        float* const dst = p.colsum_partial + static_cast<uint64_t>(blockIdx.x) * kFft4096N + t;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            dst[256 * k] = colsum[k / 2].x;
            dst[256 * (k + 1)] = colsum[k / 2].y;
        }
    }
}


}  // namespace b200
