// fft module + fused spectral chain entry points.
//   * n == 4096            : fft4096_kernel (TMA-staged, single pass, registers + smem exchange)
//   * other n = 2^k <= 16384: fft_generic_kernel (Stockham autosort radix-4/2 in shared memory,
//                             one CTA-slice per row, same fused prologue/epilogue policies)
// Replaces pocketfft::c2c (src/domains/dsp/fft/module_impl_native_cpu.cc:129-140) / cuFFT
// (src/domains/dsp/fft/module_impl_native_cuda.cc:321-333,433) and, for the chain, the module
// sequence of src/domains/dsp/spectrum_engine/block_impl.cc:120-217.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#ifndef B200_FFT4096_DEFAULT_CTAS
#define B200_FFT4096_DEFAULT_CTAS 2
#endif


#include "fft_radix.cuh"
#include "fft_twopass.cuh"
#include "fft_tile.cuh"

namespace b200 {

// ---- generic power-of-two kernel ------------------------------------------------------------
// Stockham autosort, in place with register staging: every pass reads its inputs into registers,
// barriers, writes the permuted outputs back. Pass with sub-transform size Ns, radix R:
//   j in [0, n/R), k = j mod Ns, u_t = buf[j + t n/R] * W_{R Ns}^{k t}, y = DFT_R(u),
//   buf[((j - k) * R + k) + t Ns] = y_t.
template <int MODE, int WIN, int BPT>
__global__ void __launch_bounds__(1024) fft_generic_kernel(const FftParams p, const int log2n, const int tpr,
                                                          const int rows_per_cta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* const sm = reinterpret_cast<float2*>(smem_raw);
    const uint32_t n = p.n;
    const uint32_t lane = threadIdx.x % tpr;
    const uint32_t slot = threadIdx.x / tpr;
    float2* const buf = sm + static_cast<size_t>(slot) * n;
    const uint64_t groups = (p.rows + rows_per_cta - 1) / rows_per_cta;

    for (uint64_t group = blockIdx.x; group < groups; group += gridDim.x) {
        const uint64_t row = group * rows_per_cta + slot;
        const bool valid = row < p.rows;

        if (valid) {
            const float2* const src = p.in + row * n;
            for (uint32_t j = lane; j < n; j += tpr) {
                float2 x = ldg_stream_f2(src + j);
                if constexpr (MODE == MODE_C2C) {
                    if (p.inverse) {
                        x = make_float2(x.y, x.x);
                    }
                }
                if constexpr (WIN == WIN_REAL) {
                    x = apply_window<WIN>(x, p.win_re[j], make_float2(0.f, 0.f));
                } else if constexpr (WIN == WIN_COMPLEX) {
                    x = apply_window<WIN>(x, 0.f, p.win_c[j]);
                }
                buf[j] = x;
            }
        }
        __syncthreads();

        uint32_t log2ns = 0;
        if (log2n & 1) {
            // radix-2 pass (Ns = 1: no twiddles)
            const uint32_t total = n >> 1;
            float2 r[2 * BPT][2];
#pragma unroll
            for (int b = 0; b < 2 * BPT; ++b) {
                const uint32_t j = lane + b * tpr;
                if (j < total) {
                    r[b][0] = buf[j];
                    r[b][1] = buf[j + total];
                    bfly2(r[b][0], r[b][1]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 2 * BPT; ++b) {
                const uint32_t j = lane + b * tpr;
                if (j < total) {
                    buf[2 * j] = r[b][0];
                    buf[2 * j + 1] = r[b][1];
                }
            }
            __syncthreads();
            log2ns = 1;
        }
        for (; log2ns < static_cast<uint32_t>(log2n); log2ns += 2) {
            const uint32_t total = n >> 2;
            const uint32_t ns = 1u << log2ns;
            const uint32_t shift = log2n - 2 - log2ns;  // W_{4Ns}^{e} = W_n^{e << shift}
            float2 r[BPT][4];
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                const uint32_t j = lane + b * tpr;
                if (j < total) {
                    const uint32_t k = j & (ns - 1);
                    r[b][0] = buf[j];
                    r[b][1] = buf[j + total];
                    r[b][2] = buf[j + 2 * total];
                    r[b][3] = buf[j + 3 * total];
                    if (k != 0) {
                        r[b][1] = cmul(r[b][1], p.twiddle[k << shift]);
                        r[b][2] = cmul(r[b][2], p.twiddle[(2 * k) << shift]);
                        r[b][3] = cmul(r[b][3], p.twiddle[(3 * k) << shift]);
                    }
                    bfly4(r[b][0], r[b][1], r[b][2], r[b][3]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                const uint32_t j = lane + b * tpr;
                if (j < total) {
                    const uint32_t k = j & (ns - 1);
                    const uint32_t j0 = ((j - k) << 2) + k;
                    buf[j0] = r[b][0];
                    buf[j0 + ns] = r[b][1];
                    buf[j0 + 2 * ns] = r[b][2];
                    buf[j0 + 3 * ns] = r[b][3];
                }
            }
            __syncthreads();
        }

        if (valid) {
            if constexpr (MODE == MODE_C2C) {
                float2* const dst = static_cast<float2*>(p.out) + row * n;
                for (uint32_t j = lane; j < n; j += tpr) {
                    float2 X = buf[j];
                    if (p.inverse) {
                        X = make_float2(X.y, X.x);
                    }
                    stg_stream_f2(dst + j, X);
                }
            } else {
                float* const dst = static_cast<float*>(p.out) + row * n;
                for (uint32_t j = lane; j < n; j += tpr) {
                    stg_stream_f1(dst + j, spectral_epilogue<MODE>(buf[j], p));
                }
            }
        }
        __syncthreads();
    }
}

// ---- launch logic ------------------------------------------------------------------------------

constexpr uint64_t kMaxGenericN = 16384;

static bool fft_size_supported(const uint64_t n) { return is_pow2(n) && n >= 2 && n <= kMaxGenericN; }   // fused chain kernels

// CTAs per SM for the 4096 kernel (2: 3-deep TMA ring, <=128 regs; 3: 2-deep ring, <=80 regs).
// Overridable for A/B measurements: B200_FFT4096_CTAS=2|3.
static int fft4096_ctas() {
    static const int value = [] {
        const char* env = getenv("B200_FFT4096_CTAS");
        const int v = env ? atoi(env) : 0;
        return (v == 2 || v == 3) ? v : B200_FFT4096_DEFAULT_CTAS;
    }();
    return value;
}

// ---- tensor maps (2-D TMA tiles of the column pass, fft_tile.cuh) ------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library does not link libcuda.
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static const TensorMapEncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult status = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &status) != cudaSuccess ||
            status != cudaDriverEntryPointSuccess) {
            f = nullptr;
        }
        return reinterpret_cast<TensorMapEncodeTiledFn>(f);
    }();
    return fn;
}

template <int MODE, int WIN, int CTAS>
static int launch_4096_ctas(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    auto kernel = fft4096_kernel<MODE, WIN, CTAS>;
    constexpr int smem = fft4096_smem_bytes(CTAS);
    // Set on every launch (a cheap runtime call): the attribute belongs to the CUDA *context*, and a host such as the
    // reference's CUDA backend runs this library inside its own driver context next to the primary one.
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * CTAS;
    const unsigned grid = static_cast<unsigned>(p.rows < cap ? p.rows : cap);
    kernel<<<grid, kFft4096Threads, smem, stream>>>(p);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

// Fused complex-integer ingest (cast -> window -> fft -> [agc] -> amplitude -> range in one kernel), real window only.
template <int MODE, int ITYPE, bool AGC = false, bool COLSUM = false>
static int launch_4096_int(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream, unsigned* grid_out = nullptr) {
    auto kernel = fft4096_kernel<MODE, WIN_REAL, 2, ITYPE, AGC, COLSUM>;
    constexpr int smem = ITYPE == IN_CF32 ? fft4096_smem_bytes(2) : fft4096_int_smem_bytes(ITYPE);
    // Set on every launch (a cheap runtime call): the attribute belongs to the CUDA *context*, and a host such as the
    // reference's CUDA backend runs this library inside its own driver context next to the primary one.
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * 2;
    const unsigned grid = static_cast<unsigned>(p.rows < cap ? p.rows : cap);
    if (grid_out) {
        *grid_out = grid;
    }
    kernel<<<grid, kFft4096Threads, smem, stream>>>(p);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <int MODE, int WIN>
static int launch_4096(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    if (fft4096_ctas() == 3) {
        return launch_4096_ctas<MODE, WIN, 3>(ctx, p, stream);
    }
    return launch_4096_ctas<MODE, WIN, 2>(ctx, p, stream);
}

template <int MODE, int WIN, int BPT>
static int launch_generic_bpt(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    const int log2n = ilog2(p.n);
    int tpr = static_cast<int>(p.n / (4 * BPT));
    if (tpr < 1) {
        tpr = 1;
    }
    int rows_per_cta = tpr >= 256 ? 1 : 256 / tpr;
    if (static_cast<uint64_t>(rows_per_cta) > p.rows) {
        rows_per_cta = static_cast<int>(p.rows);
    }
    const int threads = tpr * rows_per_cta;
    const size_t smem = static_cast<size_t>(rows_per_cta) * p.n * sizeof(float2);
    auto kernel = fft_generic_kernel<MODE, WIN, BPT>;
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024)));
    int per_sm = 1;
    B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) {
        per_sm = 1;
    }
    const uint64_t groups = (p.rows + rows_per_cta - 1) / rows_per_cta;
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * per_sm;
    const unsigned grid = static_cast<unsigned>(groups < cap ? groups : cap);
    kernel<<<grid, threads, smem, stream>>>(p, log2n, tpr, rows_per_cta);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

template <int LOG2N, int MODE, int WIN>
static int launch_radix(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    auto kernel = fft_radix_kernel<LOG2N, MODE, WIN>;
    constexpr int smem = radix_smem_bytes(LOG2N);
    constexpr int threads = radix_threads(LOG2N);
    constexpr int per_sm = threads > 256 ? 1 : 2;
    // Set on every launch (a cheap runtime call): the attribute belongs to the CUDA *context*, and a host such as the
    // reference's CUDA backend runs this library inside its own driver context next to the primary one.
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint64_t rows_per_block = static_cast<uint64_t>(threads) * 16 >> LOG2N;
    const uint64_t blocks = (p.rows + rows_per_block - 1) / rows_per_block;
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * per_sm;
    kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), threads, smem, stream>>>(p);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

// Register-radix kernel for the (mode, window) combinations the modules actually use; 0 = not handled here.
template <int MODE, int WIN>
static int try_launch_radix(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream, bool* handled) {
    *handled = false;
    constexpr bool kInstantiated = (MODE == MODE_C2C && WIN == WIN_NONE) || (MODE != MODE_C2C && WIN == WIN_REAL);
    if constexpr (kInstantiated) {
        if ((reinterpret_cast<uintptr_t>(p.in) & 15u) != 0) {
            return B200_SUCCESS;
        }
        *handled = true;
        switch (p.n) {
            case 16: return launch_radix<4, MODE, WIN>(ctx, p, stream);
            case 32: return launch_radix<5, MODE, WIN>(ctx, p, stream);
            case 64: return launch_radix<6, MODE, WIN>(ctx, p, stream);
            case 128: return launch_radix<7, MODE, WIN>(ctx, p, stream);
            case 256: return launch_radix<8, MODE, WIN>(ctx, p, stream);
            case 512: return launch_radix<9, MODE, WIN>(ctx, p, stream);
            case 1024: return launch_radix<10, MODE, WIN>(ctx, p, stream);
            case 2048: return launch_radix<11, MODE, WIN>(ctx, p, stream);
            case 4096: return launch_radix<12, MODE, WIN>(ctx, p, stream);
            case 8192: return launch_radix<13, MODE, WIN>(ctx, p, stream);
            default: *handled = false; return B200_SUCCESS;
        }
    }
    return B200_SUCCESS;
}

// Real rows of length 2h, 32 <= h <= 8192: the h-point kernel unpacks in its own epilogue (MODE_R2C, fft_radix.cuh).
static int launch_radix_real(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    switch (p.n) {
        case 32: return launch_radix<5, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 64: return launch_radix<6, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 128: return launch_radix<7, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 256: return launch_radix<8, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 512: return launch_radix<9, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 1024: return launch_radix<10, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 2048: return launch_radix<11, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 4096: return launch_radix<12, MODE_R2C, WIN_NONE>(ctx, p, stream);
        case 8192: return launch_radix<13, MODE_R2C, WIN_NONE>(ctx, p, stream);
        default: return fail("fused real transform: unsupported half length %u", p.n);
    }
}

static bool fft4096_use_generic() {
    static const bool value = [] {
        const char* env = getenv("B200_FFT4096_GENERIC");
        return env && atoi(env) != 0;
    }();
    return value;
}

template <int MODE, int WIN>
static int launch_fft(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    if (p.rows == 0) {
        return B200_SUCCESS;
    }
    // n == 4096: the fused chain runs the dedicated 3-stage kernel (measured 0.644 ms vs 0.670 ms per 65536 rows);
    // the plain C2C transform runs the 2-barrier radix kernel (0.659 ms vs 0.742 ms, 99 % of the measured HBM peak).
    if (p.n == kFft4096N && MODE != MODE_C2C && (reinterpret_cast<uintptr_t>(p.in) & 15u) == 0 &&
        !fft4096_use_generic()) {
        return launch_4096<MODE, WIN>(ctx, p, stream);
    }
    bool handled = false;
    const int rc = try_launch_radix<MODE, WIN>(ctx, p, stream, &handled);
    if (handled || rc != B200_SUCCESS) {
        return rc;
    }
    if (p.n <= 1024) {
        return launch_generic_bpt<MODE, WIN, 1>(ctx, p, stream);
    }
    if (p.n == 2048) {
        return launch_generic_bpt<MODE, WIN, 2>(ctx, p, stream);
    }
    return launch_generic_bpt<MODE, WIN, 4>(ctx, p, stream);
}

// ---- two-pass plan (fft_twopass.cuh) -----------------------------------------------------------------------------
constexpr uint64_t kTwoPassMinN = 16384, kTwoPassMaxN = 131072;      // n = 16 M, 1024 <= M <= 8192

// Read at plan creation (A/B measurements in one process): B200_FFT_TWOPASS=0 falls back to the four-step plan,
// B200_FFT_TWOPASS_MIN_N lowers the first two-pass length (down to 4096; default 16384), B200_FFT_TWOPASS_CHUNK_MB is the
// chunk (= L2-resident scratch) size, B200_FFT_TWOPASS_HINTS=0 drops the L2 eviction-priority hints.
static bool twopass_selected(const uint64_t n) {
    const char* env = getenv("B200_FFT_TWOPASS");
    if (env && atoi(env) == 0) {
        return false;
    }
    const char* lo = getenv("B200_FFT_TWOPASS_MIN_N");
    const uint64_t min_n = lo && atol(lo) >= 4096 ? static_cast<uint64_t>(atol(lo)) : kTwoPassMinN;
    return is_pow2(n) && n >= min_n && n <= kTwoPassMaxN;
}
static uint64_t twopass_chunk_bytes() {
    const char* env = getenv("B200_FFT_TWOPASS_CHUNK_MB");
    const long v = env ? atol(env) : 0;
    return static_cast<uint64_t>(v > 0 ? v : 64) << 20;      // measured: 64-96 MB is the optimum on a 126 MB L2
}
static int twopass_hints() {
    const char* env = getenv("B200_FFT_TWOPASS_HINTS");
    return env ? atoi(env) : 1;
}

static int launch_col16(const b200_ctx* ctx, const Col16Params& p, cudaStream_t stream) {
    const unsigned col_blocks = p.m / kCol16Threads;
    const uint64_t cap = (static_cast<uint64_t>(ctx->sms) * 2 + col_blocks - 1) / col_blocks;   // 2 CTAs per SM
    const dim3 grid(col_blocks, static_cast<unsigned>(p.rows < cap ? p.rows : cap));
    if (p.hints) {
        fft_col16_kernel<true><<<grid, kCol16Threads, 0, stream>>>(p);
    } else {
        fft_col16_kernel<false><<<grid, kCol16Threads, 0, stream>>>(p);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

static bool twopass_tiled(const uint64_t n) {
    const char* env = getenv("B200_FFT_TWOPASS_TILE");
    return !(env && atoi(env) == 0) && (n == 16384 || n == 32768 || n == 65536) && tensor_map_encoder() != nullptr;
}

// Launch with (pdl) or without the programmatic-stream-serialization attribute: with it the kernel may be scheduled while
// its predecessor on the stream still runs and orders itself with griddepcontrol.wait (fft_tile.cuh).
template <typename... KernelArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KernelArgs...), const unsigned grid, const int smem, cudaStream_t stream,
                              const bool pdl, Args... args) {
    cudaLaunchConfig_t config{};
    config.gridDim = dim3(grid);
    config.blockDim = dim3(kTileThreads);
    config.dynamicSmemBytes = static_cast<size_t>(smem);
    config.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    config.attrs = &attr;
    config.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&config, kernel, args...);
}
static bool twopass_pdl() {
    const char* env = getenv("B200_FFT_TWOPASS_PDL");
    return !(env && atoi(env) == 0);
}

static int launch_tile_cols(const b200_ctx* ctx, const TileParams& p, cudaStream_t stream, const bool pdl) {
    const uint64_t columns_per_tile = kTileElems / p.m1;
    const uint64_t tiles = p.transforms * (kTileRowLen / columns_per_tile);
    void (*kernel)(TileParams, CUtensorMap) = nullptr;
#define B200_COLS_PICK(WIN)                                                                                        \
    kernel = p.m1 == 64 ? fft_cols_kernel<6, WIN> : (p.m1 == 128 ? fft_cols_kernel<7, WIN> : fft_cols_kernel<8, WIN>)
    if (p.win_re) {
        B200_COLS_PICK(WIN_REAL);
    } else if (p.win_c) {
        B200_COLS_PICK(WIN_COMPLEX);
    } else {
        B200_COLS_PICK(WIN_NONE);
    }
#undef B200_COLS_PICK
    B200_REQUIRE(p.m1 == 64 || p.m1 == 128 || p.m1 == 256, "tiled two-pass fft: unsupported column length %u", p.m1);
    B200_REQUIRE(tensor_map_encoder() != nullptr, "tiled two-pass fft: cuTensorMapEncodeTiled is not available");
    // the chunk as [transforms * M1 lines][512 floats]; one box = one tile = M1 lines x 2 C floats
    CUtensorMap map;
    const cuuint64_t dims[2] = {2 * kTileRowLen, p.transforms * p.m1};
    const cuuint64_t strides[1] = {2 * kTileRowLen * sizeof(float)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(2 * columns_per_tile), p.m1};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult rc = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(p.in), dims, strides,
                                             box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        return fail("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(rc));
    }
    const int smem = cols_smem_bytes(static_cast<int>(p.m1));
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * 2;
    const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
    B200_CUDA_CHECK(launch_pdl(kernel, grid, smem, stream, pdl, p, map));
    return B200_SUCCESS;
}

static int launch_tile_rows(const b200_ctx* ctx, const TileParams& p, cudaStream_t stream, const bool pdl,
                            const int mode = MODE_C2C) {
    void (*kernel)(TileParams) = mode == MODE_AMP ? fft_rows256_kernel<MODE_AMP>
                                                  : (mode == MODE_AMP_RANGE ? fft_rows256_kernel<MODE_AMP_RANGE>
                                                                            : fft_rows256_kernel<MODE_C2C>);
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRows256SmemBytes));
    const uint64_t blocks = p.transforms * (p.m1 / 16);
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * 2;
    B200_CUDA_CHECK(launch_pdl(kernel, static_cast<unsigned>(blocks < cap ? blocks : cap), kRows256SmemBytes, stream, pdl, p));
    return B200_SUCCESS;
}

static int launch_radix_transposed(const b200_ctx* ctx, const FftParams& p, cudaStream_t stream) {
    switch (p.n) {
        case 256: return launch_radix<8, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        case 512: return launch_radix<9, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        case 1024: return launch_radix<10, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        case 2048: return launch_radix<11, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        case 4096: return launch_radix<12, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        case 8192: return launch_radix<13, MODE_C2C_T, WIN_NONE>(ctx, p, stream);
        default: return fail("two-pass fft: unsupported row length %u", p.n);
    }
}

static int make_twiddle_table(b200_ctx* ctx, const uint64_t n, float2** out) {
    std::vector<float2> host(n);
    const double kTwoPi = 6.283185307179586476925286766559;
    for (uint64_t j = 0; j < n; ++j) {
        const double a = -kTwoPi * static_cast<double>(j) / static_cast<double>(n);
        host[j] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
    }
    void* dev = nullptr;
    if (b200_malloc(ctx, n * sizeof(float2), &dev) != B200_SUCCESS) {
        return B200_ERROR;
    }
    const cudaError_t e = cudaMemcpy(dev, host.data(), n * sizeof(float2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return fail("twiddle upload failed: %s", cudaGetErrorString(e));
    }
    *out = static_cast<float2*>(dev);
    return B200_SUCCESS;
}

}  // namespace b200

using namespace b200;

enum FftPlanKind { FFT_DIRECT = 0, FFT_FOURSTEP = 1, FFT_BLUESTEIN = 2, FFT_TWOPASS = 3 };

struct b200_fft_plan {
    b200_ctx* ctx;
    uint64_t n;
    uint64_t batch;
    float2* twiddle;
    int kind = FFT_DIRECT;
    // four-step (n = n1 n2, power of two above the single-CTA limit): transposes + two batched sub-transforms
    uint64_t n1 = 0, n2 = 0;
    b200_fft_plan* sub1 = nullptr;
    b200_fft_plan* sub2 = nullptr;
    float2* step_twiddle = nullptr;   // [n2][n1]: W_n^(n2 k1)
    // Bluestein (any n): chirp-z through a power-of-two circular convolution of length m >= 2n - 1
    uint64_t m = 0;
    b200_fft_plan* subm = nullptr;
    float2* chirp = nullptr;          // [n]  exp(-j pi k^2 / n)
    float2* chirp_spec_fwd = nullptr; // [m]  FFT_m of the wrapped conj chirp (forward transform)
    float2* chirp_spec_inv = nullptr; // [m]  same for the inverse transform (conjugated chirps)
    float2* scratch_a = nullptr;
    float2* scratch_b = nullptr;
    // two-pass (n = 16 M, 16384 <= n <= 131072; fft_twopass.cuh): sub2 carries the W_M table, scratch_a one chunk
    uint64_t chunk_rows = 0;
    int hints = 1;
    // tiled form (n = M1 x 256, fft_tile.cuh): sub1 carries the W_M1 table, sub2 the W_256 table, step_twiddle [M1][256]
    bool tiled = false;
    float2* real_work = nullptr;      // b200_fft_exec_real: [batch, n] spectra of the even/odd-packed rows
};

namespace b200 {

// Real input of length 2h through ONE complex transform of length h (b200_fft_exec_real): z[m] = x[2m] + i x[2m+1] is the
// input array itself viewed as CF32; with Z = FFT_h(z),
//   X[k] = (Z[k] + conj(Z[h-k])) / 2  +  W_2h^k (Z[k] - conj(Z[h-k])) / (2i),   k = 0 .. h   (Z[h] = Z[0]).
// layout 0: [batch, h + 1] CF32 (pocketfft::r2c); layout 1: FFTPACK half-complex [Re X0, Re X1, Im X1, ..., Re Xh] F32.
// One thread per mirror pair (k, h - k): both outputs come from the same two inputs and one twiddle,
//   X[k] = e + w o,  X[h-k] = conj(e - w o),  e = (Z[k] + conj Z[h-k]) / 2,  o = (Z[k] - conj Z[h-k]) / (2i),  w = W_2h^k.
__device__ __forceinline__ void rfft_store(void* out, const int layout, const uint64_t row, const uint64_t h, const uint64_t k,
                                           const float2 x) {
    if (layout == 0) {
        static_cast<float2*>(out)[row * (h + 1) + k] = x;
    } else {
        float* const r = static_cast<float*>(out) + row * (2 * h);
        if (k == 0) {
            r[0] = x.x;
        } else if (k == h) {
            r[2 * h - 1] = x.x;
        } else {
            r[2 * k - 1] = x.x;
            r[2 * k] = x.y;
        }
    }
}
__global__ void rfft_unpack_kernel(const float2* __restrict__ z, void* __restrict__ out, const uint64_t batch,
                                   const uint64_t h, const int layout) {
    const uint64_t per_row = h / 2 + 1, total = batch * per_row;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / per_row, k = i - row * per_row;
        const float2* const zr = z + row * h;
        if (k == 0) {
            const float2 a = zr[0];
            rfft_store(out, layout, row, h, 0, make_float2(a.x + a.y, 0.0f));
            rfft_store(out, layout, row, h, h, make_float2(a.x - a.y, 0.0f));
            continue;
        }
        const float2 a = zr[k];
        const float2 m = zr[h - k];
        const float2 e = make_float2(0.5f * (a.x + m.x), 0.5f * (a.y - m.y));          // (a + conj m) / 2
        const float2 o = make_float2(0.5f * (a.y + m.y), -0.5f * (a.x - m.x));         // (a - conj m) / (2i)
        float sn, cs;
        sincospif(-static_cast<float>(k) / static_cast<float>(h), &sn, &cs);           // W_2h^k = exp(-i pi k / h)
        const float2 t = make_float2(cs * o.x - sn * o.y, cs * o.y + sn * o.x);
        rfft_store(out, layout, row, h, k, make_float2(e.x + t.x, e.y + t.y));
        if (2 * k != h) {
            rfft_store(out, layout, row, h, h - k, make_float2(e.x - t.x, -(e.y - t.y)));
        }
    }
}

constexpr uint64_t kMaxDirectN = 8192;      // single-CTA register-radix kernel (fft_radix.cuh); 16384 via fft_generic_kernel

// data[row][k1] *= tw[(row mod n2)][k1] (optionally conjugated): the four-step inter-stage twiddle.
__global__ void fourstep_twiddle_kernel(float2* __restrict__ data, const float2* __restrict__ tw, const uint64_t total,
                                        const uint64_t n1, const uint64_t n2, const int conj) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t k1 = i % n1, row = i / n1;
        float2 w = tw[(row % n2) * n1 + k1];
        if (conj) {
            w.y = -w.y;
        }
        data[i] = cmul(data[i], w);
    }
}

// [B][d1][d2] -> [B][d2][d1]
__global__ void transpose_kernel(const float2* __restrict__ in, float2* __restrict__ out, const uint64_t batch,
                                 const uint64_t d1, const uint64_t d2) {
    __shared__ float2 tile[32][33];
    const uint64_t tiles_x = (d2 + 31) / 32, tiles_y = (d1 + 31) / 32;
    const uint64_t tiles = batch * tiles_x * tiles_y;
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint64_t b = t / (tiles_x * tiles_y), r = t % (tiles_x * tiles_y);
        const uint64_t ty = r / tiles_x, tx = r % tiles_x;
        const float2* src = in + b * d1 * d2;
        float2* dst = out + b * d1 * d2;
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const uint64_t y = ty * 32 + j, x = tx * 32 + threadIdx.x;
            if (y < d1 && x < d2) {
                tile[j][threadIdx.x] = src[y * d2 + x];
            }
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            const uint64_t x = tx * 32 + j, y = ty * 32 + threadIdx.x;
            if (x < d2 && y < d1) {
                dst[x * d1 + y] = tile[threadIdx.x][j];
            }
        }
    }
}

// Bluestein pre: u[row][j] = x[row][j] * chirp[j] (j < n), 0 (n <= j < m); post: X[row][k] = chirp[k] * v[row][k] / m
__global__ void bluestein_pre_kernel(const float2* __restrict__ x, float2* __restrict__ u, const float2* __restrict__ chirp,
                                     const uint64_t batch, const uint64_t n, const uint64_t m, const int conj) {
    const uint64_t total = batch * m;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / m, j = i - row * m;
        float2 v = make_float2(0.f, 0.f);
        if (j < n) {
            float2 c = chirp[j];
            if (conj) {
                c.y = -c.y;
            }
            v = cmul(x[row * n + j], c);
        }
        u[i] = v;
    }
}
__global__ void bluestein_post_kernel(const float2* __restrict__ v, float2* __restrict__ out,
                                      const float2* __restrict__ chirp, const uint64_t batch, const uint64_t n,
                                      const uint64_t m, const float scale, const int conj) {
    const uint64_t total = batch * n;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t row = i / n, k = i - row * n;
        float2 c = chirp[k];
        if (conj) {
            c.y = -c.y;
        }
        const float2 r = cmul(v[row * m + k], c);
        out[i] = make_float2(r.x * scale, r.y * scale);
    }
}
__global__ void pointwise_rowbcast_mul_kernel(float2* __restrict__ data, const float2* __restrict__ spec,
                                              const uint64_t total, const uint64_t m) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        data[i] = cmul(data[i], spec[i % m]);
    }
}

}  // namespace b200

namespace b200 {
// viz.cu: row-split column sums of a [batches, columns] F32 matrix (the lineplot consumer's batch sum)
int colsum_partials(b200_ctx* ctx, const float* in, uint64_t batches, uint64_t columns, uint64_t batch_stride,
                    uint64_t col_stride, float* partial, uint64_t* splits_out, cudaStream_t s);
int colsum_reduce(const float* partial, uint64_t splits, uint64_t n, float* out, cudaStream_t s);
uint64_t colsum_max_splits(uint64_t batches);
}  // namespace b200

struct b200_chain_plan {
    b200_ctx* ctx;
    uint64_t n;
    uint64_t max_batch;
    float2* twiddle;
    float* win_re;    // non-null: window is purely real
    float2* win_c;    // non-null: general complex window
    const char* variant;
    bool composite = false;           // lengths without a single-kernel path
    bool tiled = false;               // n = 16384 / 32768 / 65536: two fused kernels over the tiled two-pass plan (fft_tile.cuh)
    b200_fft_plan* fft = nullptr;
    float2* scratch = nullptr;
    float2* cast_scratch = nullptr;   // typed input without a fused path: cast -> CF32 here, then the CF32 chain
    uint64_t cast_rows = 0;
    float* colsum_partial = nullptr;  // b200_chain_exec_colsum: per-CTA (fused) or per-split (fallback) partial sums
    uint64_t colsum_bytes = 0;
    // Host-buffer pipeline (b200_chain_exec_host): kHostSlots device staging slots, three streams.
    static constexpr int kHostSlots = 3;
    uint64_t host_chunk_rows = 0;
    float2* stage_in[kHostSlots] = {nullptr, nullptr, nullptr};
    float* stage_out[kHostSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t s_h2d = nullptr, s_exec = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[kHostSlots] = {}, ev_exec[kHostSlots] = {}, ev_out[kHostSlots] = {};
};

extern "C" {

static int fft_exec_impl(b200_fft_plan* plan, const float2* in, float2* out, int forward, cudaStream_t s);

static int upload(b200_ctx* ctx, const std::vector<float2>& host, float2** dev) {
    void* p = nullptr;
    if (b200_malloc(ctx, host.size() * sizeof(float2), &p) != B200_SUCCESS) {
        return B200_ERROR;
    }
    if (cudaMemcpy(p, host.data(), host.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p);
        return fail("device upload failed");
    }
    *dev = static_cast<float2*>(p);
    return B200_SUCCESS;
}

int b200_fft_plan_c2c(b200_ctx* ctx, uint64_t n, uint64_t batch, b200_fft_plan** plan) {
    B200_REQUIRE(ctx && plan, "b200_fft_plan_c2c: null argument");
    *plan = nullptr;
    B200_REQUIRE(n >= 1, "b200_fft_plan_c2c: transform length must be positive");
    B200_REQUIRE(n <= (1ull << 27), "b200_fft_plan_c2c: n=%llu exceeds the supported 2^27",
                 static_cast<unsigned long long>(n));
    DeviceGuard guard(ctx);
    auto* pl = new b200_fft_plan();
    pl->ctx = ctx;
    pl->n = n;
    pl->batch = batch;
    pl->twiddle = nullptr;
    const double kPi = 3.14159265358979323846;
    if (n == 1 || (is_pow2(n) && n <= kMaxDirectN && !twopass_selected(n))) {
        pl->kind = FFT_DIRECT;
        if (n > 1 && make_twiddle_table(ctx, n, &pl->twiddle) != B200_SUCCESS) {
            delete pl;
            return B200_ERROR;
        }
    } else if (twopass_selected(n)) {
        // two-pass: radix-16 column pass + 16 row transforms of length n / 16 with a stride-16 store
        pl->kind = FFT_TWOPASS;
        pl->n1 = 16;
        pl->n2 = n / 16;
        pl->hints = twopass_hints();
        const uint64_t fit = twopass_chunk_bytes() / (n * sizeof(float2));
        pl->chunk_rows = std::min<uint64_t>(std::max<uint64_t>(fit, 1), std::max<uint64_t>(batch, 1));
        pl->tiled = twopass_tiled(n);
        int rc = B200_SUCCESS;
        if (pl->tiled) {
            pl->n1 = n / kTileRowLen;
            pl->n2 = kTileRowLen;
            std::vector<float2> tw(n);                   // [k1][n2]: W_n^(k1 n2), F64-evaluated
            for (uint64_t a = 0; a < pl->n1; ++a) {
                for (uint64_t c = 0; c < pl->n2; ++c) {
                    const double ang = -2.0 * kPi * static_cast<double>((a * c) % n) / static_cast<double>(n);
                    tw[a * pl->n2 + c] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
                }
            }
            rc = upload(ctx, tw, &pl->step_twiddle);
            rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, pl->n1, 0, &pl->sub1) : rc;   // carries the W_M1 table
        } else {
            rc = make_twiddle_table(ctx, n, &pl->twiddle);
        }
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, pl->n2, 0, &pl->sub2) : rc;     // carries the W_M table
        void* a = nullptr;
        if (rc == B200_SUCCESS && batch > 0) {
            rc = b200_malloc(ctx, pl->chunk_rows * n * sizeof(float2), &a);
        }
        pl->scratch_a = static_cast<float2*>(a);
        if (rc != B200_SUCCESS) {
            b200_fft_plan_destroy(pl);
            return B200_ERROR;
        }
    } else if (is_pow2(n)) {
        // four-step: n = n1 n2 with both factors <= kMaxDirectN
        pl->kind = FFT_FOURSTEP;
        const int k = ilog2(n);
        pl->n1 = 1ull << (k / 2);
        pl->n2 = n / pl->n1;
        std::vector<float2> tw(n);
        for (uint64_t a = 0; a < pl->n2; ++a) {
            for (uint64_t c = 0; c < pl->n1; ++c) {
                const double ang = -2.0 * kPi * static_cast<double>((a * c) % n) / static_cast<double>(n);
                tw[a * pl->n1 + c] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
            }
        }
        int rc = upload(ctx, tw, &pl->step_twiddle);
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, pl->n1, batch * pl->n2, &pl->sub1) : rc;
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, pl->n2, batch * pl->n1, &pl->sub2) : rc;
        void* a = nullptr;
        void* b = nullptr;
        if (rc == B200_SUCCESS && batch > 0) {
            rc = b200_malloc(ctx, batch * n * sizeof(float2), &a);
            rc = rc == B200_SUCCESS ? b200_malloc(ctx, batch * n * sizeof(float2), &b) : rc;
        }
        pl->scratch_a = static_cast<float2*>(a);
        pl->scratch_b = static_cast<float2*>(b);
        if (rc != B200_SUCCESS) {
            b200_fft_plan_destroy(pl);
            return B200_ERROR;
        }
    } else {
        // Bluestein
        pl->kind = FFT_BLUESTEIN;
        uint64_t m = 1;
        while (m < 2 * n - 1) {
            m <<= 1;
        }
        pl->m = m;
        std::vector<float2> chirp(n), wrapped(m, make_float2(0.f, 0.f)), wrapped_inv(m, make_float2(0.f, 0.f));
        for (uint64_t j = 0; j < n; ++j) {
            const uint64_t sq = (j * j) % (2 * n);                 // exact phase reduction
            const double ang = -kPi * static_cast<double>(sq) / static_cast<double>(n);
            chirp[j] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
            const float2 b = make_float2(chirp[j].x, -chirp[j].y);   // conj chirp
            wrapped[j] = b;
            wrapped_inv[j] = chirp[j];
            if (j > 0) {
                wrapped[m - j] = b;
                wrapped_inv[m - j] = chirp[j];
            }
        }
        int rc = upload(ctx, chirp, &pl->chirp);
        rc = rc == B200_SUCCESS ? upload(ctx, wrapped, &pl->chirp_spec_fwd) : rc;
        rc = rc == B200_SUCCESS ? upload(ctx, wrapped_inv, &pl->chirp_spec_inv) : rc;
        b200_fft_plan* one = nullptr;
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, m, 1, &one) : rc;
        if (rc == B200_SUCCESS) {      // spectra of the two wrapped chirps (once)
            rc = fft_exec_impl(one, pl->chirp_spec_fwd, pl->chirp_spec_fwd, 1, nullptr);
            rc = rc == B200_SUCCESS ? fft_exec_impl(one, pl->chirp_spec_inv, pl->chirp_spec_inv, 1, nullptr) : rc;
            cudaDeviceSynchronize();
        }
        b200_fft_plan_destroy(one);
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, m, batch, &pl->subm) : rc;
        void* a = nullptr;
        if (rc == B200_SUCCESS && batch > 0) {
            rc = b200_malloc(ctx, batch * m * sizeof(float2), &a);
        }
        pl->scratch_a = static_cast<float2*>(a);
        if (rc != B200_SUCCESS) {
            b200_fft_plan_destroy(pl);
            return B200_ERROR;
        }
    }
    cudaStreamSynchronize(cudaStreamLegacy);   // uploads / zero-fills above ran on the legacy stream: settle them before a non-blocking stream executes
    *plan = pl;
    return B200_SUCCESS;
}

static int fft_exec_impl(b200_fft_plan* plan, const float2* in, float2* out, int forward, cudaStream_t s) {
    const b200_ctx* ctx = plan->ctx;
    const unsigned cap = static_cast<unsigned>(ctx->sms * 8);
    auto grid_for = [cap](uint64_t items) {
        const uint64_t g = (items + 255) / 256;
        return static_cast<unsigned>(g < cap ? (g ? g : 1) : cap);
    };
    if (plan->kind == FFT_DIRECT) {
        if (plan->n == 1) {
            if (in != out) {
                B200_CUDA_CHECK(cudaMemcpyAsync(out, in, plan->batch * sizeof(float2), cudaMemcpyDeviceToDevice, s));
            }
            return B200_SUCCESS;
        }
        FftParams p{};
        p.in = in;
        p.out = out;
        p.rows = plan->batch;
        p.n = static_cast<uint32_t>(plan->n);
        p.inverse = forward ? 0 : 1;
        p.twiddle = plan->twiddle;
        return launch_fft<MODE_C2C, WIN_NONE>(plan->ctx, p, s);
    }
    const uint64_t B = plan->batch, n = plan->n;
    if (plan->kind == FFT_TWOPASS) {
        for (uint64_t row0 = 0; row0 < B; row0 += plan->chunk_rows) {
            const uint64_t rows = std::min(plan->chunk_rows, B - row0);
            if (plan->tiled) {
                TileParams t{};
                t.in = in + row0 * n;
                if ((reinterpret_cast<uintptr_t>(t.in) & 15u) != 0) {
                    // the column tiles land by TMA (16-byte aligned sources): a view that starts on an odd sample is
                    // staged through an aligned copy of the chunk
                    if (!plan->scratch_b) {
                        void* b = nullptr;
                        if (b200_malloc(plan->ctx, plan->chunk_rows * n * sizeof(float2), &b) != B200_SUCCESS) {
                            return B200_ERROR;
                        }
                        plan->scratch_b = static_cast<float2*>(b);
                    }
                    B200_CUDA_CHECK(cudaMemcpyAsync(plan->scratch_b, t.in, rows * n * sizeof(float2), cudaMemcpyDeviceToDevice, s));
                    t.in = plan->scratch_b;
                }
                t.out = plan->scratch_a;
                t.transforms = rows;
                t.m1 = static_cast<uint32_t>(plan->n1);
                t.inverse = forward ? 0 : 1;
                t.table = plan->sub1->twiddle;
                t.stage_tw = plan->step_twiddle;
                t.hints = plan->hints;
                // the first kernel of an exec orders itself against whatever precedes it on the stream the usual way
                const bool pdl = twopass_pdl();
                if (launch_tile_cols(ctx, t, s, pdl && row0 > 0) != B200_SUCCESS) {
                    return B200_ERROR;
                }
                t.in = plan->scratch_a;
                t.out = out + row0 * n;
                t.table = plan->sub2->twiddle;
                if (launch_tile_rows(ctx, t, s, pdl) != B200_SUCCESS) {
                    return B200_ERROR;
                }
                continue;
            }
            Col16Params c{};
            c.in = in + row0 * n;
            c.scratch = plan->scratch_a;
            c.rows = rows;
            c.m = static_cast<uint32_t>(plan->n2);
            c.inverse = forward ? 0 : 1;
            c.twiddle = plan->twiddle;
            c.hints = plan->hints;
            if (launch_col16(ctx, c, s) != B200_SUCCESS) {
                return B200_ERROR;
            }
            FftParams p{};
            p.in = plan->scratch_a;
            p.out = out + row0 * n;
            p.rows = rows * 16;
            p.n = static_cast<uint32_t>(plan->n2);
            p.inverse = forward ? 0 : 1;
            p.twiddle = plan->sub2->twiddle;
            if (launch_radix_transposed(ctx, p, s) != B200_SUCCESS) {
                return B200_ERROR;
            }
        }
        return B200_SUCCESS;
    }
    if (plan->kind == FFT_FOURSTEP) {
        const uint64_t n1 = plan->n1, n2 = plan->n2;
        float2* a = plan->scratch_a;
        float2* b = plan->scratch_b;
        const dim3 tb(32, 8);
        const unsigned tgrid = static_cast<unsigned>(std::min<uint64_t>(B * ((n1 + 31) / 32) * ((n2 + 31) / 32), cap * 4ull));
        // x[b][n1][n2] -> a[b][n2][n1]
        transpose_kernel<<<tgrid, tb, 0, s>>>(in, a, B, n1, n2);
        B200_LAUNCH_CHECK();
        if (fft_exec_impl(plan->sub1, a, a, forward, s) != B200_SUCCESS) {
            return B200_ERROR;
        }
        fourstep_twiddle_kernel<<<grid_for(B * n), 256, 0, s>>>(a, plan->step_twiddle, B * n, n1, n2, forward ? 0 : 1);
        B200_LAUNCH_CHECK();
        // a[b][n2][k1] -> b[b][k1][n2]
        transpose_kernel<<<tgrid, tb, 0, s>>>(a, b, B, n2, n1);
        B200_LAUNCH_CHECK();
        if (fft_exec_impl(plan->sub2, b, b, forward, s) != B200_SUCCESS) {
            return B200_ERROR;
        }
        // b[b][k1][k2] -> out[b][k2][k1]   (k = k1 + n1 k2)
        transpose_kernel<<<tgrid, tb, 0, s>>>(b, out, B, n1, n2);
        B200_LAUNCH_CHECK();
        return B200_SUCCESS;
    }
    // Bluestein
    const uint64_t m = plan->m;
    float2* u = plan->scratch_a;
    const int conj = forward ? 0 : 1;
    bluestein_pre_kernel<<<grid_for(B * m), 256, 0, s>>>(in, u, plan->chirp, B, n, m, conj);
    B200_LAUNCH_CHECK();
    if (fft_exec_impl(plan->subm, u, u, 1, s) != B200_SUCCESS) {
        return B200_ERROR;
    }
    pointwise_rowbcast_mul_kernel<<<grid_for(B * m), 256, 0, s>>>(u, forward ? plan->chirp_spec_fwd : plan->chirp_spec_inv,
                                                                 B * m, m);
    B200_LAUNCH_CHECK();
    if (fft_exec_impl(plan->subm, u, u, 0, s) != B200_SUCCESS) {
        return B200_ERROR;
    }
    bluestein_post_kernel<<<grid_for(B * n), 256, 0, s>>>(u, out, plan->chirp, B, n, m, 1.0f / static_cast<float>(m), conj);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fft_exec(b200_fft_plan* plan, const b200_cf32* in, b200_cf32* out, int forward, b200_stream stream) {
    B200_REQUIRE(plan, "b200_fft_exec: null plan");
    if (plan->batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(in && out, "b200_fft_exec: null buffer");
    DeviceGuard guard(plan->ctx);
    return fft_exec_impl(plan, reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), forward,
                         as_stream(stream));
}

int b200_fft_exec_real(b200_fft_plan* half_plan, const float* in, void* out, int layout, b200_stream stream) {
    B200_REQUIRE(half_plan, "b200_fft_exec_real: null plan");
    B200_REQUIRE(layout == 0 || layout == 1, "b200_fft_exec_real: layout must be 0 (R2C) or 1 (FFTPACK)");
    if (half_plan->batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(in && out, "b200_fft_exec_real: null buffer");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(in) & 7u) == 0, "b200_fft_exec_real: the real input must be 8-byte aligned");
    DeviceGuard guard(half_plan->ctx);
    const uint64_t h = half_plan->n, batch = half_plan->batch;
    // Single-kernel lengths: transform and unpack in ONE kernel (4 bytes in + 4 out per real sample, no intermediate).
    const char* fused_env = getenv("B200_FFT_REAL_FUSED");
    if (half_plan->kind == FFT_DIRECT && is_pow2(h) && h >= 32 && h <= kMaxDirectN &&
        (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && !(fused_env && atoi(fused_env) == 0)) {
        FftParams p{};
        p.in = reinterpret_cast<const float2*>(in);
        p.out = nullptr;
        p.rows = batch;
        p.n = static_cast<uint32_t>(h);
        p.inverse = 0;
        p.twiddle = half_plan->twiddle;
        p.real_out = out;
        p.real_layout = layout;
        return launch_radix_real(half_plan->ctx, p, as_stream(stream));
    }
    if (!half_plan->real_work) {
        void* w = nullptr;
        if (b200_malloc(half_plan->ctx, batch * h * sizeof(float2), &w) != B200_SUCCESS) {
            return B200_ERROR;
        }
        half_plan->real_work = static_cast<float2*>(w);
        cudaStreamSynchronize(cudaStreamLegacy);
    }
    const cudaStream_t s = as_stream(stream);
    if (fft_exec_impl(half_plan, reinterpret_cast<const float2*>(in), half_plan->real_work, 1, s) != B200_SUCCESS) {
        return B200_ERROR;
    }
    const uint64_t items = batch * (h / 2 + 1);
    const uint64_t blocks = (items + 255) / 256, cap = static_cast<uint64_t>(half_plan->ctx->sms) * 8;
    rfft_unpack_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, s>>>(half_plan->real_work, out, batch, h,
                                                                                       layout);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fft_plan_destroy(b200_fft_plan* plan) {
    if (!plan) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(plan->ctx);
    b200_fft_plan_destroy(plan->sub1);
    b200_fft_plan_destroy(plan->sub2);
    b200_fft_plan_destroy(plan->subm);
    cudaFree(plan->twiddle);
    cudaFree(plan->step_twiddle);
    cudaFree(plan->chirp);
    cudaFree(plan->chirp_spec_fwd);
    cudaFree(plan->chirp_spec_inv);
    cudaFree(plan->scratch_a);
    cudaFree(plan->scratch_b);
    cudaFree(plan->real_work);
    delete plan;
    return B200_SUCCESS;
}

int b200_chain_plan_create(b200_ctx* ctx, uint64_t n, uint64_t max_batch, const b200_cf32* window_dev,
                           b200_chain_plan** plan) {
    B200_REQUIRE(ctx && plan, "b200_chain_plan_create: null argument");
    *plan = nullptr;
    B200_REQUIRE(n >= 1 && n <= (1ull << 27), "b200_chain_plan_create: n=%llu out of range",
                 static_cast<unsigned long long>(n));
    DeviceGuard guard(ctx);
    auto* pl = new b200_chain_plan();
    pl->ctx = ctx;
    pl->n = n;
    pl->max_batch = max_batch;
    pl->twiddle = nullptr;
    pl->win_re = nullptr;
    pl->win_c = nullptr;
    pl->variant = "";
    const bool tiled = twopass_selected(n) && twopass_tiled(n);
    if (!fft_size_supported(n) && !tiled) {
        // Lengths the single-kernel paths do not cover (not a power of two, or > 16384): the reference's own module
        // sequence on this provider — multiply -> fft (four-step / Bluestein plan) -> amplitude -> range — through a
        // plan-owned CF32 scratch. Same results, 4+ passes over the data instead of one.
        B200_REQUIRE(window_dev != nullptr, "b200_chain_plan_create: the composite path needs a window");
        pl->composite = true;
        void* w = nullptr;
        void* sc = nullptr;
        int rc = b200_malloc(ctx, n * sizeof(float2), &w);
        rc = rc == B200_SUCCESS ? b200_malloc(ctx, std::max<uint64_t>(1, max_batch) * n * sizeof(float2), &sc) : rc;
        rc = rc == B200_SUCCESS ? b200_fft_plan_c2c(ctx, n, max_batch, &pl->fft) : rc;
        if (rc == B200_SUCCESS &&
            cudaMemcpy(w, window_dev, n * sizeof(float2), cudaMemcpyDeviceToDevice) != cudaSuccess) {
            rc = fail("b200_chain_plan_create: window copy failed");
        }
        pl->win_c = static_cast<float2*>(w);
        pl->scratch = static_cast<float2*>(sc);
        if (rc != B200_SUCCESS) {
            b200_chain_plan_destroy(pl);
            return B200_ERROR;
        }
        pl->variant = "composite<multiply,fft(four-step|bluestein),amplitude,range>";
        cudaStreamSynchronize(cudaStreamLegacy);   // uploads / zero-fills above ran on the legacy stream: settle them before a non-blocking stream executes
        *plan = pl;
        return B200_SUCCESS;
    }
    if (tiled) {
        // window multiply fused into the column pass, amplitude / range into the row pass of the tiled two-pass plan
        pl->tiled = true;
        if (b200_fft_plan_c2c(ctx, n, std::max<uint64_t>(1, max_batch), &pl->fft) != B200_SUCCESS) {
            delete pl;
            return B200_ERROR;
        }
    } else if (make_twiddle_table(ctx, n, &pl->twiddle) != B200_SUCCESS) {
        delete pl;
        return B200_ERROR;
    }
    if (window_dev) {
        // The window is a settled STATIC_OUTPUT tensor: inspect it once.
        std::vector<float2> host(n);
        cudaError_t e = cudaMemcpy(host.data(), window_dev, n * sizeof(float2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            b200_chain_plan_destroy(pl);
            return fail("b200_chain_plan_create: window download failed: %s", cudaGetErrorString(e));
        }
        bool real = true;
        for (uint64_t i = 0; i < n; ++i) {
            real = real && host[i].y == 0.0f;
        }
        void* dev = nullptr;
        if (real) {
            std::vector<float> re(n);
            for (uint64_t i = 0; i < n; ++i) {
                re[i] = host[i].x;
            }
            if (b200_malloc(ctx, n * sizeof(float), &dev) != B200_SUCCESS) {
                b200_chain_plan_destroy(pl);
                return B200_ERROR;
            }
            e = cudaMemcpy(dev, re.data(), n * sizeof(float), cudaMemcpyHostToDevice);
            pl->win_re = static_cast<float*>(dev);
        } else {
            if (b200_malloc(ctx, n * sizeof(float2), &dev) != B200_SUCCESS) {
                b200_chain_plan_destroy(pl);
                return B200_ERROR;
            }
            e = cudaMemcpy(dev, host.data(), n * sizeof(float2), cudaMemcpyHostToDevice);
            pl->win_c = static_cast<float2*>(dev);
        }
        if (e != cudaSuccess) {
            b200_chain_plan_destroy(pl);
            return fail("b200_chain_plan_create: window upload failed: %s", cudaGetErrorString(e));
        }
    }
    pl->variant = pl->tiled ? "fft_cols_kernel<window> + fft_rows256_kernel<amplitude,range> (tiled two-pass, L2-resident scratch)"
                  : n == kFft4096N ? "fft4096_kernel<tma,radix16x3>"
                                 : (n >= 16 && n <= 8192 ? "fft_radix_kernel<tma,radix16 stockham>"
                                                         : "fft_generic_kernel<stockham4>");
    cudaStreamSynchronize(cudaStreamLegacy);   // uploads / zero-fills above ran on the legacy stream: settle them before a non-blocking stream executes
    *plan = pl;
    return B200_SUCCESS;
}

const char* b200_chain_plan_variant(const b200_chain_plan* plan) { return plan ? plan->variant : ""; }

static void chain_params(const b200_chain_plan* plan, const float2* x, float* out, uint64_t batch, float amp_coeff,
                         int enable_range, float scale, float offset, FftParams* q) {
    q->in = x;
    q->out = out;
    q->rows = batch;
    q->n = static_cast<uint32_t>(plan->n);
    q->twiddle = plan->twiddle;
    q->win_re = plan->win_re;
    q->win_c = plan->win_c;
    // dB = 20 * (Y * log10(2)) + coeff
    const double kDbPerLog2 = 20.0 * 0.3010299956639812;
    q->amp_scale = static_cast<float>(kDbPerLog2);
    q->amp_coeff = amp_coeff;
    if (enable_range) {
        // 0.5 + 0.5 tanh(z), z = 4 ((dB s + o) - 0.5)  ==  1 / (1 + 2^(-2 z log2(e)))
        const double kLog2e = 1.4426950408889634;
        const double a = -8.0 * kLog2e * static_cast<double>(scale);
        const double b = -8.0 * kLog2e * (static_cast<double>(offset) - 0.5);
        if (scale == 0.0f) {  // RangeImplNativeCpu::kernelF32: scale == 0 -> 0.5 everywhere
            q->k1 = 0.0f;
            q->k0 = 0.0f;
            q->zero_value = 0.5f;
        } else {
            q->k1 = static_cast<float>(a * kDbPerLog2);
            q->k0 = static_cast<float>(a * static_cast<double>(amp_coeff) + b);
            q->zero_value = 0.0f;
        }
    }
}

static int chain_launch(b200_chain_plan* plan, const float2* x, float* out, uint64_t batch, float amp_coeff,
                        int enable_range, float scale, float offset, cudaStream_t s) {
    if (plan->composite) {
        B200_REQUIRE(batch <= plan->max_batch, "b200_chain_exec: batch %llu exceeds the plan's max_batch %llu",
                     static_cast<unsigned long long>(batch), static_cast<unsigned long long>(plan->max_batch));
        const uint64_t shape[2] = {batch, plan->n}, sa[2] = {plan->n, 1}, sb[2] = {0, 1};
        b200_fft_plan* fft = plan->fft;
        const uint64_t saved = fft->batch;
        int rc = b200_multiply_cf32(plan->ctx, reinterpret_cast<const b200_cf32*>(x),
                                    reinterpret_cast<const b200_cf32*>(plan->win_c),
                                    reinterpret_cast<b200_cf32*>(plan->scratch), 2, shape, sa, sb, s);
        if (rc == B200_SUCCESS && batch != saved) {
            rc = fail("b200_chain_exec: the composite path runs whole plans only (batch %llu, planned %llu)",
                      static_cast<unsigned long long>(batch), static_cast<unsigned long long>(saved));
        }
        rc = rc == B200_SUCCESS ? fft_exec_impl(fft, plan->scratch, plan->scratch, 1, s) : rc;
        rc = rc == B200_SUCCESS ? b200_amplitude_cf32(plan->ctx, reinterpret_cast<const b200_cf32*>(plan->scratch), out,
                                                     batch * plan->n, amp_coeff, s) : rc;
        if (rc == B200_SUCCESS && enable_range) {
            rc = b200_range_f32(plan->ctx, out, out, batch * plan->n, scale, offset, s);
        }
        return rc;
    }
    FftParams p{};
    chain_params(plan, x, out, batch, amp_coeff, enable_range, scale, offset, &p);
    if (plan->tiled) {
        b200_fft_plan* fft = plan->fft;
        const uint64_t n = plan->n;
        const bool pdl = twopass_pdl();
        for (uint64_t row0 = 0; row0 < batch; row0 += fft->chunk_rows) {
            TileParams t{};
            t.in = x + row0 * n;
            t.out = fft->scratch_a;
            t.transforms = std::min(fft->chunk_rows, batch - row0);
            t.m1 = static_cast<uint32_t>(fft->n1);
            t.inverse = 0;
            t.table = fft->sub1->twiddle;
            t.stage_tw = fft->step_twiddle;
            t.hints = fft->hints;
            t.win_re = plan->win_re;
            t.win_c = plan->win_c;
            t.epi = p;
            if ((reinterpret_cast<uintptr_t>(t.in) & 15u) != 0) {      // TMA sources are 16-byte aligned: stage the chunk
                if (!fft->scratch_b) {
                    void* b = nullptr;
                    if (b200_malloc(plan->ctx, fft->chunk_rows * n * sizeof(float2), &b) != B200_SUCCESS) {
                        return B200_ERROR;
                    }
                    fft->scratch_b = static_cast<float2*>(b);
                }
                B200_CUDA_CHECK(cudaMemcpyAsync(fft->scratch_b, t.in, t.transforms * n * sizeof(float2),
                                                cudaMemcpyDeviceToDevice, s));
                t.in = fft->scratch_b;
            }
            if (launch_tile_cols(plan->ctx, t, s, pdl && row0 > 0) != B200_SUCCESS) {
                return B200_ERROR;
            }
            t.in = fft->scratch_a;
            t.out = reinterpret_cast<float2*>(out + row0 * n);
            t.table = fft->sub2->twiddle;
            t.win_re = nullptr;
            t.win_c = nullptr;
            if (launch_tile_rows(plan->ctx, t, s, pdl, enable_range ? MODE_AMP_RANGE : MODE_AMP) != B200_SUCCESS) {
                return B200_ERROR;
            }
        }
        return B200_SUCCESS;
    }
    const int win = plan->win_re ? WIN_REAL : (plan->win_c ? WIN_COMPLEX : WIN_NONE);
#define B200_CHAIN_DISPATCH(MODE)                                                   \
    switch (win) {                                                                  \
        case WIN_REAL: return launch_fft<MODE, WIN_REAL>(plan->ctx, p, s);          \
        case WIN_COMPLEX: return launch_fft<MODE, WIN_COMPLEX>(plan->ctx, p, s);    \
        default: return launch_fft<MODE, WIN_NONE>(plan->ctx, p, s);                \
    }
    if (enable_range) {
        B200_CHAIN_DISPATCH(MODE_AMP_RANGE)
    } else {
        B200_CHAIN_DISPATCH(MODE_AMP)
    }
#undef B200_CHAIN_DISPATCH
}

int b200_chain_exec(b200_chain_plan* plan, const b200_cf32* x, float* out, uint64_t batch, float amp_coeff,
                    int enable_range, float scale, float offset, b200_stream stream) {
    B200_REQUIRE(plan, "b200_chain_exec: null plan");
    if (batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(x && out, "b200_chain_exec: null buffer");
    DeviceGuard guard(plan->ctx);
    return chain_launch(plan, reinterpret_cast<const float2*>(x), out, batch, amp_coeff, enable_range, scale, offset,
                        as_stream(stream));
}

int b200_chain_exec_typed(b200_chain_plan* plan, const void* x, int in_dtype, float* out, uint64_t batch,
                          float amp_coeff, int enable_range, float scale, float offset, b200_stream stream) {
    B200_REQUIRE(plan, "b200_chain_exec_typed: null plan");
    if (in_dtype == B200_DTYPE_CF32) {
        return b200_chain_exec(plan, static_cast<const b200_cf32*>(x), out, batch, amp_coeff, enable_range, scale,
                               offset, stream);
    }
    B200_REQUIRE(in_dtype >= B200_DTYPE_CI8 && in_dtype <= B200_DTYPE_CU32,
                 "b200_chain_exec_typed: input dtype code %d is not CF32 or a complex integer type", in_dtype);
    if (batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(x && out, "b200_chain_exec_typed: null buffer");
    DeviceGuard guard(plan->ctx);
    const cudaStream_t s = as_stream(stream);
    const bool fused = !plan->composite && plan->n == kFft4096N && plan->win_re != nullptr &&
                       in_dtype <= B200_DTYPE_CU16 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                       !fft4096_use_generic();
    if (!fused) {
        // cast module, then the CF32 chain (one extra pass over the data)
        if (plan->cast_rows < batch) {
            cudaFree(plan->cast_scratch);
            plan->cast_scratch = nullptr;
            plan->cast_rows = 0;
            void* sc = nullptr;
            if (b200_malloc(plan->ctx, batch * plan->n * sizeof(float2), &sc) != B200_SUCCESS) {
                return B200_ERROR;
            }
            plan->cast_scratch = static_cast<float2*>(sc);
            plan->cast_rows = batch;
        }
        const int rc = b200_cast_int(plan->ctx, x, in_dtype, plan->cast_scratch, batch * plan->n, stream);
        return rc != B200_SUCCESS ? rc
                                  : chain_launch(plan, plan->cast_scratch, out, batch, amp_coeff, enable_range, scale,
                                                 offset, s);
    }
    FftParams p{};
    chain_params(plan, static_cast<const float2*>(x), out, batch, amp_coeff, enable_range, scale, offset, &p);
#define B200_INT_DISPATCH(MODE)                                                                  \
    switch (in_dtype) {                                                                          \
        case B200_DTYPE_CI8: return launch_4096_int<MODE, IN_CI8>(plan->ctx, p, s);              \
        case B200_DTYPE_CU8: return launch_4096_int<MODE, IN_CU8>(plan->ctx, p, s);              \
        case B200_DTYPE_CI16: return launch_4096_int<MODE, IN_CI16>(plan->ctx, p, s);            \
        default: return launch_4096_int<MODE, IN_CU16>(plan->ctx, p, s);                         \
    }
    if (enable_range) {
        B200_INT_DISPATCH(MODE_AMP_RANGE)
    } else {
        B200_INT_DISPATCH(MODE_AMP)
    }
#undef B200_INT_DISPATCH
}

int b200_chain_exec_agc(b200_chain_plan* plan, const void* x, int in_dtype, float* out, uint64_t batch, float amp_coeff,
                        int enable_range, float scale, float offset, double agc_reference, double agc_epsilon,
                        double agc_min_gain, double agc_max_gain, b200_stream stream) {
    B200_REQUIRE(plan, "b200_chain_exec_agc: null plan");
    B200_REQUIRE(std::isfinite(agc_reference) && agc_reference > 0.0, "[MODULE_AGC] Reference must be finite and positive.");
    B200_REQUIRE(std::isfinite(agc_epsilon) && agc_epsilon > 0.0, "[MODULE_AGC] Epsilon must be finite and positive.");
    B200_REQUIRE(std::isfinite(agc_min_gain) && agc_min_gain > 0.0, "[MODULE_AGC] Minimum gain must be finite and positive.");
    B200_REQUIRE(std::isfinite(agc_max_gain) && agc_max_gain >= agc_min_gain,
                 "[MODULE_AGC] Maximum gain must be finite and no less than minimum gain.");
    B200_REQUIRE(!plan->composite && plan->n == kFft4096N && plan->win_re != nullptr,
                 "b200_chain_exec_agc: the fused AGC path needs n = 4096 and a real window (run the modules otherwise)");
    B200_REQUIRE(in_dtype == B200_DTYPE_CF32 || (in_dtype >= B200_DTYPE_CI8 && in_dtype <= B200_DTYPE_CU16),
                 "b200_chain_exec_agc: input dtype code %d is not CF32 / CI8 / CU8 / CI16 / CU16", in_dtype);
    if (batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(x && out, "b200_chain_exec_agc: null buffer");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "b200_chain_exec_agc: input must be 16-byte aligned");
    DeviceGuard guard(plan->ctx);
    const cudaStream_t s = as_stream(stream);
    FftParams p{};
    chain_params(plan, static_cast<const float2*>(x), out, batch, amp_coeff, enable_range, scale, offset, &p);
    p.agc_reference = agc_reference;
    p.agc_epsilon = agc_epsilon;
    p.agc_min = agc_min_gain;
    p.agc_max = agc_max_gain;
#define B200_AGC_DISPATCH(MODE)                                                                  \
    switch (in_dtype) {                                                                          \
        case B200_DTYPE_CF32: return launch_4096_int<MODE, IN_CF32, true>(plan->ctx, p, s);      \
        case B200_DTYPE_CI8: return launch_4096_int<MODE, IN_CI8, true>(plan->ctx, p, s);        \
        case B200_DTYPE_CU8: return launch_4096_int<MODE, IN_CU8, true>(plan->ctx, p, s);        \
        case B200_DTYPE_CI16: return launch_4096_int<MODE, IN_CI16, true>(plan->ctx, p, s);      \
        default: return launch_4096_int<MODE, IN_CU16, true>(plan->ctx, p, s);                   \
    }
    if (enable_range) {
        B200_AGC_DISPATCH(MODE_AMP_RANGE)
    } else {
        B200_AGC_DISPATCH(MODE_AMP)
    }
#undef B200_AGC_DISPATCH
}

// spectrum_engine -> lineplot in one pass: the chain output AND its column sums (sum over the batch of every output
// column = what LineplotImplNativeCpu::computeSubmit accumulates first, lineplot/module_impl_native_cpu.cc:93-98).
// n = 4096 with a real window: the kernel's epilogue keeps the running sums of its rows in registers (COLSUM variant of
// fft4096_kernel, +8 FADD2 per thread-row, no extra pass over the 1 GiB output); everything else: the normal chain
// followed by the row-split column-sum kernel of viz.cu.
int b200_chain_exec_colsum(b200_chain_plan* plan, const void* x, int in_dtype, float* out, uint64_t batch, float amp_coeff,
                           int enable_range, float scale, float offset, float* colsum, b200_stream stream) {
    B200_REQUIRE(plan && colsum, "b200_chain_exec_colsum: null argument");
    B200_REQUIRE(in_dtype == B200_DTYPE_CF32 || (in_dtype >= B200_DTYPE_CI8 && in_dtype <= B200_DTYPE_CU32),
                 "b200_chain_exec_colsum: input dtype code %d is not CF32 or a complex integer type", in_dtype);
    DeviceGuard guard(plan->ctx);
    const cudaStream_t s = as_stream(stream);
    const uint64_t n = plan->n;
    if (batch == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(colsum, 0, n * sizeof(float), s));
        return B200_SUCCESS;
    }
    B200_REQUIRE(x && out, "b200_chain_exec_colsum: null buffer");
    const bool fused = !plan->composite && n == kFft4096N && plan->win_re != nullptr &&
                       (in_dtype == B200_DTYPE_CF32 || in_dtype <= B200_DTYPE_CU16) &&
                       (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && !fft4096_use_generic();
    const uint64_t fused_bytes = static_cast<uint64_t>(plan->ctx->sms) * 2 * kFft4096N * sizeof(float);
    const uint64_t need = fused ? fused_bytes : colsum_max_splits(batch) * n * sizeof(float) + 16;
    if (plan->colsum_bytes < need) {
        cudaFree(plan->colsum_partial);
        plan->colsum_partial = nullptr;
        plan->colsum_bytes = 0;
        B200_CUDA_CHECK(cudaMalloc(&plan->colsum_partial, need));
        plan->colsum_bytes = need;
    }
    if (!fused) {
        int rc = b200_chain_exec_typed(plan, x, in_dtype, out, batch, amp_coeff, enable_range, scale, offset, stream);
        uint64_t splits = 0;
        rc = rc == B200_SUCCESS ? colsum_partials(plan->ctx, out, batch, n, n, 1, plan->colsum_partial, &splits, s) : rc;
        return rc == B200_SUCCESS ? colsum_reduce(plan->colsum_partial, splits, n, colsum, s) : rc;
    }
    FftParams p{};
    chain_params(plan, static_cast<const float2*>(x), out, batch, amp_coeff, enable_range, scale, offset, &p);
    p.colsum_partial = plan->colsum_partial;
    unsigned grid = 0;
    int rc = B200_SUCCESS;
#define B200_COLSUM_DISPATCH(MODE)                                                                                 \
    switch (in_dtype) {                                                                                            \
        case B200_DTYPE_CF32: rc = launch_4096_int<MODE, IN_CF32, false, true>(plan->ctx, p, s, &grid); break;     \
        case B200_DTYPE_CI8: rc = launch_4096_int<MODE, IN_CI8, false, true>(plan->ctx, p, s, &grid); break;       \
        case B200_DTYPE_CU8: rc = launch_4096_int<MODE, IN_CU8, false, true>(plan->ctx, p, s, &grid); break;       \
        case B200_DTYPE_CI16: rc = launch_4096_int<MODE, IN_CI16, false, true>(plan->ctx, p, s, &grid); break;     \
        default: rc = launch_4096_int<MODE, IN_CU16, false, true>(plan->ctx, p, s, &grid); break;                  \
    }
    if (enable_range) {
        B200_COLSUM_DISPATCH(MODE_AMP_RANGE)
    } else {
        B200_COLSUM_DISPATCH(MODE_AMP)
    }
#undef B200_COLSUM_DISPATCH
    return rc == B200_SUCCESS ? colsum_reduce(plan->colsum_partial, grid, n, colsum, s) : rc;
}

static uint64_t dtype_sample_bytes(const int in_dtype) {
    switch (in_dtype) {
        case B200_DTYPE_CI8: case B200_DTYPE_CU8: return 2;
        case B200_DTYPE_CI16: case B200_DTYPE_CU16: return 4;
        default: return 8;      // CF32, CI32, CU32
    }
}

int b200_chain_exec_host(b200_chain_plan* plan, const b200_cf32* x_host, float* out_host, uint64_t batch,
                         float amp_coeff, int enable_range, float scale, float offset, uint64_t chunk_rows) {
    return b200_chain_exec_host_typed(plan, x_host, B200_DTYPE_CF32, out_host, batch, amp_coeff, enable_range, scale,
                                      offset, chunk_rows);
}

int b200_chain_exec_host_typed(b200_chain_plan* plan, const void* x_host, int in_dtype, float* out_host, uint64_t batch,
                               float amp_coeff, int enable_range, float scale, float offset, uint64_t chunk_rows) {
    B200_REQUIRE(plan, "b200_chain_exec_host: null plan");
    B200_REQUIRE(in_dtype == B200_DTYPE_CF32 || (in_dtype >= B200_DTYPE_CI8 && in_dtype <= B200_DTYPE_CU32),
                 "b200_chain_exec_host: input dtype code %d is not CF32 or a complex integer type", in_dtype);
    if (batch == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(x_host && out_host, "b200_chain_exec_host: null buffer");
    DeviceGuard guard(plan->ctx);
    const uint64_t sample_bytes = dtype_sample_bytes(in_dtype);
    if (chunk_rows == 0) {
        chunk_rows = (128ull << 20) / (plan->n * sizeof(float2));  // 128 MiB of CF32 input per chunk
        if (chunk_rows == 0) {
            chunk_rows = 1;
        }
    }
    if (chunk_rows > batch) {
        chunk_rows = batch;
    }
    constexpr int kSlots = b200_chain_plan::kHostSlots;
    if (plan->host_chunk_rows < chunk_rows) {
        for (int i = 0; i < kSlots; ++i) {
            cudaFree(plan->stage_in[i]);
            cudaFree(plan->stage_out[i]);
            plan->stage_in[i] = nullptr;
            plan->stage_out[i] = nullptr;
            B200_CUDA_CHECK(cudaMalloc(&plan->stage_in[i], chunk_rows * plan->n * sizeof(float2)));
            B200_CUDA_CHECK(cudaMalloc(&plan->stage_out[i], chunk_rows * plan->n * sizeof(float)));
        }
        plan->host_chunk_rows = chunk_rows;
    }
    if (!plan->s_exec) {
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->s_h2d, cudaStreamNonBlocking));
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->s_exec, cudaStreamNonBlocking));
        B200_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; ++i) {
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&plan->ev_in[i], cudaEventDisableTiming));
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&plan->ev_exec[i], cudaEventDisableTiming));
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&plan->ev_out[i], cudaEventDisableTiming));
        }
    }
    const uint64_t chunks = (batch + chunk_rows - 1) / chunk_rows;
    const unsigned char* src = static_cast<const unsigned char*>(x_host);
    const uint64_t row_bytes = plan->n * sample_bytes;
    for (uint64_t c = 0; c < chunks; ++c) {
        const int slot = static_cast<int>(c % kSlots);
        const uint64_t row0 = c * chunk_rows;
        const uint64_t rows = batch - row0 < chunk_rows ? batch - row0 : chunk_rows;
        // H2D: the slot's input buffer is free once the kernel of chunk c - kSlots has run.
        if (c >= kSlots) {
            B200_CUDA_CHECK(cudaStreamWaitEvent(plan->s_h2d, plan->ev_exec[slot], 0));
        }
        B200_CUDA_CHECK(cudaMemcpyAsync(plan->stage_in[slot], src + row0 * row_bytes, rows * row_bytes,
                                        cudaMemcpyHostToDevice, plan->s_h2d));
        B200_CUDA_CHECK(cudaEventRecord(plan->ev_in[slot], plan->s_h2d));
        // Kernel: needs the input, and the slot's output buffer drained by the D2H of chunk c - kSlots.
        B200_CUDA_CHECK(cudaStreamWaitEvent(plan->s_exec, plan->ev_in[slot], 0));
        if (c >= kSlots) {
            B200_CUDA_CHECK(cudaStreamWaitEvent(plan->s_exec, plan->ev_out[slot], 0));
        }
        const int rc = in_dtype == B200_DTYPE_CF32
                           ? chain_launch(plan, plan->stage_in[slot], plan->stage_out[slot], rows, amp_coeff,
                                          enable_range, scale, offset, plan->s_exec)
                           : b200_chain_exec_typed(plan, plan->stage_in[slot], in_dtype, plan->stage_out[slot], rows,
                                                   amp_coeff, enable_range, scale, offset, plan->s_exec);
        if (rc != B200_SUCCESS) {
            return B200_ERROR;
        }
        B200_CUDA_CHECK(cudaEventRecord(plan->ev_exec[slot], plan->s_exec));
        // D2H
        B200_CUDA_CHECK(cudaStreamWaitEvent(plan->s_d2h, plan->ev_exec[slot], 0));
        B200_CUDA_CHECK(cudaMemcpyAsync(out_host + row0 * plan->n, plan->stage_out[slot], rows * plan->n * sizeof(float),
                                        cudaMemcpyDeviceToHost, plan->s_d2h));
        B200_CUDA_CHECK(cudaEventRecord(plan->ev_out[slot], plan->s_d2h));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(plan->s_d2h));
    return B200_SUCCESS;
}

int b200_chain_plan_destroy(b200_chain_plan* plan) {
    if (!plan) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(plan->ctx);
    cudaFree(plan->twiddle);
    cudaFree(plan->win_re);
    cudaFree(plan->win_c);
    cudaFree(plan->scratch);
    cudaFree(plan->cast_scratch);
    cudaFree(plan->colsum_partial);
    b200_fft_plan_destroy(plan->fft);
    for (int i = 0; i < b200_chain_plan::kHostSlots; ++i) {
        cudaFree(plan->stage_in[i]);
        cudaFree(plan->stage_out[i]);
        if (plan->s_exec) {
            cudaEventDestroy(plan->ev_in[i]);
            cudaEventDestroy(plan->ev_exec[i]);
            cudaEventDestroy(plan->ev_out[i]);
        }
    }
    if (plan->s_exec) {
        cudaStreamDestroy(plan->s_h2d);
        cudaStreamDestroy(plan->s_exec);
        cudaStreamDestroy(plan->s_d2h);
    }
    delete plan;
    return B200_SUCCESS;
}

}  // extern "C"
