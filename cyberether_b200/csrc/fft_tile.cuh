// Two-pass plan, tiled form: n = M1 x 256 with M1 in {64, 128, 256} (n = 16384, 32768, 65536). Both passes are
// shared-memory FFT kernels whose global accesses are runs of >= 128 contiguous bytes, so neither side of the
// transposition that a two-factor decomposition implies is paid in scattered 8-byte stores (the radix-16 column pass +
// stride-16 row store of fft_twopass.cuh measured 26 % of the HBM roofline: the scattered stores cost one LSU sector
// cycle and one L2 fill read per element).
//
//   x[n1 * 256 + n2], n1 < M1, n2 < 256;  k = k1 + M1 k2
//   pass 1  fft_cols_kernel<log2 M1> : z[k1][n2] = W_n^(n2 k1) * sum_n1 x[n1 * 256 + n2] W_M1^(n1 k1)
//       a CTA transforms a tile of C = 4096 / M1 adjacent columns: every global load / store instruction of a warp
//       covers C * 8 >= 128 contiguous bytes (lanes run over the columns in BOTH register passes); radix 16 then radix
//       M1 / 16 through one shared-memory exchange (column pitch M1 + M1/16 + 1 elements: odd, so the 16 lanes of a
//       half-warp — 16 different columns — hit 16 different bank pairs); the stage twiddle comes from a table laid out
//       like the scratch ([k1][n2]), i.e. it is read with the store's own coalesced pattern.
//   pass 2  fft_rows256_kernel       : X[k1 + M1 k2] = sum_n2 z[k1][n2] W_256^(n2 k2)
//       16 consecutive scratch rows (k1 .. k1 + 15) land by one bulk TMA copy (2-deep ring); pass A has lanes over n2
//       (conflict-free reads of the linear landing), pass B has lanes over the ROW so that the 16 lanes of a half-warp
//       store X[k1 .. k1 + 15 + M1 k2] = 128 contiguous bytes.
// The scratch is one chunk of the batch that stays in the 126 MB L2 between the two launches (fft.cu).
#pragma once

#include <cuda.h>

#include "fft_radix.cuh"
#include "fft_twopass.cuh"

namespace b200 {

// 2-D tensor-map TMA tile load (SASS UTMALDG), completion on the mbarrier.
__device__ __forceinline__ void tma_load_tile_2d(void* smem_dst, const CUtensorMap* map, const int c0, const int c1,
                                                 uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

constexpr int kTileThreads = 256;
constexpr int kTileElems = 4096;                         // complex samples per CTA iteration
constexpr int kTileRowLen = 256;                         // the contiguous factor (pass 2 length)
__host__ __device__ constexpr int tile_pitch(const int n) { return n + n / 16 + 1; }

// Programmatic dependent launch (the chunk loop launches cols, rows, cols, rows ... on one stream): a kernel lets its
// successor be scheduled right away and the successor blocks at pdl_wait() only where it touches what the predecessor
// produces or still reads, so launch latency and the tail of one kernel overlap the head of the next. Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct TileParams {
    const float2* in;
    float2* out;
    uint64_t transforms;      // transforms in this launch (chunk)
    uint32_t m1;              // n = m1 * 256
    int inverse;
    const float2* table;      // pass 1: W_M1^j (j < M1); pass 2: W_256^j
    const float2* stage_tw;   // pass 1: [M1][256] W_n^(k1 n2)
    int hints;                // pass 1: L2 eviction-priority hints (input evict_first, scratch evict_last)
    // fused spectral chain (spectrum_engine at n = 16384 / 32768 / 65536): window multiply in the column pass's load,
    // amplitude / range epilogue (epi.amp_scale ... epi.zero_value, as fft4096_kernel) in the row pass's store
    const float* win_re;      // [n] real window (WIN_REAL)
    const float2* win_c;      // [n] complex window (WIN_COMPLEX)
    FftParams epi;
};

// ---- pass 1: column transforms of length M1 over tiles of C adjacent columns ----------------------------------------
// A tile ([M1 lines][C columns]) lands by ONE 2-D tensor-map TMA copy (box = 2 C floats x M1 lines, no swizzle: the
// landing is [n1][C], read with lanes over the columns) into a 2-deep ring, so the strided HBM reads of the next tile run
// under this tile's math. (M1 separate 1-D bulk copies of C * 8 = 128 bytes measured 1.6x slower than plain loads.)
constexpr int kColsStages = 2;
constexpr int kColsStageBytes = kTileElems * 8;                                   // 32 KiB
__host__ __device__ constexpr int cols_x1_bytes(const int n) { return (kTileElems / n) * tile_pitch(n) * 8; }
__host__ __device__ constexpr int cols_smem_bytes(const int n) {
    return kColsStages * kColsStageBytes + ((cols_x1_bytes(n) + 127) / 128) * 128 + 64;
}

template <int LOG2M1, int WIN = WIN_NONE>
__global__ void __launch_bounds__(kTileThreads, 2)
    fft_cols_kernel(const TileParams p, const __grid_constant__ CUtensorMap in_map) {
    constexpr int N = 1 << LOG2M1;               // 64, 128, 256
    constexpr int C = kTileElems / N;            // columns per tile
    constexpr int T = N / 16;                    // pass-A butterflies per column (= 256 / C)
    constexpr int R = N / 16;                    // pass-B radix (4, 8, 16); N / R = 16 butterflies per column
    constexpr int CB = 16 / R;                   // pass-B butterflies per thread
    constexpr int P = tile_pitch(N);
    static_assert(T * C == kTileThreads, "one pass-A butterfly per thread");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* const x1 = reinterpret_cast<float2*>(smem_raw + kColsStages * kColsStageBytes);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + cols_smem_bytes(N) - 64);

    const uint32_t tid = threadIdx.x;
    const uint32_t g = tid % C;                  // column inside the tile: lanes run over columns
    const uint32_t ja = tid / C;                 // pass-A butterfly, also the pass-B butterfly base (T == R)
    const uint64_t n = static_cast<uint64_t>(N) * kTileRowLen;
    constexpr uint32_t kTilesPerTransform = kTileRowLen / C;
    const uint64_t tiles = p.transforms * kTilesPerTransform;
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t my_tiles = first < tiles ? static_cast<uint32_t>((tiles - first + stride - 1) / stride) : 0u;
    pdl_launch_dependents();

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kColsStages; ++s) {
            mbar_init(&full[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t pol_in = 0, pol_out = 0;
    if (p.hints) {
        pol_in = l2_policy_evict_first();
        pol_out = l2_policy_evict_last();
    }
    // thread 0: land tile `index` (of this CTA's sequence) in stage `s`: the input is [transforms * M1 lines][512 floats]
    auto issue_tile = [&](const uint32_t index, const uint32_t s) {
        const uint64_t tile = first + static_cast<uint64_t>(index) * stride;
        const uint64_t r = tile / kTilesPerTransform;
        const uint32_t c0 = static_cast<uint32_t>(tile % kTilesPerTransform) * C;
        mbar_expect_tx(&full[s], kColsStageBytes);
        tma_load_tile_2d(smem_raw + s * kColsStageBytes, &in_map, static_cast<int>(2 * c0), static_cast<int>(r * N), &full[s]);
    };
    uint32_t issued = 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kColsStages; ++s) {
            if (issued < my_tiles) {
                issue_tile(issued, s);
                ++issued;
            }
        }
    }

    TwiddleSet tw[CB];
#pragma unroll
    for (int b = 0; b < CB; ++b) {
        tw[b] = load_twiddles_n(p.table, ja + R * b, N);              // powers of W_N^(jj), jj = ja + R b
    }

    uint32_t stage = 0, parity = 0;
    for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint64_t tile = first + static_cast<uint64_t>(i) * stride;
        const uint64_t r = tile / kTilesPerTransform;
        const uint32_t c0 = static_cast<uint32_t>(tile % kTilesPerTransform) * C;
        const float2* const land = reinterpret_cast<const float2*>(smem_raw + stage * kColsStageBytes);
        // window taps of this thread's 16 samples (fused chain): issued before the wait on the landing
        float wr[WIN == WIN_REAL ? 16 : 1];
        float2 wc[WIN == WIN_COMPLEX ? 16 : 1];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            if constexpr (WIN == WIN_REAL) {
                wr[t] = p.win_re[(ja + t * T) * kTileRowLen + c0 + g];
            } else if constexpr (WIN == WIN_COMPLEX) {
                wc[t] = p.win_c[(ja + t * T) * kTileRowLen + c0 + g];
            }
        }
        mbar_wait(&full[stage], parity);
        float2 v[16];
        // ---- pass A: radix 16 over n1 = ja + t T -------------------------------------------------------
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            float2 x = land[(ja + t * T) * C + g];
            if constexpr (WIN == WIN_REAL) {
                x = apply_window<WIN>(x, wr[t], make_float2(0.f, 0.f));
            } else if constexpr (WIN == WIN_COMPLEX) {
                x = apply_window<WIN>(x, 0.f, wc[t]);
            } else {
                if (p.inverse) {
                    x = make_float2(x.y, x.x);
                }
            }
            v[t] = x;
        }
        dft16(*reinterpret_cast<float2(*)[16]>(v));
        __syncthreads();                         // landing consumed; the previous tile's pass-B reads of x1 are complete
        if (tid == 0 && issued < my_tiles) {     // refill this stage with the tile two ahead
            fence_proxy_async();
            issue_tile(issued, stage);
            ++issued;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            x1[g * P + pad16(16 * ja + t)] = v[dft_pos<16>(t)];
        }
        // the stage twiddles of this thread's 16 outputs: issued before the barrier, consumed after pass B
        const float2* const stw = p.stage_tw + c0 + g;
        float2 sw[16];
#pragma unroll
        for (int b = 0; b < CB; ++b) {
#pragma unroll
            for (int t = 0; t < R; ++t) {
                sw[b * R + t] = stw[(ja + R * b + 16 * t) * kTileRowLen];
            }
        }
        __syncthreads();
        // ---- pass B: radix R, Ns = 16: butterfly jj reads jj + 16 t, writes k1 = jj + 16 t ---------------
        if (i == 0) {
            pdl_wait();                          // the previous chunk's row pass has finished reading the scratch
        }
        float2* const dst = p.out + r * n + c0 + g;
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const uint32_t jj = ja + R * b;
#pragma unroll
            for (int t = 0; t < R; ++t) {
                v[b * R + t] = x1[g * P + pad16(jj + 16 * t)];
            }
            twiddle_inputs<R>(v + b * R, tw[b]);
            dft_r<R>(v + b * R);
        }
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const uint32_t jj = ja + R * b;
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const uint32_t off = (jj + 16 * t) * kTileRowLen;
                const float2 y = cmul(v[b * R + dft_pos<R>(t)], sw[b * R + t]);
                if (p.hints) {
                    stg_hint_f2(dst + off, y, pol_out);
                } else {
                    dst[off] = y;
                }
            }
        }
        if (++stage == kColsStages) {
            stage = 0;
            parity ^= 1;
        }
    }
}

// ---- pass 2: 256-point row transforms, 16 consecutive rows per CTA iteration, stores transposed in 128-byte runs ------
constexpr int kRows256Stages = 2;
constexpr int kRows256StageBytes = kTileElems * 8;                                // 32 KiB: 16 rows of 256
constexpr int kRows256X1Bytes = 16 * tile_pitch(kTileRowLen) * 8;                 // 34944
constexpr int kRows256SmemBytes = kRows256Stages * kRows256StageBytes + kRows256X1Bytes + 64;

template <int MODE = MODE_C2C>
__global__ void __launch_bounds__(kTileThreads, 2) fft_rows256_kernel(const TileParams p) {
    constexpr int N = kTileRowLen;
    constexpr int P = tile_pitch(N);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* const x1 = reinterpret_cast<float2*>(smem_raw + kRows256Stages * kRows256StageBytes);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kRows256Stages * kRows256StageBytes + kRows256X1Bytes);

    const uint32_t tid = threadIdx.x;
    const uint32_t ga = tid >> 4, lt = tid & 15;             // pass A: row ga, lanes over n2
    const uint32_t gb = tid & 15, jj = tid >> 4;             // pass B: lanes over the row
    const uint64_t n = static_cast<uint64_t>(p.m1) * N;
    const uint64_t blocks_total = p.transforms * (p.m1 / 16);
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t my_blocks =
        first < blocks_total ? static_cast<uint32_t>((blocks_total - first + stride - 1) / stride) : 0u;

    pdl_launch_dependents();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRows256Stages; ++s) {
            mbar_init(&full[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();                                              // the column pass has written the scratch
    // Blocks are taken from the END of the scratch: the column pass wrote it front to back, so its most recently written
    // (most likely still L2-resident) lines are read first.
    uint32_t issued = 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRows256Stages; ++s) {
            if (issued < my_blocks) {
                const uint64_t b = blocks_total - 1 - (first + issued * stride);
                mbar_expect_tx(&full[s], kRows256StageBytes);
                tma_load_row(smem_raw + s * kRows256StageBytes, p.in + b * kTileElems, kRows256StageBytes, &full[s]);
                ++issued;
            }
        }
    }
    const TwiddleSet tw = load_twiddles_n(p.table, jj, N);   // powers of W_256^jj

    uint32_t stage = 0, parity = 0;
    for (uint32_t i = 0; i < my_blocks; ++i) {
        const uint64_t blk = blocks_total - 1 - (first + static_cast<uint64_t>(i) * stride);
        const uint64_t row0 = blk * 16;                      // scratch row = transform * M1 + k1
        const uint64_t r = row0 / p.m1;
        const uint32_t k1_0 = static_cast<uint32_t>(row0 % p.m1);
        const float2* const sbuf = reinterpret_cast<const float2*>(smem_raw + stage * kRows256StageBytes);
        mbar_wait(&full[stage], parity);

        float2 v[16];
        // ---- pass A: radix 16 over n2 = lt + 16 t (the column pass already swapped re/im for the inverse) ----
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            v[t] = sbuf[ga * N + lt + 16 * t];
        }
        dft16(*reinterpret_cast<float2(*)[16]>(v));
        __syncthreads();                                     // landing consumed; previous block's x1 reads complete
        if (tid == 0 && issued < my_blocks) {                // refill this stage with the block two ahead
            const uint64_t b = blocks_total - 1 - (first + static_cast<uint64_t>(issued) * stride);
            fence_proxy_async();
            mbar_expect_tx(&full[stage], kRows256StageBytes);
            tma_load_row(smem_raw + stage * kRows256StageBytes, p.in + b * kTileElems, kRows256StageBytes, &full[stage]);
            ++issued;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            x1[ga * P + pad16(16 * lt + t)] = v[dft_pos<16>(t)];
        }
        __syncthreads();
        // ---- pass B: butterfly jj of row gb: reads jj + 16 t, X[k2 = jj + 16 t] -> out[k1 + M1 k2] ------------
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            v[t] = x1[gb * P + pad16(jj + 16 * t)];
        }
        twiddle_inputs<16>(v, tw);
        dft16(*reinterpret_cast<float2(*)[16]>(v));
        if constexpr (MODE == MODE_C2C) {
            float2* const out = p.out + r * n + k1_0 + gb;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                float2 X = v[dft_pos<16>(t)];
                if (p.inverse) {
                    X = make_float2(X.y, X.x);
                }
                stg_stream_f2(out + static_cast<uint64_t>(jj + 16 * t) * p.m1, X);
            }
        } else {
            // fused chain: |X| -> dB (-> range), F32; the 16 lanes of a half-warp store 64 contiguous bytes
            float* const out = reinterpret_cast<float*>(p.out) + r * n + k1_0 + gb;
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
                const float2 res = spectral_epilogue2<MODE>(v[dft_pos<16>(t)], v[dft_pos<16>(t + 1)], p.epi);
                stg_stream_f1(out + static_cast<uint64_t>(jj + 16 * t) * p.m1, res.x);
                stg_stream_f1(out + static_cast<uint64_t>(jj + 16 * (t + 1)) * p.m1, res.y);
            }
        }
        if (++stage == kRows256Stages) {
            stage = 0;
            parity ^= 1;
        }
    }
}

}  // namespace b200
