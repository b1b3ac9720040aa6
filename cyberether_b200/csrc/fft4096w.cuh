// fft4096w_kernel — the fused 4096-point spectral chain with a WARP-LOCAL first exchange (CF32 input).
//
// fft4096_kernel (fft4096.cuh) synchronises its eight warps three times per row: __syncthreads (A) after the first
// exchange's stores, the split mbarrier (B) before the second exchange's stores, __syncthreads (C) after them; about a
// tenth of the warp time sits in those barriers and, because all eight warps move through the butterfly phases and the
// shared-memory phases together, the FMA pipe and the shared-memory pipe idle alternately. Here the
// first exchange never leaves a warp:
//
//   n = 256 a + 16 b + c,  k = k0 + 16 k1 + 256 k2                                   (same algebra as fft4096.cuh)
//   pass 1  lane = (b, c in {2w, 2w+1}) : sum over a -> k0, * W4096^((16b + c) k0)
//   exchange 1 : b <-> k0 for fixed c — 16 lanes of ONE warp, through a private [k0][b] tile per c (pitch 144 B:
//                STS.64 / LDS.128 at the wavefront floor), ordered by __syncwarp only
//   pass 2  lane = (k0, c)              : sum over b -> k1, * W256^(c k1)
//   exchange 2 / pass 3 / epilogue      : as fft4096_kernel ([k1][k0][c], pitch 18 / 288; thread k0 + 16 k1)
//
// Lanes that run over b read the row with a stride of 128 bytes, which a linearly landed row serves 16-way conflicted;
// so the row is landed through a 2-D tensor map ([256 lines][128 B] box, CU_TENSOR_MAP_SWIZZLE_128B — SASS UTMALDG
// instead of UBLKCP): the 16-byte chunk j of line R sits at chunk j ^ (R & 7), and the pass-1 read of 32 lanes costs the
// two wavefronts its 256 bytes need (tests/test_index_algebra.py::test_fft4096w_warp_local_exchange_scheme emulates the
// whole scheme, including every shared-memory access pattern, in numpy).
//
// Barriers per row: ONE __syncthreads (C) and two split mbarriers whose arrive and wait are far apart — p1done (a warp
// arrives once its pass-1 reads are consumed, waits just before it overwrites the row's buffer with the second exchange)
// and p3done (arrive after the pass-3 reads; thread 0 checks it early in the NEXT row and refills the stage by TMA).
// Shared memory: 2 landing / exchange-2 stages of 36 KiB + 16 exchange-1 tiles (36 KiB) = fft4096_kernel's footprint,
// 2 CTAs per SM.
#pragma once

#include <cuda.h>

#include "fft4096.cuh"

namespace b200 {

constexpr int kFft4096wStages = 2;
constexpr int kFft4096wTilePitch = 144;                                    // bytes per k0 line of an exchange-1 tile
constexpr int kFft4096wTileBytes = 16 * kFft4096wTilePitch + 64;           // 2368 = 64 (mod 128): a warp's two tiles (c, c + 1) differ in bank phase
constexpr int kFft4096wX1Bytes = 16 * kFft4096wTileBytes;                  // 37888
constexpr int kFft4096wSmemBytes =
    kFft4096wStages * kFft4096StageBytes + kFft4096wX1Bytes + 64 + 1024;   // + slack to align the base to 1024 B

__device__ __forceinline__ void tma_load_tile_2d(void* smem_dst, const CUtensorMap* map, const int c0, const int c1,
                                                 uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int MODE, int WIN, bool AGC = false, bool COLSUM = false>
__global__ void __launch_bounds__(kFft4096Threads, 2)
    fft4096w_kernel(const FftParams p, const __grid_constant__ CUtensorMap row_map) {
    static_assert(!COLSUM || MODE != MODE_C2C, "column sums exist for the amplitude outputs only");
    __shared__ float agc_partial[2][kFft4096Threads / 32];
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B repeats every 1024 bytes of shared-memory ADDRESS: align the stages to it
    unsigned char* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* const x1 = smem_raw + kFft4096wStages * kFft4096StageBytes;
    uint64_t* const full = reinterpret_cast<uint64_t*>(x1 + kFft4096wX1Bytes);
    uint64_t* const p1done = full + kFft4096wStages;
    uint64_t* const p3done = p1done + 1;

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    // lane = 16 q3 + 8 cbit + q: the 16-valued digit (b in pass 1, k0 in pass 2) is 8 q3 + q, c = 2 warp + cbit. A half-warp
    // then holds 8 digits x both c: its 64-bit accesses (pass-1 read through the swizzle, exchange-1 / exchange-2 stores)
    // cover all 32 banks once; a quarter-warp holds 8 digits of one c (LDS.128 of the tile at pitch 144 B).
    const uint32_t l16 = ((lane >> 4) << 3) | (lane & 7);
    const uint32_t c = 2 * warp + ((lane >> 3) & 1);
    const uint64_t first = blockIdx.x;
    const uint64_t stride = gridDim.x;
    const uint32_t my_rows =
        first < p.rows ? static_cast<uint32_t>((p.rows - first + stride - 1) / stride) : 0u;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kFft4096wStages; ++s) {
            mbar_init(&full[s], 1);
        }
        mbar_init(p1done, kFft4096Threads / 32);
        mbar_init(p3done, kFft4096Threads / 32);
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t next_row = first;      // thread 0: the row the next TMA load fetches (tensor line = 256 * row)
    uint32_t issued = 0;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kFft4096wStages; ++s) {
            if (issued < my_rows) {
                mbar_expect_tx(&full[s], kFft4096RowBytes);
                tma_load_tile_2d(smem_raw + s * kFft4096StageBytes, &row_map, 0, static_cast<int>(next_row * 256),
                                 &full[s]);
                next_row += stride;
                ++issued;
            }
        }
    }

    // Thread-constant operands (persistent across rows).
    const uint32_t tt = 16 * l16 + c;                                   // pass 1: n = 256 a + tt
    const TwiddleSet tw1 = load_twiddles(p.twiddle, tt);                // W4096^(tt k0)
    const TwiddleSet tw2 = load_twiddles(p.twiddle, 16 * c);            // W256^(c k1) = W4096^(16 c k1)
    float wr[16];
    if constexpr (WIN == WIN_REAL) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            wr[a] = p.win_re[tt + 256 * a];
        }
    }

    // Byte offsets. Landing: element (a, b, c) = line 16 a + b, chunk (c / 2) ^ (b & 7), half c & 1.
    const uint32_t off_p1 = l16 * 128 + ((((c >> 1) ^ (l16 & 7))) << 4) + ((c & 1) << 3);
    unsigned char* const tile = x1 + c * kFft4096wTileBytes;
    const uint32_t off_x1s = l16 * 8;                                   // + 144 k0   (b = l16)
    const uint32_t off_x1l = l16 * kFft4096wTilePitch;                  // + 16 j     (k0 = l16)
    const uint32_t off_x2s = (kFft4096X2Row * l16 + c) * 8;             // + 2304 k1  (k0 = l16)
    const uint32_t off_x2l = (kFft4096X2Row * (t & 15) + kFft4096X2Plane * (t >> 4)) * 8;

    uint32_t stage = 0, parity = 0;
    unsigned char* out_ptr = static_cast<unsigned char*>(p.out) +
                             (first * kFft4096N + t) * (MODE == MODE_C2C ? 8 : 4);
    const uint64_t out_step = stride * kFft4096N * (MODE == MODE_C2C ? 8 : 4);
    float2 colsum[COLSUM ? 8 : 1];
#pragma unroll
    for (int k = 0; k < (COLSUM ? 8 : 1); ++k) {
        colsum[k] = make_float2(0.f, 0.f);
    }

    for (uint32_t i = 0; i < my_rows; ++i) {
        unsigned char* const buf = smem_raw + stage * kFft4096StageBytes;
        const uint32_t row_parity = i & 1u;
        mbar_wait(&full[stage], parity);

        // ---- pass 1: lane = (b, c), sum over a ------------------------------------------------
        float2 v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            float2 x = *reinterpret_cast<const float2*>(buf + off_p1 + 2048 * a);
            if constexpr (MODE == MODE_C2C) {
                if (p.inverse) {
                    x = make_float2(x.y, x.x);
                }
            }
            if constexpr (WIN == WIN_REAL) {
                x = apply_window<WIN>(x, wr[a], make_float2(0.f, 0.f));
            } else if constexpr (WIN == WIN_COMPLEX) {
                x = apply_window<WIN>(x, 0.f, p.win_c[tt + 256 * a]);
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
            if (lane == 0) {
                agc_partial[row_parity][warp] = power;
            }
        }
        dft16_first_layer(v);   // consumes every loaded value: this warp's reads of the landed row have completed
        // The previous row's stage is free once every warp has finished its pass-3 reads of it (p3done, normally long
        // complete by now): land the row after this one in it. Done here rather than at the end of the previous row so
        // that thread 0 never waits for the slowest warp.
        if (t == 0 && i >= 1 && issued < my_rows) {
            unsigned char* const other = smem_raw + (stage ^ 1u) * kFft4096StageBytes;
            mbar_wait(p3done, row_parity ^ 1u);
            fence_proxy_async();
            mbar_expect_tx(&full[stage ^ 1u], kFft4096RowBytes);
            tma_load_tile_2d(other, &row_map, 0, static_cast<int>(next_row * 256), &full[stage ^ 1u]);
            next_row += stride;
            ++issued;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(p1done);
        }
        dft16_rest(v);
        apply_twiddles(v, tw1);

        // ---- exchange 1: inside the half-warp -------------------------------------------------
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            *reinterpret_cast<float2*>(tile + off_x1s + kFft4096wTilePitch * k) = v[dft16_pos(k)];
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 16; b += 2) {
            const float4 pair = *reinterpret_cast<const float4*>(tile + off_x1l + 8 * b);
            v[b] = make_float2(pair.x, pair.y);
            v[b + 1] = make_float2(pair.z, pair.w);
        }

        // ---- pass 2: lane = (k0, c), sum over b ------------------------------------------------
        dft16(v);
        apply_twiddles(v, tw2);
        // every warp has read its part of the landed row: the row's buffer may take the second exchange
        mbar_wait(p1done, row_parity);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            *reinterpret_cast<float2*>(buf + off_x2s + kFft4096X2Plane * 8 * k) = v[dft16_pos(k)];
        }
        __syncthreads();  // (C) — also orders the half-warp's tile reads before the next row's tile writes

        // ---- pass 3: thread k0 + 16 k1, sum over c --------------------------------------------
#pragma unroll
        for (int cc = 0; cc < 16; cc += 2) {
            const float4 pair = *reinterpret_cast<const float4*>(buf + off_x2l + 8 * cc);
            v[cc] = make_float2(pair.x, pair.y);
            v[cc + 1] = make_float2(pair.z, pair.w);
        }
        dft16_first_layer(v);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(p3done);
        }
        dft16_rest(v);

        // ---- epilogue: X[t + 256 k2] -----------------------------------------------------------
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
                    mean_power += agc_partial[row_parity][w];
                }
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

        if (++stage == kFft4096wStages) {
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
