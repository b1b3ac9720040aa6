// Tiled RMS automatic gain control — the `agc` module
// (src/domains/dsp/agc/module_impl_native_cpu.cc:16-160, config include/jetstream/domains/dsp/agc/module.hh:8-19).
//
// Per lane (every index except the sample axis): the samples are cut into tiles; a tile's target gain is
// clamp(reference / sqrt(mean power + epsilon), minGain, maxGain); the gain applied inside tile t runs linearly from
// the gain reached at its start to the (rate-limited) target of tile t+1; every product is limited so it stays finite.
// All gain arithmetic is F64 in the reference and here (IEEE add / mul / div / sqrt are correctly rounded on both), so
// the only difference is the ORDER of the F64 power sum (parallel tree here, sequential there): gains agree to a few
// F64 ulp and the F32 outputs are identical except for rare last-bit rounding flips.
//
// Three launches: tile gains (one CTA per tile, 8 B/sample read), the sequential per-lane rate-limit chain (one thread
// per lane, tiles steps), apply (8 B/sample read + 8 written for CF32). HBM-bound; algorithmic bytes 16 (CF32) /
// 8 (F32) per sample, moved 24 / 12 because the input is read twice.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace b200 {

struct AgcParams {
    uint64_t lanes, samples, tile, tiles;
    double reference, epsilon, min_gain, max_gain, max_change;
};

__device__ __forceinline__ double clamp_like_std(const double v, const double lo, const double hi) {
    return v < lo ? lo : (hi < v ? hi : v);        // std::clamp: NaN passes through
}
__device__ __forceinline__ double max_like_std(const double a, const double b) { return a < b ? b : a; }

__device__ __forceinline__ double sample_power(const float v) {
    const double x = v;
    return __dmul_rn(x, x);
}
__device__ __forceinline__ double sample_power(const float2 v) {
    const double re = v.x, im = v.y;
    return __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
}

// ---- 1. target gain of every tile --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) agc_tile_gain_kernel(const T* __restrict__ in, double* __restrict__ target,
                                                            const AgcParams p) {
    __shared__ double partial[8];
    const uint64_t total = p.lanes * p.tiles;
    for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const uint64_t lane = item / p.tiles, tile = item - lane * p.tiles;
        const uint64_t start = tile * p.tile;
        const uint64_t length = p.samples - start < p.tile ? p.samples - start : p.tile;
        const T* const src = in + lane * p.samples + start;
        // four independent F64 accumulators per thread and all loads of a round in flight (the single dependent chain
        // of round 1 paid one memory latency per sample); the order of the power sum was a parallel tree already
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        uint64_t s = threadIdx.x;
        for (; s + 3 * blockDim.x < length; s += 4 * blockDim.x) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = src[s + u * blockDim.x];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[u] = __dadd_rn(acc[u], sample_power(v[u]));
            }
        }
        for (; s < length; s += blockDim.x) {
            acc[0] = __dadd_rn(acc[0], sample_power(src[s]));
        }
        double sum = __dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3]));
#pragma unroll
        for (int offset = 16; offset > 0; offset >>= 1) {
            sum = __dadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, offset));
        }
        __syncthreads();                       // partial[] of the previous item has been consumed
        if ((threadIdx.x & 31) == 0) {
            partial[threadIdx.x >> 5] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double total_power = partial[0];
            for (unsigned w = 1; w < blockDim.x / 32; ++w) {
                total_power = __dadd_rn(total_power, partial[w]);
            }
            const double mean = __ddiv_rn(total_power, static_cast<double>(length));
            target[item] = clamp_like_std(__ddiv_rn(p.reference, __dsqrt_rn(__dadd_rn(mean, p.epsilon))), p.min_gain,
                                          p.max_gain);
        }
    }
}

// ---- 2. rate-limited gain at both ends of every tile (LimitGainChange, module_impl_native_cpu.cc:62-74) ----------
__global__ void agc_gain_chain_kernel(const double* __restrict__ target, double2* __restrict__ ends, const AgcParams p) {
    const uint64_t lane = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (lane >= p.lanes) {
        return;
    }
    const double* const t = target + lane * p.tiles;
    double start = t[0];
    for (uint64_t tile = 0; tile < p.tiles; ++tile) {
        double end = start;
        if (tile + 1 < p.tiles) {
            const double lowest = max_like_std(p.min_gain, __ddiv_rn(start, p.max_change));
            const double highest = start > __ddiv_rn(p.max_gain, p.max_change) ? p.max_gain
                                                                               : __dmul_rn(start, p.max_change);
            end = clamp_like_std(t[tile + 1], lowest, highest);
        }
        ends[lane * p.tiles + tile] = make_double2(start, end);
        start = end;
    }
}

// ---- 3. apply -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double limit_gain(const double magnitude, const double gain, const double limit) {
    return magnitude > __ddiv_rn(limit, gain) ? nextafter(__ddiv_rn(limit, magnitude), 0.0) : gain;
}
__device__ __forceinline__ float clamp_to_f32(const double v) {
    const double m = static_cast<double>(FLT_MAX);
    return __double2float_rn(clamp_like_std(v, -m, m));
}
// The reference limits every product so it stays finite: gain' = |x| > limit / gain ? nextafter(limit / |x|, 0) : gain
// (module_impl_native_cpu.cc:38-60). |x| gain <= limit / 2 implies |x| <= limit / gain with a factor-two margin over any
// rounding of that quotient, so the per-sample F64 division (a ~25-instruction sequence on the quarter-rate pipe) is
// only paid by samples within a factor two of overflowing F32 — results are unchanged.
__device__ __forceinline__ float apply_gain(const float v, const double gain) {
    const double x = v;
    const double limit = static_cast<double>(FLT_MAX);
    const double safe = __dmul_rn(fabs(x), gain) <= 0.5 * limit ? gain : limit_gain(fabs(x), gain, limit);
    return clamp_to_f32(__dmul_rn(x, safe));
}
__device__ __forceinline__ float2 apply_gain(const float2 v, const double gain) {
    const double re = v.x, im = v.y;
    const double limit = 3.4028232635611926e+38;       // (F64) nextafter(FLT_MAX, 0): kMaxSafeCF32Magnitude
    double safe = gain;
    // |z| <= |re| + |im|: the exact (and slow) hypot and the division are only needed when that bound, times the gain,
    // does not clear half the limit (see apply_gain(float) above)
    if (!(__dmul_rn(__dadd_rn(fabs(re), fabs(im)), gain) <= 0.5 * limit)) {
        safe = limit_gain(hypot(re, im), gain, limit);
    }
    return make_float2(clamp_to_f32(__dmul_rn(re, safe)), clamp_to_f32(__dmul_rn(im, safe)));
}

template <typename T>
__global__ void __launch_bounds__(256) agc_apply_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                        const double2* __restrict__ ends, const AgcParams p) {
    const uint64_t total = p.lanes * p.tiles;
    for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const uint64_t lane = item / p.tiles, tile = item - lane * p.tiles;
        const uint64_t start = tile * p.tile;
        const uint64_t length = p.samples - start < p.tile ? p.samples - start : p.tile;
        const double2 g = ends[item];
        const double step = __ddiv_rn(__dsub_rn(g.y, g.x), static_cast<double>(length));
        const T* const src = in + lane * p.samples + start;
        T* const dst = out + lane * p.samples + start;
        uint64_t s = threadIdx.x;
        for (; s + 3 * blockDim.x < length; s += 4 * blockDim.x) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = src[s + u * blockDim.x];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double gain = __dadd_rn(g.x, __dmul_rn(step, static_cast<double>(s + u * blockDim.x)));
                dst[s + u * blockDim.x] = apply_gain(v[u], gain);
            }
        }
        for (; s < length; s += blockDim.x) {
            const double gain = __dadd_rn(g.x, __dmul_rn(step, static_cast<double>(s)));
            dst[s] = apply_gain(src[s], gain);
        }
    }
}

// ---- fused form: one CTA per lane, the tile's samples stay in registers between the power sum and the apply ------------
// Every sample is read once and written once (16 B per CF32 sample instead of 24). A CTA walks the tiles of its lane in
// order: tile t is applied with the gain running from `start` to the rate-limited target of tile t + 1, so tile t + 1 is
// loaded and summed (registers `nxt`) before tile t (registers `cur`) is written. Tiles of up to 256 x K samples, lanes >=
// the number of CTAs the GPU holds (otherwise the three-kernel form above has more parallelism).
constexpr int kAgcFusedThreads = 256;
constexpr int kAgcFusedPerThread = 16;

template <typename T>
__device__ __forceinline__ double agc_block_sum(double sum, double* partial) {
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
        sum = __dadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, offset));
    }
    __syncthreads();                           // partial[] of the previous reduction has been consumed
    if ((threadIdx.x & 31) == 0) {
        partial[threadIdx.x >> 5] = sum;
    }
    __syncthreads();
    double total = partial[0];
#pragma unroll
    for (int w = 1; w < kAgcFusedThreads / 32; ++w) {
        total = __dadd_rn(total, partial[w]);
    }
    return total;                              // every thread holds the same value
}

// SINGLE: one tile per lane (the spectrum_engine use, and the benchmarked [16384, 4096] / tile 4096 case): no look-ahead
// tile, half the registers.
template <typename T, int K, bool SINGLE>
__global__ void __launch_bounds__(kAgcFusedThreads) agc_fused_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                     const AgcParams p) {
    __shared__ double partial[kAgcFusedThreads / 32];
    const uint32_t tid = threadIdx.x;
    for (uint64_t lane = blockIdx.x; lane < p.lanes; lane += gridDim.x) {
        const T* const src = in + lane * p.samples;
        T* const dst = out + lane * p.samples;
        auto tile_length = [&](const uint64_t tile) {
            const uint64_t start = tile * p.tile;
            return p.samples - start < p.tile ? p.samples - start : p.tile;
        };
        auto load_tile = [&](const uint64_t tile, T (&v)[K]) {
            const uint64_t length = tile_length(tile);
            const T* const s = src + tile * p.tile;
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < K; ++u) {
                const uint64_t idx = tid + static_cast<uint64_t>(u) * kAgcFusedThreads;
                if (idx < length) {
                    v[u] = s[idx];
                }
            }
#pragma unroll
            for (int u = 0; u < K; ++u) {
                const uint64_t idx = tid + static_cast<uint64_t>(u) * kAgcFusedThreads;
                if (idx < length) {
                    acc = __dadd_rn(acc, sample_power(v[u]));
                }
            }
            const double total_power = agc_block_sum<T>(acc, partial);
            const double mean = __ddiv_rn(total_power, static_cast<double>(length));
            return clamp_like_std(__ddiv_rn(p.reference, __dsqrt_rn(__dadd_rn(mean, p.epsilon))), p.min_gain, p.max_gain);
        };
        T cur[K], nxt[SINGLE ? 1 : K];
        double start = load_tile(0, cur);
        for (uint64_t tile = 0; tile < (SINGLE ? 1 : p.tiles); ++tile) {
            double end = start;
            if constexpr (!SINGLE) if (tile + 1 < p.tiles) {
                const double target = load_tile(tile + 1, nxt);
                const double lowest = max_like_std(p.min_gain, __ddiv_rn(start, p.max_change));
                const double highest = start > __ddiv_rn(p.max_gain, p.max_change) ? p.max_gain
                                                                                   : __dmul_rn(start, p.max_change);
                end = clamp_like_std(target, lowest, highest);
            }
            const uint64_t length = tile_length(tile);
            const double step = __ddiv_rn(__dsub_rn(end, start), static_cast<double>(length));
            T* const d = dst + tile * p.tile;
#pragma unroll
            for (int u = 0; u < K; ++u) {
                const uint64_t idx = tid + static_cast<uint64_t>(u) * kAgcFusedThreads;
                if (idx < length) {
                    const double gain = __dadd_rn(start, __dmul_rn(step, static_cast<double>(idx)));
                    d[idx] = apply_gain(cur[u], gain);
                }
            }
            if constexpr (!SINGLE) if (tile + 1 < p.tiles) {
#pragma unroll
                for (int u = 0; u < K; ++u) {
                    cur[u] = nxt[u];
                }
            }
            start = end;
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

// scratch: (lanes * tiles) doubles + (lanes * tiles) double2 = 24 bytes per tile, device memory, caller-owned.
int b200_agc_scratch_bytes(uint64_t lanes, uint64_t samples, uint64_t tile_size, uint64_t* bytes) {
    B200_REQUIRE(bytes, "b200_agc_scratch_bytes: null argument");
    B200_REQUIRE(tile_size > 0, "[MODULE_AGC] Tile size must be greater than zero.");
    const uint64_t tiles = samples == 0 ? 0 : 1 + (samples - 1) / tile_size;
    *bytes = lanes * tiles * 24 + 8;     // + padding that keeps the double2 region 16-byte aligned
    return B200_SUCCESS;
}

int b200_agc(b200_ctx* ctx, const void* in, void* out, int is_complex, uint64_t lanes, uint64_t samples,
             uint64_t tile_size, double reference, double epsilon, double min_gain, double max_gain,
             double max_gain_change, void* scratch, b200_stream stream) {
    B200_REQUIRE(ctx, "b200_agc: null context");
    B200_REQUIRE(tile_size > 0, "[MODULE_AGC] Tile size must be greater than zero.");
    B200_REQUIRE(std::isfinite(reference) && reference > 0.0, "[MODULE_AGC] Reference must be finite and positive.");
    B200_REQUIRE(std::isfinite(epsilon) && epsilon > 0.0, "[MODULE_AGC] Epsilon must be finite and positive.");
    B200_REQUIRE(std::isfinite(min_gain) && min_gain > 0.0, "[MODULE_AGC] Minimum gain must be finite and positive.");
    B200_REQUIRE(std::isfinite(max_gain) && max_gain >= min_gain,
                 "[MODULE_AGC] Maximum gain must be finite and no less than minimum gain.");
    B200_REQUIRE(std::isfinite(max_gain_change) && max_gain_change >= 1.0,
                 "[MODULE_AGC] Maximum gain change must be finite and at least one.");
    if (lanes == 0 || samples == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(in && out && scratch, "b200_agc: null buffer");
    DeviceGuard guard(ctx);
    AgcParams p{};
    p.lanes = lanes;
    p.samples = samples;
    p.tile = tile_size;
    p.tiles = 1 + (samples - 1) / tile_size;
    p.reference = reference;
    p.epsilon = epsilon;
    p.min_gain = min_gain;
    p.max_gain = max_gain;
    p.max_change = max_gain_change;
    double* const target = static_cast<double*>(scratch);
    double2* const ends = reinterpret_cast<double2*>(target + lanes * p.tiles + ((lanes * p.tiles) & 1));
    const uint64_t items = lanes * p.tiles;
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(items, static_cast<uint64_t>(ctx->sms) * 8));
    const cudaStream_t s = as_stream(stream);
    // Fused form (every sample read once): tiles that fit the registers of one CTA and at least one lane per resident CTA.
    const char* fused_env = getenv("B200_AGC_FUSED");
    // Measured on [16384, 4096] CF32: tile 1024 (4 samples per thread, 4 CTAs/SM) 0.358 -> 0.303 ms; tile 4096 (16 per
    // thread: 80 registers alone, 122 with a look-ahead tile) 0.284 -> 0.318 ms, so long tiles keep the three-kernel form
    // (B200_AGC_FUSED=1 forces the fused form up to 16 samples per thread, =0 disables it).
    const uint64_t per_thread = (std::min(tile_size, samples) + kAgcFusedThreads - 1) / kAgcFusedThreads;
    const uint64_t fused_limit = fused_env ? (atoi(fused_env) != 0 ? kAgcFusedPerThread : 0) : 4;
    if (per_thread <= fused_limit && lanes >= static_cast<uint64_t>(ctx->sms) * 2) {
        const unsigned fgrid = static_cast<unsigned>(std::min<uint64_t>(lanes, static_cast<uint64_t>(ctx->sms) * 4));
        const int k = static_cast<int>((std::min(tile_size, samples) + kAgcFusedThreads - 1) / kAgcFusedThreads);
#define B200_AGC_FUSED(T, K)                                                                                              \
    if (p.tiles == 1) {                                                                                                   \
        agc_fused_kernel<T, K, true><<<fgrid, kAgcFusedThreads, 0, s>>>(static_cast<const T*>(in), static_cast<T*>(out), p);  \
    } else {                                                                                                              \
        agc_fused_kernel<T, K, false><<<fgrid, kAgcFusedThreads, 0, s>>>(static_cast<const T*>(in), static_cast<T*>(out), p); \
    }
        if (is_complex) {
            if (k <= 4) { B200_AGC_FUSED(float2, 4); } else if (k <= 8) { B200_AGC_FUSED(float2, 8); } else { B200_AGC_FUSED(float2, 16); }
        } else {
            if (k <= 4) { B200_AGC_FUSED(float, 4); } else if (k <= 8) { B200_AGC_FUSED(float, 8); } else { B200_AGC_FUSED(float, 16); }
        }
#undef B200_AGC_FUSED
        B200_LAUNCH_CHECK();
        return B200_SUCCESS;
    }
    if (is_complex) {
        agc_tile_gain_kernel<float2><<<grid, 256, 0, s>>>(static_cast<const float2*>(in), target, p);
    } else {
        agc_tile_gain_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(in), target, p);
    }
    B200_LAUNCH_CHECK();
    agc_gain_chain_kernel<<<static_cast<unsigned>((lanes + 63) / 64), 64, 0, s>>>(target, ends, p);
    B200_LAUNCH_CHECK();
    if (is_complex) {
        agc_apply_kernel<float2><<<grid, 256, 0, s>>>(static_cast<const float2*>(in), static_cast<float2*>(out), ends, p);
    } else {
        agc_apply_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(in), static_cast<float*>(out), ends, p);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

}  // extern "C"
