// FM quadrature demodulation — replaces FmImplNativeCpu::computeSubmit
// (src/domains/dsp/fm/module_impl_native_cpu.cc:43-175), narrow mode:
//     d[n] = arg(conj(x[n-1]) * x[n]) * ref,   ref = 1 / (2 pi (100e3 / sampleRate))      (module_impl.cc:108-111)
//     first sample ever -> 0; non-finite x[n] or x[n-1] -> NaN (and the de-emphasis state is not touched)
//     optional one-pole de-emphasis y += alpha (d - y), state carried per lane across frames and cycles.
// Tensor layout [frames, lanes, frame_len]: frames (the batch axis) are consecutive in time, lanes
// (channel axis, e.g. filter heads) are independent. The discriminator is elementwise (12 B/sample);
// the de-emphasis recurrence is evaluated as a blocked linear-recurrence scan: per-chunk (gain, offset)
// pairs -> serial carry propagation over chunks -> each chunk replayed sequentially from its carry in the
// reference's own operation order.
#include <cmath>

#include "common.cuh"

namespace b200 {

struct FmState {          // one per lane, device resident
    float2 previous;
    float deemphasis;
    int has_previous;
};

__device__ __forceinline__ bool finite2(const float2 v) { return isfinite(v.x) && isfinite(v.y); }

// std::arg(std::conj(p) * c) * ref with the reference's F32 rounding sequence (no FMA contraction).
__device__ __forceinline__ float discriminate(const float2 p, const float2 c, const float ref) {
    const float re = __fadd_rn(__fmul_rn(p.x, c.x), __fmul_rn(p.y, c.y));
    const float im = __fsub_rn(__fmul_rn(p.x, c.y), __fmul_rn(p.y, c.x));
    return __fmul_rn(atan2f(im, re), ref);
}

__global__ void fm_discriminator_kernel(const float2* __restrict__ x, float* __restrict__ out,
                                        const FmState* __restrict__ state, const uint64_t frames,
                                        const uint64_t lanes, const uint64_t frame_len, const float ref) {
    const uint64_t total = frames * lanes * frame_len;
    for (uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t s = e % frame_len;
        const uint64_t row = e / frame_len;          // frame * lanes + lane
        const uint64_t lane = row % lanes;
        const uint64_t frame = row / lanes;
        const float2 cur = x[e];
        float2 prev;
        bool has_prev = true;
        if (s > 0) {
            prev = x[e - 1];
        } else if (frame > 0) {
            prev = x[((frame - 1) * lanes + lane) * frame_len + frame_len - 1];
        } else {
            prev = state[lane].previous;
            has_prev = state[lane].has_previous != 0;
        }
        float d;
        if (!has_prev) {
            d = 0.0f;
        } else if (finite2(cur) && finite2(prev)) {
            d = discriminate(prev, cur, ref);
        } else {
            d = __int_as_float(0x7fc00000);
        }
        out[e] = d;
    }
}

// ---- de-emphasis: y[n] = y[n-1] + alpha (d[n] - y[n-1]) over finite d, blocked scan ------------------
constexpr int kFmChunk = 256;

// Phase 1: per chunk, zero-state response at the chunk end (offset) and the state gain (1-alpha)^#finite.
__global__ void fm_deemph_reduce_kernel(const float* __restrict__ d, float2* __restrict__ chunk_coeff,
                                        const uint64_t frames, const uint64_t lanes, const uint64_t frame_len,
                                        const uint64_t chunks_per_lane, const float alpha) {
    const uint64_t total_chunks = chunks_per_lane * lanes;
    const uint64_t lane_len = frames * frame_len;
    for (uint64_t c = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; c < total_chunks;
         c += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t lane = c / chunks_per_lane;
        const uint64_t n0 = (c % chunks_per_lane) * kFmChunk;
        float y = 0.0f, gain = 1.0f;
        const float keep = 1.0f - alpha;
        for (uint64_t n = n0; n < n0 + kFmChunk && n < lane_len; ++n) {
            const uint64_t frame = n / frame_len, s = n % frame_len;
            const float v = d[(frame * lanes + lane) * frame_len + s];
            if (isfinite(v)) {
                y = __fadd_rn(y, __fmul_rn(alpha, __fsub_rn(v, y)));
                gain *= keep;
            }
        }
        chunk_coeff[c] = make_float2(gain, y);
    }
}

// Phase 2: carries. One thread per lane walks its chunks: carry[c+1] = gain_c * carry[c] + offset_c.
__global__ void fm_deemph_carry_kernel(float2* __restrict__ chunk_coeff, const FmState* __restrict__ state,
                                       const uint64_t lanes, const uint64_t chunks_per_lane) {
    const uint64_t lane = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (lane >= lanes) {
        return;
    }
    float carry = state[lane].deemphasis;
    for (uint64_t c = 0; c < chunks_per_lane; ++c) {
        const float2 k = chunk_coeff[lane * chunks_per_lane + c];
        chunk_coeff[lane * chunks_per_lane + c].x = carry;       // carry-in of chunk c
        carry = fmaf(k.x, carry, k.y);
    }
}

// Phase 3: replay each chunk from its carry in the reference's operation order; the last chunk of each
// lane publishes the new state.
__global__ void fm_deemph_apply_kernel(float* __restrict__ d, const float2* __restrict__ chunk_coeff,
                                       FmState* __restrict__ state, const uint64_t frames, const uint64_t lanes,
                                       const uint64_t frame_len, const uint64_t chunks_per_lane, const float alpha) {
    const uint64_t total_chunks = chunks_per_lane * lanes;
    const uint64_t lane_len = frames * frame_len;
    for (uint64_t c = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; c < total_chunks;
         c += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t lane = c / chunks_per_lane;
        const uint64_t ci = c % chunks_per_lane;
        const uint64_t n0 = ci * kFmChunk;
        float y = chunk_coeff[c].x;
        for (uint64_t n = n0; n < n0 + kFmChunk && n < lane_len; ++n) {
            const uint64_t frame = n / frame_len, s = n % frame_len;
            float* const slot = d + (frame * lanes + lane) * frame_len + s;
            const float v = *slot;
            if (isfinite(v)) {
                y = __fadd_rn(y, __fmul_rn(alpha, __fsub_rn(v, y)));
                *slot = y;
            }
        }
        if (ci == chunks_per_lane - 1) {
            state[lane].deemphasis = y;
        }
    }
}

__global__ void fm_state_update_kernel(const float2* __restrict__ x, FmState* __restrict__ state,
                                       const uint64_t frames, const uint64_t lanes, const uint64_t frame_len) {
    const uint64_t lane = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (lane < lanes) {
        state[lane].previous = x[((frames - 1) * lanes + lane) * frame_len + frame_len - 1];
        state[lane].has_previous = 1;
    }
}

}  // namespace b200

using namespace b200;

struct b200_fm_plan {
    b200_ctx* ctx;
    uint64_t lanes;
    float ref;
    float alpha;          // 1.0 = de-emphasis disabled
    bool deemphasis;
    FmState* state;
    float2* chunk_coeff;
    uint64_t chunk_capacity;
};

extern "C" {

int b200_fm_plan_create(b200_ctx* ctx, uint64_t lanes, float sample_rate, int wide, int deemphasis_us,
                        b200_fm_plan** plan) {
    B200_REQUIRE(ctx && plan, "b200_fm_plan_create: null argument");
    *plan = nullptr;
    B200_REQUIRE(lanes >= 1, "b200_fm_plan_create: lanes must be positive");
    B200_REQUIRE(std::isfinite(sample_rate) && sample_rate > 0.0f && sample_rate <= 20e6f,
                 "b200_fm_plan_create: sample rate must be finite, positive and <= 20 MHz");
    B200_REQUIRE(deemphasis_us == 0 || deemphasis_us == 50 || deemphasis_us == 75,
                 "b200_fm_plan_create: de-emphasis must be 0 (none), 50 or 75 us");
    B200_REQUIRE(!wide, "b200_fm_plan_create: wideband stereo mode is not implemented by this provider yet");
    DeviceGuard guard(ctx);
    auto* pl = new b200_fm_plan();
    pl->ctx = ctx;
    pl->lanes = lanes;
    // FmImpl::updateCoefficients (src/domains/dsp/fm/module_impl.cc:108-124), F32 arithmetic as written there.
    const float kPi = 3.14159265358979323846f;   // JST_PI is a double literal; the products below round to F32
    const float deviation = 100e3f;
    const float kf = deviation / sample_rate;
    pl->ref = static_cast<float>(1.0f / (2.0 * 3.14159265358979323846 * kf));
    (void)kPi;
    pl->deemphasis = deemphasis_us != 0;
    if (pl->deemphasis) {
        const double tau = deemphasis_us == 50 ? 50e-6 : 75e-6;
        pl->alpha = static_cast<float>(1.0 - std::exp(-1.0 / (static_cast<double>(sample_rate) * tau)));
    } else {
        pl->alpha = 1.0f;
    }
    void* st = nullptr;
    if (b200_malloc(ctx, lanes * sizeof(FmState), &st) != B200_SUCCESS) {
        delete pl;
        return B200_ERROR;
    }
    pl->state = static_cast<FmState*>(st);
    pl->chunk_coeff = nullptr;
    pl->chunk_capacity = 0;
    *plan = pl;
    return B200_SUCCESS;
}

int b200_fm_reset(b200_fm_plan* plan, b200_stream stream) {
    B200_REQUIRE(plan, "b200_fm_reset: null plan");
    DeviceGuard guard(plan->ctx);
    B200_CUDA_CHECK(cudaMemsetAsync(plan->state, 0, plan->lanes * sizeof(FmState), as_stream(stream)));
    return B200_SUCCESS;
}

int b200_fm_exec(b200_fm_plan* plan, const b200_cf32* x, float* out, uint64_t frames, uint64_t frame_len,
                 b200_stream stream) {
    B200_REQUIRE(plan, "b200_fm_exec: null plan");
    const uint64_t total = frames * plan->lanes * frame_len;
    if (total == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(x && out, "b200_fm_exec: null buffer");
    DeviceGuard guard(plan->ctx);
    const cudaStream_t s = as_stream(stream);
    const uint64_t blocks = (total + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(plan->ctx->sms) * 8;
    fm_discriminator_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, s>>>(
        reinterpret_cast<const float2*>(x), out, plan->state, frames, plan->lanes, frame_len, plan->ref);
    B200_LAUNCH_CHECK();
    if (plan->deemphasis) {
        const uint64_t lane_len = frames * frame_len;
        const uint64_t chunks_per_lane = (lane_len + kFmChunk - 1) / kFmChunk;
        const uint64_t total_chunks = chunks_per_lane * plan->lanes;
        if (total_chunks > plan->chunk_capacity) {
            cudaFree(plan->chunk_coeff);
            plan->chunk_coeff = nullptr;
            plan->chunk_capacity = 0;
            B200_CUDA_CHECK(cudaMalloc(&plan->chunk_coeff, total_chunks * sizeof(float2)));
            plan->chunk_capacity = total_chunks;
        }
        const unsigned cgrid = static_cast<unsigned>(std::min<uint64_t>((total_chunks + 127) / 128, cap));
        fm_deemph_reduce_kernel<<<cgrid, 128, 0, s>>>(out, plan->chunk_coeff, frames, plan->lanes, frame_len,
                                                     chunks_per_lane, plan->alpha);
        B200_LAUNCH_CHECK();
        fm_deemph_carry_kernel<<<static_cast<unsigned>((plan->lanes + 63) / 64), 64, 0, s>>>(
            plan->chunk_coeff, plan->state, plan->lanes, chunks_per_lane);
        B200_LAUNCH_CHECK();
        fm_deemph_apply_kernel<<<cgrid, 128, 0, s>>>(out, plan->chunk_coeff, plan->state, frames, plan->lanes,
                                                    frame_len, chunks_per_lane, plan->alpha);
        B200_LAUNCH_CHECK();
    }
    fm_state_update_kernel<<<static_cast<unsigned>((plan->lanes + 63) / 64), 64, 0, s>>>(
        reinterpret_cast<const float2*>(x), plan->state, frames, plan->lanes, frame_len);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_fm_plan_destroy(b200_fm_plan* plan) {
    if (!plan) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(plan->ctx);
    cudaFree(plan->state);
    cudaFree(plan->chunk_coeff);
    delete plan;
    return B200_SUCCESS;
}

}  // extern "C"
