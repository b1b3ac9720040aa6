// lineplot / waterfall consumer compute: the step right after the spectral chain.
//   lineplot   LineplotImplNativeCpu::computeSubmit   src/domains/visualization/lineplot/module_impl_native_cpu.cc:80-122
//              (CUDA counterpart lineplot/module_impl_native_cuda.cc:21-60: one thread per element looping over batches)
//   waterfall  WaterfallImplNativeCpu::computeSubmit  src/domains/visualization/waterfall/module_impl_native_cpu.cc:53-78
//              ring bookkeeping                        src/domains/visualization/waterfall/ring_state.hh:18-44
//
// lineplot = column sums over the batch (with decimation) -> normalise -> clamp -> EMA into the module's persistent
// average. The column sum is the only part that touches the [batches, extent] input (1 GiB for BASELINE configs[1]):
// it runs as a row-split partial-sum kernel (HBM-bound, 4 B per input element) + an O(n) finalize kernel. The split is
// a fixed function of (batches, elements), so results are reproducible run to run; they differ from the reference's
// strictly sequential F32 accumulation by reassociation only (identical when one CTA covers all rows, i.e. small
// batches). The spectral chain kernel can also deliver the column sums from its own epilogue
// (b200_chain_exec_colsum), in which case b200_lineplot_update_from_colsum skips the big read entirely.
#include "common.cuh"

namespace b200 {

namespace {

constexpr int kColsumThreads = 256;
constexpr int kRowsUnroll = 8;

// partial[split][e] = sum over the split's rows b of in[b * batch_stride + e * col_stride]   (e < elements)
// VEC: col_stride == 1 and 16-byte aligned rows: four columns per thread with float4 loads.
template <bool VEC>
__global__ void __launch_bounds__(kColsumThreads) colsum_partial_kernel(const float* __restrict__ in,
                                                                        float* __restrict__ partial,
                                                                        const uint64_t batches, const uint64_t elements,
                                                                        const uint64_t batch_stride,
                                                                        const uint64_t col_stride,
                                                                        const uint64_t rows_per_split) {
    const uint64_t split = blockIdx.y;
    const uint64_t row0 = split * rows_per_split;
    const uint64_t row1 = min(batches, row0 + rows_per_split);
    if constexpr (VEC) {
        const uint64_t quad = blockIdx.x * static_cast<uint64_t>(kColsumThreads) + threadIdx.x;   // columns 4q..4q+3
        if (quad * 4 >= elements) {
            return;
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* base = reinterpret_cast<const float4*>(in) + quad;
        const uint64_t pitch4 = batch_stride / 4;
        uint64_t b = row0;
        for (; b + kRowsUnroll <= row1; b += kRowsUnroll) {
            float4 v[kRowsUnroll];
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {
                v[u] = ldg_stream_f4(base + (b + u) * pitch4);
            }
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {        // row order inside a split = the reference's b order
                acc.x = __fadd_rn(acc.x, v[u].x);
                acc.y = __fadd_rn(acc.y, v[u].y);
                acc.z = __fadd_rn(acc.z, v[u].z);
                acc.w = __fadd_rn(acc.w, v[u].w);
            }
        }
        for (; b < row1; ++b) {
            const float4 v = ldg_stream_f4(base + b * pitch4);
            acc.x = __fadd_rn(acc.x, v.x);
            acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z);
            acc.w = __fadd_rn(acc.w, v.w);
        }
        *reinterpret_cast<float4*>(partial + split * elements + quad * 4) = acc;
    } else {
        const uint64_t e = blockIdx.x * static_cast<uint64_t>(kColsumThreads) + threadIdx.x;
        if (e >= elements) {
            return;
        }
        float acc = 0.f;
        const float* base = in + e * col_stride;
        uint64_t b = row0;
        for (; b + kRowsUnroll <= row1; b += kRowsUnroll) {
            float v[kRowsUnroll];
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {
                v[u] = __ldg(base + (b + u) * batch_stride);
            }
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {
                acc = __fadd_rn(acc, v[u]);
            }
        }
        for (; b < row1; ++b) {
            acc = __fadd_rn(acc, __ldg(base + b * batch_stride));
        }
        partial[split * elements + e] = acc;
    }
}

// sums[e] = partial[0][e * pick] + partial[1][e * pick] + ... (fixed order), then lineplot/module_impl_native_cpu.cc:
// 103-113: amplitude = fmin(fmax(sum * normalization - 1, -1), 1); average -= average / averaging;
// average += amplitude / averaging; points[2 e + 1] = average. Every step individually rounded (no FMA).
__global__ void lineplot_finalize_kernel(const float* __restrict__ partial, const uint64_t splits,
                                         const uint64_t partial_pitch, const uint64_t pick, const uint64_t elements,
                                         const float normalization, const float averaging,
                                         float* __restrict__ average, float* __restrict__ points) {
    const uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (e >= elements) {
        return;
    }
    float sum = 0.f;
    for (uint64_t s = 0; s < splits; ++s) {
        sum = __fadd_rn(sum, partial[s * partial_pitch + e * pick]);
    }
    const float amplitude = fminf(fmaxf(__fsub_rn(__fmul_rn(sum, normalization), 1.0f), -1.0f), 1.0f);
    float avg = average[e];
    avg = __fsub_rn(avg, __fdiv_rn(avg, averaging));
    avg = __fadd_rn(avg, __fdiv_rn(amplitude, averaging));
    average[e] = avg;
    points[2 * e + 1] = avg;
}

// signalPoints x coordinates, lineplot/module_impl_native_cpu.cc:66-70: i * 2.0f / (n - 1) - 1.0f, y = 0.
__global__ void lineplot_init_kernel(float* __restrict__ points, float* __restrict__ average, const uint64_t elements) {
    const uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (e >= elements) {
        return;
    }
    const float x = __fsub_rn(__fdiv_rn(__fmul_rn(static_cast<float>(e), 2.0f), static_cast<float>(elements - 1)), 1.0f);
    points[2 * e] = x;
    points[2 * e + 1] = 0.0f;
    average[e] = 0.0f;
}

// ring[(dest + r) % height][c] = in[(source + r) * batch_stride + c * col_stride]
template <bool VEC>
__global__ void waterfall_write_kernel(const float* __restrict__ in, float* __restrict__ ring, const uint64_t rows,
                                       const uint64_t elements, const uint64_t batch_stride, const uint64_t col_stride,
                                       const uint64_t source, const uint64_t dest, const uint64_t height) {
    const uint64_t per_row = VEC ? elements / 4 : elements;
    const uint64_t total = rows * per_row;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / per_row, c = i - r * per_row;
        const uint64_t out_row = (dest + r) % height;
        if constexpr (VEC) {
            const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(in + (source + r) * batch_stride) + c);
            *(reinterpret_cast<float4*>(ring + out_row * elements) + c) = v;
        } else {
            ring[out_row * elements + c] = __ldg(in + (source + r) * batch_stride + c * col_stride);
        }
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Row splits of the column-sum kernel: enough CTAs to fill the GPU (>= ~4 per SM across the column blocks), at least
// 64 rows per split so the O(n * splits) finalize stays negligible. Depends only on (batches, column blocks, SMs).
uint64_t colsum_splits(const b200_ctx* ctx, const uint64_t batches, const uint64_t column_blocks) {
    const uint64_t want = (static_cast<uint64_t>(ctx->sms) * 4 + column_blocks - 1) / column_blocks;
    const uint64_t cap = (batches + 63) / 64;
    uint64_t splits = want < cap ? want : cap;
    return splits == 0 ? 1 : splits;
}

}  // namespace

// Column sums of a [batches, columns] F32 matrix into partials (scratch) — shared by b200_lineplot_update and the
// unfused fallback of b200_chain_exec_colsum. Returns the number of splits written.
int colsum_partials(b200_ctx* ctx, const float* in, uint64_t batches, uint64_t columns, uint64_t batch_stride,
                    uint64_t col_stride, float* partial, uint64_t* splits_out, cudaStream_t s) {
    const bool vec = col_stride == 1 && columns % 4 == 0 && batch_stride % 4 == 0 && aligned16(in) && aligned16(partial);
    const uint64_t per_block = vec ? kColsumThreads * 4 : kColsumThreads;
    const uint64_t blocks = (columns + per_block - 1) / per_block;
    const uint64_t splits = colsum_splits(ctx, batches, blocks);
    const uint64_t rows_per_split = (batches + splits - 1) / splits;
    const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(splits));
    if (vec) {
        colsum_partial_kernel<true><<<grid, kColsumThreads, 0, s>>>(in, partial, batches, columns, batch_stride,
                                                                   col_stride, rows_per_split);
    } else {
        colsum_partial_kernel<false><<<grid, kColsumThreads, 0, s>>>(in, partial, batches, columns, batch_stride,
                                                                    col_stride, rows_per_split);
    }
    B200_LAUNCH_CHECK();
    *splits_out = splits;
    return B200_SUCCESS;
}

// out[e] = partial[0][e] + partial[1][e] + ... in split order (fixed, reproducible)
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, const uint64_t splits, const uint64_t n,
                                     float* __restrict__ out) {
    const uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (e >= n) {
        return;
    }
    float sum = 0.f;
    for (uint64_t s = 0; s < splits; ++s) {
        sum = __fadd_rn(sum, partial[s * n + e]);
    }
    out[e] = sum;
}

int colsum_reduce(const float* partial, uint64_t splits, uint64_t n, float* out, cudaStream_t s) {
    colsum_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(partial, splits, n, out);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

uint64_t colsum_max_splits(uint64_t batches) { return (batches + 63) / 64 == 0 ? 1 : (batches + 63) / 64; }

}  // namespace b200

using namespace b200;

extern "C" {

int b200_lineplot_scratch_bytes(uint64_t batches, uint64_t elements, uint64_t decimation, uint64_t* bytes) {
    B200_REQUIRE(bytes != nullptr && decimation >= 1, "b200_lineplot_scratch_bytes: null argument or zero decimation");
    *bytes = colsum_max_splits(batches) * elements * decimation * sizeof(float) + 16;
    return B200_SUCCESS;
}

int b200_lineplot_init(b200_ctx* ctx, float* points, float* average, uint64_t elements, b200_stream stream) {
    B200_REQUIRE(ctx && points && average, "b200_lineplot_init: null argument");
    B200_REQUIRE(elements >= 2, "b200_lineplot_init: need at least 2 elements (lineplot/module_impl.cc:130-134)");
    DeviceGuard guard(ctx);
    lineplot_init_kernel<<<static_cast<unsigned>((elements + 255) / 256), 256, 0, as_stream(stream)>>>(points, average,
                                                                                                      elements);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_lineplot_update(b200_ctx* ctx, const float* in, uint64_t batches, uint64_t elements, uint64_t batch_stride,
                         uint64_t element_stride, uint64_t decimation, float normalization, uint64_t averaging,
                         float* average, float* points, void* scratch, b200_stream stream) {
    B200_REQUIRE(ctx && in && average && points && scratch, "b200_lineplot_update: null argument");
    B200_REQUIRE(decimation >= 1 && averaging >= 1 && elements >= 2 && batches >= 1,
                 "b200_lineplot_update: decimation, averaging >= 1, elements >= 2, batches >= 1");
    DeviceGuard guard(ctx);
    const cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(scratch);
    uint64_t splits = 0;
    // decimation picks every `decimation`-th column: fold it into the column stride (LineplotInputIndex,
    // lineplot/module_impl.hh:24-30) unless the rows are dense, where reading whole rows with float4 and picking in
    // the finalize step is faster than a strided gather.
    const bool dense = element_stride == 1 && decimation <= 4 && decimation > 1;
    uint64_t columns = elements, col_stride = element_stride * decimation, pick = 1;
    if (dense) {
        columns = elements * decimation;     // <= extent
        col_stride = 1;
        pick = decimation;
    }
    const int rc = colsum_partials(ctx, in, batches, columns, batch_stride, col_stride, partial, &splits, s);
    if (rc != B200_SUCCESS) {
        return rc;
    }
    lineplot_finalize_kernel<<<static_cast<unsigned>((elements + 255) / 256), 256, 0, s>>>(
        partial, splits, columns, pick, elements, normalization, static_cast<float>(averaging), average, points);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_lineplot_update_from_colsum(b200_ctx* ctx, const float* colsum, uint64_t elements, uint64_t decimation,
                                     float normalization, uint64_t averaging, float* average, float* points,
                                     b200_stream stream) {
    B200_REQUIRE(ctx && colsum && average && points, "b200_lineplot_update_from_colsum: null argument");
    B200_REQUIRE(decimation >= 1 && averaging >= 1 && elements >= 2,
                 "b200_lineplot_update_from_colsum: decimation, averaging >= 1, elements >= 2");
    DeviceGuard guard(ctx);
    lineplot_finalize_kernel<<<static_cast<unsigned>((elements + 255) / 256), 256, 0, as_stream(stream)>>>(
        colsum, 1, 0, decimation, elements, normalization, static_cast<float>(averaging), average, points);
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_waterfall_update(b200_ctx* ctx, const float* in, uint64_t batches, uint64_t elements, uint64_t batch_stride,
                          uint64_t element_stride, float* ring, uint64_t height, uint64_t write_index,
                          b200_stream stream) {
    B200_REQUIRE(ctx && in && ring, "b200_waterfall_update: null argument");
    B200_REQUIRE(height >= 1 && write_index < height, "b200_waterfall_update: height >= 1 and write_index < height");
    if (batches == 0 || elements == 0) {
        return B200_SUCCESS;
    }
    // PlanWaterfallWrite (ring_state.hh:18-29): only the newest min(batches, height) rows are kept
    const uint64_t retained = batches < height ? batches : height;
    const uint64_t source = batches - retained;
    const uint64_t dest = (write_index + (source % height)) % height;
    DeviceGuard guard(ctx);
    const bool vec = element_stride == 1 && elements % 4 == 0 && batch_stride % 4 == 0 && aligned16(in) && aligned16(ring);
    const uint64_t work = retained * (vec ? elements / 4 : elements);
    uint64_t blocks = (work + 255) / 256;
    const uint64_t cap = static_cast<uint64_t>(ctx->sms) * 8;
    blocks = blocks > cap ? cap : blocks;
    if (vec) {
        waterfall_write_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
            in, ring, retained, elements, batch_stride, element_stride, source, dest, height);
    } else {
        waterfall_write_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
            in, ring, retained, elements, batch_stride, element_stride, source, dest, height);
    }
    B200_LAUNCH_CHECK();
    return B200_SUCCESS;
}

int b200_waterfall_advance(uint64_t* write_index, uint64_t batches, uint64_t height) {
    B200_REQUIRE(write_index && height >= 1, "b200_waterfall_advance: null argument or zero height");
    *write_index = (*write_index + (batches % height)) % height;     // WaterfallRingState::advance, ring_state.hh:41-44
    return B200_SUCCESS;
}

}  // extern "C"
