// Backend + memory + stream entry points of libb200dsp.
// Replaces: src/backend/devices/cuda/base.cc:9-149 (device context), src/memory/buffer_cuda.cc:31-124
// (device allocation) and the stream ownership of src/runtime/native/cuda/impl.cc:35-118 in the
// reference. One b200_ctx per device; any number may coexist in a process (the reference has a
// single-device singleton, include/jetstream/backend/base.hh:125-136).
#include <cmath>

#include <cstring>

#include "common.cuh"

namespace b200 {

std::string& last_error() {
    thread_local std::string text;
    return text;
}

int fail(const char* fmt, ...) {
    char buffer[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buffer, sizeof(buffer), fmt, args);
    va_end(args);
    last_error() = buffer;
    return B200_ERROR;
}

}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_version(void) { return "b200dsp 0.1.0 (sm_100a)"; }

const char* b200_last_error(void) { return last_error().c_str(); }

int b200_amplitude_scaling_coeff(uint64_t n, float* coeff) {
    B200_REQUIRE(coeff != nullptr && n > 0, "b200_amplitude_scaling_coeff: bad argument");
    *coeff = 20.0f * std::log10(1.0f / static_cast<float>(n));
    return B200_SUCCESS;
}

int b200_range_coefficients(float min, float max, float* scale, float* offset) {
    B200_REQUIRE(scale && offset, "b200_range_coefficients: null output");
    const float lower = min < max ? min : max;
    const float upper = min < max ? max : min;
    if (lower == upper) {
        *scale = 0.0f;
        *offset = 0.5f;
        return B200_SUCCESS;
    }
    *scale = 1.0f / (upper - lower);
    *offset = -lower * *scale;
    return B200_SUCCESS;
}

int b200_device_count(int* count) {
    B200_REQUIRE(count != nullptr, "b200_device_count: null output");
    *count = 0;
    B200_CUDA_CHECK(cudaGetDeviceCount(count));
    return B200_SUCCESS;
}

int b200_ctx_create(int device, b200_ctx** out) {
    B200_REQUIRE(out != nullptr, "b200_ctx_create: null output");
    *out = nullptr;
    int count = 0;
    B200_CUDA_CHECK(cudaGetDeviceCount(&count));
    B200_REQUIRE(device >= 0 && device < count, "b200_ctx_create: device %d out of range (%d present)",
                 device, count);
    cudaDeviceProp prop;
    B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    B200_REQUIRE(prop.major == 10, "b200_ctx_create: device %d is sm_%d%d; this library ships sm_100a code only",
                 device, prop.major, prop.minor);
    auto* ctx = new b200_ctx();
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    ctx->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    {
        DeviceGuard guard(ctx);
        B200_CUDA_CHECK(cudaFree(nullptr));  // force primary-context creation on this device
    }
    *out = ctx;
    return B200_SUCCESS;
}

int b200_ctx_destroy(b200_ctx* ctx) {
    delete ctx;
    return B200_SUCCESS;
}

int b200_ctx_device(const b200_ctx* ctx, int* device) {
    B200_REQUIRE(ctx && device, "b200_ctx_device: null argument");
    *device = ctx->device;
    return B200_SUCCESS;
}

int b200_ctx_sm_count(const b200_ctx* ctx, int* sms) {
    B200_REQUIRE(ctx && sms, "b200_ctx_sm_count: null argument");
    *sms = ctx->sms;
    return B200_SUCCESS;
}

int b200_malloc(b200_ctx* ctx, uint64_t bytes, void** ptr) {
    B200_REQUIRE(ctx && ptr, "b200_malloc: null argument");
    *ptr = nullptr;
    if (bytes == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaMalloc(ptr, bytes));
    const cudaError_t e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return fail("b200_malloc: cudaMemset failed: %s", cudaGetErrorString(e));
    }
    // The zero-fill ran on the legacy stream, which non-blocking streams (the runtime's, PyTorch's) do not order
    // against: settle it here, allocation is not on the per-cycle path.
    B200_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    return B200_SUCCESS;
}

int b200_free(b200_ctx* ctx, void* ptr) {
    B200_REQUIRE(ctx != nullptr, "b200_free: null context");
    if (!ptr) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaFree(ptr));
    return B200_SUCCESS;
}

int b200_host_alloc(b200_ctx* ctx, uint64_t bytes, void** ptr) {
    B200_REQUIRE(ctx && ptr, "b200_host_alloc: null argument");
    *ptr = nullptr;
    if (bytes == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return B200_SUCCESS;
}

int b200_host_free(b200_ctx* ctx, void* ptr) {
    B200_REQUIRE(ctx != nullptr, "b200_host_free: null context");
    if (!ptr) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaFreeHost(ptr));
    return B200_SUCCESS;
}

int b200_memcpy(b200_ctx* ctx, void* dst, const void* src, uint64_t bytes, int kind, b200_stream stream) {
    B200_REQUIRE(ctx != nullptr, "b200_memcpy: null context");
    B200_REQUIRE(kind >= 0 && kind <= 3, "b200_memcpy: kind must be 0 (h2d), 1 (d2h), 2 (d2d) or 3 (by address, UVA)");
    if (bytes == 0) {
        return B200_SUCCESS;
    }
    B200_REQUIRE(dst && src, "b200_memcpy: null pointer");
    DeviceGuard guard(ctx);
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice
                             : (kind == 1 ? cudaMemcpyDeviceToHost : (kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyDefault));
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, k, as_stream(stream)));
    return B200_SUCCESS;
}

int b200_memset(b200_ctx* ctx, void* dst, int value, uint64_t bytes, b200_stream stream) {
    B200_REQUIRE(ctx != nullptr, "b200_memset: null context");
    if (bytes == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaMemsetAsync(dst, value, bytes, as_stream(stream)));
    return B200_SUCCESS;
}

int b200_stream_create(b200_ctx* ctx, b200_stream* stream) {
    B200_REQUIRE(ctx && stream, "b200_stream_create: null argument");
    DeviceGuard guard(ctx);
    cudaStream_t s;
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return B200_SUCCESS;
}

int b200_stream_destroy(b200_ctx* ctx, b200_stream stream) {
    B200_REQUIRE(ctx != nullptr, "b200_stream_destroy: null context");
    if (!stream) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaStreamDestroy(as_stream(stream)));
    return B200_SUCCESS;
}

int b200_stream_synchronize(b200_ctx* ctx, b200_stream stream) {
    B200_REQUIRE(ctx != nullptr, "b200_stream_synchronize: null context");
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    return B200_SUCCESS;
}

int b200_malloc_managed(b200_ctx* ctx, uint64_t bytes, void** ptr) {
    B200_REQUIRE(ctx && ptr, "b200_malloc_managed: null argument");
    *ptr = nullptr;
    if (bytes == 0) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaMallocManaged(ptr, bytes));
    const cudaError_t e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return fail("b200_malloc_managed: cudaMemset failed: %s", cudaGetErrorString(e));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    return B200_SUCCESS;
}

int b200_host_register(b200_ctx* ctx, void* host, uint64_t bytes, int* registered) {
    B200_REQUIRE(ctx && host && registered, "b200_host_register: null argument");
    *registered = 0;
    DeviceGuard guard(ctx);
    cudaPointerAttributes attributes = {};
    const cudaError_t query = cudaPointerGetAttributes(&attributes, host);
    if (query == cudaSuccess && attributes.type != cudaMemoryTypeUnregistered) {
        return B200_SUCCESS;                 // already pinned / mapped by someone else: nothing to own
    }
    if (query != cudaSuccess) {
        cudaGetLastError();
    }
    B200_CUDA_CHECK(cudaHostRegister(host, bytes, cudaHostRegisterPortable));
    *registered = 1;
    return B200_SUCCESS;
}

int b200_host_unregister(b200_ctx* ctx, void* host) {
    B200_REQUIRE(ctx && host, "b200_host_unregister: null argument");
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaHostUnregister(host));
    return B200_SUCCESS;
}

int b200_event_create(b200_ctx* ctx, b200_event* event) {
    B200_REQUIRE(ctx && event, "b200_event_create: null argument");
    DeviceGuard guard(ctx);
    cudaEvent_t e;
    B200_CUDA_CHECK(cudaEventCreate(&e));
    *event = e;
    return B200_SUCCESS;
}

int b200_event_record(b200_ctx* ctx, b200_event event, b200_stream stream) {
    B200_REQUIRE(ctx && event, "b200_event_record: null argument");
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(event), as_stream(stream)));
    return B200_SUCCESS;
}

int b200_event_elapsed_ms(b200_ctx* ctx, b200_event start, b200_event end, float* ms) {
    B200_REQUIRE(ctx && start && end && ms, "b200_event_elapsed_ms: null argument");
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(end)));
    return B200_SUCCESS;
}

int b200_event_destroy(b200_ctx* ctx, b200_event event) {
    B200_REQUIRE(ctx != nullptr, "b200_event_destroy: null context");
    if (!event) {
        return B200_SUCCESS;
    }
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
    return B200_SUCCESS;
}

/* The last asynchronous CUDA error of the calling thread, cleared (what NativeCudaRuntime checks after every submit,
 * src/runtime/native/cuda/impl.cc:228-231). */
int b200_check_async_error(b200_ctx* ctx) {
    B200_REQUIRE(ctx != nullptr, "b200_check_async_error: null context");
    DeviceGuard guard(ctx);
    B200_CUDA_CHECK(cudaGetLastError());
    return B200_SUCCESS;
}

int b200_ctx_info(const b200_ctx* ctx, b200_device_info* info) {
    B200_REQUIRE(ctx && info, "b200_ctx_info: null argument");
    DeviceGuard guard(ctx);
    cudaDeviceProp prop{};
    B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
    memset(info, 0, sizeof(*info));
    strncpy(info->name, prop.name, sizeof(info->name) - 1);
    info->device = ctx->device;
    info->compute_capability_major = prop.major;
    info->compute_capability_minor = prop.minor;
    info->sm_count = prop.multiProcessorCount;
    info->total_memory_bytes = prop.totalGlobalMem;
    info->integrated = prop.integrated;
    info->can_map_host_memory = prop.canMapHostMemory;
    info->can_use_host_pointer_for_registered_memory = prop.canUseHostPointerForRegisteredMem;
    info->shared_memory_per_block_optin = static_cast<uint64_t>(prop.sharedMemPerBlockOptin);
    int runtime = 0;
    B200_CUDA_CHECK(cudaRuntimeGetVersion(&runtime));
    info->runtime_version = runtime;
    return B200_SUCCESS;
}

}  // extern "C"
