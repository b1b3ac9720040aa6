// SURVEY.md §8 row a15 / north-star "src/backend (CUDA device)": the reference's Backend::CUDA (declared in
// include/jetstream/backend/devices/cuda/base.hh, defined in src/backend/devices/cuda/base.cc:9-191) re-implemented over
// libb200dsp's b200_ctx_*. shim/build_shim.sh links THIS translation unit in place of the reference's.
//
// One deliberate difference: the reference creates a NEW driver context per backend (cuCtxCreate, base.cc:35-43) and makes
// it current in activate(). This implementation retains the device's PRIMARY context — the one the CUDA runtime, cuFFT,
// NCCL, PyTorch and libb200dsp itself use — so a host application's own device buffers, streams and events are valid
// inside the flowgraph without a second context (and per-context state such as kernel attributes exists once).
// getContext() returns that primary context; everything else (device table, capabilities, banner) is as the reference.
#include <cuda.h>
#include <cuda_runtime.h>

#include "jetstream/backend/devices/cuda/base.hh"
#include "jetstream/logger.hh"
#include "jetstream/macros.hh"

#include "b200_provider.hh"

namespace Jetstream::Backend {

CUDA::CUDA(const Config& config) : config(config), cache({}) {
    b200_ctx* ctx = nullptr;
    b200_device_info info;
    if (b200_ctx_create(static_cast<int>(config.deviceId), &ctx) != B200_SUCCESS ||
        b200_ctx_info(ctx, &info) != B200_SUCCESS) {
        JST_FATAL("[CUDA:B200] Cannot get desired device ID ({}): {}", config.deviceId, b200_last_error());
        JST_CHECK_THROW(Result::ERROR);
    }
    b200_ctx_destroy(ctx);      // modules obtain theirs from B200::Ctx(); this one was only used for the device table
    if (cuInit(0) != CUDA_SUCCESS || cuDeviceGet(&device, static_cast<int>(config.deviceId)) != CUDA_SUCCESS ||
        cudaSetDevice(static_cast<int>(config.deviceId)) != cudaSuccess ||
        cuDevicePrimaryCtxRetain(&context, device) != CUDA_SUCCESS) {
        JST_FATAL("[CUDA:B200] Cannot retain the primary context of device ID ({}).", config.deviceId);
        JST_CHECK_THROW(Result::ERROR);
    }
    _isAvailable = true;

    cache.deviceName = info.name;
    cache.computeCapability = jst::fmt::format("{}{}", info.compute_capability_major, info.compute_capability_minor);
    cache.apiVersion = jst::fmt::format("{}.{}.{}", info.runtime_version / 1000, info.runtime_version % 1000 / 10,
                                        info.runtime_version % 10);
    cache.physicalDeviceType = info.integrated ? PhysicalDeviceType::INTEGRATED : PhysicalDeviceType::DISCRETE;
    cache.hasUnifiedMemory = info.integrated != 0;
    cache.physicalMemory = info.total_memory_bytes;
    cache.canImportDeviceMemory = false;     // CUDA <-> Vulkan memory exchange belongs to the render path (out of scope)
    cache.canExportDeviceMemory = false;
    cache.canImportHostMemory = info.can_use_host_pointer_for_registered_memory != 0 || info.can_map_host_memory != 0;

    JST_INFO("-----------------------------------------------------");
    JST_INFO("Jetstream Heterogeneous Backend [CUDA / libb200dsp {}]", b200_version());
    JST_INFO("-----------------------------------------------------");
    JST_INFO("Device ID:          {}", getDeviceId());
    JST_INFO("Device Name:        {}", getDeviceName());
    JST_INFO("Device Type:        {}", getPhysicalDeviceType());
    JST_INFO("API Version:        {}", getApiVersion());
    JST_INFO("Compute Capability: {}", getComputeCapability());
    JST_INFO("SMs:                {}", info.sm_count);
    JST_INFO("Device Memory:      {:.2f} GB", static_cast<F32>(getPhysicalMemory()) / (1024 * 1024 * 1024));
    JST_INFO("Context:            primary (shared with the CUDA runtime)");
    JST_INFO("  - Can Import Host Memory:   {}", canImportHostMemory() ? "YES" : "NO");
    JST_INFO("-----------------------------------------------------");
}

CUDA::~CUDA() {
    if (_isAvailable) {
        cuDevicePrimaryCtxRelease(device);
    }
}

Result CUDA::activate() const {
    if (cudaSetDevice(static_cast<int>(config.deviceId)) != cudaSuccess || cuCtxSetCurrent(context) != CUDA_SUCCESS) {
        JST_ERROR("[CUDA:B200] Cannot activate device ID ({}).", config.deviceId);
        return Result::ERROR;
    }
    return Result::SUCCESS;
}

bool CUDA::isAvailable() const { return _isAvailable; }
std::string CUDA::getDeviceName() const { return cache.deviceName; }
std::string CUDA::getApiVersion() const { return cache.apiVersion; }
std::string CUDA::getComputeCapability() const { return cache.computeCapability; }
PhysicalDeviceType CUDA::getPhysicalDeviceType() const { return cache.physicalDeviceType; }
bool CUDA::hasUnifiedMemory() const { return cache.hasUnifiedMemory; }
bool CUDA::canExportDeviceMemory() const { return cache.canExportDeviceMemory; }
bool CUDA::canImportDeviceMemory() const { return cache.canImportDeviceMemory; }
bool CUDA::canImportHostMemory() const { return cache.canImportHostMemory; }
U64 CUDA::getPhysicalMemory() const { return cache.physicalMemory; }

}  // namespace Jetstream::Backend
