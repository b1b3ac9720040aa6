// SURVEY.md §8 row a14 — the CUDA Tensor allocation of the reference, re-implemented behind the reference's own
// interface: detail::Backend (src/memory/buffer_backend.hh:12-30) + CudaBufferBackend
// (include/jetstream/memory/devices/cuda/buffer.hh) over libb200dsp's b200_malloc / b200_malloc_managed /
// b200_host_register / b200_memcpy. shim/build_shim.sh links THIS translation unit in place of the reference's
// src/memory/buffer_cuda.cc (both define detail::CreateCudaBackend()), so every CUDA tensor of a flowgraph — module
// outputs, the lineplot / waterfall state, TestContext's mapped host inputs — is allocated through the C ABI.
//
// Behaviour kept (file:line of the reference): zero-filled allocations (:118-121), managed memory when the caller asks
// for host access (:49-58), zero-copy mapping of page-aligned CPU buffers with ownership of the pinning (:140-200),
// borrowed raw pointers rejected (:126-131), copyFrom on the caller's stream or synchronously (:284-306).
// Dropped: the exportable VMM path (cuMemCreate with a POSIX fd, :59-107) — it exists for CUDA -> Vulkan render
// interop, which is out of scope; exportableDeviceMemory() answers false.
#include <cstdint>
#include <memory>

#include "jetstream/backend/base.hh"
#include "jetstream/logger.hh"
#include "jetstream/memory/devices/cuda/buffer.hh"
#include "jetstream/memory/macros.hh"

#include "memory/buffer_backend.hh"

#include "b200_provider.hh"

namespace Jetstream::detail {

namespace {

class B200CudaBackend final : public CudaBufferBackend, public Backend {
 public:
    ~B200CudaBackend() override { destroy(); }

    DeviceType device() const override { return DeviceType::CUDA; }

    Result create(const U64& bytes, const Buffer::Config& config) override {
        destroy();
        sizeBytes = bytes;
        if (bytes == 0) {
            return Result::SUCCESS;
        }
        if (!Jetstream::Backend::State<DeviceType::CUDA>()->isAvailable()) {
            JST_ERROR("[MEMORY:BUFFER:B200] CUDA is not available.");
            return Result::ERROR;
        }
        U64 rounded = 0;
        if (!CheckedPageAlignedSize(bytes, rounded)) {
            JST_ERROR("[MEMORY:BUFFER:B200] Allocation size {} is too large.", bytes);
            return Result::ERROR;
        }
        const int code = config.hostAccessible ? b200_malloc_managed(B200::Ctx(), rounded, &buffer)
                                               : b200_malloc(B200::Ctx(), rounded, &buffer);
        if (code != B200_SUCCESS) {
            JST_ERROR("[MEMORY:BUFFER:B200] {}", b200_last_error());
            sizeBytes = 0;
            return Result::ERROR;
        }
        kind = config.hostAccessible ? Kind::Managed : Kind::Device;
        return Result::SUCCESS;
    }

    Result create(void*, const U64&) override {
        JST_ERROR("[MEMORY:BUFFER:B200] Borrowed raw-pointer create is not supported.");
        return Result::ERROR;
    }

    Result create(const Backend& source) override {
        destroy();
        if (source.size() == 0) {
            return Result::SUCCESS;
        }
        if (source.device() != DeviceType::CPU) {
            JST_ERROR("[MEMORY:BUFFER:B200] Cannot mirror from device {}.", source.device());
            return Result::ERROR;
        }
        void* host = const_cast<void*>(source.rawHandle());
        U64 rounded = 0;
        if (!host || !JST_IS_ALIGNED(host) || !CheckedPageAlignedSize(source.size(), rounded)) {
            JST_ERROR("[MEMORY:BUFFER:B200] CPU source must be a page-aligned, importable allocation.");
            return Result::ERROR;
        }
        int registered = 0;
        if (b200_host_register(B200::Ctx(), host, rounded, &registered) != B200_SUCCESS) {
            JST_ERROR("[MEMORY:BUFFER:B200] {}", b200_last_error());
            return Result::ERROR;
        }
        buffer = host;
        sizeBytes = source.size();
        ownsRegistration = registered != 0;
        kind = Kind::MappedHost;
        mappedLocation = source.location();
        return Result::SUCCESS;
    }

    Result copyFrom(const Backend& source, void* context) override {
        if (sizeBytes == 0) {
            return Result::SUCCESS;
        }
        if (b200_memcpy(B200::Ctx(), buffer, source.rawHandle(), source.size(), 3, context) != B200_SUCCESS ||
            (!context && b200_stream_synchronize(B200::Ctx(), nullptr) != B200_SUCCESS)) {
            JST_ERROR("[MEMORY:BUFFER:B200] {}", b200_last_error());
            return Result::ERROR;
        }
        return Result::SUCCESS;
    }

    void destroy() override {
        if (buffer && kind == Kind::MappedHost && ownsRegistration) {
            if (b200_host_unregister(B200::Ctx(), buffer) != B200_SUCCESS) {
                JST_WARN("[MEMORY:BUFFER:B200] {}", b200_last_error());
            }
        } else if (buffer && (kind == Kind::Device || kind == Kind::Managed)) {
            b200_free(B200::Ctx(), buffer);
        }
        buffer = nullptr;
        sizeBytes = 0;
        ownsRegistration = false;
        kind = Kind::None;
        mappedLocation = Location::None;
    }

    void* rawHandle() override { return buffer; }
    const void* rawHandle() const override { return buffer; }
    bool isBorrowed() const override { return kind == Kind::MappedHost; }
    Location location() const override {
        switch (kind) {
            case Kind::Device: return Location::Device;
            case Kind::Managed: return Location::Unified;
            case Kind::MappedHost: return mappedLocation;
            default: return Location::None;
        }
    }
    U64 size() const override { return sizeBytes; }

    bool hostAccessible() const override { return kind == Kind::Managed || kind == Kind::MappedHost; }
    bool deviceNative() const override { return kind == Kind::Device || kind == Kind::Managed; }
    bool exportableDeviceMemory() const override { return false; }
    CUmemGenericAllocationHandle allocationHandle() const override { return 0; }

 private:
    enum class Kind { None, Device, Managed, MappedHost };
    void* buffer = nullptr;
    U64 sizeBytes = 0;
    bool ownsRegistration = false;
    Kind kind = Kind::None;
    Location mappedLocation = Location::None;
};

}  // namespace

std::unique_ptr<Backend> CreateCudaBackend() { return std::make_unique<B200CudaBackend>(); }

}  // namespace Jetstream::detail
