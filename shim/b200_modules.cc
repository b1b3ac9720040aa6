// Reference-side binding for the "b200" provider: Module::Impl subclasses that derive the reference's own
// shared module implementations (validate/define/create/reconfigure are the reference's code, unchanged) and
// forward computeSubmit(stream) to libb200dsp through the C ABI (include/b200dsp.h).
//
// This is the file a CyberEther maintainer would add (e.g. as src/domains/b200/modules.cc, or as a `.cep`
// plugin target with `device: cuda`, docs/plugins.md:89-112). It is compiled here against the reference
// headers where they lie (shim/build_shim.sh) to prove the boundary; it contains no kernels.
//
//   registration     JST_REGISTER_MODULE(Impl, DeviceType::CUDA, RuntimeType::NATIVE, "b200")
//                                                                     include/jetstream/registry.hh:174-175
//   compute hooks    NativeCudaRuntimeContext::{computeInitialize, computeSubmit(const cudaStream_t&),
//                    computeDeinitialize}                              include/jetstream/runtime_context_native_cuda.hh:13-39
//   selection        blockCreate(name, type, config, inputs, DeviceType::CUDA, RuntimeType::NATIVE, "b200")
//                    or `device: cuda / runtime: native / provider: b200` in a flowgraph YAML
#include <cstdint>
#include <vector>

#include <jetstream/backend/base.hh>
#include <jetstream/logger.hh>
#include <jetstream/module_context.hh>
#include <jetstream/registry.hh>
#include <jetstream/runtime_context_native_cuda.hh>
#include <jetstream/scheduler_context.hh>

#include "domains/core/cast/module_impl.hh"
#include "domains/core/multiply/module_impl.hh"
#include "domains/core/range/module_impl.hh"
#include "domains/core/reshape/module_impl.hh"
#include "domains/dsp/agc/module_impl.hh"
#include "domains/dsp/amplitude/module_impl.hh"
#include "domains/dsp/fft/module_impl.hh"
#include "domains/dsp/invert/module_impl.hh"
#include "domains/dsp/window/module_impl.hh"

#include "b200dsp.h"

namespace Jetstream::Modules {

namespace {

// One b200_ctx for the device the reference's CUDA backend singleton selected.
b200_ctx* B200Ctx() {
    static b200_ctx* ctx = [] {
        b200_ctx* created = nullptr;
        int device = 0;
        cudaGetDevice(&device);
        if (b200_ctx_create(device, &created) != B200_SUCCESS) {
            JST_ERROR("[B200] {}", b200_last_error());
        }
        return created;
    }();
    return ctx;
}

Result Check(const int code, const char* module) {
    if (code == B200_SUCCESS) {
        return Result::SUCCESS;
    }
    JST_ERROR("[MODULE_{}_B200] {}", module, b200_last_error());
    return static_cast<Result>(code);
}

template<typename T>
T* DevicePtr(Tensor& tensor) {
    return reinterpret_cast<T*>(static_cast<std::uint8_t*>(tensor.buffer().data()) + tensor.offsetBytes());
}
template<typename T>
const T* DevicePtr(const Tensor& tensor) {
    return reinterpret_cast<const T*>(static_cast<const std::uint8_t*>(tensor.buffer().data()) +
                                      tensor.offsetBytes());
}

}  // namespace

// ---- window ---------------------------------------------------------------------------------------
struct WindowImplB200 : public WindowImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t& stream) override {
        return Check(b200_window_blackman_cf32(B200Ctx(), DevicePtr<b200_cf32>(output), size, stream), "WINDOW");
    }
};
JST_REGISTER_MODULE(WindowImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- invert ---------------------------------------------------------------------------------------
struct InvertImplB200 : public InvertImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result create() final {
        JST_CHECK(InvertImpl::create());
        if (!input.contiguous() || input.dtype() != DataType::CF32) {
            JST_ERROR("[MODULE_INVERT_B200] Only contiguous CF32 inputs are supported by this provider.");
            return Result::ERROR;
        }
        outer = 1;
        inner = 1;
        for (Index i = 0; i < resolvedAxis; ++i) outer *= input.shape(i);
        for (Index i = resolvedAxis + 1; i < input.rank(); ++i) inner *= input.shape(i);
        return Result::SUCCESS;
    }
    Result computeSubmit(const cudaStream_t& stream) override {
        return Check(b200_invert_cf32(B200Ctx(), DevicePtr<b200_cf32>(input), DevicePtr<b200_cf32>(output), outer,
                                      input.shape(resolvedAxis), inner, stream), "INVERT");
    }
    U64 outer = 1, inner = 1;
};
JST_REGISTER_MODULE(InvertImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- reshape / cast (views; cast bypasses when the dtype already matches) ----------------------------
struct ReshapeImplB200 : public ReshapeImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t&) override { return Result::SUCCESS; }
};
JST_REGISTER_MODULE(ReshapeImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

struct CastImplB200 : public CastImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t& stream) override {
        if (bypass) {
            return Result::SUCCESS;
        }
        if (input.dtype() == DataType::F32 && outputDtype == DataType::CF32 && input.contiguous()) {
            return Check(b200_cast_f32_cf32(B200Ctx(), DevicePtr<float>(input), DevicePtr<b200_cf32>(output),
                                            input.size(), stream), "CAST");
        }
        // integer -> F32 and complex integer -> CF32 (module_impl_native_cpu.cc:163-330); dtype codes of b200dsp.h
        int code = -1;
        switch (input.dtype()) {
            case DataType::I8: code = B200_DTYPE_I8; break;
            case DataType::U8: code = B200_DTYPE_U8; break;
            case DataType::I16: code = B200_DTYPE_I16; break;
            case DataType::U16: code = B200_DTYPE_U16; break;
            case DataType::I32: code = B200_DTYPE_I32; break;
            case DataType::U32: code = B200_DTYPE_U32; break;
            case DataType::CI8: code = B200_DTYPE_CI8; break;
            case DataType::CU8: code = B200_DTYPE_CU8; break;
            case DataType::CI16: code = B200_DTYPE_CI16; break;
            case DataType::CU16: code = B200_DTYPE_CU16; break;
            case DataType::CI32: code = B200_DTYPE_CI32; break;
            case DataType::CU32: code = B200_DTYPE_CU32; break;
            default: break;
        }
        const bool complexIn = code >= B200_DTYPE_CI8;
        if (code < 0 || !input.contiguous() || outputDtype != (complexIn ? DataType::CF32 : DataType::F32)) {
            JST_ERROR("[MODULE_CAST_B200] Unsupported conversion '{}' -> '{}' (contiguous inputs only).",
                      input.dtype(), outputDtype);
            return Result::ERROR;
        }
        return Check(b200_cast_int(B200Ctx(), DevicePtr<void>(input), code, DevicePtr<void>(output), input.size(),
                                   stream), "CAST");
    }
};
JST_REGISTER_MODULE(CastImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- multiply ---------------------------------------------------------------------------------------
struct MultiplyImplB200 : public MultiplyImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result create() final {
        JST_CHECK(MultiplyImpl::create());
        const Index rank = c.rank();
        if (rank > 8) {
            JST_ERROR("[MODULE_MULTIPLY_B200] Rank {} exceeds the supported 8.", rank);
            return Result::ERROR;
        }
        shape.assign(c.shape().begin(), c.shape().end());
        strideA.assign(a.stride().begin(), a.stride().end());   // broadcast views: stride 0 on broadcast dims
        strideB.assign(b.stride().begin(), b.stride().end());
        return Result::SUCCESS;
    }
    Result computeSubmit(const cudaStream_t& stream) override {
        const int rank = static_cast<int>(shape.size());
        if (a.dtype() == DataType::CF32) {
            return Check(b200_multiply_cf32(B200Ctx(), DevicePtr<b200_cf32>(a), DevicePtr<b200_cf32>(b),
                                            DevicePtr<b200_cf32>(c), rank, shape.data(), strideA.data(),
                                            strideB.data(), stream), "MULTIPLY");
        }
        return Check(b200_multiply_f32(B200Ctx(), DevicePtr<float>(a), DevicePtr<float>(b), DevicePtr<float>(c), rank,
                                       shape.data(), strideA.data(), strideB.data(), stream), "MULTIPLY");
    }
    std::vector<U64> shape, strideA, strideB;
};
JST_REGISTER_MODULE(MultiplyImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- fft ----------------------------------------------------------------------------------------------
struct FftImplB200 : public FftImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result create() final {
        JST_CHECK(FftImpl::create());
        if (input.dtype() != DataType::CF32 || !input.contiguous() || resolvedAxis + 1 != input.rank()) {
            JST_ERROR("[MODULE_FFT_B200] This provider implements contiguous CF32 transforms along the innermost axis.");
            return Result::ERROR;
        }
        return Result::SUCCESS;
    }
    Result computeInitialize() override {
        const U64 n = input.shape(resolvedAxis);
        return Check(b200_fft_plan_c2c(B200Ctx(), n, input.size() / n, &plan), "FFT");
    }
    Result computeSubmit(const cudaStream_t& stream) override {
        return Check(b200_fft_exec(plan, DevicePtr<b200_cf32>(input), DevicePtr<b200_cf32>(output), forward ? 1 : 0,
                                   stream), "FFT");
    }
    Result computeDeinitialize() override {
        const auto result = Check(b200_fft_plan_destroy(plan), "FFT");
        plan = nullptr;
        return result;
    }
    b200_fft_plan* plan = nullptr;
};
JST_REGISTER_MODULE(FftImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- agc ---------------------------------------------------------------------------------------------------
struct AgcImplB200 : public AgcImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result create() final {
        JST_CHECK(AgcImpl::create());
        if (!input.contiguous() || sampleAxis + 1 != input.rank() ||
            (input.dtype() != DataType::CF32 && input.dtype() != DataType::F32)) {
            JST_ERROR("[MODULE_AGC_B200] This provider implements contiguous F32 / CF32 inputs with the sample axis innermost.");
            return Result::ERROR;
        }
        return Result::SUCCESS;
    }
    Result computeSubmit(const cudaStream_t& stream) override {
        const U64 samples = input.shape(sampleAxis);
        uint64_t need = 0;
        JST_CHECK(Check(b200_agc_scratch_bytes(laneCount, samples, tileSize, &need), "AGC"));
        if (need > scratchBytes) {                       // tileSize may be reconfigured between cycles
            b200_free(B200Ctx(), scratch);
            scratch = nullptr;
            scratchBytes = 0;
            JST_CHECK(Check(b200_malloc(B200Ctx(), need, &scratch), "AGC"));
            scratchBytes = need;
        }
        return Check(b200_agc(B200Ctx(), DevicePtr<void>(input), DevicePtr<void>(output),
                              input.dtype() == DataType::CF32 ? 1 : 0, laneCount, samples, tileSize, reference, epsilon,
                              minGain, maxGain, maxGainChange, scratch, stream), "AGC");
    }
    Result computeDeinitialize() override {
        b200_free(B200Ctx(), scratch);
        scratch = nullptr;
        scratchBytes = 0;
        return Result::SUCCESS;
    }
    void* scratch = nullptr;
    uint64_t scratchBytes = 0;
};
JST_REGISTER_MODULE(AgcImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- amplitude -------------------------------------------------------------------------------------------
struct AmplitudeImplB200 : public AmplitudeImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t& stream) override {
        if (!input.contiguous()) {
            JST_ERROR("[MODULE_AMPLITUDE_B200] Strided inputs are not supported by this provider yet.");
            return Result::ERROR;
        }
        if (input.dtype() == DataType::CF32) {
            return Check(b200_amplitude_cf32(B200Ctx(), DevicePtr<b200_cf32>(input), DevicePtr<float>(output),
                                             input.size(), scalingCoeff, stream), "AMPLITUDE");
        }
        return Check(b200_amplitude_f32(B200Ctx(), DevicePtr<float>(input), DevicePtr<float>(output), input.size(),
                                        scalingCoeff, stream), "AMPLITUDE");
    }
};
JST_REGISTER_MODULE(AmplitudeImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

// ---- range -------------------------------------------------------------------------------------------------
struct RangeImplB200 : public RangeImpl, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t& stream) override {
        if (!input.contiguous() || input.dtype() != DataType::F32) {
            JST_ERROR("[MODULE_RANGE_B200] Only contiguous F32 inputs are supported by this provider.");
            return Result::ERROR;
        }
        return Check(b200_range_f32(B200Ctx(), DevicePtr<float>(input), DevicePtr<float>(output), input.size(),
                                    scalingCoeff, offsetCoeff, stream), "RANGE");
    }
};
JST_REGISTER_MODULE(RangeImplB200, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

}  // namespace Jetstream::Modules
