// Shared pieces of the reference-side binding for provider "b200": the per-device libb200dsp context, the
// Result/last-error bridge, raw device pointers of reference Tensors, and the Module::Config records of the
// module types this provider ADDS to the registry (fused replacements of whole per-cycle module chains):
//
//   spectral_chain  multiply(window) -> fft -> [agc] -> amplitude -> [range]   (spectrum_engine block, one kernel)
//   fir_filter      pad -> fft -> multiply -> fold -> ifft -> normalize -> phase_correction -> unpad -> overlap_add
//                   (filter block, one time-domain polyphase kernel)
// and implements for the reference's own types `lineplot` / `waterfall` (computeSubmit: batch sum + decimate +
// normalise + clamp + EMA; ring write), fed by the chain's fused column sums when the producer publishes them.
//
// The config records follow the reference's own declaration style (include/jetstream/module.hh JST_MODULE_TYPE /
// JST_MODULE_PARAMS) so flowgraph YAML, Parser::Map and block reconfigure work on them unchanged.
#ifndef B200_PROVIDER_HH
#define B200_PROVIDER_HH

#include <cstdint>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include <jetstream/logger.hh>
#include <jetstream/memory/tensor.hh>
#include <jetstream/module.hh>

#include "b200dsp.h"

namespace Jetstream::Modules {

struct SpectralChain : public Module::Config {
    bool enableScale = false;
    F32 rangeMin = -120.0f;
    F32 rangeMax = 0.0f;
    bool enableAgc = false;
    F64 agcReference = 1.0;        // the `agc` module's own defaults, include/jetstream/domains/dsp/agc/module.hh:9-14
    F64 agcEpsilon = 1e-12;
    F64 agcMinGain = 0.01;
    F64 agcMaxGain = 100.0;
    // The kernel's epilogue also accumulates sum-over-batch of every output column and the output tensor carries them
    // as attribute "b200.columnSums" (Tensor [n] F32): a `lineplot` on this provider downstream then skips its own
    // pass over the [batch, n] spectra (b200_chain_exec_colsum -> b200_lineplot_update_from_colsum).
    bool publishColumnSums = true;

    JST_MODULE_TYPE(spectral_chain);
    JST_MODULE_PARAMS(enableScale, rangeMin, rangeMax, enableAgc, agcReference, agcEpsilon, agcMinGain, agcMaxGain,
                      publishColumnSums);
};

struct FirFilter : public Module::Config {
    U64 decimation = 1;
    std::vector<F64> centerBins = {};   // per head, integer-valued bins of the block's (T + taps - 1)-point spectrum

    JST_MODULE_TYPE(fir_filter);
    JST_MODULE_PARAMS(decimation, centerBins);
};

}  // namespace Jetstream::Modules

namespace Jetstream::B200 {

inline constexpr const char* kColumnSumsAttribute = "b200.columnSums";

// One b200_ctx per CUDA device, keyed by the device that is current on the calling thread (the reference's CUDA
// backend activates its device before every create / compute, src/runtime/native/cuda/impl.cc:36,186).
inline b200_ctx* Ctx() {
    static std::mutex mutex;
    static std::unordered_map<int, b200_ctx*> contexts;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        JST_ERROR("[B200] cudaGetDevice failed.");
        return nullptr;
    }
    std::lock_guard<std::mutex> guard(mutex);
    const auto it = contexts.find(device);
    if (it != contexts.end()) {
        return it->second;
    }
    b200_ctx* created = nullptr;
    if (b200_ctx_create(device, &created) != B200_SUCCESS) {
        JST_ERROR("[B200] {}", b200_last_error());
        return nullptr;
    }
    contexts.emplace(device, created);
    return created;
}

inline Result Check(const int code, const char* module) {
    if (code == B200_SUCCESS) {
        return Result::SUCCESS;
    }
    JST_ERROR("[MODULE_{}_B200] {}", module, b200_last_error());
    return static_cast<Result>(code);
}

template<typename T>
inline T* DevicePtr(Tensor& tensor) {
    return reinterpret_cast<T*>(static_cast<std::uint8_t*>(tensor.buffer().data()) + tensor.offsetBytes());
}
template<typename T>
inline const T* DevicePtr(const Tensor& tensor) {
    return reinterpret_cast<const T*>(static_cast<const std::uint8_t*>(tensor.buffer().data()) +
                                      tensor.offsetBytes());
}

// b200dsp.h dtype code of a reference DataType (-1: not an integer / complex-integer type).
inline int IntegerDtypeCode(const DataType dtype) {
    switch (dtype) {
        case DataType::I8: return B200_DTYPE_I8;
        case DataType::U8: return B200_DTYPE_U8;
        case DataType::I16: return B200_DTYPE_I16;
        case DataType::U16: return B200_DTYPE_U16;
        case DataType::I32: return B200_DTYPE_I32;
        case DataType::U32: return B200_DTYPE_U32;
        case DataType::CI8: return B200_DTYPE_CI8;
        case DataType::CU8: return B200_DTYPE_CU8;
        case DataType::CI16: return B200_DTYPE_CI16;
        case DataType::CU16: return B200_DTYPE_CU16;
        case DataType::CI32: return B200_DTYPE_CI32;
        case DataType::CU32: return B200_DTYPE_CU32;
        default: return -1;
    }
}

}  // namespace Jetstream::B200

#endif  // B200_PROVIDER_HH
