// C-ABI driver over the reference's own Flowgraph / scheduler_synchronous / runtimes with BOTH providers linked in:
// the reference CPU modules ("generic") and the b200 binding (shim/b200_modules.cc, shim/b200_blocks.cc ->
// libb200dsp.so). It is how tests/ and bench.py run one and the same reference flowgraph on either target:
//
//   session = jst_shim_create(); add_source(..., device, provider); add_block(name, type, config, inputs, device,
//   provider); write_source(); compute(); output_read()
//
// Everything between the caller-filled source and the output read is reference code (block wiring, topological
// order, runtime segments, NativeCudaRuntime stream + per-cycle synchronise) plus, on the b200 target, the binding.
// This file is harness, not product: it holds no kernels and is not part of libb200dsp.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "jetstream/backend/base.hh"
#include "jetstream/block.hh"
#include "jetstream/detail/block_impl.hh"
#include "jetstream/detail/module_impl.hh"
#include "jetstream/flowgraph.hh"
#include "jetstream/flowgraph_view.hh"
#include "jetstream/logger.hh"
#include "jetstream/module.hh"
#include "jetstream/module_context.hh"
#include "jetstream/registry.hh"
#include "jetstream/runtime_context_native_cpu.hh"
#include "jetstream/runtime_context_native_cuda.hh"
#include "jetstream/scheduler_context.hh"

#include "domains/visualization/lineplot/module_impl.hh"
#include "domains/visualization/waterfall/module_impl.hh"

namespace Jetstream::Headless {       // shim/viz_headless.cc: lineplot / waterfall modules by name
std::map<std::string, Modules::LineplotImpl*>& Lineplots();
std::map<std::string, Modules::WaterfallImpl*>& Waterfalls();
}  // namespace Jetstream::Headless

namespace Jetstream {

namespace Modules {

struct ShimSource : public Module::Config {
    std::string shape = "1";      // comma separated dims
    std::string dataType = "CF32";
    I64 sampleAxis = -1;
    I64 batchAxis = -1;
    I64 channelAxis = -1;
    bool mapped = false;          // CUDA only: a CPU tensor mapped onto the device (what TestContext hands a CUDA module)

    JST_MODULE_TYPE(shim_source);
    JST_MODULE_PARAMS(shape, dataType, sampleAxis, batchAxis, channelAxis, mapped);
};

struct ShimSourceBase : public Module::Impl, public DynamicConfig<ShimSource> {
    Result define() override { return defineInterfaceOutput("signal"); }

    Result create() override {
        const DataType dtype = NameToDataType(dataType);
        if (dtype == DataType::None) {
            JST_ERROR("[SHIM_SOURCE] Unknown data type '{}'.", dataType);
            return Result::ERROR;
        }
        Shape dims;
        std::stringstream ss(shape);
        std::string item;
        while (std::getline(ss, item, ',')) {
            if (!item.empty()) {
                dims.push_back(static_cast<U64>(std::stoull(item)));
            }
        }
        if (mapped && device() == DeviceType::CUDA) {
            // src/testing.cc:133-136: Tensor(deviceType, cpuTensor) = zero-copy host mapping (buffer backend create(source))
            JST_CHECK(host.create(DeviceType::CPU, dtype, dims));
            signal = Tensor(DeviceType::CUDA, host);
            if (!signal.validShape() || signal.device() != DeviceType::CUDA) {
                JST_ERROR("[SHIM_SOURCE] Mapping the host tensor onto the device failed.");
                return Result::ERROR;
            }
        } else {
            JST_CHECK(signal.create(device(), dtype, dims));
        }
        if (sampleAxis >= 0) JST_CHECK(signal.setAttribute("sampleAxis", Index{static_cast<U64>(sampleAxis)}));
        if (batchAxis >= 0) JST_CHECK(signal.setAttribute("batchAxis", Index{static_cast<U64>(batchAxis)}));
        if (channelAxis >= 0) JST_CHECK(signal.setAttribute("channelAxis", Index{static_cast<U64>(channelAxis)}));
        outputs()["signal"].produced(name(), "signal", signal);
        return Result::SUCCESS;
    }

    Tensor signal;
    Tensor host;
};

struct ShimSourceCpu : public ShimSourceBase, public NativeCpuRuntimeContext, public Scheduler::Context {
    Result computeSubmit() override { return Result::SUCCESS; }             // the caller already wrote the bytes
};
struct ShimSourceCuda : public ShimSourceBase, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t&) override { return Result::SUCCESS; }
};
JST_REGISTER_MODULE(ShimSourceCpu, DeviceType::CPU, RuntimeType::NATIVE, "generic");
JST_REGISTER_MODULE(ShimSourceCuda, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

}  // namespace Modules

namespace Blocks {

struct ShimSource : public Block::Config {
    std::string shape = "1";
    std::string dataType = "CF32";
    I64 sampleAxis = -1;
    I64 batchAxis = -1;
    I64 channelAxis = -1;
    bool mapped = false;

    JST_BLOCK_TYPE(shim_source);
    JST_BLOCK_DOMAIN("Test");
    JST_BLOCK_PARAMS(shape, dataType, sampleAxis, batchAxis, channelAxis, mapped);
    JST_BLOCK_DESCRIPTION("Shim Source", "Caller-filled source tensor.", "Harness source block of the b200 shim.");
};

struct ShimSourceBlock : public Block::Impl, public DynamicConfig<Blocks::ShimSource> {
    Result configure() override {
        moduleConfig->shape = shape;
        moduleConfig->dataType = dataType;
        moduleConfig->sampleAxis = sampleAxis;
        moduleConfig->batchAxis = batchAxis;
        moduleConfig->channelAxis = channelAxis;
        moduleConfig->mapped = mapped;
        return Result::SUCCESS;
    }
    Result define() override { return defineInterfaceOutput("signal", "Output", "Caller-filled tensor."); }
    Result create() override {
        JST_CHECK(moduleCreate("source", moduleConfig, {}));
        return moduleExposeOutput("signal", {"source", "signal"});
    }
    std::shared_ptr<Modules::ShimSource> moduleConfig = std::make_shared<Modules::ShimSource>();
};
JST_REGISTER_BLOCK(ShimSourceBlock, {"shim_source"});

}  // namespace Blocks

}  // namespace Jetstream

using namespace Jetstream;

namespace {

struct Session {
    std::unique_ptr<Flowgraph> flowgraph;
    double lastComputeSeconds = 0.0;
};

thread_local std::string g_error;

int Fail(const std::string& what, const Result result) {
    std::ostringstream os;
    os << what << ": " << result << " | " << JST_LOG_LAST_ERROR();
    g_error = os.str();
    return static_cast<int>(result) == 0 ? -1 : static_cast<int>(result);
}

Parser::Map ParseKv(const char* text) {
    Parser::Map map;
    if (!text) {
        return map;
    }
    std::stringstream ss(text);
    std::string line;
    while (std::getline(ss, line)) {
        const auto eq = line.find('=');
        if (eq != std::string::npos) {
            map[line.substr(0, eq)] = line.substr(eq + 1);
        }
    }
    return map;
}

Result FindOutput(Session* s, const char* block, const char* port, Tensor& out) {
    TensorMap outputs;
    JST_CHECK(s->flowgraph->view().outputs(block, outputs));
    const auto it = outputs.find(port);
    if (it == outputs.end()) {
        JST_ERROR("[SHIM] Block '{}' has no output port '{}'.", block, port);
        return Result::ERROR;
    }
    out = it->second.tensor;
    return Result::SUCCESS;
}

I64 AxisOrMinusOne(const Tensor& tensor, const char* key) {
    if (!tensor.hasAttribute(key)) {
        return -1;
    }
    try {
        return static_cast<I64>(std::any_cast<Index>(tensor.attribute(key)));
    } catch (...) {
        return -2;
    }
}

// The reference's CUDA backend makes ITS driver context current on the calling thread (Backend::State<CUDA>::activate).
// A host process that also uses the primary context (PyTorch, libb200dsp through ctypes) must get its own context back
// when a harness call returns, or its next runtime-API call would run in the wrong context.
struct ContextRestore {
    CUcontext previous = nullptr;
    bool valid = false;
    ContextRestore() { valid = cuCtxGetCurrent(&previous) == CUDA_SUCCESS; }
    ~ContextRestore() {
        if (valid) {
            cuCtxSetCurrent(previous);
        }
    }
};

DeviceType Device(const int code) { return code == 1 ? DeviceType::CUDA : DeviceType::CPU; }

std::uint8_t* RawPointer(const Tensor& tensor) {
    return static_cast<std::uint8_t*>(const_cast<void*>(tensor.buffer().data())) + tensor.offsetBytes();
}

}  // namespace

extern "C" {

const char* jst_shim_last_error() { return g_error.c_str(); }

// Selects the CUDA device of the reference's backend singleton (Backend::Config::deviceId, default 0) — one process per
// GPU calls this with its LOCAL_RANK before the first session. A no-op when the backend already runs on that device.
int jst_shim_set_cuda_device(int device) {
    const ContextRestore restore;
    if (Backend::Initialized<DeviceType::CUDA>()) {
        if (static_cast<int>(Backend::State<DeviceType::CUDA>()->getDeviceId()) == device) {
            return 0;
        }
        g_error = "set_cuda_device: the CUDA backend is already initialised on another device";
        return -1;
    }
    Backend::Config config;
    config.deviceId = static_cast<U64>(device);
    const auto result = Backend::Initialize<DeviceType::CUDA>(config);
    return result == Result::SUCCESS ? 0 : Fail("set_cuda_device", result);
}

// "deviceId|name|computeCapability|apiVersion|memoryBytes|contextIsPrimary" of the CUDA backend singleton (initialising it
// on the default device when nobody has): Backend::CUDA is shim/b200_backend.cc in this library.
int jst_shim_backend_info(char* buffer, uint64_t capacity) {
    const ContextRestore restore;
    const auto& state = Backend::State<DeviceType::CUDA>();
    if (!state || !state->isAvailable()) {
        g_error = "backend_info: CUDA backend unavailable";
        return -1;
    }
    CUcontext primary = nullptr;
    unsigned flags = 0;
    int active = 0;
    cuDevicePrimaryCtxGetState(state->getDevice(), &flags, &active);
    cuDevicePrimaryCtxRetain(&primary, state->getDevice());
    const bool isPrimary = primary == state->getContext();
    cuDevicePrimaryCtxRelease(state->getDevice());
    std::ostringstream os;
    os << state->getDeviceId() << '|' << state->getDeviceName() << '|' << state->getComputeCapability() << '|'
       << state->getApiVersion() << '|' << state->getPhysicalMemory() << '|' << (isPrimary ? 1 : 0);
    const std::string text = os.str();
    if (buffer && capacity > 0) {
        const auto n = std::min<uint64_t>(capacity - 1, text.size());
        std::memcpy(buffer, text.data(), n);
        buffer[n] = 0;
    }
    return static_cast<int>(text.size());
}

void* jst_shim_create(int logLevel) {
    const ContextRestore restore;
    JST_LOG_SET_DEBUG_LEVEL(logLevel);
    auto* s = new Session();
    s->flowgraph = std::make_unique<Flowgraph>();
    const auto result = s->flowgraph->create({}, nullptr, nullptr, nullptr);
    if (result != Result::SUCCESS) {
        Fail("flowgraph create", result);
        delete s;
        return nullptr;
    }
    return s;
}

void jst_shim_destroy(void* handle) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    if (!s) {
        return;
    }
    std::vector<std::string> names;
    if (s->flowgraph->view().keys(names) == Result::SUCCESS) {
        for (auto it = names.rbegin(); it != names.rend(); ++it) {
            (void)s->flowgraph->blockDestroy(*it, false);
        }
    }
    (void)s->flowgraph->destroy();
    delete s;
}

// device: 0 CPU, 1 CUDA. dtype: the codes of include/b200dsp.h. axes: -1 = attribute absent.
int jst_shim_add_source(void* handle, const char* name, int dtype, int rank, const uint64_t* shape, int64_t sampleAxis,
                        int64_t batchAxis, int64_t channelAxis, int device, const char* provider, int mapped) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    static const char* const kNames[] = {"F32", "CF32", "I8", "U8", "I16", "U16", "I32", "U32",
                                         "CI8", "CU8", "CI16", "CU16", "CI32", "CU32"};
    if (dtype < 0 || dtype > 13) {
        g_error = "add_source: unknown dtype code";
        return -1;
    }
    std::string dims;
    for (int i = 0; i < rank; ++i) {
        dims += (i ? "," : "") + std::to_string(shape[i]);
    }
    Parser::Map config;
    config["shape"] = dims;
    config["dataType"] = std::string(kNames[dtype]);
    config["sampleAxis"] = std::to_string(sampleAxis);
    config["batchAxis"] = std::to_string(batchAxis);
    config["channelAxis"] = std::to_string(channelAxis);
    config["mapped"] = std::string(mapped ? "true" : "false");
    const auto result = s->flowgraph->blockCreate(name, "shim_source", config, {}, Device(device), RuntimeType::NATIVE,
                                                  provider);
    return result == Result::SUCCESS ? 0 : Fail("add_source", result);
}

// Host bytes -> the source tensor (memcpy on the CPU target, a synchronous H2D copy on the CUDA target).
int jst_shim_write_source(void* handle, const char* name, const void* data, uint64_t bytes) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, name, "signal", tensor);
    if (result != Result::SUCCESS) {
        return Fail("write_source", result);
    }
    if (bytes != tensor.size() * tensor.elementSize()) {
        g_error = "write_source: size mismatch";
        return -1;
    }
    if (tensor.device() == DeviceType::CUDA) {
        const auto err = cudaMemcpy(RawPointer(tensor), data, bytes, cudaMemcpyDefault);
        if (err != cudaSuccess) {
            g_error = std::string("write_source: ") + cudaGetErrorString(err);
            return -1;
        }
        return 0;
    }
    std::memcpy(RawPointer(tensor), data, bytes);
    return 0;
}

// config: "key=value\n..." (the strings a flowgraph YAML would carry). inputs: "port=block.port\n...".
int jst_shim_add_block(void* handle, const char* name, const char* type, const char* config, const char* inputs,
                       int device, const char* provider) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    TensorMap links;
    for (const auto& [port, value] : ParseKv(inputs)) {
        const std::string endpoint = std::any_cast<std::string>(value);
        const auto dot = endpoint.find('.');
        if (dot == std::string::npos) {
            g_error = "add_block: input must be block.port";
            return -1;
        }
        links[port].requested(endpoint.substr(0, dot), endpoint.substr(dot + 1));
    }
    const auto result = s->flowgraph->blockCreate(name, type, ParseKv(config), links, Device(device),
                                                  RuntimeType::NATIVE, provider);
    return result == Result::SUCCESS ? 0 : Fail(std::string("add_block ") + type, result);
}

int jst_shim_reconfigure(void* handle, const char* name, const char* config) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    const auto result = s->flowgraph->blockReconfigure(name, ParseKv(config));
    return result == Result::SUCCESS ? 0 : Fail("reconfigure", result);
}

int jst_shim_compute(void* handle) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    const auto t0 = std::chrono::steady_clock::now();
    const auto result = s->flowgraph->compute();
    const auto t1 = std::chrono::steady_clock::now();
    s->lastComputeSeconds = std::chrono::duration<double>(t1 - t0).count();
    return result == Result::SUCCESS ? 0 : Fail("compute", result);
}

double jst_shim_last_compute_seconds(void* handle) { return static_cast<Session*>(handle)->lastComputeSeconds; }

// info[0]=dtype (0 F32, 1 CF32, -1 other), info[1]=rank, info[2..9]=shape, info[10..12]=sample/batch/channel axis,
// info[13]=contiguous, info[14]=size (elements), info[15]=device (0 CPU, 1 CUDA)
int jst_shim_output_info(void* handle, const char* block, const char* port, int64_t* info) {
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, block, port, tensor);
    if (result != Result::SUCCESS) {
        return Fail("output_info", result);
    }
    info[0] = tensor.dtype() == DataType::F32 ? 0 : (tensor.dtype() == DataType::CF32 ? 1 : -1);
    info[1] = static_cast<int64_t>(tensor.rank());
    for (Index i = 0; i < 8; ++i) {
        info[2 + i] = i < tensor.rank() ? static_cast<int64_t>(tensor.shape(i)) : 0;
    }
    info[10] = AxisOrMinusOne(tensor, "sampleAxis");
    info[11] = AxisOrMinusOne(tensor, "batchAxis");
    info[12] = AxisOrMinusOne(tensor, "channelAxis");
    info[13] = tensor.contiguous() ? 1 : 0;
    info[14] = static_cast<int64_t>(tensor.size());
    info[15] = tensor.device() == DeviceType::CUDA ? 1 : 0;
    return 0;
}

// F32 attribute of an output tensor (e.g. "sampleRate"); returns 1 when absent or of another type.
int jst_shim_output_attribute_f32(void* handle, const char* block, const char* port, const char* key, float* value) {
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, block, port, tensor);
    if (result != Result::SUCCESS) {
        return Fail("output_attribute", result);
    }
    if (!tensor.hasAttribute(key)) {
        return 1;
    }
    const std::any attribute = tensor.attribute(key);
    if (const auto* f = std::any_cast<F32>(&attribute)) {
        *value = *f;
        return 0;
    }
    return 1;
}

// Raw address of an output tensor (a device pointer on the CUDA target): lets a caller that owns device data
// (bench.py) fill a source / read a result without a host round trip.
void* jst_shim_output_pointer(void* handle, const char* block, const char* port) {
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    if (FindOutput(s, block, port, tensor) != Result::SUCCESS || !tensor.contiguous()) {
        return nullptr;
    }
    return RawPointer(tensor);
}

int jst_shim_output_read(void* handle, const char* block, const char* port, void* dst, uint64_t bytes) {
    const ContextRestore restore;
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, block, port, tensor);
    if (result != Result::SUCCESS) {
        return Fail("output_read", result);
    }
    if (bytes != tensor.size() * tensor.elementSize() || !tensor.contiguous()) {
        g_error = "output_read: size mismatch or strided output";
        return -1;
    }
    if (tensor.device() == DeviceType::CUDA) {
        const auto err = cudaMemcpy(dst, RawPointer(tensor), bytes, cudaMemcpyDefault);
        if (err != cudaSuccess) {
            g_error = std::string("output_read: ") + cudaGetErrorString(err);
            return -1;
        }
        return 0;
    }
    std::memcpy(dst, RawPointer(tensor), bytes);
    return 0;
}

// "metricName cycles computeTimeMs\n" per module of a block (Module::Timing, include/jetstream/module.hh:25-31).
int jst_shim_metrics(void* handle, const char* block, char* buffer, uint64_t capacity) {
    auto* s = static_cast<Session*>(handle);
    std::vector<Flowgraph::View::MetricEntry> metrics;
    const auto result = s->flowgraph->view().metrics(block, metrics);
    if (result != Result::SUCCESS) {
        return -Fail("metrics", result);
    }
    std::ostringstream os;
    for (const auto& metric : metrics) {
        if (const auto* timing = std::any_cast<Module::Timing>(&metric.value)) {
            os << metric.name << ' ' << timing->cycles << ' ' << timing->computeTime << '\n';
        }
    }
    const std::string text = os.str();
    if (buffer && capacity > 0) {
        const auto n = std::min<uint64_t>(capacity - 1, text.size());
        std::memcpy(buffer, text.data(), n);
        buffer[n] = 0;
    }
    return static_cast<int>(text.size());
}

// ---- lineplot / waterfall state (the tensors the reference hands to its renderer), by MODULE name --------------------
// A block named "lp" of type lineplot holds the module "lp-lineplot" (Block::Impl::moduleCreate naming).

namespace {

struct LineplotPeek : Modules::LineplotImpl {
    static auto points() { return &LineplotPeek::signalPoints; }
};
struct WaterfallPeek : Modules::WaterfallImpl {
    static auto bins() { return &WaterfallPeek::frequencyBins; }
    static auto ring() { return &WaterfallPeek::ringState; }
};

int64_t CopyOut(const Tensor& tensor, float* dst, const uint64_t capacity) {
    const uint64_t n = std::min<uint64_t>(capacity, tensor.size());
    if (n != 0 && tensor.device() == DeviceType::CUDA) {
        if (cudaMemcpy(dst, RawPointer(tensor), n * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
            g_error = "viz read: device copy failed";
            return -1;
        }
    } else if (n != 0) {
        std::memcpy(dst, RawPointer(tensor), n * sizeof(float));
    }
    return static_cast<int64_t>(tensor.size());
}

}  // namespace

// Lists "lineplot:<name>\n" / "waterfall:<name>\n" for every live module (to discover the naming).
int jst_shim_viz_list(char* buffer, uint64_t capacity) {
    std::string text;
    for (const auto& [name, _] : Headless::Lineplots()) text += "lineplot:" + name + "\n";
    for (const auto& [name, _] : Headless::Waterfalls()) text += "waterfall:" + name + "\n";
    if (buffer && capacity > 0) {
        const auto n = std::min<uint64_t>(capacity - 1, text.size());
        std::memcpy(buffer, text.data(), n);
        buffer[n] = 0;
    }
    return static_cast<int>(text.size());
}

// signalPoints [n, 2] of a lineplot module / frequencyBins [height, n] of a waterfall module. Returns the element count.
int64_t jst_shim_viz_read(const char* module, float* dst, uint64_t capacity) {
    const ContextRestore restore;
    if (const auto it = Headless::Lineplots().find(module); it != Headless::Lineplots().end()) {
        return CopyOut(it->second->*LineplotPeek::points(), dst, capacity);
    }
    if (const auto it = Headless::Waterfalls().find(module); it != Headless::Waterfalls().end()) {
        return CopyOut(it->second->*WaterfallPeek::bins(), dst, capacity);
    }
    g_error = std::string("viz_read: no lineplot / waterfall module named '") + module + "'";
    return -1;
}

int64_t jst_shim_viz_write_index(const char* module) {
    const auto it = Headless::Waterfalls().find(module);
    return it == Headless::Waterfalls().end() ? -1 : static_cast<int64_t>((it->second->*WaterfallPeek::ring()).writeIndex);
}

}  // extern "C"
