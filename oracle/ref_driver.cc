// TEST INFRASTRUCTURE — C-ABI driver over the UNMODIFIED reference (CyberEther/Jetstream
// v1.9.1) CPU compute path. Compiled together with the reference's own sources by
// oracle/build_ref.sh into oracle/_ref/libjst_ref.so. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load that library.
//
// The driver is our code written against the reference's PUBLIC plugin API:
//   * Flowgraph::{create,blockCreate,compute,view}       include/jetstream/flowgraph.hh:40-110
//   * Module::Impl / Block::Impl + JST_REGISTER_{MODULE,BLOCK}  include/jetstream/registry.hh:138-205
// It registers one extra block/module pair, "oracle_source": a non-static source whose output
// tensor the caller fills before each compute cycle (the reference's own sources are either
// STATIC_OUTPUT (ones_tensor) or seeded from std::random_device (signal_generator), so neither
// can deliver caller-chosen bytes every cycle).
//
// Everything downstream of that source is reference code: block wiring (block_impl.cc),
// scheduler_synchronous, NativeCpuRuntime and each module's computeSubmit().

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <unordered_set>
#include <string>
#include <vector>

#include "jetstream/block.hh"
#include "jetstream/detail/block_impl.hh"
#include "jetstream/detail/module_impl.hh"
#include "jetstream/flowgraph.hh"
#include "jetstream/flowgraph_view.hh"
#include "jetstream/logger.hh"
#include "jetstream/module.hh"
#include "jetstream/module_context.hh"
#include "jetstream/registry.hh"
#include "jetstream/runtime.hh"
#include "jetstream/runtime_context_native_cpu.hh"
#include "jetstream/scheduler_context.hh"

#include "domains/visualization/lineplot/module_impl.hh"
#include "domains/visualization/waterfall/module_impl.hh"

namespace Jetstream {

namespace Modules {

struct OracleSource : public Module::Config {
    std::string shape = "1";      // comma separated dims
    std::string dataType = "CF32";
    I64 sampleAxis = -1;
    I64 batchAxis = -1;
    I64 channelAxis = -1;

    JST_MODULE_TYPE(oracle_source);
    JST_MODULE_PARAMS(shape, dataType, sampleAxis, batchAxis, channelAxis);
};

static Shape ParseDims(const std::string& text) {
    Shape dims;
    std::stringstream ss(text);
    std::string item;
    while (std::getline(ss, item, ',')) {
        if (!item.empty()) {
            dims.push_back(static_cast<U64>(std::stoull(item)));
        }
    }
    return dims;
}

struct OracleSourceImpl : public Module::Impl,
                          public DynamicConfig<OracleSource>,
                          public NativeCpuRuntimeContext,
                          public Scheduler::Context {
    Result define() override {
        return defineInterfaceOutput("signal");
    }

    Result create() override {
        const DataType dtype = NameToDataType(dataType);
        if (dtype == DataType::None) {
            return Result::ERROR;
        }
        JST_CHECK(signal.create(device(), dtype, ParseDims(shape)));
        if (sampleAxis >= 0) {
            JST_CHECK(signal.setAttribute("sampleAxis", Index{static_cast<U64>(sampleAxis)}));
        }
        if (batchAxis >= 0) {
            JST_CHECK(signal.setAttribute("batchAxis", Index{static_cast<U64>(batchAxis)}));
        }
        if (channelAxis >= 0) {
            JST_CHECK(signal.setAttribute("channelAxis", Index{static_cast<U64>(channelAxis)}));
        }
        outputs()["signal"].produced(name(), "signal", signal);
        return Result::SUCCESS;
    }

    Result computeSubmit() override {
        return Result::SUCCESS;  // caller has already written the bytes
    }

    Tensor signal;
};

JST_REGISTER_MODULE(OracleSourceImpl, DeviceType::CPU, RuntimeType::NATIVE, "generic");

}  // namespace Modules

namespace Blocks {

struct OracleSource : public Block::Config {
    std::string shape = "1";
    std::string dataType = "CF32";
    I64 sampleAxis = -1;
    I64 batchAxis = -1;
    I64 channelAxis = -1;

    JST_BLOCK_TYPE(oracle_source);
    JST_BLOCK_DOMAIN("Test");
    JST_BLOCK_PARAMS(shape, dataType, sampleAxis, batchAxis, channelAxis);
    JST_BLOCK_DESCRIPTION("Oracle Source", "Caller-filled source tensor.",
                          "Parity-oracle source block (test infrastructure).");
};

struct OracleSourceBlockImpl : public Block::Impl, public DynamicConfig<Blocks::OracleSource> {
    Result configure() override {
        moduleConfig->shape = shape;
        moduleConfig->dataType = dataType;
        moduleConfig->sampleAxis = sampleAxis;
        moduleConfig->batchAxis = batchAxis;
        moduleConfig->channelAxis = channelAxis;
        return Result::SUCCESS;
    }

    Result define() override {
        return defineInterfaceOutput("signal", "Output", "Caller-filled tensor.");
    }

    Result create() override {
        JST_CHECK(moduleCreate("source", moduleConfig, {}));
        return moduleExposeOutput("signal", {"source", "signal"});
    }

    std::shared_ptr<Modules::OracleSource> moduleConfig = std::make_shared<Modules::OracleSource>();
};

JST_REGISTER_BLOCK(OracleSourceBlockImpl, {"oracle_source"});

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
        if (eq == std::string::npos) {
            continue;
        }
        map[line.substr(0, eq)] = line.substr(eq + 1);
    }
    return map;
}

Result FindOutput(Session* s, const char* block, const char* port, Tensor& out) {
    TensorMap outputs;
    JST_CHECK(s->flowgraph->view().outputs(block, outputs));
    const auto it = outputs.find(port);
    if (it == outputs.end()) {
        JST_ERROR("[ORACLE] Block '{}' has no output port '{}'.", block, port);
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

}  // namespace

extern "C" {

const char* jst_ref_last_error() { return g_error.c_str(); }

const char* jst_ref_version() { return JETSTREAM_VERSION_STR; }

void* jst_ref_create(int logLevel) {
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

void jst_ref_destroy(void* handle) {
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

// dtype: the codes of include/b200dsp.h (0 F32, 1 CF32, 2 I8, 3 U8, 4 I16, 5 U16, 6 I32, 7 U32, 8 CI8 ... 13 CU32).
// axes: -1 = attribute absent.
int jst_ref_add_source(void* handle, const char* name, int dtype, int rank, const uint64_t* shape,
                       int64_t sampleAxis, int64_t batchAxis, int64_t channelAxis) {
    auto* s = static_cast<Session*>(handle);
    std::string dims;
    for (int i = 0; i < rank; ++i) {
        dims += (i ? "," : "") + std::to_string(shape[i]);
    }
    Parser::Map config;
    config["shape"] = dims;
    static const char* const kNames[] = {"F32", "CF32", "I8", "U8", "I16", "U16", "I32", "U32",
                                         "CI8", "CU8", "CI16", "CU16", "CI32", "CU32"};
    if (dtype < 0 || dtype > 13) {
        g_error = "add_source: unknown dtype code";
        return -1;
    }
    config["dataType"] = std::string(kNames[dtype]);
    config["sampleAxis"] = std::to_string(sampleAxis);
    config["batchAxis"] = std::to_string(batchAxis);
    config["channelAxis"] = std::to_string(channelAxis);
    const auto result = s->flowgraph->blockCreate(name, "oracle_source", config, {});
    return result == Result::SUCCESS ? 0 : Fail("add_source", result);
}

int jst_ref_write_source(void* handle, const char* name, const void* data, uint64_t bytes) {
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, name, "signal", tensor);
    if (result != Result::SUCCESS) {
        return Fail("write_source", result);
    }
    if (bytes != tensor.sizeBytes()) {
        g_error = "write_source: size mismatch";
        return -1;
    }
    std::memcpy(tensor.data(), data, bytes);
    return 0;
}

// config: "key=value\n..." (values are the same strings a flowgraph YAML would carry).
// inputs: "port=block.port\n...".
int jst_ref_add_block(void* handle, const char* name, const char* type, const char* config,
                      const char* inputs) {
    auto* s = static_cast<Session*>(handle);
    TensorMap links;
    const Parser::Map wiring = ParseKv(inputs);
    for (const auto& [port, value] : wiring) {
        const std::string endpoint = std::any_cast<std::string>(value);
        const auto dot = endpoint.find('.');
        if (dot == std::string::npos) {
            g_error = "add_block: input must be block.port";
            return -1;
        }
        links[port].requested(endpoint.substr(0, dot), endpoint.substr(dot + 1));
    }
    const auto result = s->flowgraph->blockCreate(name, type, ParseKv(config), links);
    return result == Result::SUCCESS ? 0 : Fail(std::string("add_block ") + type, result);
}

int jst_ref_reconfigure(void* handle, const char* name, const char* config) {
    auto* s = static_cast<Session*>(handle);
    const auto result = s->flowgraph->blockReconfigure(name, ParseKv(config));
    return result == Result::SUCCESS ? 0 : Fail("reconfigure", result);
}

int jst_ref_compute(void* handle) {
    auto* s = static_cast<Session*>(handle);
    const auto t0 = std::chrono::steady_clock::now();
    const auto result = s->flowgraph->compute();
    const auto t1 = std::chrono::steady_clock::now();
    s->lastComputeSeconds = std::chrono::duration<double>(t1 - t0).count();
    return result == Result::SUCCESS ? 0 : Fail("compute", result);
}

double jst_ref_last_compute_seconds(void* handle) {
    return static_cast<Session*>(handle)->lastComputeSeconds;
}

// info[0]=dtype (0 F32, 1 CF32, -1 other), info[1]=rank, info[2..9]=shape, info[10..12]=sample/batch/channel axis,
// info[13]=contiguous, info[14]=size (elements)
int jst_ref_output_info(void* handle, const char* block, const char* port, int64_t* info) {
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
    return 0;
}

int jst_ref_output_read(void* handle, const char* block, const char* port, void* dst, uint64_t bytes) {
    auto* s = static_cast<Session*>(handle);
    Tensor tensor;
    const auto result = FindOutput(s, block, port, tensor);
    if (result != Result::SUCCESS) {
        return Fail("output_read", result);
    }
    if (bytes != tensor.size() * tensor.elementSize()) {
        g_error = "output_read: size mismatch";
        return -1;
    }
    if (tensor.contiguous()) {
        std::memcpy(dst, static_cast<const uint8_t*>(tensor.buffer().data()) + tensor.offsetBytes(), bytes);
        return 0;
    }
    // Strided view: gather row-major.
    const Index rank = tensor.rank();
    std::vector<U64> coord(rank, 0);
    const U64 es = tensor.elementSize();
    const auto* base = static_cast<const uint8_t*>(tensor.buffer().data());
    auto* out = static_cast<uint8_t*>(dst);
    for (U64 i = 0; i < tensor.size(); ++i) {
        U64 off = tensor.offset();
        for (Index d = 0; d < rank; ++d) {
            off += coord[d] * tensor.stride(d);
        }
        std::memcpy(out + i * es, base + off * es, es);
        for (Index d = rank; d-- > 0;) {
            if (++coord[d] < tensor.shape(d)) {
                break;
            }
            coord[d] = 0;
        }
    }
    return 0;
}

// Writes "metricName cycles computeTimeMs\n" lines for one block (Module::Timing metrics,
// include/jetstream/module.hh:25-31). Returns bytes needed.
int jst_ref_metrics(void* handle, const char* block, char* buffer, uint64_t capacity) {
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

// ---- lineplot / waterfall: the reference modules driven directly through Registry::BuildModule +
// Runtime, the way the reference's own module_tests.cc do (lineplot/module_tests.cc:63-94), with their internal
// state read through Module::getImpl<> (module_tests.cc:25-46). Compute is the reference's module_impl_native_cpu.cc;
// the non-compute halves are shim/viz_headless.cc.

namespace {

struct LineplotPeek : Modules::LineplotImpl {
    static auto points() { return &LineplotPeek::signalPoints; }
};
struct WaterfallPeek : Modules::WaterfallImpl {
    static auto bins() { return &WaterfallPeek::frequencyBins; }
    static auto ring() { return &WaterfallPeek::ringState; }
};

struct VizSession {
    std::string type;
    Tensor input;
    std::shared_ptr<Module> module;
    std::unique_ptr<Runtime> runtime;
};

}  // namespace

void* jst_ref_viz_create(const char* type, const char* config, int rank, const uint64_t* shape, int64_t sampleAxis,
                         int64_t batchAxis, int64_t channelAxis) {
    auto v = std::make_unique<VizSession>();
    v->type = type;
    Shape dims(shape, shape + rank);
    if (v->input.create(DeviceType::CPU, DataType::F32, dims) != Result::SUCCESS) {
        Fail("viz input", Result::ERROR);
        return nullptr;
    }
    if (sampleAxis >= 0) v->input.setAttribute("sampleAxis", Index{static_cast<U64>(sampleAxis)});
    if (batchAxis >= 0) v->input.setAttribute("batchAxis", Index{static_cast<U64>(batchAxis)});
    if (channelAxis >= 0) v->input.setAttribute("channelAxis", Index{static_cast<U64>(channelAxis)});
    TensorMap inputs;
    inputs["signal"].requested("source", "signal");
    inputs["signal"].tensor = v->input;
    auto result = Registry::BuildModule(type, DeviceType::CPU, RuntimeType::NATIVE, "generic", v->module);
    if (result == Result::SUCCESS) {
        result = v->module->create(type, ParseKv(config), inputs);
    }
    if (result != Result::SUCCESS) {
        Fail(std::string("viz create ") + type, result);
        return nullptr;
    }
    v->runtime = std::make_unique<Runtime>(type, DeviceType::CPU, RuntimeType::NATIVE);
    result = v->runtime->create({{type, v->module}});
    if (result != Result::SUCCESS) {
        Fail("viz runtime", result);
        return nullptr;
    }
    return v.release();
}

int jst_ref_viz_compute(void* handle, const float* data, uint64_t count) {
    auto* v = static_cast<VizSession*>(handle);
    if (count != v->input.size()) {
        g_error = "viz_compute: size mismatch";
        return -1;
    }
    std::memcpy(v->input.data(), data, count * sizeof(float));
    std::unordered_set<std::string> skipped, failed;
    const auto result = v->runtime->compute({}, skipped, failed);
    return result == Result::SUCCESS ? 0 : Fail("viz compute", result);
}

int jst_ref_viz_reconfigure(void* handle, const char* config) {
    auto* v = static_cast<VizSession*>(handle);
    const auto result = v->module->reconfigure(ParseKv(config));
    return result == Result::SUCCESS ? 0 : Fail("viz reconfigure", result);
}

// lineplot: signalPoints [n, 2] (x, averaged amplitude); waterfall: frequencyBins [height, n] ring. Returns the
// element count; copies min(count, capacity) floats.
int64_t jst_ref_viz_read(void* handle, float* dst, uint64_t capacity) {
    auto* v = static_cast<VizSession*>(handle);
    const Tensor* tensor = nullptr;
    if (v->type == "lineplot") {
        const auto* impl = v->module->getImpl<Modules::LineplotImpl>();
        tensor = impl ? &(impl->*LineplotPeek::points()) : nullptr;
    } else {
        const auto* impl = v->module->getImpl<Modules::WaterfallImpl>();
        tensor = impl ? &(impl->*WaterfallPeek::bins()) : nullptr;
    }
    if (!tensor) {
        g_error = "viz_read: implementation unavailable";
        return -1;
    }
    const uint64_t n = std::min<uint64_t>(capacity, tensor->size());
    std::memcpy(dst, tensor->data(), n * sizeof(float));
    return static_cast<int64_t>(tensor->size());
}

int64_t jst_ref_viz_write_index(void* handle) {
    auto* v = static_cast<VizSession*>(handle);
    const auto* impl = v->module->getImpl<Modules::WaterfallImpl>();
    return impl ? static_cast<int64_t>((impl->*WaterfallPeek::ring()).writeIndex) : -1;
}

void jst_ref_viz_destroy(void* handle) {
    auto* v = static_cast<VizSession*>(handle);
    if (!v) {
        return;
    }
    (void)v->runtime->destroy();
    (void)v->module->destroy();
    delete v;
}

}  // extern "C"
