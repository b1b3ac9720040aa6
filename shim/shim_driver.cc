// TEST HARNESS for the shim (not product): runs the reference's own `spectrum_engine` block on the reference's
// own Flowgraph + scheduler_synchronous + NativeCudaRuntime with every module resolved from provider "b200"
// (shim/b200_modules.cc -> libb200dsp.so), next to the same block on the reference CPU provider, and prints the
// difference. The input is a device tensor produced by a caller-filled source module registered for
// (CUDA, NATIVE, "b200") and (CPU, NATIVE, "generic").
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "jetstream/block.hh"
#include "jetstream/detail/block_impl.hh"
#include "jetstream/detail/module_impl.hh"
#include "jetstream/flowgraph.hh"
#include "jetstream/flowgraph_view.hh"
#include "jetstream/logger.hh"
#include "jetstream/module_context.hh"
#include "jetstream/registry.hh"
#include "jetstream/runtime_context_native_cpu.hh"
#include "jetstream/runtime_context_native_cuda.hh"
#include "jetstream/scheduler_context.hh"

namespace Jetstream {
namespace Modules {

struct ShimSource : public Module::Config {
    U64 rows = 1;
    U64 size = 4096;
    JST_MODULE_TYPE(shim_source);
    JST_MODULE_PARAMS(rows, size);
};

struct ShimSourceBase : public Module::Impl, public DynamicConfig<ShimSource> {
    Result define() override { return defineInterfaceOutput("signal"); }
    Result create() override {
        JST_CHECK(signal.create(device(), DataType::CF32, {rows, size}));
        JST_CHECK(signal.setAttribute("sampleAxis", Index{1}));
        JST_CHECK(signal.setAttribute("batchAxis", Index{0}));
        outputs()["signal"].produced(name(), "signal", signal);
        return Result::SUCCESS;
    }
    Tensor signal;
};
struct ShimSourceCpu : public ShimSourceBase, public NativeCpuRuntimeContext, public Scheduler::Context {
    Result computeSubmit() override { return Result::SUCCESS; }
};
struct ShimSourceCuda : public ShimSourceBase, public NativeCudaRuntimeContext, public Scheduler::Context {
    Result computeSubmit(const cudaStream_t&) override { return Result::SUCCESS; }
};
JST_REGISTER_MODULE(ShimSourceCpu, DeviceType::CPU, RuntimeType::NATIVE, "generic");
JST_REGISTER_MODULE(ShimSourceCuda, DeviceType::CUDA, RuntimeType::NATIVE, "b200");

}  // namespace Modules

namespace Blocks {
struct ShimSource : public Block::Config {
    U64 rows = 1;
    U64 size = 4096;
    JST_BLOCK_TYPE(shim_source);
    JST_BLOCK_DOMAIN("Test");
    JST_BLOCK_PARAMS(rows, size);
    JST_BLOCK_DESCRIPTION("Shim Source", "Caller-filled source.", "Shim test harness source block.");
};
struct ShimSourceBlock : public Block::Impl, public DynamicConfig<Blocks::ShimSource> {
    Result configure() override {
        moduleConfig->rows = rows;
        moduleConfig->size = size;
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

static Tensor OutputOf(Flowgraph& fg, const char* block, const char* port) {
    TensorMap outputs;
    fg.view().outputs(block, outputs);
    return outputs.at(port).tensor;
}

static bool RunChain(const DeviceType device, const char* provider, const std::vector<std::complex<float>>& x,
                     const U64 rows, const U64 n, const bool scale, const bool agc, std::vector<float>& result) {
    Flowgraph fg;
    if (fg.create({}, nullptr, nullptr, nullptr) != Result::SUCCESS) return false;
    Parser::Map srcConfig;
    srcConfig["rows"] = std::to_string(rows);
    srcConfig["size"] = std::to_string(n);
    if (fg.blockCreate("src", "shim_source", srcConfig, {}, device, RuntimeType::NATIVE, provider) != Result::SUCCESS) {
        std::printf("source create failed: %s\n", JST_LOG_LAST_ERROR().c_str());
        return false;
    }
    Tensor source = OutputOf(fg, "src", "signal");
    if (device == DeviceType::CUDA) {
        cudaMemcpy(source.buffer().data(), x.data(), x.size() * sizeof(x[0]), cudaMemcpyHostToDevice);
    } else {
        std::memcpy(source.data(), x.data(), x.size() * sizeof(x[0]));
    }
    Parser::Map cfg;
    cfg["enableScale"] = std::string(scale ? "true" : "false");
    cfg["enableAgc"] = std::string(agc ? "true" : "false");
    cfg["rangeMin"] = std::string("-120");
    cfg["rangeMax"] = std::string("0");
    TensorMap inputs;
    inputs["buffer"].requested("src", "signal");
    if (fg.blockCreate("spec", "spectrum_engine", cfg, inputs, device, RuntimeType::NATIVE, provider) != Result::SUCCESS) {
        std::printf("spectrum_engine create failed on provider %s: %s\n", provider, JST_LOG_LAST_ERROR().c_str());
        return false;
    }
    for (int cycle = 0; cycle < 2; ++cycle) {
        if (fg.compute() != Result::SUCCESS) {
            std::printf("compute failed: %s\n", JST_LOG_LAST_ERROR().c_str());
            return false;
        }
    }
    Tensor out = OutputOf(fg, "spec", "buffer");
    result.resize(out.size());
    if (device == DeviceType::CUDA) {
        cudaDeviceSynchronize();
        cudaMemcpy(result.data(), static_cast<const std::uint8_t*>(out.buffer().data()) + out.offsetBytes(),
                   result.size() * sizeof(float), cudaMemcpyDeviceToHost);
    } else {
        std::memcpy(result.data(), out.data(), result.size() * sizeof(float));
    }
    std::vector<Flowgraph::View::MetricEntry> metrics;
    fg.view().metrics("spec", metrics);
    std::printf("  [%s/%s] modules:", device == DeviceType::CUDA ? "cuda" : "cpu", provider);
    for (const auto& m : metrics) {
        if (const auto* timing = std::any_cast<Module::Timing>(&m.value)) {
            std::printf(" %s(x%llu)", m.name.c_str(), static_cast<unsigned long long>(timing->cycles));
        }
    }
    std::printf("\n");
    std::vector<std::string> names;
    fg.view().keys(names);
    for (auto it = names.rbegin(); it != names.rend(); ++it) fg.blockDestroy(*it, false);
    fg.destroy();
    return true;
}

int main() {
    JST_LOG_SET_DEBUG_LEVEL(1);
    const U64 rows = 64, n = 4096;
    std::vector<std::complex<float>> x(rows * n);
    std::mt19937 rng(7);
    std::normal_distribution<float> noise(0.0f, 1e-3f);
    for (U64 r = 0; r < rows; ++r) {
        for (U64 i = 0; i < n; ++i) {
            const double ph = 2.0 * M_PI * static_cast<double>((97 * r) % n) * i / n;
            x[r * n + i] = {static_cast<float>(0.5 * std::cos(ph)) + noise(rng),
                            static_cast<float>(0.5 * std::sin(ph)) + noise(rng)};
        }
    }
    int failures = 0;
    for (const int variant : {0, 1, 3}) {                 // bit 0: enableScale, bit 1: enableAgc
        const bool scale = variant & 1, agc = variant & 2;
        std::vector<float> cpu, gpu;
        if (!RunChain(DeviceType::CPU, "generic", x, rows, n, scale, agc, cpu) ||
            !RunChain(DeviceType::CUDA, "b200", x, rows, n, scale, agc, gpu)) {
            return 2;
        }
        double worst = 0.0;
        U64 bad = 0;
        for (U64 i = 0; i < cpu.size(); ++i) {
            if (std::isinf(cpu[i]) && std::isinf(gpu[i])) continue;
            const double d = std::fabs(static_cast<double>(cpu[i]) - gpu[i]);
            worst = std::max(worst, d);
            bad += d > (scale ? 2e-3 : 0.25);   // noise-floor bins: see DESIGN.md §2 (tests/ hold the tight bound)
        }
        std::printf("spectrum_engine(enableScale=%d, enableAgc=%d) reference-CPU vs b200 provider through the reference "
                    "Flowgraph: max |diff| = %.3e, out-of-bound = %llu of %zu\n",
                    scale ? 1 : 0, agc ? 1 : 0, worst, static_cast<unsigned long long>(bad), cpu.size());
        failures += bad != 0;
    }
    std::printf(failures ? "SHIM FAIL\n" : "SHIM OK\n");
    return failures ? 1 : 0;
}
