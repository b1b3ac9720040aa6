// SURVEY.md §8 row a12 — the per-segment compute dispatch of the reference's CUDA runtime, re-implemented behind the
// reference's own interface: Runtime::Impl (include/jetstream/detail/runtime_impl.hh) over libb200dsp's
// b200_stream_* / b200_event_*. shim/build_shim.sh links THIS translation unit in place of the reference's
// src/runtime/native/cuda/impl.cc (both define NativeCudaRuntimeFactory()), so scheduler_synchronous drives every
// CUDA segment of a flowgraph through it.
//
// Contract kept (src/runtime/native/cuda/impl.cc:35-272): one non-blocking stream per segment; computeInitialize on
// create and computeDeinitialize on destroy, in reverse order, with clean-up of a partially created runtime; modules
// whose inputs were skipped are skipped; YIELD / TIMEOUT abort the cycle quietly, any other failure marks the module
// failed and ends the cycle; ONE stream synchronisation per cycle; Module::Timing (cycles, accumulated ms) from an
// event pair per module, read after the synchronisation.
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include <jetstream/backend/base.hh>
#include <jetstream/detail/runtime_impl.hh>
#include <jetstream/module.hh>
#include <jetstream/module_context.hh>
#include <jetstream/runtime_context_native_cuda.hh>

#include "b200_provider.hh"

namespace Jetstream {

namespace {

struct B200CudaRuntime : public Runtime::Impl {
    Result create(const Runtime::Modules& modules) override {
        JST_CHECK(Backend::State<DeviceType::CUDA>()->activate());
        if (b200_stream_create(B200::Ctx(), &stream) != B200_SUCCESS) {
            JST_ERROR("[RUNTIME_B200] Can't create stream: {}", b200_last_error());
            return Result::ERROR;
        }
        for (const auto& [moduleName, module] : modules) {
            if (module->device() != DeviceType::CUDA || module->runtime() != RuntimeType::NATIVE) {
                JST_ERROR("[RUNTIME_B200] Module '{}' is incompatible (DeviceType::{}, RuntimeType::{}).", moduleName,
                          module->device(), module->runtime());
                destroy();
                return Result::ERROR;
            }
            const auto result = hooks(module)->computeInitialize();
            if (result != Result::SUCCESS && result != Result::RELOAD) {
                hooks(module)->computeDeinitialize();
                destroy();
                return result;
            }
            Module::Timing timing;
            timing.runtime = this->name;
            timing.device = GetDevicePrettyName(device);
            timing.backend = GetRuntimePrettyName(backend);
            module->timing(timing);
            Entry entry{module, nullptr, nullptr};
            if (b200_event_create(B200::Ctx(), &entry.start) != B200_SUCCESS ||
                b200_event_create(B200::Ctx(), &entry.end) != B200_SUCCESS) {
                JST_ERROR("[RUNTIME_B200] Can't create timing events for '{}': {}", moduleName, b200_last_error());
                b200_event_destroy(B200::Ctx(), entry.start);
                hooks(module)->computeDeinitialize();
                destroy();
                return Result::ERROR;
            }
            entries.emplace(moduleName, entry);
            order.push_back(moduleName);
        }
        return Result::SUCCESS;
    }

    Result destroy() override {
        JST_CHECK(Backend::State<DeviceType::CUDA>()->activate());
        Result result = Result::SUCCESS;
        for (auto it = order.rbegin(); it != order.rend(); ++it) {
            auto& entry = entries.at(*it);
            const auto deinitialized = hooks(entry.module)->computeDeinitialize();
            if (result == Result::SUCCESS && deinitialized != Result::SUCCESS && deinitialized != Result::RELOAD) {
                result = deinitialized;
            }
            if (b200_event_destroy(B200::Ctx(), entry.start) != B200_SUCCESS ||
                b200_event_destroy(B200::Ctx(), entry.end) != B200_SUCCESS) {
                JST_ERROR("[RUNTIME_B200] Can't destroy timing events for '{}': {}", *it, b200_last_error());
                result = result == Result::SUCCESS ? Result::ERROR : result;
            }
        }
        if (stream && b200_stream_destroy(B200::Ctx(), stream) != B200_SUCCESS) {
            JST_ERROR("[RUNTIME_B200] Can't destroy stream: {}", b200_last_error());
            result = result == Result::SUCCESS ? Result::ERROR : result;
        }
        stream = nullptr;
        entries.clear();
        order.clear();
        return result;
    }

    Result compute(const std::vector<std::string>& modules, std::unordered_set<std::string>& skippedModules,
                   std::unordered_set<std::string>& failedModules) override {
        JST_CHECK(Backend::State<DeviceType::CUDA>()->activate());
        const auto& targets = modules.empty() ? order : modules;
        std::vector<Entry*> executed;
        executed.reserve(targets.size());
        const auto cudaStream = static_cast<cudaStream_t>(stream);
        for (const auto& moduleName : targets) {
            const auto found = entries.find(moduleName);
            if (found == entries.end()) {
                failedModules.insert(moduleName);
                JST_ERROR("[RUNTIME_B200] Context for module '{}' not found.", moduleName);
                return Result::ERROR;
            }
            Entry& entry = found->second;
            if (skippedModules.contains(moduleName) || hasSkippedInputs(entry.module, skippedModules)) {
                skippedModules.insert(moduleName);
                continue;
            }
            if (b200_event_record(B200::Ctx(), entry.start, stream) != B200_SUCCESS) {
                failedModules.insert(moduleName);
                JST_ERROR("[RUNTIME_B200] Can't record start event for '{}': {}", moduleName, b200_last_error());
                return Result::ERROR;
            }
            const auto result = hooks(entry.module)->computeSubmit(cudaStream);
            if (result == Result::YIELD || result == Result::TIMEOUT) {
                return result;
            }
            if (result != Result::SUCCESS && result != Result::RELOAD && result != Result::SKIP) {
                failedModules.insert(moduleName);
                return result;
            }
            if (b200_event_record(B200::Ctx(), entry.end, stream) != B200_SUCCESS ||
                b200_check_async_error(B200::Ctx()) != B200_SUCCESS) {
                failedModules.insert(moduleName);
                JST_ERROR("[RUNTIME_B200] Module kernel execution failed for '{}': {}", moduleName, b200_last_error());
                return Result::ERROR;
            }
            executed.push_back(&entry);
            if (result == Result::SKIP) {
                skippedModules.insert(moduleName);
            }
        }
        // one synchronisation per cycle, then the timings
        if (b200_stream_synchronize(B200::Ctx(), stream) != B200_SUCCESS ||
            b200_check_async_error(B200::Ctx()) != B200_SUCCESS) {
            JST_ERROR("[RUNTIME_B200] Runtime execution failed: {}", b200_last_error());
            return Result::ERROR;
        }
        for (Entry* entry : executed) {
            float elapsed = 0.0f;
            if (b200_event_elapsed_ms(B200::Ctx(), entry->start, entry->end, &elapsed) != B200_SUCCESS) {
                failedModules.insert(entry->module->name());
                JST_ERROR("[RUNTIME_B200] Can't get elapsed time for '{}': {}", entry->module->name(), b200_last_error());
                return Result::ERROR;
            }
            auto timing = entry->module->timing();
            timing.cycles += 1;
            timing.computeTime += elapsed;
            entry->module->timing(timing);
        }
        return Result::SUCCESS;
    }

 private:
    struct Entry {
        std::shared_ptr<Module> module;
        b200_event start;
        b200_event end;
    };

    static std::shared_ptr<NativeCudaRuntimeContext> hooks(const std::shared_ptr<Module>& module) {
        return std::dynamic_pointer_cast<NativeCudaRuntimeContext>(module->context()->runtime());
    }

    b200_stream stream = nullptr;
    std::unordered_map<std::string, Entry> entries;
    std::vector<std::string> order;
};

}  // namespace

std::shared_ptr<Runtime::Impl> NativeCudaRuntimeFactory() { return std::make_shared<B200CudaRuntime>(); }

}  // namespace Jetstream
