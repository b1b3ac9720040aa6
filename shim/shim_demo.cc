// Stand-alone self-check of the b200 binding (no Python): runs the reference's own blocks — spectrum_engine,
// filter, fm — in the reference's own Flowgraph on the reference CPU provider and on provider "b200", three cycles
// of fresh input each, and prints which modules each target created and how far the outputs are apart.
// tests/test_gpu_shim.py holds the tight tolerances; the bounds here only catch a broken build.
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

extern "C" {
const char* jst_shim_last_error();
void* jst_shim_create(int logLevel);
void jst_shim_destroy(void* handle);
int jst_shim_add_source(void* handle, const char* name, int dtype, int rank, const uint64_t* shape, int64_t sampleAxis,
                        int64_t batchAxis, int64_t channelAxis, int device, const char* provider, int mapped);
int jst_shim_write_source(void* handle, const char* name, const void* data, uint64_t bytes);
int jst_shim_add_block(void* handle, const char* name, const char* type, const char* config, const char* inputs,
                       int device, const char* provider);
int jst_shim_compute(void* handle);
int jst_shim_output_info(void* handle, const char* block, const char* port, int64_t* info);
int jst_shim_output_read(void* handle, const char* block, const char* port, void* dst, uint64_t bytes);
int jst_shim_metrics(void* handle, const char* block, char* buffer, uint64_t capacity);
}

namespace {

using Cf = std::complex<float>;

struct Case {
    const char* title;
    const char* type;
    const char* config;
    const char* inputPort;
    const char* outputPort;
    uint64_t rows, cols;
    double bound;       // max |cpu - b200| relative to max |cpu| (absolute for F32 outputs in [0,1] / dB)
    bool absolute;
};

bool Run(const Case& c, const int device, const char* provider, const std::vector<std::vector<Cf>>& cycles,
         std::vector<std::vector<float>>& results, std::string& modules) {
    void* s = jst_shim_create(1);
    if (!s) {
        std::printf("create failed: %s\n", jst_shim_last_error());
        return false;
    }
    const uint64_t shape[2] = {c.rows, c.cols};
    bool ok = jst_shim_add_source(s, "src", 1, 2, shape, 1, 0, -1, device, provider, 0) == 0;
    const std::string inputs = std::string(c.inputPort) + "=src.signal";
    ok = ok && jst_shim_add_block(s, "dut", c.type, c.config, inputs.c_str(), device, provider) == 0;
    for (const auto& x : cycles) {
        ok = ok && jst_shim_write_source(s, "src", x.data(), x.size() * sizeof(Cf)) == 0;
        ok = ok && jst_shim_compute(s) == 0;
        int64_t info[16] = {};
        ok = ok && jst_shim_output_info(s, "dut", c.outputPort, info) == 0;
        if (!ok) {
            break;
        }
        std::vector<float> out(static_cast<size_t>(info[14]) * (info[0] == 1 ? 2 : 1));
        ok = jst_shim_output_read(s, "dut", c.outputPort, out.data(), out.size() * sizeof(float)) == 0;
        results.push_back(std::move(out));
    }
    if (!ok) {
        std::printf("  [%s] %s failed: %s\n", provider, c.title, jst_shim_last_error());
    } else {
        char buffer[2048] = {};
        jst_shim_metrics(s, "dut", buffer, sizeof(buffer));
        modules.clear();
        std::string line;
        for (const char* p = buffer; *p; ++p) {
            if (*p == '\n') {
                modules += line.substr(0, line.find(' ')) + " ";
                line.clear();
            } else {
                line += *p;
            }
        }
    }
    jst_shim_destroy(s);
    return ok;
}

}  // namespace

int main() {
    const Case cases[] = {
        {"spectrum_engine", "spectrum_engine", "enableScale=false\nenableAgc=false", "buffer", "buffer", 64, 4096, 0.25, true},
        {"spectrum_engine +scale", "spectrum_engine", "enableScale=true\nenableAgc=false\nrangeMin=-120\nrangeMax=0",
         "buffer", "buffer", 64, 4096, 2e-3, true},
        {"spectrum_engine +scale +agc", "spectrum_engine", "enableScale=true\nenableAgc=true\nrangeMin=-120\nrangeMax=0",
         "buffer", "buffer", 64, 4096, 2e-3, true},
        {"filter 129 taps R=8, 3 heads", "filter",
         "sampleRate=8000000\nbandwidth=1000000\ntaps=129\nheads=3\ncenter=[0, 1000000, -2000000]", "signal", "buffer", 8,
         4096, 2e-5, false},
        {"fm narrow 75us", "fm", "mode=narrow\ndeemphasis=75us\nsampleRate=250000", "signal", "signal", 4, 8192, 2e-5,
         true},
        {"fm wide 50us", "fm", "mode=wide\ndeemphasis=50us\nsampleRate=250000", "signal", "signal", 2, 8192, 1e-4, true},
    };
    int failures = 0;
    for (const auto& c : cases) {
        std::mt19937 rng(7);
        std::normal_distribution<float> noise(0.0f, 1e-2f);
        std::vector<std::vector<Cf>> cycles(3, std::vector<Cf>(c.rows * c.cols));
        double phase = 0.0;
        for (auto& x : cycles) {
            for (uint64_t i = 0; i < x.size(); ++i) {
                // FM-ish tone: slowly swept phase, continuous across rows and cycles
                phase += 0.3 + 0.25 * std::sin(2.0 * M_PI * static_cast<double>(i % 4096) / 4096.0);
                x[i] = Cf(static_cast<float>(0.5 * std::cos(phase)) + noise(rng),
                          static_cast<float>(0.5 * std::sin(phase)) + noise(rng));
            }
        }
        std::vector<std::vector<float>> cpu, gpu;
        std::string cpuModules, gpuModules;
        if (!Run(c, 0, "generic", cycles, cpu, cpuModules) || !Run(c, 1, "b200", cycles, gpu, gpuModules)) {
            ++failures;
            continue;
        }
        double worst = 0.0, peak = 0.0;
        bool shapes = cpu.size() == gpu.size();
        for (size_t k = 0; shapes && k < cpu.size(); ++k) {
            shapes = cpu[k].size() == gpu[k].size();
            for (size_t i = 0; shapes && i < cpu[k].size(); ++i) {
                if (!std::isfinite(cpu[k][i]) && !std::isfinite(gpu[k][i])) {
                    continue;
                }
                worst = std::max(worst, std::fabs(static_cast<double>(cpu[k][i]) - gpu[k][i]));
                peak = std::max(peak, std::fabs(static_cast<double>(cpu[k][i])));
            }
        }
        const double measure = c.absolute ? worst : worst / std::max(peak, 1e-30);
        const bool ok = shapes && measure <= c.bound;
        std::printf("%-32s cpu/generic: %s\n%-32s cuda/b200:   %s\n%-32s 3 cycles, max %s diff %.3e (bound %.1e) %s\n",
                    c.title, cpuModules.c_str(), "", gpuModules.c_str(), "", c.absolute ? "abs" : "rel", measure,
                    c.bound, ok ? "ok" : "FAIL");
        failures += ok ? 0 : 1;
    }
    std::printf(failures ? "SHIM FAIL\n" : "SHIM OK\n");
    return failures ? 1 : 0;
}
