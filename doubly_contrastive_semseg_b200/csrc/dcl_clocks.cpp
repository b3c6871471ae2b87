// Measurement aid (bench.py): SM clock and throttle-reason samples taken DURING a timed region by a native thread
// through NVML.  Why not `nvidia-smi -lms`: while that process polls - even at 100 ms - every step of a 0.4 ms
// workload gets 0.25 ms slower (it re-queries every GPU of the box and keeps driver locks busy), and its start-up
// can stall a step for 100+ ms; a Python thread fights the interpreter lock of the thread that issues the launches.
// This thread asks for three numbers of ONE device every few milliseconds and sleeps in between.  libnvidia-ml is
// dlopen'ed: the library still loads where NVML is absent and the sampler then reports zero samples.
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>
#include "dcl_common.cuh"

namespace {

typedef int (*fn_init)();
typedef int (*fn_handle_by_bus)(const char*, void**);
typedef int (*fn_clock)(void*, int, unsigned int*);
typedef int (*fn_reasons)(void*, unsigned long long*);
typedef int (*fn_power)(void*, unsigned int*);

struct Nvml {
    void* lib = nullptr;
    fn_init init = nullptr;
    fn_handle_by_bus by_bus = nullptr;
    fn_clock clock = nullptr, max_clock = nullptr;
    fn_reasons reasons = nullptr;
    fn_power power = nullptr;
    bool ok = false;
};

Nvml& nvml() {
    static Nvml n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    n.lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
    if (!n.lib) return n;
    n.init = reinterpret_cast<fn_init>(dlsym(n.lib, "nvmlInit_v2"));
    n.by_bus = reinterpret_cast<fn_handle_by_bus>(dlsym(n.lib, "nvmlDeviceGetHandleByPciBusId_v2"));
    n.clock = reinterpret_cast<fn_clock>(dlsym(n.lib, "nvmlDeviceGetClockInfo"));
    n.max_clock = reinterpret_cast<fn_clock>(dlsym(n.lib, "nvmlDeviceGetMaxClockInfo"));
    n.reasons = reinterpret_cast<fn_reasons>(dlsym(n.lib, "nvmlDeviceGetCurrentClocksThrottleReasons"));
    n.power = reinterpret_cast<fn_power>(dlsym(n.lib, "nvmlDeviceGetPowerUsage"));
    n.ok = n.init && n.by_bus && n.clock && n.max_clock && n.reasons && n.init() == 0;
    return n;
}

struct Sampler {
    std::thread th;
    std::atomic<bool> stop{false};
    std::vector<unsigned int> sm;
    unsigned long long reasons = 0;
    unsigned int max_mhz = 0, power_mw = 0;
    bool running = false;
};
Sampler g_s;
std::mutex g_mu;

}  // namespace

// Start sampling the CURRENT CUDA device every period_us microseconds (>= 1000).  Returns 0, or DCL_ERR_ARG when
// NVML is unavailable (the caller then reports zero samples) or a sampler is already running.
extern "C" int dcl_clock_sampler_start(int period_us) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_s.running) return dcl::fail(DCL_ERR_ARG, "clock sampler already running");
    Nvml& n = nvml();
    if (!n.ok) return dcl::fail(DCL_ERR_ARG, "NVML is not available");
    int dev = 0;
    char bus[32] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) {
        cudaGetLastError();
        return dcl::fail(DCL_ERR_ARG, "no current CUDA device");
    }
    void* h = nullptr;
    if (n.by_bus(bus, &h) != 0 || !h) return dcl::fail(DCL_ERR_ARG, "NVML does not know device %s", bus);
    if (period_us < 1000) period_us = 1000;
    g_s.sm.clear();
    g_s.sm.reserve(4096);
    g_s.reasons = 0;
    g_s.power_mw = 0;
    n.max_clock(h, /* NVML_CLOCK_SM */ 1, &g_s.max_mhz);
    g_s.stop.store(false);
    g_s.running = true;
    g_s.th = std::thread([h, period_us] {
        Nvml& nv = nvml();
        while (!g_s.stop.load(std::memory_order_relaxed)) {
            unsigned int mhz = 0, mw = 0;
            unsigned long long r = 0;
            if (nv.clock(h, 1, &mhz) == 0) g_s.sm.push_back(mhz);
            if (nv.reasons(h, &r) == 0) g_s.reasons |= r;
            if (nv.power && nv.power(h, &mw) == 0 && mw > g_s.power_mw) g_s.power_mw = mw;
            std::this_thread::sleep_for(std::chrono::microseconds(period_us));
        }
    });
    return 0;
}

// Stop and summarise: out[0] = samples, out[1] = median SM MHz, out[2] = max SM MHz (device limit), out[3] = OR of the
// throttle-reason bit masks seen (NVML: 0x4 sw power cap, 0x8 hw slowdown, 0x20 sw thermal, 0x40 hw thermal),
// out[4] = highest power draw in W.
extern "C" int dcl_clock_sampler_stop(double* out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!out) return dcl::fail(DCL_ERR_ARG, "null pointer argument");
    for (int i = 0; i < 5; ++i) out[i] = 0.0;
    if (!g_s.running) return 0;
    g_s.stop.store(true);
    g_s.th.join();
    g_s.running = false;
    std::vector<unsigned int> v = g_s.sm;
    out[0] = static_cast<double>(v.size());
    if (!v.empty()) {
        std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
        out[1] = v[v.size() / 2];
    }
    out[2] = g_s.max_mhz;
    out[3] = static_cast<double>(g_s.reasons);
    out[4] = g_s.power_mw / 1000.0;
    return 0;
}
