// profile.cu -- see profile.cuh
#include "profile.cuh"

#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace splacu {

    namespace {
        struct Sample {
            const char* label;
            cudaEvent_t start, stop;
            int         device;
            double      host_ms;
        };
        struct Total {
            uint64_t calls = 0;
            double   device_ms = 0.0, host_ms = 0.0;
        };
        bool                         g_enabled = false;
        std::mutex                   g_mutex;
        std::vector<Sample>          g_samples;
        std::map<std::string, Total> g_totals;

        long long now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

        void drain_locked() {
            int prev = 0;
            cudaGetDevice(&prev);
            for (Sample& s : g_samples) {
                cudaSetDevice(s.device);
                float ms = 0.f;
                if (cudaEventSynchronize(s.stop) == cudaSuccess && cudaEventElapsedTime(&ms, s.start, s.stop) == cudaSuccess) {
                    Total& t = g_totals[s.label];
                    t.calls += 1;
                    t.device_ms += ms;
                    t.host_ms += s.host_ms;
                }
                cudaEventDestroy(s.start);
                cudaEventDestroy(s.stop);
            }
            g_samples.clear();
            cudaSetDevice(prev);
            cudaGetLastError();
        }
    }// namespace

    ProfScope::ProfScope(const char* label, cudaStream_t s) : stream(s) {
        nvtxRangePushA(label);
        if (!g_enabled) return;
        Sample smp;
        smp.label   = label;
        smp.host_ms = 0.0;
        cudaGetDevice(&smp.device);
        if (cudaEventCreate(&smp.start) != cudaSuccess || cudaEventCreate(&smp.stop) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        cudaEventRecord(smp.start, s);
        t0 = now_ns();
        std::lock_guard<std::mutex> lock(g_mutex);
        slot = (int) g_samples.size();
        g_samples.push_back(smp);
    }

    ProfScope::~ProfScope() {
        if (slot >= 0) {
            std::lock_guard<std::mutex> lock(g_mutex);
            if (slot < (int) g_samples.size()) {
                Sample& smp = g_samples[slot];
                cudaEventRecord(smp.stop, stream);
                smp.host_ms = (double) (now_ns() - t0) * 1e-6;
            }
            if (g_samples.size() > 4096) drain_locked();// bounded: long traversals call thousands of ops
        }
        nvtxRangePop();
    }

}// namespace splacu

using namespace splacu;

extern "C" {

int splacu_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!on && g_enabled) drain_locked();
    g_enabled = on != 0;
    return SPLACU_OK;
}

int splacu_profile_reset(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    drain_locked();
    g_totals.clear();
    return SPLACU_OK;
}

int splacu_profile_dump(char* buffer, int length) {
    SPLACU_REQUIRE(buffer && length > 0, "bad buffer");
    std::lock_guard<std::mutex> lock(g_mutex);
    drain_locked();
    std::string out = "label, calls, device_ms, host_ms\n";
    for (auto& kv : g_totals) {
        char line[256];
        snprintf(line, sizeof(line), "%s, %llu, %.4f, %.4f\n", kv.first.c_str(), (unsigned long long) kv.second.calls, kv.second.device_ms, kv.second.host_ms);
        out += line;
    }
    snprintf(buffer, (size_t) length, "%s", out.c_str());
    return SPLACU_OK;
}

}// extern "C"
