// profile.cuh -- profiling hooks of the backend (the reference's counterpart is the scope-based TimeProfiler,
// src/profiling/time_profiler.hpp:84-100, whose labels carry host "nano" and device "queued / executed" times filled from OpenCL
// events by the CL backend). Every C-ABI entry point opens a scope:
//   * always: an NVTX range "splacu/<entry>" (visible in Nsight Systems / Compute; a no-op without a tool attached)
//   * with splacu_profile_enable(1): a cudaEvent pair on the launching stream; splacu_profile_dump() waits for the events and
//     reports calls, device milliseconds and host milliseconds per label, splacu_profile_reset() clears them.
#pragma once

#include "common.cuh"

namespace splacu {

    struct ProfScope {
        ProfScope(const char* label, cudaStream_t stream);
        ~ProfScope();
        int          slot = -1;
        cudaStream_t stream;
        long long    t0 = 0;
    };

}// namespace splacu

#define SPLACU_PROFILE(label, stream) ::splacu::ProfScope _splacu_prof_scope(label, stream)
