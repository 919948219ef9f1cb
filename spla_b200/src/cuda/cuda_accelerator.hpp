// cuda_accelerator.hpp -- spla Accelerator implementation for one CUDA device (B200, sm_100a).
//
// Plugs into the reference's backend interface (reference src/core/accelerator.hpp:58-69) next to
// CLAccelerator (reference src/opencl/cl_accelerator.hpp:62-136). It owns nothing but a handle to the
// splacu C-ABI runtime (include/splacu.h): host C++ never sees a CUDA header.
#ifndef SPLA_CUDA_ACCELERATOR_HPP
#define SPLA_CUDA_ACCELERATOR_HPP

#include <core/accelerator.hpp>
#include <core/logger.hpp>
#include <spla/library.hpp>

#include <splacu.h>

#include <stdexcept>
#include <string>

namespace spla {

    /** Value a maintainer adds to `enum class AcceleratorType` (reference include/spla/config.hpp:89-94) as `Cuda = 2`,
     *  mirrored in the C API as SPLA_ACCELERATOR_TYPE_CUDA (reference include/spla.h:83-86). Kept here so that the
     *  public headers of the reference compile unmodified. */
    constexpr AcceleratorType ACCELERATOR_TYPE_CUDA = static_cast<AcceleratorType>(2);

#define GPU_CUDA_SUFFIX                      "__cuda"
#define MAKE_KEY_CUDA_0(name, type)          MAKE_KEY_0(name, type) + GPU_CUDA_SUFFIX
#define MAKE_KEY_CUDA_1(name, op)            MAKE_KEY_1(name, op) + GPU_CUDA_SUFFIX
#define MAKE_KEY_CUDA_2(name, op1, op2)      MAKE_KEY_2(name, op1, op2) + GPU_CUDA_SUFFIX
#define MAKE_KEY_CUDA_3(name, op1, op2, op3) MAKE_KEY_3(name, op1, op2, op3) + GPU_CUDA_SUFFIX

// every splacu call site: a failing device call becomes a C++ exception, which the dispatcher turns into
// Status::Error (reference src/core/dispatcher.cpp:62-80). There is no CPU fallback on this path.
#define SPLACU_CALL(expr)                                                                                    \
    do {                                                                                                     \
        int _rc = (expr);                                                                                    \
        if (_rc != 0) throw std::runtime_error(std::string("cuda backend: " #expr " failed: ") + splacu_last_error()); \
    } while (0)

    /**
     * @class CudaAccelerator
     * @brief CUDA acceleration backend: one home device, optionally a group of devices for the two products
     *
     * The vectors of a program always live on the home device (set_device). With more than one device requested --
     * Library::set_queues_count(n), the knob the reference reserves for parallel queues (src/core/accelerator.hpp:58-69,
     * src/library.cpp:136-138), or the environment variable SPLA_CUDA_DEVICES=n read at init -- mxv_masked / vxm_masked run
     * sharded over n devices of the box (splacu_dist_*: rows nnz-balanced for the pull, columns for the push, v broadcast
     * over NVLink with NCCL); everything else is unchanged. SPLA_CUDA_SHARE_DEVICES=1 lets the shards wrap around the
     * devices present (n shards on fewer GPUs: for testing).
     */
    class CudaAccelerator final : public Accelerator {
    public:
        ~CudaAccelerator() override;

        Status             init() override;
        Status             set_platform(int index) override;
        Status             set_device(int index) override;
        Status             set_queues_count(int count) override;
        const std::string& get_name() override;
        const std::string& get_description() override;
        const std::string& get_suffix() override;

        /** scratch shared by vxm / compaction / reductions; one in-order stream => one workspace */
        splacu_workspace get_workspace() { return m_workspace; }
        /** the backend's in-order stream (NULL selects it on the C-ABI side) */
        void* get_stream() { return nullptr; }
        /** the device group of the sharded products, or nullptr when a single device is used */
        splacu_dist get_group() { return m_group; }

    private:
        std::string      m_name        = "CUDA";
        std::string      m_description = "no device";
        std::string      m_suffix      = GPU_CUDA_SUFFIX;
        splacu_workspace m_workspace   = nullptr;
        splacu_dist      m_group       = nullptr;
        int              m_device      = 0;
        int              m_n_devices   = 1;

        Status rebuild_group();
    };

    /** @return the active CUDA accelerator; throws when another (or no) accelerator is installed */
    static inline CudaAccelerator* get_acc_cuda() {
        auto* acc = dynamic_cast<CudaAccelerator*>(Library::get()->get_accelerator());
        if (!acc) throw std::runtime_error("cuda backend: CudaAccelerator is not the active accelerator");
        return acc;
    }

}// namespace spla

#endif//SPLA_CUDA_ACCELERATOR_HPP
