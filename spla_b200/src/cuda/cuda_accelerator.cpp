#include "cuda_accelerator.hpp"

#include <cuda/cuda_storage.hpp>

#include <cstdlib>
#include <vector>

namespace spla {

    CudaAccelerator::~CudaAccelerator() {
        // decorations may outlive the accelerator (reference src/library.cpp:97-104); device buffers are
        // released by their owners through splacu_free, which stays valid after the runtime is finalised
        if (m_group) splacu_dist_destroy(m_group);
        if (m_workspace) splacu_workspace_destroy(m_workspace);
        splacu_finalize();
    }

    Status CudaAccelerator::init() {
        int count = 0;
        splacu_device_count(&count);
        if (count == 0) {
            LOG_MSG(Status::DeviceNotFound, "no cuda device found");
            return Status::DeviceNotFound;
        }
        // device formats: constructors / validators / converters in the Acc* slots of the storage managers
        register_formats_cuda();
        if (const char* env = std::getenv("SPLA_CUDA_DEVICES")) m_n_devices = std::max(1, std::atoi(env));
        return set_device(0);
    }

    Status CudaAccelerator::set_platform(int) {
        return Status::Ok;// a single CUDA platform exists
    }

    Status CudaAccelerator::set_device(int index) {
        int count = 0;
        splacu_device_count(&count);
        if (index < 0 || index >= count) {
            LOG_MSG(Status::DeviceNotFound, "cuda device index " << index << " out of range, devices: " << count);
            return Status::DeviceNotFound;
        }
        if (m_group) {
            splacu_dist_destroy(m_group);
            m_group = nullptr;
        }
        if (m_workspace) {
            splacu_workspace_destroy(m_workspace);
            m_workspace = nullptr;
        }
        if (splacu_init(index) != 0) {
            LOG_MSG(Status::Error, "failed to init cuda device " << index << ": " << splacu_last_error());
            return Status::Error;
        }
        if (splacu_workspace_create(&m_workspace) != 0) {
            LOG_MSG(Status::Error, "failed to create device workspace: " << splacu_last_error());
            return Status::Error;
        }
        m_device = index;
        char name[256];
        splacu_device_name(name, sizeof(name));
        m_description = std::string("CUDA device ") + std::to_string(index) + ": " + name;
        LOG_MSG(Status::Ok, "select " << m_description);
        return rebuild_group();
    }

    Status CudaAccelerator::set_queues_count(int count) {
        // the reference's knob for parallel queues (it only ever uses one, cl_accelerator.hpp:81); here: devices for the two products
        m_n_devices = count < 1 ? 1 : count;
        return rebuild_group();
    }

    Status CudaAccelerator::rebuild_group() {
        if (m_group) {
            splacu_dist_destroy(m_group);
            m_group = nullptr;
        }
        if (m_n_devices <= 1) return Status::Ok;
        int count = 0;
        splacu_device_count(&count);
        const bool share = std::getenv("SPLA_CUDA_SHARE_DEVICES") != nullptr;
        int        n     = m_n_devices;
        if (!share && n > count) {
            LOG_MSG(Status::Ok, "cuda backend: " << n << " devices requested, " << count << " present");
            n = count;
        }
        if (n > SPLACU_MAX_PEERS) n = SPLACU_MAX_PEERS;
        if (n <= 1) return Status::Ok;
        std::vector<int> ids(n);
        for (int p = 0; p < n; ++p) ids[p] = (m_device + p) % count;
        if (splacu_dist_create(&m_group, n, ids.data()) != 0) {
            LOG_MSG(Status::Error, "failed to create the device group: " << splacu_last_error());
            m_group = nullptr;
            return Status::Error;
        }
        int uses_nccl = 0;
        splacu_dist_info(m_group, nullptr, &uses_nccl);
        m_description += " + " + std::to_string(n - 1) + " more shard(s) for mxv / vxm (" + (uses_nccl ? "NCCL broadcast" : "peer copies") + ")";
        LOG_MSG(Status::Ok, "cuda backend: products sharded over " << n << " shards");
        return Status::Ok;
    }

    const std::string& CudaAccelerator::get_name() { return m_name; }
    const std::string& CudaAccelerator::get_description() { return m_description; }
    const std::string& CudaAccelerator::get_suffix() { return m_suffix; }

}// namespace spla
