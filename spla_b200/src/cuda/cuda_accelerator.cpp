#include "cuda_accelerator.hpp"

#include <cuda/cuda_storage.hpp>

namespace spla {

    CudaAccelerator::~CudaAccelerator() {
        // decorations may outlive the accelerator (reference src/library.cpp:97-104); device buffers are
        // released by their owners through splacu_free, which stays valid after the runtime is finalised
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
        return Status::Ok;
    }

    Status CudaAccelerator::set_queues_count(int) {
        return Status::Ok;// one in-order stream, like the single queue the reference ever uses (cl_accelerator.hpp:81)
    }

    const std::string& CudaAccelerator::get_name() { return m_name; }
    const std::string& CudaAccelerator::get_description() { return m_description; }
    const std::string& CudaAccelerator::get_suffix() { return m_suffix; }

}// namespace spla
