// cuda_algo_registry.hpp -- registration of the CUDA algorithms in spla's registry under the "__cuda" key suffix,
// beside register_algo_cpu / register_algo_cl (reference src/cpu/cpu_algo_registry.hpp, src/opencl/cl_algo_registry.hpp).
#ifndef SPLA_CUDA_ALGO_REGISTRY_HPP
#define SPLA_CUDA_ALGO_REGISTRY_HPP

#include <core/registry.hpp>

namespace spla {

    /** @brief Register all cuda algorithms; called from Library::Library() next to register_algo_cpu (reference src/library.cpp:80-87) */
    void register_algo_cuda(class Registry* g_registry);

}// namespace spla

#endif//SPLA_CUDA_ALGO_REGISTRY_HPP
