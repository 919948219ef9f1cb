// jit.cuh -- user-defined ops: NVRTC-compiled generic kernels, see jit.cu
#pragma once

#include "common.cuh"

#include <string>

namespace splacu { namespace jit {

    enum Kernel { K_MXV_SEQ = 0, K_VXM_PAIRS, K_VXM_FOLD, K_ASSIGN_DENSE, K_ASSIGN_SPARSE, K_EADD_DENSE, K_EADD_FDB_DENSE, K_EADD_FDB_SPARSE, K_COUNT };

    struct Module {
        void* module = nullptr;// CUmodule
        void* fn[K_COUNT] = {};// CUfunction
    };

    inline bool is_user(const splacu_op* op) { return op && op->id < 0; }

    // compile (or find in the cache) the module for (dtype, mult, add, select); a null op = the neutral built-in of its slot
    int get_module(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel, const Module** out);
    // compile only (no device needed): what the CPU-side tests call
    int compile_only(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel, size_t* image_bytes);
    int launch(const Module* m, int kernel, size_t work_items, void** args, cudaStream_t s);
    std::string make_key(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel);

}}// namespace splacu::jit
