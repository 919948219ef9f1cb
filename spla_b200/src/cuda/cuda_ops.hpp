// cuda_ops.hpp -- maps spla's built-in op objects (reference src/op.cpp:194-266, include/spla/op.hpp) to the compiled
// device functors of the splacu library (splacu_binop / splacu_selop). The reference's OpenCL backend JIT-compiles the
// op's source text per kernel (src/opencl/cl_program_builder.cpp:65-120); here the built-ins are AOT-specialised and a
// user-defined op (OpBinary::make_*, src/op.cpp:294-342) has no device code: the algorithm reports
// Status::NotImplemented -- it does NOT silently run on the CPU.
#ifndef SPLA_CUDA_OPS_HPP
#define SPLA_CUDA_OPS_HPP

#include <spla/op.hpp>

#include <splacu.h>

namespace spla {

    /** @return splacu_binop of a built-in binary op, or -1 for a user-defined one (identity by object, not by name) */
    int cuda_find_binop(const OpBinary* op);

    /** @return splacu_selop of a built-in select op, or -1 for a user-defined one */
    int cuda_find_selop(const OpSelect* op);

    /** Neighbour vector ops only (NOT mxv / vxm): an op without device code is handed to spla's own CPU algorithm of the same
     *  task, exactly what Dispatcher::dispatch does for a key the accelerator does not provide (reference
     *  src/core/dispatcher.cpp:57-60); the storage manager moves the operands back to host formats. */
    Status cuda_defer_to_cpu(const struct DispatchContext& ctx);

#define SPLA_CUDA_OP_OR_CPU(id, ctx) \
    if ((id) < 0) return cuda_defer_to_cpu(ctx);

#define SPLA_CUDA_REQUIRE_OP(id, op)                                                                              \
    if ((id) < 0) {                                                                                               \
        LOG_MSG(Status::NotImplemented, "cuda backend: op " << (op)->get_name() << " is user-defined: no device code"); \
        return Status::NotImplemented;                                                                            \
    }

}// namespace spla

#endif//SPLA_CUDA_OPS_HPP
