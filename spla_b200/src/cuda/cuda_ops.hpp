// cuda_ops.hpp -- maps spla's built-in op objects (reference src/op.cpp:194-266, include/spla/op.hpp) to the compiled
// device functors of the splacu library (splacu_binop / splacu_selop). The reference's OpenCL backend JIT-compiles the
// op's source text per kernel (src/opencl/cl_program_builder.cpp:65-120); here the built-ins are AOT-specialised and a
// user-defined op (OpBinary::make_*, src/op.cpp:294-342) travels as its name + source text (splacu_op) and is compiled with
// NVRTC at first use inside libsplacu (csrc/jit.cu). It never runs on the CPU on the mxv / vxm path: a source that does not
// compile gives Status::CompilationError, a machine without NVRTC Status::NotImplemented.
#ifndef SPLA_CUDA_OPS_HPP
#define SPLA_CUDA_OPS_HPP

#include <spla/op.hpp>

#include <splacu.h>

#include <stdexcept>
#include <string>

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

    /** A binary / select op as the C ABI takes it: the built-in id, or id = -1 with the op's name and source text
     *  (reference Op::get_name / get_source_cl, include/spla/op.hpp:49-53). Keeps the strings alive for the call. */
    class CudaOpDesc {
    public:
        explicit CudaOpDesc(OpBinary* op) : m_name(op->get_name()), m_source(op->get_source_cl()) { m_desc.id = cuda_find_binop(op); }
        explicit CudaOpDesc(OpSelect* op) : m_name(op->get_name()), m_source(op->get_source_cl()) { m_desc.id = cuda_find_selop(op); }
        CudaOpDesc(const CudaOpDesc&)            = delete;
        CudaOpDesc& operator=(const CudaOpDesc&) = delete;
        const splacu_op* get() {
            m_desc.name   = m_name.c_str();
            m_desc.source = m_source.c_str();
            return &m_desc;
        }
        bool user_defined() const { return m_desc.id < 0; }

    private:
        splacu_op   m_desc{};
        std::string m_name, m_source;
    };

/** like SPLACU_CALL, but the two outcomes a user-defined op can have are spla statuses, not exceptions */
#define SPLACU_CALL_OPS(expr)                                                                                       \
    do {                                                                                                            \
        int _rc = (expr);                                                                                           \
        if (_rc == SPLACU_E_COMPILE) {                                                                              \
            LOG_MSG(Status::CompilationError, "cuda backend: " << splacu_last_error());                             \
            return Status::CompilationError;                                                                        \
        }                                                                                                           \
        if (_rc == SPLACU_E_NOT_IMPLEMENTED) {                                                                      \
            LOG_MSG(Status::NotImplemented, "cuda backend: " << splacu_last_error());                               \
            return Status::NotImplemented;                                                                          \
        }                                                                                                           \
        if (_rc != 0) throw std::runtime_error(std::string("cuda backend: " #expr " failed: ") + splacu_last_error()); \
    } while (0)

#define SPLA_CUDA_REQUIRE_OP(id, op)                                                                              \
    if ((id) < 0) {                                                                                               \
        LOG_MSG(Status::NotImplemented, "cuda backend: op " << (op)->get_name() << " is user-defined: no device code"); \
        return Status::NotImplemented;                                                                            \
    }

}// namespace spla

#endif//SPLA_CUDA_OPS_HPP
