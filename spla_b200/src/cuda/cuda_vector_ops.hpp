// cuda_vector_ops.hpp -- the vector operations that sit between consecutive mxv / vxm calls inside spla's bfs / sssp / pr
// loops (reference src/algorithm.cpp:91,102,213-214,313-317), on the CUDA device, so that a whole traversal stays in
// device formats: v_assign_masked, v_count_mf, v_eadd_fdb, v_eadd, v_reduce. Semantics follow the CPU algorithms
// (reference src/cpu/cpu_v_assign.hpp:55-133, cpu_v_count_mf.hpp:55-111, cpu_v_eadd_fdb.hpp:55-139, cpu_v_eadd.hpp:55-155,
// cpu_v_reduce.hpp:55-118); format dispatch follows their OpenCL twins (src/opencl/cl_v_*.hpp).
#ifndef SPLA_CUDA_VECTOR_OPS_HPP
#define SPLA_CUDA_VECTOR_OPS_HPP

#include <schedule/schedule_tasks.hpp>

#include <core/dispatcher.hpp>
#include <core/registry.hpp>
#include <core/top.hpp>
#include <core/tscalar.hpp>
#include <core/ttype.hpp>
#include <core/tvector.hpp>
#include <cuda/cuda_formats.hpp>
#include <cuda/cuda_ops.hpp>
#include <profiling/time_profiler.hpp>

namespace spla {

    /** r[i] = select(mask[i]) ? assign(r[i], value) : r[i] */
    template<typename T>
    class Algo_v_assign_masked_cuda final : public RegistryAlgo {
    public:
        std::string get_name() override { return "v_assign_masked"; }
        std::string get_description() override { return "parallel vector masked assignment on cuda device"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/vector_assign");

            auto t         = ctx.task.template cast_safe<ScheduleTask_v_assign_masked>();
            auto r         = t->r.template cast_safe<TVector<T>>();
            auto mask      = t->mask.template cast_safe<TVector<T>>();
            auto value     = t->value.template cast_safe<TScalar<T>>();
            auto op_assign = t->op_assign.template cast_safe<TOpBinary<T, T, T>>();
            auto op_select = t->op_select.template cast_safe<TOpSelect<T>>();

            CudaOpDesc d_assign(op_assign.get()), d_sel(op_select.get());// user-defined ops: NVRTC (csrc/jit.cu)

            auto* acc = get_acc_cuda();
            // a dense mask is used as such; anything else (AccCoo, CpuCoo, CpuDok ...) goes through the sparse form
            const bool dense_mask = !mask->is_valid(FormatVector::AccCoo) &&
                                    (mask->is_valid(FormatVector::AccDense) || mask->is_valid(FormatVector::CpuDense));

            r->validate_rwd(FormatVector::AccDense);
            auto* p_r = r->template get<CudaDenseVec<T>>();

            if (dense_mask) {
                mask->validate_rw(FormatVector::AccDense);
                const auto* p_mask = mask->template get<CudaDenseVec<T>>();
                SPLACU_CALL_OPS(splacu_v_assign_masked_dense_ops(cuda_dtype<T>(), d_assign.get(), d_sel.get(), r->get_n_rows(), p_r->Ax.get(),
                                                                 p_mask->Ax.get(), cuda_bits(value->get_value()), acc->get_stream()));
            } else {
                mask->validate_rw(FormatVector::AccCoo);
                const auto* p_mask = mask->template get<CudaCooVec<T>>();
                SPLACU_CALL_OPS(splacu_v_assign_masked_sparse_ops(cuda_dtype<T>(), d_assign.get(), d_sel.get(), p_r->Ax.get(), p_mask->values,
                                                                  static_cast<const uint32_t*>(p_mask->Ai.get()), p_mask->Ax.get(),
                                                                  cuda_bits(value->get_value()), acc->get_stream()));
            }
            return Status::Ok;
        }
    };

    /** number of meaningful (stored / non-fill) entries -> host scalar; the one synchronisation point of a traversal level */
    template<typename T>
    class Algo_v_count_mf_cuda final : public RegistryAlgo {
    public:
        std::string get_name() override { return "v_count_mf"; }
        std::string get_description() override { return "parallel vector count of meaningful entries on cuda device"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/v_count_mf");

            auto t = ctx.task.template cast_safe<ScheduleTask_v_count_mf>();
            auto v = t->v.template cast_safe<TVector<T>>();

            if (!v->is_valid(FormatVector::AccCoo) && v->is_valid(FormatVector::AccDense)) {
                auto*       acc   = get_acc_cuda();
                const auto* p_v   = v->template get<CudaDenseVec<T>>();
                uint32_t    count = 0;
                SPLACU_CALL(splacu_v_count_mf_dense(cuda_dtype<T>(), v->get_n_rows(), p_v->Ax.get(), cuda_bits(v->get_fill_value()),
                                                    acc->get_workspace(), &count, acc->get_stream()));
                t->r->set_uint(count);
                return Status::Ok;
            }
            v->validate_rw(FormatVector::AccCoo);
            t->r->set_uint(v->template get<CudaCooVec<T>>()->values);// structural count (reference cpu_v_count_mf.hpp:81-90)
            return Status::Ok;
        }
    };

    /** r = op(r, v); fdb = the entries of r that changed */
    template<typename T>
    class Algo_v_eadd_fdb_cuda final : public RegistryAlgo {
    public:
        std::string get_name() override { return "v_eadd_fdb"; }
        std::string get_description() override { return "parallel vector element-wise add with feedback on cuda device"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/vector_eadd_fdb");

            auto t   = ctx.task.template cast_safe<ScheduleTask_v_eadd_fdb>();
            auto r   = t->r.template cast_safe<TVector<T>>();
            auto v   = t->v.template cast_safe<TVector<T>>();
            auto fdb = t->fdb.template cast_safe<TVector<T>>();
            auto op  = t->op.template cast_safe<TOpBinary<T, T, T>>();

            CudaOpDesc d_op(op.get());// user-defined op: NVRTC (csrc/jit.cu)

            auto*      acc     = get_acc_cuda();
            const bool dense_v = !v->is_valid(FormatVector::AccCoo) && !v->is_valid(FormatVector::CpuCoo) &&
                                 (v->is_valid(FormatVector::AccDense) || v->is_valid(FormatVector::CpuDense));

            r->validate_rwd(FormatVector::AccDense);
            auto* p_r = r->template get<CudaDenseVec<T>>();

            if (dense_v) {
                v->validate_rw(FormatVector::AccDense);
                fdb->validate_wd(FormatVector::AccDense);
                const auto* p_v   = v->template get<CudaDenseVec<T>>();
                auto*       p_fdb = fdb->template get<CudaDenseVec<T>>();
                SPLACU_CALL_OPS(splacu_v_eadd_fdb_dense_op(cuda_dtype<T>(), d_op.get(), r->get_n_rows(), p_r->Ax.get(), p_v->Ax.get(), p_fdb->Ax.get(),
                                                           cuda_bits(fdb->get_fill_value()), acc->get_stream()));
            } else {
                v->validate_rw(FormatVector::AccCoo);
                fdb->validate_wd(FormatVector::AccCoo);
                const auto* p_v   = v->template get<CudaCooVec<T>>();
                auto*       p_fdb = fdb->template get<CudaCooVec<T>>();
                uint32_t    nf    = 0;
                SPLACU_CALL_OPS(splacu_v_eadd_fdb_sparse_begin_op(cuda_dtype<T>(), d_op.get(), p_r->Ax.get(), p_v->values,
                                                                  static_cast<const uint32_t*>(p_v->Ai.get()), p_v->Ax.get(), acc->get_workspace(), &nf,
                                                                  acc->get_stream()));
                cuda_coo_vec_resize(nf, *p_fdb);
                SPLACU_CALL(splacu_v_eadd_fdb_sparse_emit(acc->get_workspace(), p_fdb->Ai.as_index(), p_fdb->Ax.get(), acc->get_stream()));
            }
            return Status::Ok;
        }
    };

    /** r[i] = op(u[i], v[i]) on dense vectors (the only form the accelerated reference has, src/opencl/cl_v_eadd.hpp:62-71) */
    template<typename T>
    class Algo_v_eadd_cuda final : public RegistryAlgo {
    public:
        std::string get_name() override { return "v_eadd"; }
        std::string get_description() override { return "parallel vector element-wise add on cuda device"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/vector_eadd");

            auto t  = ctx.task.template cast_safe<ScheduleTask_v_eadd>();
            auto r  = t->r.template cast_safe<TVector<T>>();
            auto u  = t->u.template cast_safe<TVector<T>>();
            auto v  = t->v.template cast_safe<TVector<T>>();
            auto op = t->op.template cast_safe<TOpBinary<T, T, T>>();

            CudaOpDesc d_op(op.get());// user-defined op: NVRTC (csrc/jit.cu)

            u->validate_rw(FormatVector::AccDense);
            v->validate_rw(FormatVector::AccDense);
            r->validate_wd(FormatVector::AccDense);// after u / v: r may alias one of them

            const auto* p_u = u->template get<CudaDenseVec<T>>();
            const auto* p_v = v->template get<CudaDenseVec<T>>();
            auto*       p_r = r->template get<CudaDenseVec<T>>();

            SPLACU_CALL_OPS(splacu_v_eadd_dense_op(cuda_dtype<T>(), d_op.get(), r->get_n_rows(), p_r->Ax.get(), p_u->Ax.get(), p_v->Ax.get(),
                                                   get_acc_cuda()->get_stream()));
            return Status::Ok;
        }
    };

    /** r = fold(op, s, v) -> host scalar. The device fold is a tree, so op must be one of the associative + commutative
     *  built-ins; any other op is folded by spla's sequential CPU algorithm. FLOAT PLUS / MULT differ from the sequential fold
     *  by summation order. */
    template<typename T>
    class Algo_v_reduce_cuda final : public RegistryAlgo {
    public:
        std::string get_name() override { return "v_reduce"; }
        std::string get_description() override { return "parallel vector reduction on cuda device"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/vector_reduce");

            auto t  = ctx.task.template cast_safe<ScheduleTask_v_reduce>();
            auto r  = t->r.template cast_safe<TScalar<T>>();
            auto s  = t->s.template cast_safe<TScalar<T>>();
            auto v  = t->v.template cast_safe<TVector<T>>();
            auto op = t->op_reduce.template cast_safe<TOpBinary<T, T, T>>();

            const int id_op = cuda_find_binop(op.get());
            switch (id_op) {
                case SPLACU_PLUS: case SPLACU_MULT: case SPLACU_MIN: case SPLACU_MAX: case SPLACU_LOR:
                case SPLACU_LAND: case SPLACU_BOR: case SPLACU_BAND: case SPLACU_BXOR: break;
                default: return cuda_defer_to_cpu(ctx);// user-defined or order-dependent op: spla's sequential fold
            }

            auto*       acc  = get_acc_cuda();
            const void* data = nullptr;
            uint        n    = 0;
            if (v->is_valid(FormatVector::AccCoo) || (!v->is_valid(FormatVector::AccDense) && !v->is_valid(FormatVector::CpuDense))) {
                v->validate_rw(FormatVector::AccCoo);// sparse: fold the stored values (reference cpu_v_reduce.hpp:70-88)
                const auto* p_v = v->template get<CudaCooVec<T>>();
                data            = p_v->Ax.get();
                n               = p_v->values;
            } else {
                v->validate_rw(FormatVector::AccDense);
                data = v->template get<CudaDenseVec<T>>()->Ax.get();
                n    = v->get_n_rows();
            }
            uint32_t bits = 0;
            SPLACU_CALL(splacu_v_reduce_dense(cuda_dtype<T>(), id_op, n, data, cuda_bits(s->get_value()), acc->get_workspace(), &bits, acc->get_stream()));
            T result;
            std::memcpy(&result, &bits, sizeof(result));
            r->get_value() = result;
            return Status::Ok;
        }
    };

}// namespace spla

#endif//SPLA_CUDA_VECTOR_OPS_HPP
