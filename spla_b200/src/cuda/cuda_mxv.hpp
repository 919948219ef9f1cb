// cuda_mxv.hpp -- masked matrix-vector product (pull) on the CUDA device.
// Registered as "mxv_masked_{I,U,F}__cuda" beside Algo_mxv_masked_cpu (reference src/cpu/cpu_mxv.hpp:56-106, the semantics)
// and in place of Algo_mxv_masked_cl (reference src/opencl/cl_mxv.hpp:53-278). One C-ABI call; no host synchronisation.
#ifndef SPLA_CUDA_MXV_HPP
#define SPLA_CUDA_MXV_HPP

#include <schedule/schedule_tasks.hpp>

#include <core/dispatcher.hpp>
#include <core/registry.hpp>
#include <core/tmatrix.hpp>
#include <core/top.hpp>
#include <core/tscalar.hpp>
#include <core/ttype.hpp>
#include <core/tvector.hpp>
#include <cuda/cuda_formats.hpp>
#include <cuda/cuda_ops.hpp>
#include <profiling/time_profiler.hpp>

namespace spla {

    template<typename T>
    class Algo_mxv_masked_cuda final : public RegistryAlgo {
    public:
        ~Algo_mxv_masked_cuda() override = default;

        std::string get_name() override { return "mxv_masked"; }
        std::string get_description() override { return "parallel matrix-vector masked product on cuda device (sm_100a)"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/mxv");

            auto t = ctx.task.template cast_safe<ScheduleTask_mxv_masked>();

            auto r           = t->r.template cast_safe<TVector<T>>();
            auto mask        = t->mask.template cast_safe<TVector<T>>();
            auto M           = t->M.template cast_safe<TMatrix<T>>();
            auto v           = t->v.template cast_safe<TVector<T>>();
            auto op_multiply = t->op_multiply.template cast_safe<TOpBinary<T, T, T>>();
            auto op_add      = t->op_add.template cast_safe<TOpBinary<T, T, T>>();
            auto op_select   = t->op_select.template cast_safe<TOpSelect<T>>();
            auto init        = t->init.template cast_safe<TScalar<T>>();

            // built-ins by id (ahead-of-time specialised kernels), user-defined ops by source text (NVRTC, csrc/jit.cu)
            CudaOpDesc d_mult(op_multiply.get()), d_add(op_add.get()), d_sel(op_select.get());

            r->validate_wd(FormatVector::AccDense);
            mask->validate_rw(FormatVector::AccDense);
            M->validate_rw(FormatMatrix::AccCsr);
            v->validate_rw(FormatVector::AccDense);

            auto*       p_r    = r->template get<CudaDenseVec<T>>();
            const auto* p_mask = mask->template get<CudaDenseVec<T>>();
            auto*       p_M    = M->template get<CudaCsr<T>>();
            const auto* p_v    = v->template get<CudaDenseVec<T>>();

            const bool early_exit = t->get_desc_or_default()->get_early_exit();

            // several devices (CudaAccelerator::set_queues_count / SPLA_CUDA_DEVICES): rows sharded, v broadcast over NVLink
            if (!d_mult.user_defined() && !d_add.user_defined() && !d_sel.user_defined()) {
                if (splacu_dcsr sharded = p_M->sharded()) {
                    SPLACU_CALL(splacu_dist_mxv_masked(sharded, cuda_dtype<T>(), d_mult.get()->id, d_add.get()->id, d_sel.get()->id,
                                                       p_v->Ax.get(), p_mask->Ax.get(), p_r->Ax.get(), cuda_bits(init->get_value()),
                                                       early_exit ? 1 : 0, get_acc_cuda()->get_stream()));
                    return Status::Ok;
                }
            }

            SPLACU_CALL_OPS(splacu_mxv_masked_ops(p_M->handle, cuda_dtype<T>(), d_mult.get(), d_add.get(), d_sel.get(),
                                                  p_v->Ax.get(), p_mask->Ax.get(), p_r->Ax.get(),
                                                  cuda_bits(init->get_value()), early_exit ? 1 : 0, get_acc_cuda()->get_stream()));
            return Status::Ok;
        }
    };

}// namespace spla

#endif//SPLA_CUDA_MXV_HPP
