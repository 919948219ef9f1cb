// cuda_vxm.hpp -- masked sparse-vector matrix product (push, SpMSpV) on the CUDA device.
// Registered as "vxm_masked_{I,U,F}__cuda" beside Algo_vxm_masked_cpu (reference src/cpu/cpu_vxm.hpp:58-128, the semantics)
// and in place of Algo_vxm_masked_cl (reference src/opencl/cl_vxm.hpp:56-208). The result count is a host `uint`
// (TDecoration::values, reference src/core/tdecoration.hpp:58), hence the two-phase call with ONE 4-byte device->host read.
#ifndef SPLA_CUDA_VXM_HPP
#define SPLA_CUDA_VXM_HPP

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
    class Algo_vxm_masked_cuda final : public RegistryAlgo {
    public:
        ~Algo_vxm_masked_cuda() override = default;

        std::string get_name() override { return "vxm_masked"; }
        std::string get_description() override { return "parallel vector-matrix masked product on cuda device (sm_100a)"; }

        Status execute(const DispatchContext& ctx) override {
            TIME_PROFILE_SCOPE("cuda/vxm");

            auto t = ctx.task.template cast_safe<ScheduleTask_vxm_masked>();

            auto r           = t->r.template cast_safe<TVector<T>>();
            auto mask        = t->mask.template cast_safe<TVector<T>>();
            auto v           = t->v.template cast_safe<TVector<T>>();
            auto M           = t->M.template cast_safe<TMatrix<T>>();
            auto op_multiply = t->op_multiply.template cast_safe<TOpBinary<T, T, T>>();
            auto op_add      = t->op_add.template cast_safe<TOpBinary<T, T, T>>();
            auto op_select   = t->op_select.template cast_safe<TOpSelect<T>>();
            auto init        = t->init.template cast_safe<TScalar<T>>();// read, never used: reference src/cpu/cpu_vxm.hpp:72
            (void) init;

            // built-ins by id (ahead-of-time specialised kernels), user-defined ops by source text (NVRTC, csrc/jit.cu)
            CudaOpDesc d_mult(op_multiply.get()), d_add(op_add.get()), d_sel(op_select.get());

            r->validate_wd(FormatVector::AccCoo);
            mask->validate_rw(FormatVector::AccDense);
            M->validate_rw(FormatMatrix::AccCsr);
            v->validate_rw(FormatVector::AccCoo);

            auto*       p_r    = r->template get<CudaCooVec<T>>();
            const auto* p_mask = mask->template get<CudaDenseVec<T>>();
            auto*       p_M    = M->template get<CudaCsr<T>>();
            const auto* p_v    = v->template get<CudaCooVec<T>>();

            auto*            acc = get_acc_cuda();
            splacu_workspace ws  = acc->get_workspace();
            uint32_t         nr  = 0;

            // several devices: columns sharded, every shard expands the whole frontier against its slice
            if (!d_mult.user_defined() && !d_add.user_defined() && !d_sel.user_defined()) {
                if (splacu_dcsr sharded = p_M->sharded()) {
                    SPLACU_CALL(splacu_dist_vxm_masked_begin(sharded, cuda_dtype<T>(), d_mult.get()->id, d_add.get()->id, d_sel.get()->id, p_v->values,
                                                             static_cast<const uint32_t*>(p_v->Ai.get()), p_v->Ax.get(), p_mask->Ax.get(), &nr,
                                                             acc->get_stream()));
                    cuda_coo_vec_resize(nr, *p_r);
                    SPLACU_CALL(splacu_dist_vxm_masked_emit(sharded, p_r->Ai.as_index(), p_r->Ax.get(), acc->get_stream()));
                    return Status::Ok;
                }
            }

            SPLACU_CALL_OPS(splacu_vxm_masked_begin_ops(p_M->handle, cuda_dtype<T>(), d_mult.get(), d_add.get(), d_sel.get(),
                                                        p_v->values, static_cast<const uint32_t*>(p_v->Ai.get()), p_v->Ax.get(),
                                                        p_mask->Ax.get(), ws, &nr, acc->get_stream()));
            cuda_coo_vec_resize(nr, *p_r);
            SPLACU_CALL(splacu_vxm_masked_emit(ws, p_r->Ai.as_index(), p_r->Ax.get(), acc->get_stream()));
            return Status::Ok;
        }
    };

}// namespace spla

#endif//SPLA_CUDA_VXM_HPP
