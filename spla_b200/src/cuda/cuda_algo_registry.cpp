// cuda_algo_registry.cpp -- see cuda_algo_registry.hpp. Key grammar: task name + "_" + type code + "__cuda"
// (reference src/core/registry.hpp:40-55, src/schedule/schedule_tasks.cpp:131-175); Dispatcher::dispatch looks the
// accelerator's suffix up first and only falls back to "__cpu" for keys that are absent (src/core/dispatcher.cpp:51-59),
// so all six hot-path keys and their neighbours are present here and none of them defers to the CPU internally.
#include "cuda_algo_registry.hpp"

#include <cuda/cuda_mxv.hpp>
#include <cuda/cuda_ops.hpp>
#include <cuda/cuda_storage.hpp>
#include <cuda/cuda_vector_ops.hpp>
#include <cuda/cuda_vxm.hpp>

#include <unordered_map>

namespace spla {

    void register_algo_cuda(Registry* g_registry) {
        // the hot path
        g_registry->add(MAKE_KEY_CUDA_0("mxv_masked", INT), std::make_shared<Algo_mxv_masked_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("mxv_masked", UINT), std::make_shared<Algo_mxv_masked_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("mxv_masked", FLOAT), std::make_shared<Algo_mxv_masked_cuda<T_FLOAT>>());

        g_registry->add(MAKE_KEY_CUDA_0("vxm_masked", INT), std::make_shared<Algo_vxm_masked_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("vxm_masked", UINT), std::make_shared<Algo_vxm_masked_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("vxm_masked", FLOAT), std::make_shared<Algo_vxm_masked_cuda<T_FLOAT>>());

        // its neighbours inside the bfs / sssp / pr loops
        g_registry->add(MAKE_KEY_CUDA_0("v_assign_masked", INT), std::make_shared<Algo_v_assign_masked_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_assign_masked", UINT), std::make_shared<Algo_v_assign_masked_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_assign_masked", FLOAT), std::make_shared<Algo_v_assign_masked_cuda<T_FLOAT>>());

        g_registry->add(MAKE_KEY_CUDA_0("v_count_mf", INT), std::make_shared<Algo_v_count_mf_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_count_mf", UINT), std::make_shared<Algo_v_count_mf_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_count_mf", FLOAT), std::make_shared<Algo_v_count_mf_cuda<T_FLOAT>>());

        g_registry->add(MAKE_KEY_CUDA_0("v_eadd_fdb", INT), std::make_shared<Algo_v_eadd_fdb_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_eadd_fdb", UINT), std::make_shared<Algo_v_eadd_fdb_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_eadd_fdb", FLOAT), std::make_shared<Algo_v_eadd_fdb_cuda<T_FLOAT>>());

        g_registry->add(MAKE_KEY_CUDA_0("v_eadd", INT), std::make_shared<Algo_v_eadd_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_eadd", UINT), std::make_shared<Algo_v_eadd_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_eadd", FLOAT), std::make_shared<Algo_v_eadd_cuda<T_FLOAT>>());

        g_registry->add(MAKE_KEY_CUDA_0("v_reduce", INT), std::make_shared<Algo_v_reduce_cuda<T_INT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_reduce", UINT), std::make_shared<Algo_v_reduce_cuda<T_UINT>>());
        g_registry->add(MAKE_KEY_CUDA_0("v_reduce", FLOAT), std::make_shared<Algo_v_reduce_cuda<T_FLOAT>>());
    }

    void register_formats_cuda() {
        static bool done = false;
        if (done) return;
        done = true;
        register_formats_vector_cuda(*TVector<T_INT>::get_storage_manager());
        register_formats_vector_cuda(*TVector<T_UINT>::get_storage_manager());
        register_formats_vector_cuda(*TVector<T_FLOAT>::get_storage_manager());
        register_formats_matrix_cuda(*TMatrix<T_INT>::get_storage_manager());
        register_formats_matrix_cuda(*TMatrix<T_UINT>::get_storage_manager());
        register_formats_matrix_cuda(*TMatrix<T_FLOAT>::get_storage_manager());
    }

    // ---- built-in op lookup (identity of the op OBJECT, so a user op that reuses a built-in name is never mistaken) ----

    namespace {
        struct OpTables {
            std::unordered_map<const void*, int> bin;
            std::unordered_map<const void*, int> sel;
            OpTables() {
#define SPLA_CUDA_BIN3(NAME, ID)       \
    bin[NAME##_INT.get()]   = (ID);    \
    bin[NAME##_UINT.get()]  = (ID);    \
    bin[NAME##_FLOAT.get()] = (ID);
#define SPLA_CUDA_BIN2(NAME, ID)      \
    bin[NAME##_INT.get()]  = (ID);    \
    bin[NAME##_UINT.get()] = (ID);
#define SPLA_CUDA_SEL3(NAME, ID)       \
    sel[NAME##_INT.get()]   = (ID);    \
    sel[NAME##_UINT.get()]  = (ID);    \
    sel[NAME##_FLOAT.get()] = (ID);
                SPLA_CUDA_BIN3(PLUS, SPLACU_PLUS)
                SPLA_CUDA_BIN3(MINUS, SPLACU_MINUS)
                SPLA_CUDA_BIN3(MULT, SPLACU_MULT)
                SPLA_CUDA_BIN3(DIV, SPLACU_DIV)
                SPLA_CUDA_BIN3(MINUS_POW2, SPLACU_MINUS_POW2)
                SPLA_CUDA_BIN3(FIRST, SPLACU_FIRST)
                SPLA_CUDA_BIN3(SECOND, SPLACU_SECOND)
                SPLA_CUDA_BIN3(BONE, SPLACU_BONE)
                SPLA_CUDA_BIN3(MIN, SPLACU_MIN)
                SPLA_CUDA_BIN3(MAX, SPLACU_MAX)
                SPLA_CUDA_BIN3(LOR, SPLACU_LOR)
                SPLA_CUDA_BIN3(LAND, SPLACU_LAND)
                SPLA_CUDA_BIN2(BOR, SPLACU_BOR)
                SPLA_CUDA_BIN2(BAND, SPLACU_BAND)
                SPLA_CUDA_BIN2(BXOR, SPLACU_BXOR)
                SPLA_CUDA_SEL3(EQZERO, SPLACU_EQZERO)
                SPLA_CUDA_SEL3(NQZERO, SPLACU_NQZERO)
                SPLA_CUDA_SEL3(GTZERO, SPLACU_GTZERO)
                SPLA_CUDA_SEL3(GEZERO, SPLACU_GEZERO)
                SPLA_CUDA_SEL3(LTZERO, SPLACU_LTZERO)
                SPLA_CUDA_SEL3(LEZERO, SPLACU_LEZERO)
                SPLA_CUDA_SEL3(ALWAYS, SPLACU_ALWAYS)
                SPLA_CUDA_SEL3(NEVER, SPLACU_NEVER)
#undef SPLA_CUDA_BIN3
#undef SPLA_CUDA_BIN2
#undef SPLA_CUDA_SEL3
            }
        };
        const OpTables& op_tables() {
            static OpTables tables;// built on first use, after register_ops() (reference src/op.cpp:157)
            return tables;
        }
    }// namespace

    Status cuda_defer_to_cpu(const DispatchContext& ctx) {
        auto algo = Library::get()->get_registry()->find(ctx.task->get_key() + CPU_SUFFIX);
        if (!algo) return Status::NotImplemented;
        return algo->execute(ctx);
    }

    int cuda_find_binop(const OpBinary* op) {
        auto it = op_tables().bin.find(op);
        return it == op_tables().bin.end() ? -1 : it->second;
    }
    int cuda_find_selop(const OpSelect* op) {
        auto it = op_tables().sel.find(op);
        return it == op_tables().sel.end() ? -1 : it->second;
    }

}// namespace spla
