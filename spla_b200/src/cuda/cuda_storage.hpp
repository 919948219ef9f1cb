// cuda_storage.hpp -- registers the device decorations with spla's storage managers: constructors, validators and the
// converter edges CpuDense<->AccDense, CpuCoo<->AccCoo, AccCoo<->AccDense, CpuCsr<->AccCsr. This is the CUDA counterpart
// of the SPLA_BUILD_OPENCL blocks in reference src/storage/storage_manager_vector.hpp:101-171 and
// storage_manager_matrix.hpp:133-159; it uses the managers' public registration API (storage_manager.hpp:64-68), so it
// can be called from those two functions (in-tree) or once from CudaAccelerator::init (what this build does).
#ifndef SPLA_CUDA_STORAGE_HPP
#define SPLA_CUDA_STORAGE_HPP

#include <core/tmatrix.hpp>
#include <core/tvector.hpp>
#include <cpu/cpu_format_coo_vec.hpp>
#include <cpu/cpu_format_csr.hpp>
#include <cpu/cpu_format_dense_vec.hpp>
#include <cpu/cpu_formats.hpp>
#include <cuda/cuda_formats.hpp>

#include <vector>

namespace spla {

    template<typename T>
    void register_formats_vector_cuda(StorageManagerVector<T>& manager) {
        using Storage = typename StorageManagerVector<T>::Storage;

        manager.register_constructor(FormatVector::AccCoo, [](Storage& s) {
            s.get_ref(FormatVector::AccCoo) = make_ref<CudaCooVec<T>>();
        });
        manager.register_constructor(FormatVector::AccDense, [](Storage& s) {
            s.get_ref(FormatVector::AccDense) = make_ref<CudaDenseVec<T>>();
            cuda_dense_vec_resize(s.get_n_rows(), *s.template get<CudaDenseVec<T>>());
        });

        manager.register_validator_discard(FormatVector::AccCoo, [](Storage& s) {
            cuda_coo_vec_clear(*s.template get<CudaCooVec<T>>());
        });
        manager.register_validator(FormatVector::AccDense, [](Storage& s) {
            cuda_dense_vec_fill(s.get_n_rows(), s.get_fill_value(), *s.template get<CudaDenseVec<T>>());
        });

        manager.register_converter(FormatVector::CpuDense, FormatVector::AccDense, [](Storage& s) {
            auto* host = s.template get<CpuDenseVec<T>>();
            cuda_dense_vec_init(s.get_n_rows(), host->Ax.data(), *s.template get<CudaDenseVec<T>>());
        });
        manager.register_converter(FormatVector::AccDense, FormatVector::CpuDense, [](Storage& s) {
            auto* host = s.template get<CpuDenseVec<T>>();
            cpu_dense_vec_resize(s.get_n_rows(), *host);
            cuda_dense_vec_read(s.get_n_rows(), host->Ax.data(), *s.template get<CudaDenseVec<T>>());
        });
        manager.register_converter(FormatVector::CpuCoo, FormatVector::AccCoo, [](Storage& s) {
            auto* host = s.template get<CpuCooVec<T>>();
            cuda_coo_vec_init(host->values, host->Ai.data(), host->Ax.data(), *s.template get<CudaCooVec<T>>());
        });
        manager.register_converter(FormatVector::AccCoo, FormatVector::CpuCoo, [](Storage& s) {
            auto* dev  = s.template get<CudaCooVec<T>>();
            auto* host = s.template get<CpuCooVec<T>>();
            cpu_coo_vec_resize(dev->values, *host);
            cuda_coo_vec_read(dev->values, host->Ai.data(), host->Ax.data(), *dev);
        });
        manager.register_converter(FormatVector::AccCoo, FormatVector::AccDense, [](Storage& s) {
            cuda_coo_vec_to_dense(s.get_n_rows(), s.get_fill_value(), *s.template get<CudaCooVec<T>>(), *s.template get<CudaDenseVec<T>>());
        });
        manager.register_converter(FormatVector::AccDense, FormatVector::AccCoo, [](Storage& s) {
            cuda_dense_vec_to_coo(s.get_n_rows(), s.get_fill_value(), *s.template get<CudaDenseVec<T>>(), *s.template get<CudaCooVec<T>>());
        });
    }

    template<typename T>
    void register_formats_matrix_cuda(StorageManagerMatrix<T>& manager) {
        using Storage = typename StorageManagerMatrix<T>::Storage;

        manager.register_constructor(FormatMatrix::AccCsr, [](Storage& s) {
            s.get_ref(FormatMatrix::AccCsr) = make_ref<CudaCsr<T>>();
        });
        manager.register_converter(FormatMatrix::CpuCsr, FormatMatrix::AccCsr, [](Storage& s) {
            auto* host = s.template get<CpuCsr<T>>();
            cuda_csr_init(s.get_n_rows(), s.get_n_cols(), host->values, host->Ap.data(), host->Aj.data(), host->Ax.data(), *s.template get<CudaCsr<T>>());
        });
        // Direct edges from the formats Matrix::set_* (CpuLil) and Matrix::build (CpuCoo) produce: one hop instead of two, no
        // intermediate host CSR copy, and the entry count is taken from the data itself rather than from TDecoration::values
        // (cpu_coo_to_lil leaves it untouched, reference src/cpu/cpu_format_coo.hpp:58-76, which cpu_lil_to_csr then trusts).
        manager.register_converter(FormatMatrix::CpuLil, FormatMatrix::AccCsr, [](Storage& s) {
            auto*             host = s.template get<CpuLil<T>>();
            const uint        n    = s.get_n_rows();
            std::vector<uint> Ap(std::size_t(n) + 1, 0);
            for (uint i = 0; i < n; ++i) Ap[i + 1] = Ap[i] + uint(host->Ar[i].size());
            std::vector<uint> Aj(Ap[n]);
            std::vector<T>    Ax(Ap[n]);
            for (uint i = 0, k = 0; i < n; ++i)
                for (const auto& entry : host->Ar[i]) {
                    Aj[k] = entry.first;
                    Ax[k] = entry.second;
                    ++k;
                }
            cuda_csr_init(n, s.get_n_cols(), Ap[n], Ap.data(), Aj.data(), Ax.data(), *s.template get<CudaCsr<T>>());
        });
        manager.register_converter(FormatMatrix::CpuCoo, FormatMatrix::AccCsr, [](Storage& s) {
            // device-side ingest (splacu_coo_to_csr): the triplets are uploaded as they are; row-sorted input (a loader, a sorted
            // Matrix::build) needs only the row extents, anything else is sorted stably by row on the device -- the entry order inside
            // a row is the input order, exactly what the host counting sort of the reference chain gives
            auto*             host = s.template get<CpuCoo<T>>();
            auto*             dev  = s.template get<CudaCsr<T>>();
            const uint        n    = s.get_n_rows();
            const std::size_t nnz  = host->Ai.size();
            cuda_csr_init_from_coo(n, s.get_n_cols(), uint(nnz), host->Ai.data(), host->Aj.data(), host->Ax.data(), *dev);
        });
        manager.register_converter(FormatMatrix::AccCsr, FormatMatrix::CpuCsr, [](Storage& s) {
            auto* dev  = s.template get<CudaCsr<T>>();
            auto* host = s.template get<CpuCsr<T>>();
            cpu_csr_resize(s.get_n_rows(), dev->values, *host);
            cuda_csr_read(s.get_n_rows(), dev->values, host->Ap.data(), host->Aj.data(), host->Ax.data(), *dev);
        });
    }

    /** @brief Registers the CUDA formats for the three value types; idempotent */
    void register_formats_cuda();

}// namespace spla

#endif//SPLA_CUDA_STORAGE_HPP
