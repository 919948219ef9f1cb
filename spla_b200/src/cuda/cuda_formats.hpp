// cuda_formats.hpp -- device-resident storage decorations of the CUDA backend and their host<->device / format
// conversions. They occupy the accelerator slots the reference reserves in its format enums
// (FormatVector::AccDense / AccCoo, FormatMatrix::AccCsr, reference include/spla/config.hpp:102-142) and play the role of
// CLDenseVec / CLCooVec / CLCsr (reference src/opencl/cl_formats.hpp:60-103, cl_format_{dense_vec,coo_vec,csr}.hpp).
// Only the C ABI of include/splacu.h is used: no CUDA header is visible to spla's host code.
#ifndef SPLA_CUDA_FORMATS_HPP
#define SPLA_CUDA_FORMATS_HPP

#include <spla/config.hpp>

#include <core/tdecoration.hpp>
#include <cuda/cuda_accelerator.hpp>

#include <cstring>
#include <type_traits>

namespace spla {

    /** @brief spla value type -> splacu_dtype */
    template<typename T>
    constexpr int cuda_dtype() {
        static_assert(sizeof(T) == 4, "spla device values are 4 bytes");
        if constexpr (std::is_same<T, T_INT>::value) return SPLACU_INT;
        else if constexpr (std::is_same<T, T_UINT>::value)
            return SPLACU_UINT;
        else
            return SPLACU_FLOAT;
    }

    /** @brief bit pattern of a value, the form scalars cross the C ABI in */
    template<typename T>
    inline uint32_t cuda_bits(T value) {
        uint32_t bits;
        std::memcpy(&bits, &value, sizeof(bits));
        return bits;
    }

    /** @brief Owning device allocation (RAII); safe to destroy after Library::finalize (reference src/library.cpp:97-104) */
    class CudaBuffer {
    public:
        CudaBuffer() = default;
        ~CudaBuffer() { release(); }
        CudaBuffer(const CudaBuffer&)            = delete;
        CudaBuffer& operator=(const CudaBuffer&) = delete;

        /** make room for `count` 4-byte elements; contents are NOT preserved when the buffer grows. Growth is geometric: a
         *  traversal front that gains a few entries per step (a grid wave front) must not pay a cudaFree + cudaMalloc -- an
         *  implicit device synchronisation -- on every step */
        void reserve(std::size_t count) {
            if (count <= m_capacity && m_ptr) return;
            std::size_t want = m_capacity + m_capacity / 2;
            if (want < count) want = count;
            if (want < 256) want = 256;
            release();
            SPLACU_CALL(splacu_malloc(&m_ptr, want * 4));
            m_capacity = want;
        }
        void release() {
            if (m_ptr) splacu_free(m_ptr);
            m_ptr      = nullptr;
            m_capacity = 0;
        }
        void swap(CudaBuffer& other) {
            std::swap(m_ptr, other.m_ptr);
            std::swap(m_capacity, other.m_capacity);
        }
        void*       get() { return m_ptr; }
        const void* get() const { return m_ptr; }
        uint32_t*   as_index() { return static_cast<uint32_t*>(m_ptr); }

    private:
        void*       m_ptr      = nullptr;
        std::size_t m_capacity = 0;
    };

    /** @brief Dense device vector: T Ax[n_rows] */
    template<typename T>
    class CudaDenseVec : public TDecoration<T> {
    public:
        static constexpr FormatVector FORMAT = FormatVector::AccDense;
        ~CudaDenseVec() override             = default;
        CudaBuffer Ax;
    };

    /** @brief Sparse device vector: uint Ai[values] ascending, T Ax[values] */
    template<typename T>
    class CudaCooVec : public TDecoration<T> {
    public:
        static constexpr FormatVector FORMAT = FormatVector::AccCoo;
        ~CudaCooVec() override               = default;
        CudaBuffer Ai;
        CudaBuffer Ax;
    };

    /** @brief Device CSR: uint Ap[n_rows+1], uint Aj[values], T Ax[values] + the backend's load-balancing / hub metadata */
    template<typename T>
    class CudaCsr : public TDecoration<T> {
    public:
        static constexpr FormatMatrix FORMAT = FormatMatrix::AccCsr;
        ~CudaCsr() override {
            if (dhandle) splacu_dcsr_destroy(dhandle);
            if (handle) splacu_csr_destroy(handle);
        }
        CudaBuffer  Ap;
        CudaBuffer  Aj;
        CudaBuffer  Ax;
        splacu_csr  handle  = nullptr;
        splacu_dcsr dhandle = nullptr;// the matrix sharded over the accelerator's device group, built at the first product that uses it
        uint        n_rows = 0, n_cols = 0;

        /** the sharded handle when the accelerator drives several devices, else nullptr */
        splacu_dcsr sharded() {
            splacu_dist group = get_acc_cuda()->get_group();
            if (!group) return nullptr;
            if (!dhandle)
                SPLACU_CALL(splacu_dcsr_create(&dhandle, group, n_rows, n_cols, this->values, Ap.as_index(), Aj.as_index(), Ax.get(), nullptr));
            return dhandle;
        }
    };

    // ---- dense vector ------------------------------------------------------------------------------------------------

    template<typename T>
    void cuda_dense_vec_resize(uint n_rows, CudaDenseVec<T>& storage) {
        storage.Ax.reserve(n_rows);
    }
    template<typename T>
    void cuda_dense_vec_fill(uint n_rows, T value, CudaDenseVec<T>& storage) {
        storage.Ax.reserve(n_rows);
        SPLACU_CALL(splacu_fill(storage.Ax.get(), cuda_bits(value), n_rows, nullptr));
    }
    template<typename T>
    void cuda_dense_vec_init(uint n_rows, const T* values, CudaDenseVec<T>& storage) {
        storage.Ax.reserve(n_rows);
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ax.get(), values, std::size_t(n_rows) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));// the source is pageable host memory owned by the CPU decoration
    }
    template<typename T>
    void cuda_dense_vec_read(uint n_rows, T* values, const CudaDenseVec<T>& storage) {
        SPLACU_CALL(splacu_memcpy_d2h(values, storage.Ax.get(), std::size_t(n_rows) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));
    }

    // ---- sparse vector -----------------------------------------------------------------------------------------------

    template<typename T>
    void cuda_coo_vec_clear(CudaCooVec<T>& storage) {
        storage.values = 0;
    }
    template<typename T>
    void cuda_coo_vec_resize(uint n_values, CudaCooVec<T>& storage) {
        storage.Ai.reserve(n_values);
        storage.Ax.reserve(n_values);
        storage.values = n_values;
    }
    template<typename T>
    void cuda_coo_vec_init(uint n_values, const uint* Ai, const T* Ax, CudaCooVec<T>& storage) {
        cuda_coo_vec_resize(n_values, storage);
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ai.get(), Ai, std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ax.get(), Ax, std::size_t(n_values) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));
    }
    template<typename T>
    void cuda_coo_vec_read(uint n_values, uint* Ai, T* Ax, const CudaCooVec<T>& storage) {
        SPLACU_CALL(splacu_memcpy_d2h(Ai, storage.Ai.get(), std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_d2h(Ax, storage.Ax.get(), std::size_t(n_values) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));
    }
    /** sparse -> dense on the device; missing entries take the vector's fill value (reference storage_manager_vector.hpp:159-164) */
    template<typename T>
    void cuda_coo_vec_to_dense(uint n_rows, T fill_value, const CudaCooVec<T>& in, CudaDenseVec<T>& out) {
        out.Ax.reserve(n_rows);
        SPLACU_CALL(splacu_coo_to_dense(n_rows, cuda_bits(fill_value), in.values, static_cast<const uint32_t*>(in.Ai.get()), in.Ax.get(), out.Ax.get(), nullptr));
    }
    /** dense -> sparse on the device, ascending indices like the CPU converter (reference src/cpu/cpu_format_dense_vec.hpp:53-67) */
    template<typename T>
    void cuda_dense_vec_to_coo(uint n_rows, T fill_value, const CudaDenseVec<T>& in, CudaCooVec<T>& out) {
        splacu_workspace ws = get_acc_cuda()->get_workspace();
        uint32_t         count = 0;
        SPLACU_CALL(splacu_dense_to_coo_count(cuda_dtype<T>(), n_rows, cuda_bits(fill_value), in.Ax.get(), ws, &count, nullptr));
        cuda_coo_vec_resize(count, out);
        SPLACU_CALL(splacu_dense_to_coo_emit(cuda_dtype<T>(), n_rows, cuda_bits(fill_value), in.Ax.get(), ws, out.Ai.as_index(), out.Ax.get(), nullptr));
    }

    // ---- csr matrix ---------------------------------------------------------------------------------------------------

    template<typename T>
    void cuda_csr_init(uint n_rows, uint n_cols, uint n_values, const uint* Ap, const uint* Aj, const T* Ax, CudaCsr<T>& storage) {
        if (storage.dhandle) {
            splacu_dcsr_destroy(storage.dhandle);
            storage.dhandle = nullptr;
        }
        if (storage.handle) {
            splacu_csr_destroy(storage.handle);
            storage.handle = nullptr;
        }
        storage.n_rows = n_rows;
        storage.n_cols = n_cols;
        storage.Ap.reserve(std::size_t(n_rows) + 1);
        storage.Aj.reserve(n_values);
        storage.Ax.reserve(n_values);
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ap.get(), Ap, (std::size_t(n_rows) + 1) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_h2d(storage.Aj.get(), Aj, std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ax.get(), Ax, std::size_t(n_values) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));
        storage.values = n_values;
        SPLACU_CALL(splacu_csr_create(&storage.handle, n_rows, n_cols, n_values, storage.Ap.as_index(), storage.Aj.as_index(), storage.Ax.get(), nullptr));
    }
    /** triplets on the host -> CSR on the device (device-side ingest, splacu_coo_to_csr) */
    template<typename T>
    void cuda_csr_init_from_coo(uint n_rows, uint n_cols, uint n_values, const uint* Ai, const uint* Aj, const T* Ax, CudaCsr<T>& storage) {
        if (storage.dhandle) {
            splacu_dcsr_destroy(storage.dhandle);
            storage.dhandle = nullptr;
        }
        if (storage.handle) {
            splacu_csr_destroy(storage.handle);
            storage.handle = nullptr;
        }
        storage.n_rows = n_rows;
        storage.n_cols = n_cols;
        storage.Ap.reserve(std::size_t(n_rows) + 1);
        storage.Aj.reserve(n_values);
        storage.Ax.reserve(n_values);
        CudaBuffer rows;
        rows.reserve(n_values);
        SPLACU_CALL(splacu_memcpy_h2d(rows.get(), Ai, std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_h2d(storage.Aj.get(), Aj, std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_h2d(storage.Ax.get(), Ax, std::size_t(n_values) * sizeof(T), nullptr));
        splacu_workspace ws     = get_acc_cuda()->get_workspace();
        int              sorted = 1;
        // try in place first: row-sorted triplets only need their row extents
        int rc = splacu_coo_to_csr(n_rows, n_values, rows.as_index(), storage.Aj.as_index(), storage.Ax.get(), storage.Ap.as_index(), storage.Aj.as_index(),
                                   storage.Ax.get(), ws, &sorted, nullptr);
        if (rc != 0) {
            // unsorted rows: stable sort into fresh buffers
            CudaBuffer Aj2, Ax2;
            Aj2.reserve(n_values);
            Ax2.reserve(n_values);
            SPLACU_CALL(splacu_coo_to_csr(n_rows, n_values, rows.as_index(), storage.Aj.as_index(), storage.Ax.get(), storage.Ap.as_index(), Aj2.as_index(),
                                          Ax2.get(), ws, &sorted, nullptr));
            storage.Aj.swap(Aj2);
            storage.Ax.swap(Ax2);
        }
        SPLACU_CALL(splacu_sync(nullptr));// the sources are pageable host memory owned by the CPU decoration
        storage.values = n_values;
        SPLACU_CALL(splacu_csr_create(&storage.handle, n_rows, n_cols, n_values, storage.Ap.as_index(), storage.Aj.as_index(), storage.Ax.get(), nullptr));
    }

    template<typename T>
    void cuda_csr_read(uint n_rows, uint n_values, uint* Ap, uint* Aj, T* Ax, const CudaCsr<T>& storage) {
        SPLACU_CALL(splacu_memcpy_d2h(Ap, storage.Ap.get(), (std::size_t(n_rows) + 1) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_d2h(Aj, storage.Aj.get(), std::size_t(n_values) * sizeof(uint), nullptr));
        SPLACU_CALL(splacu_memcpy_d2h(Ax, storage.Ax.get(), std::size_t(n_values) * sizeof(T), nullptr));
        SPLACU_CALL(splacu_sync(nullptr));
    }

}// namespace spla

#endif//SPLA_CUDA_FORMATS_HPP
