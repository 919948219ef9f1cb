// ingest.cu -- device-side COO -> CSR ("Matrix::build -> device CSR without LIL").
//
// Matrix::build leaves the user's triplets in a CpuCoo decoration (reference src/core/tmatrix.hpp:220-253); the reference then
// reaches a device CSR through host conversions (CpuCoo -> CpuLil -> CpuCsr -> AccCsr, src/storage/storage_manager_matrix.hpp:
// 133-159, a vector-of-vectors rebuild that takes seconds at RMAT-24). Here the triplets are uploaded as they are and turned into
// CSR on the device:
//   rows already non-decreasing (what a loader or a sorted build gives)  Ap from the row boundaries, Aj / Ax ARE the CSR arrays
//                                                                        (no data movement at all)
//   otherwise                                                            stable radix sort of (row, position) pairs, gather of Aj / Ax
// In both cases the order of the entries inside a row is the input order, exactly what the reference's stable host counting sort
// (src/cpu/cpu_format_coo.hpp:58-76) produces, so every fold sees the same sequence.
#include "common.cuh"
#include "profile.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace splacu {
    namespace {
        constexpr int kBlock = 256;

        __global__ void __launch_bounds__(kBlock) coo_check_sorted_kernel(const uint32_t* __restrict__ Ai, uint32_t nnz, uint32_t n_rows, uint32_t* __restrict__ flags) {
            const uint32_t stride = gridDim.x * blockDim.x;
            uint32_t       f      = 0;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
                const uint32_t i = Ai[k];
                if (i >= n_rows) f |= 2u;// row id out of range
                if (k > 0 && Ai[k - 1] > i) f |= 1u;
            }
            if (f) atomicOr(flags, f);
        }
        // sorted rows: Ap[row] = first position of a row >= row
        __global__ void __launch_bounds__(kBlock) coo_boundaries_kernel(const uint32_t* __restrict__ Ai, uint32_t nnz, uint32_t n_rows, uint32_t* __restrict__ Ap) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k <= nnz; k += stride) {
                const uint32_t lo = k == 0 ? 0u : Ai[k - 1] + 1u;        // rows (previous row, this row] start at k
                const uint32_t hi = k == nnz ? n_rows : Ai[k];            // the sentinel position closes the trailing empty rows
                for (uint32_t row = lo; row <= hi && row <= n_rows; ++row) Ap[row] = k;
            }
        }
        __global__ void __launch_bounds__(kBlock) coo_hist_kernel(const uint32_t* __restrict__ Ai, uint32_t nnz, uint32_t* __restrict__ cnt) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&cnt[Ai[k]], 1u);
        }
        __global__ void __launch_bounds__(kBlock) iota_kernel(uint32_t* __restrict__ x, uint32_t n) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) x[k] = k;
        }
        __global__ void __launch_bounds__(kBlock) gather2_kernel(const uint32_t* __restrict__ idx, uint32_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                                 uint32_t* __restrict__ oa, uint32_t* __restrict__ ob) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
                const uint32_t q = idx[k];
                oa[k]            = a[q];
                ob[k]            = b[q];
            }
        }
    }// namespace
}// namespace splacu

using namespace splacu;

extern "C" int splacu_coo_to_csr(uint32_t n_rows, uint32_t nnz, const uint32_t* d_Ai, const uint32_t* d_Aj, const void* d_Ax, uint32_t* d_Ap, uint32_t* d_Aj_out,
                                 void* d_Ax_out, splacu_workspace wsh, int* was_sorted, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/coo_to_csr", resolve_stream(stream));
    SPLACU_REQUIRE(d_Ap && wsh, "null pointer");
    SPLACU_REQUIRE(nnz == 0 || (d_Ai && d_Aj && d_Ax && d_Aj_out && d_Ax_out), "null coo pointers");
    Workspace*   ws = reinterpret_cast<Workspace*>(wsh);
    cudaStream_t s  = resolve_stream(stream);
    if (was_sorted) *was_sorted = 1;
    if (nnz == 0) {
        SPLACU_CUDA(cudaMemsetAsync(d_Ap, 0, ((size_t) n_rows + 1) * 4, s));
        return SPLACU_OK;
    }
    uint32_t* flags = ws->d_scalars + 4;
    SPLACU_CUDA(cudaMemsetAsync(flags, 0, 4, s));
    coo_check_sorted_kernel<<<grid_for(nnz, kBlock, 8), kBlock, 0, s>>>(d_Ai, nnz, n_rows, flags);
    SPLACU_LAUNCH_CHECK();
    SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars + 4, flags, 4, cudaMemcpyDeviceToHost, s));
    SPLACU_CUDA(cudaStreamSynchronize(s));
    const uint32_t f = ws->h_scalars[4];
    SPLACU_REQUIRE(!(f & 2u), "a row index of the triplets is outside the matrix");
    if (!(f & 1u)) {
        coo_boundaries_kernel<<<grid_for((size_t) nnz + 1, kBlock, 8), kBlock, 0, s>>>(d_Ai, nnz, n_rows, d_Ap);
        SPLACU_LAUNCH_CHECK();
        if (d_Aj_out != d_Aj) SPLACU_CUDA(cudaMemcpyAsync(d_Aj_out, d_Aj, (size_t) nnz * 4, cudaMemcpyDeviceToDevice, s));
        if (d_Ax_out != d_Ax) SPLACU_CUDA(cudaMemcpyAsync(d_Ax_out, d_Ax, (size_t) nnz * 4, cudaMemcpyDeviceToDevice, s));
        return SPLACU_OK;
    }
    if (was_sorted) *was_sorted = 0;
    SPLACU_REQUIRE(d_Aj_out != d_Aj && d_Ax_out != d_Ax, "unsorted triplets cannot be converted in place");
    // row extents: histogram + exclusive scan
    int rc;
    SPLACU_CUDA(cudaMemsetAsync(d_Ap, 0, ((size_t) n_rows + 1) * 4, s));
    coo_hist_kernel<<<grid_for(nnz, kBlock, 8), kBlock, 0, s>>>(d_Ai, nnz, d_Ap);
    SPLACU_LAUNCH_CHECK();
    if ((rc = scan_exclusive_u32(ws, d_Ap, d_Ap, n_rows + 1, nullptr, s))) return rc;
    // stable sort of (row, position): the sorted positions are the gather list
    uint32_t *keys_out = nullptr, *idx = nullptr, *idx_out = nullptr;
    void*     tmp = nullptr;
    auto      cleanup = [&]() {
        cudaFree(keys_out);
        cudaFree(idx);
        cudaFree(idx_out);
        cudaFree(tmp);
    };
#define ING_CUDA(expr)                                                   \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) {                                         \
            cleanup();                                                   \
            return ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);   \
        }                                                                \
    } while (0)
    ING_CUDA(cudaMalloc(&keys_out, (size_t) nnz * 4));
    ING_CUDA(cudaMalloc(&idx, (size_t) nnz * 4));
    ING_CUDA(cudaMalloc(&idx_out, (size_t) nnz * 4));
    iota_kernel<<<grid_for(nnz, kBlock, 8), kBlock, 0, s>>>(idx, nnz);
    count_launch();
    int end_bit = 1;
    while (end_bit < 32 && ((n_rows - 1u) >> end_bit) != 0u) ++end_bit;
    size_t tmp_bytes = 0;
    ING_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_Ai, keys_out, idx, idx_out, (int) nnz, 0, end_bit, s));
    ING_CUDA(cudaMalloc(&tmp, tmp_bytes));
    ING_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_Ai, keys_out, idx, idx_out, (int) nnz, 0, end_bit, s));
    count_launch(8);
    gather2_kernel<<<grid_for(nnz, kBlock, 8), kBlock, 0, s>>>(idx_out, nnz, d_Aj, static_cast<const uint32_t*>(d_Ax), d_Aj_out, static_cast<uint32_t*>(d_Ax_out));
    count_launch();
    ING_CUDA(cudaStreamSynchronize(s));
#undef ING_CUDA
    cleanup();
    return SPLACU_OK;
}
