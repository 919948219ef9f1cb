// mxv_pull.cu -- masked semiring pull product r = M x v (spla exec_mxv_masked).
//
// Semantics: reference src/cpu/cpu_mxv.hpp:88-103 (see include/splacu.h). Replaces the OpenCL kernels
// mxv_vector / mxv_config / mxv_config_scalar (reference src/opencl/kernels/mxv.cl:43-170), which keep only
// 512 work-groups x 32 lanes in flight and need a compaction pass + blocking read for the early-exit case.
//
// Kernels
//   mxv_rows_kernel<LANES>  LANES (1..32) lanes per row, rows of a warp consecutive; the mask is tested before
//                           Ap / Aj / Ax of a row are touched; shuffle tree with the add functor; for associative
//                           + commutative adds: r = add(init, reduce(products)), rows without products get init.
//   mxv_seq_kernel          one thread per row, strict left-to-right fold: the exact path for early_exit and for
//                           non-associative adds (MINUS, DIV, FIRST, SECOND, BONE, MINUS_POW2).
#include "common.cuh"
#include "ops.cuh"

namespace splacu {

    static constexpr int kBlock = 256;

    int csr_build_metadata(Csr* M, cudaStream_t s) {
        (void) s;
        M->avg_row_nnz = M->n_rows ? (float) M->nnz / (float) M->n_rows : 0.f;
        return 0;
    }

    template<typename T, typename S, int LANES>
    __global__ void __launch_bounds__(kBlock) mxv_rows_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj,
                                                              const T* __restrict__ Ax, const T* __restrict__ v, const T* __restrict__ mask,
                                                              T* __restrict__ r, T init, uint32_t n_rows) {
        constexpr uint32_t G       = 32 / LANES;// rows per warp per step
        const uint32_t     lane    = threadIdx.x & 31u;
        const uint32_t     sub     = lane % LANES;
        const uint32_t     g       = lane / LANES;
        const uint32_t     warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const uint32_t     n_warps = (gridDim.x * blockDim.x) >> 5;

        // `first` is warp-uniform, so every lane of a warp runs the same number of iterations and the
        // shuffles below are always executed by the full warp.
        for (uint32_t first = warp * G; first < n_rows; first += n_warps * G) {
            const uint32_t row  = first + g;
            bool           take = row < n_rows;
            if (take) take = sel.reads_mask ? sel.test(mask[row]) : (sel.classes != 0u);

            T        acc = sr.identity();
            uint32_t k0 = 0, k1 = 0;
            if (take) {
                k0 = Ap[row];
                k1 = Ap[row + 1];
#pragma unroll 4
                for (uint32_t k = k0 + sub; k < k1; k += LANES) acc = sr.add(acc, sr.mult(Ax[k], v[Aj[k]]));
            }
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));

            if (sub == 0 && row < n_rows) r[row] = (k1 > k0) ? sr.add(init, acc) : init;
        }
    }

    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_seq_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj,
                                                             const T* __restrict__ Ax, const T* __restrict__ v, const T* __restrict__ mask,
                                                             T* __restrict__ r, T init, uint32_t n_rows, int early_exit) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
            T          sum  = init;
            const bool take = sel.reads_mask ? sel.test(mask[row]) : (sel.classes != 0u);
            if (take) {
                const uint32_t k1 = Ap[row + 1];
                for (uint32_t k = Ap[row]; k < k1; ++k) {
                    sum = sr.add(sum, sr.mult(Ax[k], v[Aj[k]]));
                    if (early_exit && value_neq(sum, init)) break;
                }
            }
            r[row] = sum;
        }
    }

    template<typename T, typename S>
    static int launch_rows(S sr, Select sel, const Csr* M, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        const float avg = M->avg_row_nnz;
#define SPLACU_ROWS(L)                                                                                              \
    do {                                                                                                            \
        const size_t threads = (size_t) ((M->n_rows + (32 / L) - 1) / (32 / L)) * 32;                               \
        mxv_rows_kernel<T, S, L><<<grid_for(threads, kBlock, 8), kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, reinterpret_cast<const T*>(M->Ax), \
                                                                                v, mask, r, init, M->n_rows);      \
    } while (0)
        if (avg <= 2.f) SPLACU_ROWS(2);
        else if (avg <= 6.f) SPLACU_ROWS(4);
        else if (avg <= 12.f) SPLACU_ROWS(8);
        else if (avg <= 48.f) SPLACU_ROWS(16);
        else SPLACU_ROWS(32);
#undef SPLACU_ROWS
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

}// namespace splacu

using namespace splacu;

extern "C" int splacu_mxv_masked(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select,
                                 const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits, int early_exit, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(handle, "null matrix handle");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    SPLACU_REQUIRE(op_valid_for(dtype, op_mult), "op_mult not defined for dtype");
    SPLACU_REQUIRE(op_valid_for(dtype, op_add), "op_add not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (M->n_rows == 0) return SPLACU_OK;
    const Select sel = make_select(op_select);
    SPLACU_REQUIRE(d_r, "null result pointer");
    SPLACU_REQUIRE(d_mask || !sel.reads_mask, "null mask pointer");
    SPLACU_REQUIRE(d_v || M->nnz == 0, "null vector pointer");
    cudaStream_t s = resolve_stream(stream);

    return dispatch_dtype(dtype, [&](auto tag) {
        using T       = decltype(tag);
        const T* v    = static_cast<const T*>(d_v);
        const T* mask = static_cast<const T*>(d_mask);
        T*       r    = static_cast<T*>(d_r);
        const T  init = from_bits<T>(init_bits);
        return dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
            using S = decltype(sr);
            if (!early_exit && is_assoc_commutative(op_add)) return launch_rows<T, S>(sr, sel, M, v, mask, r, init, s);
            mxv_seq_kernel<T, S><<<grid_for(M->n_rows, kBlock, 8), kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, reinterpret_cast<const T*>(M->Ax), v, mask,
                                                                                  r, init, M->n_rows, early_exit);
            SPLACU_LAUNCH_CHECK();
            return 0;
        });
    });
}
