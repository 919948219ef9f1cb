// mxv_pull.cu -- masked semiring pull product r = M x v (spla exec_mxv_masked).
//
// Semantics: reference src/cpu/cpu_mxv.hpp:88-103 (see include/splacu.h). Replaces the OpenCL kernels
// mxv_vector / mxv_config / mxv_config_scalar (reference src/opencl/kernels/mxv.cl:43-170), which keep only
// 512 work-groups x 32 lanes in flight and need a compaction pass + blocking read for the early-exit case.
//
// Kernels
//   mxv_tile_kernel     THE streaming kernel (associative + commutative op_add, no early exit). The nnz range is cut
//                       into equal tiles of kTile entries (nnz-split, merge-path style load balance: a power-law hub
//                       row simply spans many tiles, a run of short rows shares one). A CTA streams its tile of
//                       Aj / Ax with 128-bit evict-first loads, gathers v through the read-only path, and parks the
//                       products in shared memory; the rows of the tile are then folded from shared memory:
//                       thread-per-row for short segments, warp-per-row for long ones. Rows that cross a tile border
//                       leave a deterministic partial (tail / head) that mxv_fixup_kernel chains left to right.
//                       With a mask-reading select the selected rows first mark the 4-entry groups they need in a
//                       shared bitmap, so that Aj / Ax / v of unselected rows are never touched.
//   mxv_fixup_kernel    one thread per tile: r[row] of the (at most one) row that starts in the tile and ends later.
//   mxv_seq_kernel      one thread per row, strict left-to-right fold: the exact path for early_exit and for
//                       non-associative adds (MINUS, DIV, FIRST, SECOND, BONE, MINUS_POW2).
#include "common.cuh"
#include "ops.cuh"

namespace splacu {

    static constexpr int kBlock = 256;

    // ---- streaming-load helpers ---------------------------------------------------------------
    // CSR arrays are read exactly once per product: evict-first in L2, no L1 allocation, so that L1 / L2 keep v.
    __device__ __forceinline__ uint64_t policy_evict_first() {
        uint64_t p;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
        return p;
    }
    __device__ __forceinline__ uint4 ld_stream_u4(const uint4* p, uint64_t pol) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                     : "l"(p), "l"(pol));
        return r;
    }
    __device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol) {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
        return r;
    }
    template<typename T> __device__ __forceinline__ T ld_gather(const T* p) { return __ldg(p); }

    // ---- tile metadata ---------------------------------------------------------------------------
    // tile_row[t] = first row r with Ap[r] >= t * tile  (rows that START in tile t are [tile_row[t], tile_row[t+1]));
    // tile_row[n_tiles] = n_rows, so trailing empty rows belong to the last tile.
    __global__ void __launch_bounds__(kBlock) tile_rows_kernel(const uint32_t* __restrict__ Ap, uint32_t n_rows, uint32_t tile, uint32_t n_tiles,
                                                               uint32_t* __restrict__ tile_row) {
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t > n_tiles) return;
        if (t == n_tiles) {
            tile_row[t] = n_rows;
            return;
        }
        const uint64_t target = (uint64_t) t * tile;
        uint32_t       lo = 0, hi = n_rows + 1;// search in Ap[0 .. n_rows]
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (Ap[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        tile_row[t] = lo < n_rows ? lo : n_rows;
    }

    int csr_build_metadata(Csr* M, cudaStream_t s) {
        M->avg_row_nnz = M->n_rows ? (float) M->nnz / (float) M->n_rows : 0.f;
        M->vec_ok      = ((((uintptr_t) M->Aj) | ((uintptr_t) M->Ax)) & 15u) == 0;
        M->n_tiles     = 0;
        if (M->nnz == 0 || M->n_rows == 0) return 0;
        M->tile    = kMxvTile;
        M->n_tiles = (uint32_t) (((uint64_t) M->nnz + M->tile - 1) / M->tile);
        SPLACU_CUDA(cudaMalloc(&M->tile_row, ((size_t) M->n_tiles + 1) * sizeof(uint32_t)));
        SPLACU_CUDA(cudaMalloc(&M->carry, (size_t) M->n_tiles * 2 * sizeof(uint32_t)));
        tile_rows_kernel<<<(M->n_tiles + 1 + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->Ap, M->n_rows, M->tile, M->n_tiles, M->tile_row);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    // ---- the streaming kernel ------------------------------------------------------------------------
    static constexpr int kItems   = kMxvTile / kBlock;// entries per thread per tile (16)
    static constexpr int kGroups  = kItems / 4;       // 128-bit groups per thread per tile (4)
    static constexpr int kShort   = 32;               // segments up to this length are folded by one thread
    static constexpr int kMaxLong = kMxvTile / kShort;// more long segments than this cannot exist in a tile
    static_assert(kMxvTile % (kBlock * 4) == 0, "tile must be a whole number of 128-bit groups per thread");

    template<typename T, typename S, bool MASKED>
    __global__ void __launch_bounds__(kBlock, 4)
            mxv_tile_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj, const T* __restrict__ Ax,
                            const T* __restrict__ v, const T* __restrict__ mask, T* __restrict__ r, T init, uint32_t n_rows, uint32_t nnz,
                            uint32_t n_tiles, const uint32_t* __restrict__ tile_row, T* __restrict__ carry, int vec_ok) {
        __shared__ __align__(16) T s_prod[kMxvTile];
        __shared__ uint32_t        s_need[kMxvTile / 128];// one bit per 4-entry group (MASKED only)
        __shared__ uint32_t        s_long[kMaxLong];      // row ids of long segments
        __shared__ uint32_t        s_nlong;

        const uint32_t tid  = threadIdx.x;
        const uint32_t lane = tid & 31u;
        const uint32_t warp = tid >> 5;
        const bool     all  = !MASKED && (sel.classes != 0u);// ALWAYS; (NEVER never gets here)
        const uint64_t pol  = policy_evict_first();

        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t lo     = tile * (uint32_t) kMxvTile;
            const uint32_t hi     = (nnz - lo > (uint32_t) kMxvTile) ? lo + (uint32_t) kMxvTile : nnz;
            const uint32_t row_lo = tile_row[tile];
            const uint32_t row_hi = tile_row[tile + 1];
            // the row before row_lo reaches into this tile iff the first owned row starts after lo
            const uint32_t first_start = Ap[row_lo];// row_lo <= n_rows, Ap[n_rows] == nnz
            const uint32_t row_first   = (first_start > lo) ? row_lo - 1 : row_lo;

            if (tid == 0) s_nlong = 0;
            if (MASKED) {
                if (tid < kMxvTile / 128) s_need[tid] = 0u;
                __syncthreads();
                // selected rows mark the groups they need: mask is tested before Aj / Ax / v are touched
                for (uint32_t row = row_first + tid; row < row_hi; row += kBlock) {
                    if (!sel.test(mask[row])) continue;
                    const uint32_t a = Ap[row], b = Ap[row + 1];
                    const uint32_t s = max(a, lo), e = min(b, hi);
                    if (e <= s) continue;
                    const uint32_t g0 = (s - lo) >> 2, g1 = (e - 1 - lo) >> 2;
                    const uint32_t w0 = g0 >> 5, w1 = g1 >> 5;
                    for (uint32_t w = w0; w <= w1; ++w) {
                        uint32_t m = 0xffffffffu;
                        if (w == w0) m &= 0xffffffffu << (g0 & 31u);
                        if (w == w1) m &= 0xffffffffu >> (31u - (g1 & 31u));
                        atomicOr(&s_need[w], m);
                    }
                }
            }
            __syncthreads();

            // ---- phase A: stream the tile, gather, multiply, park products in shared memory ----
            if (vec_ok && hi - lo == (uint32_t) kMxvTile) {
                uint4 j[kGroups], a[kGroups];
                bool  need[kGroups];
#pragma unroll
                for (int c = 0; c < kGroups; ++c) {
                    const uint32_t g = c * kBlock + tid;
                    need[c]          = MASKED ? ((s_need[g >> 5] >> (g & 31u)) & 1u) != 0u : true;
                    if (need[c]) {
                        j[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Aj + lo) + g, pol);
                        a[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Ax + lo) + g, pol);
                    }
                }
#pragma unroll
                for (int c = 0; c < kGroups; ++c) {
                    if (need[c]) {
                        const T x0 = ld_gather(v + j[c].x), x1 = ld_gather(v + j[c].y), x2 = ld_gather(v + j[c].z), x3 = ld_gather(v + j[c].w);
                        uint4   p;
                        p.x = to_bits(sr.mult(from_bits<T>(a[c].x), x0));
                        p.y = to_bits(sr.mult(from_bits<T>(a[c].y), x1));
                        p.z = to_bits(sr.mult(from_bits<T>(a[c].z), x2));
                        p.w = to_bits(sr.mult(from_bits<T>(a[c].w), x3));
                        reinterpret_cast<uint4*>(s_prod)[c * kBlock + tid] = p;
                    }
                }
            } else {
                for (uint32_t k = lo + tid; k < hi; k += kBlock) {
                    const uint32_t g    = (k - lo) >> 2;
                    const bool     need = MASKED ? ((s_need[g >> 5] >> (g & 31u)) & 1u) != 0u : true;
                    if (need) s_prod[k - lo] = sr.mult(from_bits<T>(ld_stream_u32(reinterpret_cast<const uint32_t*>(Ax) + k, pol)), ld_gather(v + ld_stream_u32(Aj + k, pol)));
                }
            }
            __syncthreads();

            // ---- phase B: fold the rows of the tile from shared memory ----
            for (uint32_t row = row_first + tid; row < row_hi; row += kBlock) {
                const bool     take = all ? true : (MASKED ? sel.test(mask[row]) : false);
                const uint32_t a = Ap[row], b = Ap[row + 1];
                const bool     head = a < lo, tail = b > hi;
                if (!take) {
                    if (!head && !tail) r[row] = init;
                    continue;// partial segments of unselected rows are never read by the fix-up
                }
                const uint32_t s = max(a, lo) - lo, e = min(b, hi) - lo;
                if (e - s > (uint32_t) kShort) {
                    s_long[atomicAdd(&s_nlong, 1u)] = row;
                    continue;
                }
                T acc = sr.identity();
                for (uint32_t k = s; k < e; ++k) acc = sr.add(acc, s_prod[k]);
                if (head) carry[2 * tile] = acc;
                else if (tail) carry[2 * tile + 1] = acc;
                else r[row] = (e > s) ? sr.add(init, acc) : init;
            }
            __syncthreads();
            const uint32_t nlong = s_nlong;
            for (uint32_t q = warp; q < nlong; q += kBlock / 32) {
                const uint32_t row = s_long[q];
                const uint32_t a = Ap[row], b = Ap[row + 1];
                const uint32_t s = max(a, lo) - lo, e = min(b, hi) - lo;
                T              acc = sr.identity();
                for (uint32_t k = s + lane; k < e; k += 32) acc = sr.add(acc, s_prod[k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                if (lane == 0) {
                    if (a < lo) carry[2 * tile] = acc;
                    else if (b > hi) carry[2 * tile + 1] = acc;
                    else r[row] = sr.add(init, acc);
                }
            }
            __syncthreads();// s_prod / s_long are reused by the next tile
        }
    }

    // r[row] of rows that start in tile t and end in a later tile: tail(t) + head(t+1) + ... chained left to right
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_fixup_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const T* __restrict__ mask,
                                                               T* __restrict__ r, T init, uint32_t nnz, uint32_t n_tiles,
                                                               const uint32_t* __restrict__ tile_row, const T* __restrict__ carry) {
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n_tiles) return;
        const uint32_t row_lo = tile_row[t], row_hi = tile_row[t + 1];
        if (row_hi == row_lo) return;
        const uint32_t row = row_hi - 1;
        const uint64_t end = Ap[row + 1];
        if (end <= (uint64_t) (t + 1) * kMxvTile) return;// ends inside its own tile (for the last tile: end <= nnz)
        const bool take = sel.reads_mask ? sel.test(mask[row]) : (sel.classes != 0u);
        if (!take) {
            r[row] = init;
            return;
        }
        T acc = carry[2 * t + 1];
        for (uint32_t u = t + 1; u < n_tiles; ++u) {
            acc = sr.add(acc, carry[2 * u]);
            if (end <= (uint64_t) (u + 1) * kMxvTile) break;
        }
        r[row] = sr.add(init, acc);
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
    static int launch_tiles(S sr, Select sel, const Csr* M, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        const int grid = (int) min((uint64_t) M->n_tiles, (uint64_t) sm_count() * 4);
        T*        carry = reinterpret_cast<T*>(M->carry);
        const T*  Ax    = reinterpret_cast<const T*>(M->Ax);
        if (sel.reads_mask)
            mxv_tile_kernel<T, S, true><<<grid, kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, Ax, v, mask, r, init, M->n_rows, M->nnz, M->n_tiles, M->tile_row, carry, (int) M->vec_ok);
        else
            mxv_tile_kernel<T, S, false><<<grid, kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, Ax, v, mask, r, init, M->n_rows, M->nnz, M->n_tiles, M->tile_row, carry, (int) M->vec_ok);
        SPLACU_LAUNCH_CHECK();
        mxv_fixup_kernel<T, S><<<(M->n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, sel, M->Ap, mask, r, init, M->nnz, M->n_tiles, M->tile_row, carry);
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

    // nothing can be selected / nothing stored: r = init everywhere (reference src/cpu/cpu_mxv.hpp:89,102)
    if (M->nnz == 0 || (!sel.reads_mask && sel.classes == 0u)) return splacu_fill(d_r, init_bits, M->n_rows, stream);

    return dispatch_dtype(dtype, [&](auto tag) {
        using T       = decltype(tag);
        const T* v    = static_cast<const T*>(d_v);
        const T* mask = static_cast<const T*>(d_mask);
        T*       r    = static_cast<T*>(d_r);
        const T  init = from_bits<T>(init_bits);
        return dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
            using S = decltype(sr);
            if (!early_exit && is_assoc_commutative(op_add)) return launch_tiles<T, S>(sr, sel, M, v, mask, r, init, s);
            mxv_seq_kernel<T, S><<<grid_for(M->n_rows, kBlock, 8), kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, reinterpret_cast<const T*>(M->Ax), v, mask,
                                                                                  r, init, M->n_rows, early_exit);
            SPLACU_LAUNCH_CHECK();
            return 0;
        });
    });
}
