// mxv_scat.cu -- the ROW classes of the pull product's tail: scatter into shared memory while v streams.
//
// Semantics: reference src/cpu/cpu_mxv.hpp:88-103 for associative + commutative op_add without early exit (as mxv_seg.cu).
//
// Why: after the column classes (mxv_pull.cu) the tail class -- entries whose column is not among the hub columns -- gathers v from
// L2, one L1->L2 request per entry, and that request port (~1 per clock and SM, ~285 G/s on the GPU) bounds the pass: 4.1 ps per
// entry against ~1 ps for a streamed one. A power-law matrix is skewed in its ROWS as well: the 45 K rows with the most tail
// entries hold ~40 % of them (RMAT-24). For those entries the product is turned around. They are stored in COLUMN order, so v is
// read as an ascending stream (once per column segment), and the random access goes to a table of per-row partial results in
// SHARED memory (16-bit row slots, atomics on shared memory; no request leaves the SM). Every CTA ends with a private table; a
// merge kernel folds the tables in CTA order and adds the totals onto r under the mask. The rows of a row class have no entry in
// the tail class any more.
//
// Order of the additions: entries of a row arrive in the order the CTA's warps reach them, so FLOAT PLUS / MULT sums are reproducible
// only up to rounding (well inside the 1e-5 bar; integer, MIN / MAX, logical and bitwise results are exact and unaffected).
#include "common.cuh"
#include "ops.cuh"
#include "profile.cuh"

#include <cub/device/device_scan.cuh>

namespace splacu {

    namespace {
        constexpr int      kBlock     = 256;
        constexpr int      kScatWarps = 20;
        constexpr uint32_t kSmemMax   = 227u * 1024u;
        constexpr uint32_t kNone      = 0xffffffffu;

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

        // ---- build ------------------------------------------------------------------------------------
        // a warp per row of the class: count / place its tail-column entries by column
        __global__ void __launch_bounds__(kBlock) scat_count_kernel(const uint32_t* __restrict__ rows, uint32_t n_slots, const uint32_t* __restrict__ Ap,
                                                                    const uint32_t* __restrict__ Aj, const uint32_t* __restrict__ col_slot,
                                                                    uint32_t* __restrict__ col_count) {
            const uint32_t lane    = threadIdx.x & 31u;
            const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
            for (uint32_t sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < n_slots; sl += n_warps) {
                const uint32_t row = rows[sl];
                for (uint32_t k = Ap[row] + lane, k1 = Ap[row + 1]; k < k1; k += 32) {
                    const uint32_t col = Aj[k];
                    if (col_slot[col] == kNone) atomicAdd(&col_count[col], 1u);
                }
            }
        }
        __global__ void __launch_bounds__(kBlock) scat_fill_kernel(const uint32_t* __restrict__ rows, uint32_t n_slots, const uint32_t* __restrict__ Ap,
                                                                   const uint32_t* __restrict__ Aj, const uint32_t* __restrict__ Ax,
                                                                   const uint32_t* __restrict__ col_slot, uint32_t* __restrict__ cursor,
                                                                   uint16_t* __restrict__ out_slot, uint32_t* __restrict__ out_val) {
            const uint32_t lane    = threadIdx.x & 31u;
            const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
            for (uint32_t sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < n_slots; sl += n_warps) {
                const uint32_t row = rows[sl];
                for (uint32_t k = Ap[row] + lane, k1 = Ap[row + 1]; k < k1; k += 32) {
                    const uint32_t col = Aj[k];
                    if (col_slot[col] != kNone) continue;
                    const uint32_t pos     = atomicAdd(&cursor[col], 1u);// position in the column order of the class
                    out_slot[seg_pos16(pos)] = (uint16_t) sl;
                    out_val[seg_pos32(pos)]  = Ax[k];
                }
            }
        }

        // ---- the kernel ---------------------------------------------------------------------------------
        template<typename T, typename S, int WARPS>
        __global__ void __launch_bounds__(WARPS * 32, 1)
                mxv_scat_kernel(S sr, const uint32_t* __restrict__ slot16, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ flags,
                                const uint32_t* __restrict__ seg_base, const uint32_t* __restrict__ seg_col, const T* __restrict__ v, uint32_t nnz,
                                uint32_t n_tiles, uint32_t n_segs, uint32_t n_slots, uint32_t* __restrict__ partial, const uint32_t* __restrict__ gate,
                                uint32_t gate_min) {
            extern __shared__ __align__(16) uint32_t smem[];
            if (gate && *gate < gate_min) return;// sparse mask: the CSR kernel runs instead (as the column classes)
            const uint32_t tid  = threadIdx.x;
            const uint32_t lane = tid & 31u;
            const uint32_t warp = tid >> 5;
            T*             s_v   = reinterpret_cast<T*>(smem) + warp * 512;// v of the column segments of this warp's tile
            T*             s_tab = reinterpret_cast<T*>(smem) + WARPS * 512;// partial result of every row slot
            const T        ident = sr.identity();
            for (uint32_t i = tid; i < n_slots; i += WARPS * 32) s_tab[i] = ident;
            __syncthreads();

            const uint64_t pol     = policy_evict_first();
            const uint32_t n_warps = gridDim.x * WARPS;
            const uint32_t first   = blockIdx.x * WARPS + warp;
            const uint4*   idx4    = reinterpret_cast<const uint4*>(slot16);
            const uint4*   val4    = reinterpret_cast<const uint4*>(vals);

            // one tile ahead: slices, flags, segment range, the columns of the first 64 segments
            uint4    xv[4], xi[2];
            uint32_t fw = 0, sb0 = 0, sb1 = 0, c0 = 0, c1 = 0;
            auto     prefetch = [&](uint32_t t) {
                if (t >= n_tiles) return;
                sb0 = __ldg(seg_base + t);
                sb1 = __ldg(seg_base + t + 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) xi[h] = ld_stream_u4(idx4 + (size_t) t * 64 + h * 32 + lane, pol);
#pragma unroll
                for (int q = 0; q < 4; ++q) xv[q] = ld_stream_u4(val4 + (size_t) t * 128 + q * 32 + lane, pol);
                fw = __ldg(flags + t * 16u + (lane >> 1));
                c0 = __ldg(seg_col + sb0 + lane);// the list is padded by 32 words
                c1 = (sb0 + 32u + lane <= n_segs) ? __ldg(seg_col + sb0 + 32u + lane) : 0u;
            };
            prefetch(first);

            for (uint32_t tile = first; tile < n_tiles; tile += n_warps) {
                // segments that end in the tile, plus the one its last entries open (it ends in a later tile); at most 512
                uint32_t nseg = sb1 - sb0 + 1u;
                if (nseg > 512u) nseg = 512u;
                if (sb0 + nseg > n_segs) nseg = n_segs - sb0;
                const uint32_t base = sb0;
                if (lane < nseg) s_v[lane] = v[c0];
                if (lane + 32u < nseg) s_v[lane + 32u] = v[c1];
                for (uint32_t o = 64u + lane; o < nseg; o += 32) s_v[o] = v[__ldg(seg_col + base + o)];
                const uint32_t fl = (fw >> ((lane & 1u) * 16u)) & 0xffffu;
                uint32_t       sl[16];
                T              a[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w[4] = {xi[h].x, xi[h].y, xi[h].z, xi[h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        sl[8 * h + 2 * k]     = w[k] & 0xffffu;
                        sl[8 * h + 2 * k + 1] = w[k] >> 16;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[4 * q + 0] = from_bits<T>(xv[q].x);
                    a[4 * q + 1] = from_bits<T>(xv[q].y);
                    a[4 * q + 2] = from_bits<T>(xv[q].z);
                    a[4 * q + 3] = from_bits<T>(xv[q].w);
                }
                prefetch(tile + n_warps);// the slice registers are free again

                // segment of the lane's first entry = flags in the lanes below
                const uint32_t cnt  = __popc(fl);
                uint32_t       incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if ((int) lane >= d) incl += t;
                }
                uint32_t       k     = incl - cnt;
                const uint32_t e0    = tile * 512u + lane * 16u;
                const uint32_t valid = e0 >= nnz ? 0u : (nnz - e0 < 16u ? nnz - e0 : 16u);// only the last tile is ragged
                __syncwarp();// s_v is complete
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if ((uint32_t) i < valid) {
                        const T prod = sr.mult(a[i], s_v[k]);
                        T*      dst  = s_tab + sl[i];
                        atomic_combine<T>(sr.add_op(), dst, prod, *dst);
                    }
                    k += (fl >> i) & 1u;
                }
                __syncwarp();// s_v is reused by the next tile
            }
            __syncthreads();
            uint32_t* out = partial + (size_t) blockIdx.x * n_slots;
            for (uint32_t i = tid; i < n_slots; i += WARPS * 32) out[i] = to_bits(s_tab[i]);
        }

        // r[row of slot] = add(r[row], table of CTA 0 + table of CTA 1 + ...) for the selected rows
        template<typename T, typename S>
        __global__ void __launch_bounds__(kBlock) mxv_scat_merge_kernel(S sr, const uint32_t* __restrict__ partial, uint32_t grid, uint32_t n_slots,
                                                                        const uint32_t* __restrict__ rows, const uint32_t* __restrict__ sel_bits, T* r,
                                                                        const uint32_t* __restrict__ gate, uint32_t gate_min) {
            if (gate && *gate < gate_min) return;
            const uint32_t sl = blockIdx.x * blockDim.x + threadIdx.x;
            if (sl >= n_slots) return;
            const uint32_t row = rows[sl];
            if (sel_bits && !((sel_bits[row >> 5] >> (row & 31u)) & 1u)) return;
            T acc = from_bits<T>(partial[sl]);
            for (uint32_t c = 1; c < grid; ++c) acc = sr.add(acc, from_bits<T>(partial[(size_t) c * n_slots + sl]));
            r[row] = sr.add(r[row], acc);
        }
    }// namespace

    void scat_free(Csr* M) {
        for (int q = 0; q < M->n_scat; ++q) {
            CsrScat& sc = M->scat[q];
            cudaFree(sc.rows); cudaFree(sc.slot); cudaFree(sc.Ax); cudaFree(sc.flags); cudaFree(sc.seg_base); cudaFree(sc.seg_col); cudaFree(sc.partial);
            sc = CsrScat();
        }
        M->n_scat = 0;
    }

    int scat_build(Csr* M, const uint32_t* d_col_slot, const uint32_t* d_rows_sorted, uint32_t n_hub_rows, uint32_t slots_per_class, cudaStream_t s) {
        uint32_t *col_count = nullptr, *col_ptr = nullptr, *cursor = nullptr;
        void*     tmp       = nullptr;
        int       rc        = 0;
#define SC_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);          \
            goto done;                                                        \
        }                                                                     \
    } while (0)
        {
            const size_t nc = (size_t) M->n_cols + 1;
            size_t       tmp_bytes = 0;
            SC_CUDA(cudaMalloc(&col_count, nc * 4));
            SC_CUDA(cudaMalloc(&col_ptr, nc * 4));
            SC_CUDA(cudaMalloc(&cursor, nc * 4));
            SC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, col_count, col_ptr, (int) nc, s));
            SC_CUDA(cudaMalloc(&tmp, tmp_bytes));
            const int n_classes = (int) ((n_hub_rows + slots_per_class - 1) / slots_per_class);
            for (int q = 0; q < n_classes && q < kMaxScat; ++q) {
                CsrScat& sc = M->scat[q];
                M->n_scat   = q + 1;
                sc.n_slots  = n_hub_rows - q * slots_per_class < slots_per_class ? n_hub_rows - q * slots_per_class : slots_per_class;
                SC_CUDA(cudaMalloc(&sc.rows, (size_t) sc.n_slots * 4));
                SC_CUDA(cudaMemcpyAsync(sc.rows, d_rows_sorted + (size_t) q * slots_per_class, (size_t) sc.n_slots * 4, cudaMemcpyDeviceToDevice, s));
                SC_CUDA(cudaMemsetAsync(col_count, 0, nc * 4, s));
                scat_count_kernel<<<grid_for((size_t) sc.n_slots * 32, kBlock, 8), kBlock, 0, s>>>(sc.rows, sc.n_slots, M->Ap, M->Aj, d_col_slot, col_count);
                SC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, col_count, col_ptr, (int) nc, s));
                SC_CUDA(cudaMemcpyAsync(&sc.nnz, col_ptr + M->n_cols, 4, cudaMemcpyDeviceToHost, s));
                SC_CUDA(cudaMemcpyAsync(cursor, col_ptr, nc * 4, cudaMemcpyDeviceToDevice, s));
                SC_CUDA(cudaStreamSynchronize(s));
                count_launch(2);
                sc.n_tiles = (uint32_t) (((uint64_t) sc.nnz + kMxvTile - 1) / kMxvTile);
                if (sc.n_tiles == 0) continue;
                const size_t padded = (size_t) sc.n_tiles * kMxvTile;
                SC_CUDA(cudaMalloc(&sc.slot, padded * 2));
                SC_CUDA(cudaMalloc(&sc.Ax, padded * 4));
                SC_CUDA(cudaMemsetAsync(sc.slot, 0, padded * 2, s));
                SC_CUDA(cudaMemsetAsync(sc.Ax, 0, padded * 4, s));
                scat_fill_kernel<<<grid_for((size_t) sc.n_slots * 32, kBlock, 8), kBlock, 0, s>>>(sc.rows, sc.n_slots, M->Ap, M->Aj, M->Ax, d_col_slot, cursor,
                                                                                                   static_cast<uint16_t*>(sc.slot), sc.Ax);
                count_launch(1);
                if ((rc = seg_structure(col_ptr, col_count, M->n_cols, sc.n_tiles, &sc.flags, &sc.seg_base, &sc.seg_col, &sc.n_segs, s))) goto done;
                const uint32_t want = (sc.n_tiles + kScatWarps - 1) / kScatWarps;
                sc.grid             = want < (uint32_t) sm_count() ? want : (uint32_t) sm_count();
                SC_CUDA(cudaMalloc(&sc.partial, (size_t) sc.grid * sc.n_slots * 4));
            }
            SC_CUDA(cudaStreamSynchronize(s));
            SC_CUDA(cudaGetLastError());
        }
    done:
#undef SC_CUDA
        cudaFree(col_count);
        cudaFree(col_ptr);
        cudaFree(cursor);
        cudaFree(tmp);
        if (rc) scat_free(M);
        return rc;
    }

    int scat_mxv(const Csr* M, int dtype, int op_mult, int op_add, const Select& sel, const void* d_v, void* d_r, const uint32_t* gate, uint32_t gate_min,
                 cudaStream_t s) {
        if (M->n_scat == 0) return 0;
        const uint32_t* sel_bits = (sel.reads_mask && gate) ? M->sel_bits : nullptr;
        return dispatch_dtype(dtype, [&](auto tag) {
            using T    = decltype(tag);
            const T* v = static_cast<const T*>(d_v);
            T*       r = static_cast<T*>(d_r);
            return dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
                using S = decltype(sr);
                auto            kern      = mxv_scat_kernel<T, S, kScatWarps>;
                static uint64_t attr_done = 0;// per instantiation, one bit per device
                const int       dev_bit   = current_device() & 63;
                if (!((attr_done >> dev_bit) & 1u)) {
                    SPLACU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemMax));
                    attr_done |= (uint64_t) 1 << dev_bit;
                }
                for (int q = 0; q < M->n_scat; ++q) {
                    const CsrScat& sc = M->scat[q];
                    if (sc.n_tiles == 0) continue;
                    const uint32_t smem = kScatWarps * 512u * 4u + ((sc.n_slots + 3u) & ~3u) * 4u;
                    if (smem > kSmemMax) {
                        set_error("mxv: row class of %u slots does not fit in shared memory", sc.n_slots);
                        return (int) SPLACU_E_INVALID;
                    }
                    SPLACU_PROFILE("splacu/mxv/row_class", s);
                    const uint32_t cap  = persistent_grid_cap();
                    const uint32_t grid = sc.grid < cap ? sc.grid : cap;// the partial tables are sized for sc.grid CTAs
                    kern<<<grid, kScatWarps * 32, smem, s>>>(sr, static_cast<const uint32_t*>(sc.slot), sc.Ax, sc.flags, sc.seg_base, sc.seg_col, v, sc.nnz,
                                                             sc.n_tiles, sc.n_segs, sc.n_slots, sc.partial, gate, gate_min);
                    SPLACU_LAUNCH_CHECK();
                    mxv_scat_merge_kernel<T, S><<<(sc.n_slots + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, sc.partial, grid, sc.n_slots, sc.rows, sel_bits, r,
                                                                                                     gate, gate_min);
                    SPLACU_LAUNCH_CHECK();
                }
                return 0;
            });
        });
    }

}// namespace splacu
