// mxv_pull.cu -- masked semiring pull product r = M x v (spla exec_mxv_masked).
//
// Semantics: reference src/cpu/cpu_mxv.hpp:88-103 (see include/splacu.h). Replaces the OpenCL kernels
// mxv_vector / mxv_config / mxv_config_scalar (reference src/opencl/kernels/mxv.cl:43-170), which keep only
// 512 work-groups x 32 lanes in flight and need a compaction pass + blocking read for the early-exit case.
//
// Kernels
//   mxv_wtile_kernel    THE streaming kernel (associative + commutative op_add, no early exit). The nnz range is cut
//                       into equal WARP tiles of kMxvTile = 512 entries (nnz-split, merge-path style load balance: a
//                       power-law hub row simply spans many tiles, a run of short rows shares one). One warp owns one
//                       tile at a time and never meets a CTA barrier: it streams its slice of Aj / Ax with 128-bit
//                       evict-first loads, gathers v, parks the 512 products in its private shared-memory slice and
//                       folds the rows of the tile from there (lane-per-row for short segments, whole warp for long
//                       ones). Rows that cross a tile border leave a deterministic partial (head / tail) that
//                       mxv_fixup_kernel chains left to right.
//                       Mask-reading selects: the selected rows first build the bitmap of 4-entry groups they need
//                       (warp OR-reduction in registers), so Aj / Ax / v of unselected rows are never touched.
//                       HUB variant: random 4-byte gathers of v are bounded by the SM's L1-miss path (~1 sector per
//                       clock per SM, measured: tools/gather_bench.cu), not by HBM. Power-law graphs concentrate a large
//                       share of the gathers on few columns, so the matrix handle keeps a second index array in which
//                       the most referenced columns are replaced by slots of a hub table; every persistent CTA holds
//                       the hub values of v in shared memory (packed once per call by mxv_hub_pack_kernel) and serves
//                       those gathers on chip.
//   mxv_fixup_kernel    one thread per tile: r[row] of the (at most one) row that starts in the tile and ends later;
//                       long chains (hub rows) are summed by the whole warp.
//   mxv_early_kernel    early_exit (BFS bottom-up): first qualifying entry of each row, lane-serial for the first 8 entries,
//                       then warp-cooperative with a ballot (exact for any op pair).
//   mxv_seq_kernel      one thread per row, strict left-to-right fold: the exact path for non-associative adds
//                       (MINUS, DIV, FIRST, SECOND, BONE, MINUS_POW2) without early_exit.
#include "common.cuh"
#include "ops.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cstring>

namespace splacu {

    static constexpr int      kBlock     = 256;
    static constexpr int      kWarps     = 24;                   // warps per persistent CTA of the streaming kernel
    static constexpr int      kThreads   = kWarps * 32;          // 768 threads (80 registers each), one CTA per SM
    static constexpr int      kItems     = kMxvTile / 32;        // entries per lane per tile (16)
    static constexpr int      kGroups    = kItems / 4;           // 128-bit groups per lane per tile (4)
    static constexpr int      kShort     = 32;                   // segments up to this length are folded by one lane
    static constexpr uint32_t kHubFlag   = 0x80000000u;          // Aj_hub entry = kHubFlag | slot
    static constexpr uint32_t kSmemLimit = 227u * 1024u;         // opt-in dynamic shared memory per CTA on sm_100
    static constexpr uint32_t kProdBytes = kWarps * kMxvTile * 4;// 48 KB of per-warp product slices
    static constexpr uint32_t kHubCap    = (kSmemLimit - kProdBytes - 1024u) / 4u & ~3u;// hub slots per CTA (~41 K)
    static_assert(kMxvTile == 512, "warp tile = 32 lanes x 4 groups x 4 entries");

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

    __device__ __forceinline__ uint32_t ld_gather_plain(const uint32_t* p) {
        uint32_t r;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
        return r;
    }
    // one gather of the HUB variant without divergence (two predicated loads): j = kHubFlag | slot -> shared memory when
    // slot < n_smem, else the dense packed hub table through L1; plain column id -> v.
    // (Measured: L1::no_allocate on the cold gathers halves the throughput -- L1 lines are the miss buffers -- so they allocate.)
    __device__ __forceinline__ uint32_t gather_hub(uint32_t j, const uint32_t* __restrict__ v, const uint32_t* __restrict__ hub_vals,
                                                   uint32_t s_hub_addr, uint32_t n_smem) {
        const uint32_t  slot    = j & 0x7fffffffu;
        const bool      hub     = (j >> 31) != 0u;
        const uint32_t  in_smem = (hub && slot < n_smem) ? 1u : 0u;
        const uint32_t* gp      = hub ? hub_vals + slot : v + j;
        const uint32_t  sp      = s_hub_addr + slot * 4u;
        uint32_t        r;
        asm volatile("{\n\t"
                     ".reg .pred ps;\n\t"
                     "setp.ne.u32 ps, %3, 0;\n\t"
                     "@ps ld.shared.u32 %0, [%1];\n\t"
                     "@!ps ld.global.nc.u32 %0, [%2];\n\t"
                     "}"
                     : "=r"(r)
                     : "r"(sp), "l"(gp), "r"(in_smem));
        return r;
    }

    // ---- tile metadata ---------------------------------------------------------------------------
    // Rows that START in tile t are [row_lo(t), row_lo(t+1)) with row_lo(t) = first row r with Ap[r] >= t * tile and
    // row_lo(n_tiles) = n_rows (trailing empty rows belong to the last tile). tile_rows[t] = (row_first, row_hi):
    // row_first = row_lo(t) - 1 when that row reaches into the tile (head segment), else row_lo(t); row_hi = row_lo(t+1).
    __device__ __forceinline__ uint32_t row_lower_bound(const uint32_t* __restrict__ Ap, uint32_t n_rows, uint64_t target) {
        uint32_t lo = 0, hi = n_rows + 1;// search in Ap[0 .. n_rows]
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (Ap[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        return lo < n_rows ? lo : n_rows;
    }
    __global__ void __launch_bounds__(kBlock) tile_rows_kernel(const uint32_t* __restrict__ Ap, uint32_t n_rows, uint32_t tile, uint32_t n_tiles,
                                                               uint2* __restrict__ tile_rows) {
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n_tiles) return;
        const uint64_t lo     = (uint64_t) t * tile;
        const uint32_t row_lo = row_lower_bound(Ap, n_rows, lo);
        const uint32_t row_hi = (t + 1 == n_tiles) ? n_rows : row_lower_bound(Ap, n_rows, lo + tile);
        tile_rows[t]          = make_uint2((Ap[row_lo] > lo) ? row_lo - 1 : row_lo, row_hi);
    }

    // ---- hub metadata (built once per matrix) ------------------------------------------------------
    __global__ void __launch_bounds__(kBlock) col_count_kernel(const uint32_t* __restrict__ Aj, uint32_t nnz, uint32_t* __restrict__ count) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&count[Aj[k]], 1u);
    }
    __global__ void __launch_bounds__(kBlock) hub_keys_kernel(const uint32_t* __restrict__ count, uint32_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) {
            keys[i] = ~count[i];// ascending sort of ~count == descending count; the sort is stable => ties keep ascending column order
            ids[i]  = i;
        }
    }
    __global__ void __launch_bounds__(kBlock) hub_slots_kernel(const uint32_t* __restrict__ hub_cols, uint32_t n_hub, uint32_t* __restrict__ slot) {
        const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
        if (s < n_hub) slot[hub_cols[s]] = s;
    }
    __global__ void __launch_bounds__(kBlock) hub_encode_kernel(const uint32_t* __restrict__ Aj, uint32_t nnz, const uint32_t* __restrict__ slot,
                                                                uint32_t* __restrict__ Aj_hub) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
            const uint32_t j = Aj[k];
            const uint32_t s = slot[j];
            Aj_hub[k]        = (s != 0xffffffffu) ? (kHubFlag | s) : j;
        }
    }
    // per call: hub_vals[s] = v[hub_cols[s]]
    __global__ void __launch_bounds__(kBlock) mxv_hub_pack_kernel(const uint32_t* __restrict__ hub_cols, uint32_t n_hub, const uint32_t* __restrict__ v,
                                                                  uint32_t* __restrict__ hub_vals) {
        const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
        if (s < n_hub) hub_vals[s] = __ldg(v + hub_cols[s]);
    }

    static int build_hub(Csr* M, cudaStream_t s) {
        const int mode = (int) get_option(OPT_MXV_HUB);// 0 off, 1 auto, 2 force
        if (mode == 0 || !M->vec_ok || M->n_cols >= kHubFlag) return 0;
        if (mode == 1 && (M->nnz < (1u << 22) || M->n_cols < 4 * kHubCap)) return 0;// v already fits on chip / too little work
        const uint32_t n = M->n_cols;
        uint32_t *     count = nullptr, *keys = nullptr, *ids = nullptr, *keys_out = nullptr, *ids_out = nullptr;
        void*          tmp   = nullptr;
        auto           cleanup = [&]() {
            cudaFree(count); cudaFree(keys); cudaFree(ids); cudaFree(keys_out); cudaFree(ids_out); cudaFree(tmp);
        };
#define HUB_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            cleanup();                                                        \
            return ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);        \
        }                                                                     \
    } while (0)
        HUB_CUDA(cudaMalloc(&count, (size_t) n * 4));
        HUB_CUDA(cudaMalloc(&keys, (size_t) n * 4));
        HUB_CUDA(cudaMalloc(&ids, (size_t) n * 4));
        HUB_CUDA(cudaMalloc(&keys_out, (size_t) n * 4));
        HUB_CUDA(cudaMalloc(&ids_out, (size_t) n * 4));
        HUB_CUDA(cudaMemsetAsync(count, 0, (size_t) n * 4, s));
        col_count_kernel<<<grid_for(M->nnz, kBlock, 8), kBlock, 0, s>>>(M->Aj, M->nnz, count);
        hub_keys_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(count, n, keys, ids);
        size_t tmp_bytes = 0;
        HUB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, ids, ids_out, (int) n, 0, 32, s));
        HUB_CUDA(cudaMalloc(&tmp, tmp_bytes));
        HUB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, ids, ids_out, (int) n, 0, 32, s));
        count_launch(6);
        // how many of the top kHubCap columns are referenced often enough to pay for their slot
        uint32_t cap = (uint32_t) get_option(OPT_MXV_HUB_TOTAL);
        if (cap > n) cap = n;
        if (cap < 4) cap = 4;
        const uint32_t min_count = (uint32_t) get_option(OPT_MXV_HUB_MIN_COUNT);
        uint32_t*      h_keys    = (uint32_t*) malloc((size_t) cap * 4);
        HUB_CUDA(cudaMemcpyAsync(h_keys, keys_out, (size_t) cap * 4, cudaMemcpyDeviceToHost, s));
        HUB_CUDA(cudaStreamSynchronize(s));
        uint32_t n_hub = 0;
        while (n_hub < cap && ~h_keys[n_hub] >= min_count) ++n_hub;
        free(h_keys);
        if (n_hub >= 64 || mode == 2) {
            if (n_hub == 0) n_hub = cap < 4 ? cap : 4;
            HUB_CUDA(cudaMalloc(&M->hub_cols, (size_t) n_hub * 4));
            HUB_CUDA(cudaMalloc(&M->hub_vals, ((size_t) n_hub + 4) * 4));
            HUB_CUDA(cudaMalloc(&M->Aj_hub, (size_t) M->nnz * 4));
            HUB_CUDA(cudaMemcpyAsync(M->hub_cols, ids_out, (size_t) n_hub * 4, cudaMemcpyDeviceToDevice, s));
            HUB_CUDA(cudaMemsetAsync(count, 0xff, (size_t) n * 4, s));// reuse as the slot map
            hub_slots_kernel<<<(n_hub + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->hub_cols, n_hub, count);
            hub_encode_kernel<<<grid_for(M->nnz, kBlock, 8), kBlock, 0, s>>>(M->Aj, M->nnz, count, M->Aj_hub);
            count_launch(2);
            M->n_hub      = n_hub;
            M->n_hub_smem = (uint32_t) get_option(OPT_MXV_HUB_SMEM) & ~3u;
            if (M->n_hub_smem > kHubCap) M->n_hub_smem = kHubCap;
            if (M->n_hub_smem > n_hub) M->n_hub_smem = n_hub & ~3u;
        }
        HUB_CUDA(cudaStreamSynchronize(s));
#undef HUB_CUDA
        cleanup();
        return 0;
    }

    int csr_build_metadata(Csr* M, cudaStream_t s) {
        M->avg_row_nnz = M->n_rows ? (float) M->nnz / (float) M->n_rows : 0.f;
        M->vec_ok      = ((((uintptr_t) M->Aj) | ((uintptr_t) M->Ax)) & 15u) == 0;
        M->n_tiles     = 0;
        if (M->nnz == 0 || M->n_rows == 0) return 0;
        M->tile    = kMxvTile;
        M->n_tiles = (uint32_t) (((uint64_t) M->nnz + M->tile - 1) / M->tile);
        SPLACU_CUDA(cudaMalloc(&M->tile_rows, (size_t) M->n_tiles * sizeof(uint2)));
        SPLACU_CUDA(cudaMalloc(&M->carry, (size_t) M->n_tiles * 2 * sizeof(uint32_t)));
        tile_rows_kernel<<<(M->n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->Ap, M->n_rows, M->tile, M->n_tiles, M->tile_rows);
        SPLACU_LAUNCH_CHECK();
        return build_hub(M, s);
    }

    // ---- the streaming kernel ------------------------------------------------------------------------
    template<typename T, typename S, bool MASKED, bool HUB>
    __global__ void __launch_bounds__(kThreads, 1)
            mxv_wtile_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj, const T* __restrict__ Ax,
                             const T* __restrict__ v, const T* __restrict__ mask, T* __restrict__ r, T init, uint32_t nnz, uint32_t n_tiles,
                             const uint2* __restrict__ tile_rows, T* __restrict__ carry, int vec_ok, const uint32_t* __restrict__ hub_vals,
                             uint32_t n_hub_smem) {
        extern __shared__ __align__(16) uint32_t smem[];
        const uint32_t tid  = threadIdx.x;
        const uint32_t lane = tid & 31u;
        const uint32_t warp = tid >> 5;
        T*             s_prod = reinterpret_cast<T*>(smem) + warp * kMxvTile;// this warp's product slice
        const T*       s_hub  = reinterpret_cast<const T*>(smem) + kWarps * kMxvTile;

        if (HUB) {// hub values of v -> shared memory, 128-bit coalesced
            uint4*       dst = reinterpret_cast<uint4*>(smem + kWarps * kMxvTile);
            const uint4* src = reinterpret_cast<const uint4*>(hub_vals);
            for (uint32_t i = tid; i < n_hub_smem / 4; i += kThreads) dst[i] = __ldg(src + i);
            __syncthreads();
        }

        const bool     all = !MASKED && (sel.classes != 0u);// ALWAYS (NEVER never gets here)
        const uint64_t pol = policy_evict_first();
        // hub slots below n_hub_smem live in shared memory, the rest of the (dense, packed) hub table is served by L1.
        // Branch-free: predicated loads.
        const uint32_t s_hub_addr = (uint32_t) __cvta_generic_to_shared(s_hub);
        auto gather = [&](uint32_t j) -> T {
            if (HUB) return from_bits<T>(gather_hub(j, reinterpret_cast<const uint32_t*>(v), hub_vals, s_hub_addr, n_hub_smem));
            return from_bits<T>(ld_gather_plain(reinterpret_cast<const uint32_t*>(v) + j));
        };

        // Software pipeline: everything a tile needs from HBM (its row range and, unless the mask is sparse, its Aj / Ax slices)
        // is requested one tile ahead, before the row folds of the current tile, so that on entry a warp only waits for
        // its gathers.
        const uint32_t n_warps = gridDim.x * kWarps;
        uint4          j[kGroups], a[kGroups];
        uint2          rows     = make_uint2(0u, 0u);
        bool           streamed = false;
        bool           dense    = !MASKED;// masked variant: stream ahead only while the mask keeps selecting most of a tile
        auto           prefetch = [&](uint32_t t) {
            streamed = false;
            if (t >= n_tiles) return;
            rows = __ldg(tile_rows + t);
            if (!dense || !vec_ok || nnz - t * (uint32_t) kMxvTile < (uint32_t) kMxvTile) return;
#pragma unroll
            for (int c = 0; c < kGroups; ++c) j[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Aj + t * (uint32_t) kMxvTile) + c * 32 + lane, pol);
#pragma unroll
            for (int c = 0; c < kGroups; ++c) a[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Ax + t * (uint32_t) kMxvTile) + c * 32 + lane, pol);
            streamed = true;
        };
        prefetch(blockIdx.x * kWarps + warp);

        for (uint32_t tile = blockIdx.x * kWarps + warp; tile < n_tiles; tile += n_warps) {
            const uint32_t lo        = tile * (uint32_t) kMxvTile;
            const uint32_t hi        = (nnz - lo > (uint32_t) kMxvTile) ? lo + (uint32_t) kMxvTile : nnz;
            const uint32_t row_first = rows.x;// first row with entries in this tile (the head row if one reaches in)
            const uint32_t row_hi    = rows.y;// one past the last row that starts in this tile

            // row extents of the first 32 rows of the tile: requested now, consumed by the folds after the gathers
            const uint32_t row0   = row_first + lane;
            const bool     valid0 = row0 < row_hi;
            uint32_t       a0 = 0, b0 = 0;
            bool           take0 = false;
            if (valid0) {
                a0    = __ldg(Ap + row0);
                b0    = __ldg(Ap + row0 + 1);
                take0 = all ? true : (MASKED ? sel.test(mask[row0]) : false);
            }

            uint32_t need[kGroups];
#pragma unroll
            for (int c = 0; c < kGroups; ++c) need[c] = 0xffffffffu;
            if (MASKED) {
                // selected rows mark the 4-entry groups they need: the mask is tested before Aj / Ax / v are touched
#pragma unroll
                for (int c = 0; c < kGroups; ++c) need[c] = 0u;
                for (uint32_t row = row0; row < row_hi; row += 32) {
                    uint32_t ra, rb;
                    bool     tk;
                    if (row == row0) {
                        ra = a0, rb = b0, tk = take0;
                    } else {
                        tk = sel.test(mask[row]);
                        ra = __ldg(Ap + row), rb = __ldg(Ap + row + 1);
                    }
                    const uint32_t s = max(ra, lo), e = min(rb, hi);
                    if (!tk || e <= s) continue;
                    const uint32_t g0 = (s - lo) >> 2, g1 = (e - 1 - lo) >> 2;
#pragma unroll
                    for (int c = 0; c < kGroups; ++c) {
                        const uint32_t b0g = max(g0, (uint32_t) c * 32u), b1g = min(g1, (uint32_t) c * 32u + 31u);
                        if (b0g <= b1g) need[c] |= (0xffffffffu << (b0g & 31u)) & (0xffffffffu >> (31u - (b1g & 31u)));
                    }
                }
                uint32_t n_need = 0;
#pragma unroll
                for (int c = 0; c < kGroups; ++c) {
                    need[c] = __reduce_or_sync(0xffffffffu, need[c]);
                    n_need += __popc(need[c]);
                }
                dense = n_need >= (uint32_t) (kMxvTile / 8);// at least half of the 128 groups
            }

            // ---- phase A: stream the tile, gather, multiply, park products in shared memory ----
            if (vec_ok && hi - lo == (uint32_t) kMxvTile) {
                T x[kGroups][4];
                if (!streamed) {
#pragma unroll
                    for (int c = 0; c < kGroups; ++c)
                        if ((need[c] >> lane) & 1u) j[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Aj + lo) + c * 32 + lane, pol);
#pragma unroll
                    for (int c = 0; c < kGroups; ++c)
                        if ((need[c] >> lane) & 1u) a[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Ax + lo) + c * 32 + lane, pol);
                }
#pragma unroll
                for (int c = 0; c < kGroups; ++c)
                    if ((need[c] >> lane) & 1u) {
                        x[c][0] = gather(j[c].x);
                        x[c][1] = gather(j[c].y);
                        x[c][2] = gather(j[c].z);
                        x[c][3] = gather(j[c].w);
                    }
#pragma unroll
                for (int c = 0; c < kGroups; ++c)
                    if ((need[c] >> lane) & 1u) {
                        uint4 p;
                        p.x = to_bits(sr.mult(from_bits<T>(a[c].x), x[c][0]));
                        p.y = to_bits(sr.mult(from_bits<T>(a[c].y), x[c][1]));
                        p.z = to_bits(sr.mult(from_bits<T>(a[c].z), x[c][2]));
                        p.w = to_bits(sr.mult(from_bits<T>(a[c].w), x[c][3]));
                        reinterpret_cast<uint4*>(s_prod)[c * 32 + lane] = p;
                    }
            } else {
                for (uint32_t k = lo + lane; k < hi; k += 32) {
                    const uint32_t g  = (k - lo) >> 2;
                    bool           nd = true;
#pragma unroll
                    for (int c = 0; c < kGroups; ++c)
                        if ((g >> 5) == (uint32_t) c) nd = ((need[c] >> (g & 31u)) & 1u) != 0u;
                    if (nd) s_prod[k - lo] = sr.mult(from_bits<T>(ld_stream_u32(reinterpret_cast<const uint32_t*>(Ax) + k, pol)), gather(ld_stream_u32(Aj + k, pol)));
                }
            }
            prefetch(tile + n_warps);
            __syncwarp();

            // ---- phase B: fold the rows of the tile from shared memory ----
            for (uint32_t base = row_first; base < row_hi; base += 32) {
                const uint32_t row   = base + lane;
                const bool     valid = row < row_hi;
                bool           take = false, head = false, tail = false;
                uint32_t       s = 0, e = 0;
                if (valid) {
                    uint32_t ra, rb;
                    if (base == row_first) {
                        ra = a0, rb = b0, take = take0;
                    } else {
                        take = all ? true : (MASKED ? sel.test(mask[row]) : false);
                        ra = __ldg(Ap + row), rb = __ldg(Ap + row + 1);
                    }
                    head = ra < lo;
                    tail = rb > hi;
                    s    = max(ra, lo) - lo;
                    e    = min(rb, hi) - lo;
                    if (!take && !head && !tail) r[row] = init;// partial segments of unselected rows are never read by the fix-up
                }
                const bool is_long = valid && take && (e - s > (uint32_t) kShort);
                if (valid && take && !is_long) {
                    T acc = sr.identity();
                    for (uint32_t k = s; k < e; ++k) acc = sr.add(acc, s_prod[k]);
                    if (head) carry[2 * tile] = acc;
                    else if (tail) carry[2 * tile + 1] = acc;
                    else r[row] = (e > s) ? sr.add(init, acc) : init;
                }
                uint32_t long_mask = __ballot_sync(0xffffffffu, is_long);
                while (long_mask) {
                    const int      src = __ffs(long_mask) - 1;
                    long_mask &= long_mask - 1;
                    const uint32_t ls = __shfl_sync(0xffffffffu, s, src), le = __shfl_sync(0xffffffffu, e, src);
                    T              acc = sr.identity();
                    for (uint32_t k = ls + lane; k < le; k += 32) acc = sr.add(acc, s_prod[k]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                    if ((int) lane == src) {
                        if (head) carry[2 * tile] = acc;
                        else if (tail) carry[2 * tile + 1] = acc;
                        else r[row] = sr.add(init, acc);
                    }
                }
            }
            __syncwarp();// the product slice is reused by this warp's next tile
        }
    }

    // r[row] of rows that start in tile t and end in a later tile: tail(t) + head(t+1) + ... chained left to right.
    // One thread per tile; chains longer than 4 tiles (hub rows) are summed by the whole warp in a fixed order.
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_fixup_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const T* __restrict__ mask,
                                                               T* __restrict__ r, T init, uint32_t n_tiles, const uint2* __restrict__ tile_rows,
                                                               const T* __restrict__ carry) {
        const uint32_t t    = blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t       row = 0, chain = 0;// chain = number of later tiles the row reaches into
        if (t < n_tiles) {
            const uint32_t row_hi = tile_rows[t].y;
            if (row_hi > 0) {
                row                = row_hi - 1;// the last row with entries in the tile
                const uint64_t end = Ap[row + 1], own = (uint64_t) (t + 1) * kMxvTile;
                if (end > own && (uint64_t) Ap[row] >= own - kMxvTile) {// starts in this tile, ends later (last tile: end <= nnz <= own)
                    const bool take = sel.reads_mask ? sel.test(mask[row]) : (sel.classes != 0u);
                    if (take) chain = (uint32_t) ((end - own + kMxvTile - 1) / kMxvTile);
                    else r[row] = init;
                }
            }
        }
        if (chain > 0 && chain <= 4) {
            T acc = carry[2 * t + 1];
            for (uint32_t u = 1; u <= chain; ++u) acc = sr.add(acc, carry[2 * (t + u)]);
            r[row] = sr.add(init, acc);
        }
        uint32_t long_mask = __ballot_sync(0xffffffffu, chain > 4);
        while (long_mask) {
            const int      src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            const uint32_t t0 = __shfl_sync(0xffffffffu, t, src), len = __shfl_sync(0xffffffffu, chain, src);
            T              acc = sr.identity();
            for (uint32_t u = 1 + lane; u <= len; u += 32) acc = sr.add(acc, carry[2 * (t0 + u)]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if ((int) lane == src) r[row] = sr.add(init, sr.add(carry[2 * t + 1], acc));
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

    // early_exit (BFS bottom-up): r[row] = add(init, p_k*) for the FIRST entry k* of the row whose single-term fold differs
    // from init, else init -- exactly the state of the reference's sequential fold when it breaks (SURVEY 8a note E), for
    // any op pair. A warp takes 32 rows: every lane scans the first kEarlySerial entries of its own row (most rows of a
    // dense frontier stop there); rows that are still undecided are then scanned by the whole warp, 128 entries per step,
    // with a ballot picking the lowest qualifying position -- a hub row costs nnz/128 steps instead of nnz.
    static constexpr int kEarlySerial = 8;
    static constexpr int kEarlyUnroll = 4;

    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_early_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj,
                                                               const T* __restrict__ Ax, const T* __restrict__ v, const T* __restrict__ mask,
                                                               T* __restrict__ r, T init, uint32_t n_rows) {
        const uint32_t lane    = threadIdx.x & 31u;
        const uint32_t warp    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
        for (uint32_t base = warp * 32u; base < n_rows; base += n_warps * 32u) {
            const uint32_t row   = base + lane;
            const bool     valid = row < n_rows;
            const bool     take  = valid && (sel.reads_mask ? sel.test(mask[row]) : (sel.classes != 0u));
            uint32_t       k0 = 0, k1 = 0;
            if (take) {
                k0 = Ap[row];
                k1 = Ap[row + 1];
            }
            T    res  = init;
            bool done = true;
            if (k1 > k0) {
                const uint32_t kend = min(k1, k0 + (uint32_t) kEarlySerial);
                done                = (kend == k1);
                for (uint32_t k = k0; k < kend; ++k) {
                    const T s = sr.add(init, sr.mult(Ax[k], v[Aj[k]]));
                    if (value_neq(s, init)) {
                        res  = s;
                        done = true;
                        break;
                    }
                }
            }
            uint32_t pending = __ballot_sync(0xffffffffu, !done);
            while (pending) {
                const int      src = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint32_t ks = __shfl_sync(0xffffffffu, k0, src) + (uint32_t) kEarlySerial;
                const uint32_t ke = __shfl_sync(0xffffffffu, k1, src);
                bool           hit   = false;
                T              found = init;
                for (uint32_t kb = ks; kb < ke && !hit; kb += 32u * kEarlyUnroll) {
                    T    s[kEarlyUnroll];
                    bool q[kEarlyUnroll];
#pragma unroll
                    for (int u = 0; u < kEarlyUnroll; ++u) {
                        const uint32_t k = kb + u * 32u + lane;
                        q[u]             = k < ke;
                        s[u]             = q[u] ? sr.add(init, sr.mult(Ax[k], v[Aj[k]])) : init;
                    }
#pragma unroll
                    for (int u = 0; u < kEarlyUnroll; ++u) {
                        const uint32_t m = __ballot_sync(0xffffffffu, q[u] && value_neq(s[u], init));
                        if (m && !hit) {
                            found = __shfl_sync(0xffffffffu, s[u], __ffs(m) - 1);
                            hit   = true;
                        }
                    }
                }
                if ((int) lane == src && hit) res = found;
            }
            if (valid) r[row] = res;
        }
    }

    template<typename T, typename S, bool MASKED, bool HUB>
    static int launch_wtile(S sr, Select sel, const Csr* M, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        auto           kern = mxv_wtile_kernel<T, S, MASKED, HUB>;
        const uint32_t smem = kProdBytes + (HUB ? M->n_hub_smem * 4u : 0u);
        static bool    attr_done = false;// per instantiation
        if (!attr_done) {
            SPLACU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemLimit));
            attr_done = true;
        }
        const uint32_t want = (M->n_tiles + kWarps - 1) / kWarps;
        const int      grid = (int) (want < (uint32_t) sm_count() ? want : (uint32_t) sm_count());
        kern<<<grid, kThreads, smem, s>>>(sr, sel, M->Ap, HUB ? M->Aj_hub : M->Aj, reinterpret_cast<const T*>(M->Ax), v, mask, r, init, M->nnz,
                                         M->n_tiles, M->tile_rows, reinterpret_cast<T*>(M->carry), (int) M->vec_ok, M->hub_vals, M->n_hub_smem);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    // keep v resident in L2 while the CSR arrays stream through it (access-policy window on the launching stream)
    static int set_persisting_window(const void* base, size_t bytes, cudaStream_t s) {
        static const void*  cur_base  = nullptr;
        static size_t       cur_bytes = 0;
        static cudaStream_t cur_s     = nullptr;
        static size_t       max_win = 0, max_persist = 0;
        static bool         probed = false;
        if (!probed) {
            probed = true;
            int dev = 0, v1 = 0, v2 = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&v1, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            cudaDeviceGetAttribute(&v2, cudaDevAttrMaxPersistingL2CacheSize, dev);
            max_win     = (size_t) v1;
            max_persist = (size_t) v2;
            if (max_persist) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist);
            cudaGetLastError();
        }
        if (!max_win || !max_persist) return 0;
        if (bytes > max_win) bytes = max_win;
        if (base == cur_base && bytes == cur_bytes && s == cur_s) return 0;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr  = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio  = bytes <= max_persist ? 1.0f : (float) max_persist / (float) bytes;
        attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
        SPLACU_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
        cur_base = base, cur_bytes = bytes, cur_s = s;
        return 0;
    }

    template<typename T, typename S>
    static int launch_tiles(S sr, Select sel, const Csr* M, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        int rc;
        if (get_option(OPT_MXV_L2_PERSIST) && (size_t) M->nnz * 8 > (size_t) 64 << 20)
            if ((rc = set_persisting_window(v, (size_t) M->n_cols * 4, s))) return rc;
        if (M->n_hub) {
            mxv_hub_pack_kernel<<<(M->n_hub + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->hub_cols, M->n_hub, reinterpret_cast<const uint32_t*>(v), M->hub_vals);
            SPLACU_LAUNCH_CHECK();
            rc = sel.reads_mask ? launch_wtile<T, S, true, true>(sr, sel, M, v, mask, r, init, s) : launch_wtile<T, S, false, true>(sr, sel, M, v, mask, r, init, s);
        } else {
            rc = sel.reads_mask ? launch_wtile<T, S, true, false>(sr, sel, M, v, mask, r, init, s) : launch_wtile<T, S, false, false>(sr, sel, M, v, mask, r, init, s);
        }
        if (rc) return rc;
        mxv_fixup_kernel<T, S><<<(M->n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, sel, M->Ap, mask, r, init, M->n_tiles, M->tile_rows,
                                                                                    reinterpret_cast<const T*>(M->carry));
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
            if (early_exit) {
                mxv_early_kernel<T, S><<<grid_for((size_t) M->n_rows, kBlock, 8), kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, reinterpret_cast<const T*>(M->Ax), v,
                                                                                                   mask, r, init, M->n_rows);
                SPLACU_LAUNCH_CHECK();
                return 0;
            }
            mxv_seq_kernel<T, S><<<grid_for(M->n_rows, kBlock, 8), kBlock, 0, s>>>(sr, sel, M->Ap, M->Aj, reinterpret_cast<const T*>(M->Ax), v, mask,
                                                                                  r, init, M->n_rows, early_exit);
            SPLACU_LAUNCH_CHECK();
            return 0;
        });
    });
}
