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
#include "profile.cuh"
#include "ops.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstring>
#include <vector>

namespace splacu {

    static constexpr int      kBlock     = 256;
    static constexpr int      kWarps     = 24;                   // warps per persistent CTA of the streaming kernel: 768 threads
                                                                 // (80 registers each), one CTA per SM
    static constexpr int      kWarpsHub  = 24;                   // warps per CTA of the hub-class phases (no gather latency to hide)
    static constexpr int      kItems     = kMxvTile / 32;        // entries per lane per tile (16)
    static constexpr int      kGroups    = kItems / 4;           // 128-bit groups per lane per tile (4)
    static constexpr uint32_t kHubFlag   = 0x80000000u;          // Aj_hub entry = kHubFlag | slot
    static constexpr uint32_t kSmemLimit = 227u * 1024u;         // opt-in dynamic shared memory per CTA on sm_100
    static constexpr uint32_t kSliceWords = kMxvTile + 16;       // per-warp slice: 512 products + 512 row-end flag bits
    static constexpr uint32_t kProdBytes = kWarps * kSliceWords * 4;// ~50 KB of per-warp slices
    static constexpr uint32_t kHubCap    = (kSmemLimit - kProdBytes - 1024u) / 4u & ~3u;// hub slots per CTA (~41 K)
    static constexpr uint32_t kPhaseCap  = (kSmemLimit - kWarpsHub * kSliceWords * 4u) / 4u & ~3u;// slots of one hub class (45 K)
    static_assert(kWarpsHub * 32 <= 1024 && kPhaseCap <= 65536u, "hub-class slots are 16-bit");
    enum { MODE_PLAIN = 0, MODE_HUB = 1, MODE_SMEM16 = 2 };
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
    __device__ __forceinline__ uint2 ld_stream_u2(const uint2* p, uint64_t pol) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
        return r;
    }
    __device__ __forceinline__ uint32_t ld_stream_u16(const uint16_t* p, uint64_t pol) {
        uint16_t r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
        return r;
    }
    __device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol) {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
        return r;
    }

    __device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

    // The product slice is written in the streaming layout (lane owns the 16-byte groups c * 32 + lane) and read back in the
    // blocked layout (lane owns groups 4 * lane .. 4 * lane + 3): group g lives at g ^ ((g >> 3) & 3), which keeps the 8 lanes
    // of every 128-bit shared-memory phase on 8 different bank groups in both directions.
    __device__ __forceinline__ uint32_t swz_group(uint32_t g) { return g ^ ((g >> 3) & 3u); }
    __device__ __forceinline__ uint32_t swz_entry(uint32_t e) { return (swz_group(e >> 2) << 2) | (e & 3u); }

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

    // ---- column-class phases (built once per matrix) ---------------------------------------------------
    // Random 4-byte gathers that leave the SM cost one L1->L2 request each (~290 G/s on the whole GPU), streaming the
    // matrix does not. So the entries are split by the popularity class of their column: class p holds the entries whose
    // column ranks in [p * S, (p + 1) * S) of the most referenced columns (S <= 49152 values = one shared-memory table),
    // as its own CSR over all rows with 16-bit slots instead of column ids (6 bytes per entry); the tail class keeps the
    // rest with 32-bit column ids. Each class is one pass of the streaming kernel; only the tail pass gathers from L2.
    struct PhasePtrs {
        uint32_t* Ap[kMaxHubPhases + 1];
        void*     Aj[kMaxHubPhases + 1];
        uint32_t* Ax[kMaxHubPhases + 1];
    };
    // Tail entries are further split by column RANGE (2^range_shift columns each) once v is larger than what the L2 keeps:
    // a pass then gathers from one L2-resident window of v instead of all of it (measured on RMAT-25 / 26, where v is 128 /
    // 256 MB: gathers that miss the L2 run at 67-96 G/s instead of ~280 G/s).
    __device__ __forceinline__ uint32_t phase_of(uint32_t slot, uint32_t col, uint32_t slots_per_phase, uint32_t n_hub_phases, uint32_t range_shift) {
        return slot == 0xffffffffu ? n_hub_phases + (col >> range_shift) : slot / slots_per_phase;
    }
    // a warp per row: cnt[p][row] = entries of the row in class p
    __global__ void __launch_bounds__(kBlock) phase_count_kernel(const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj, uint32_t n_rows,
                                                                 const uint32_t* __restrict__ slot, uint32_t slots_per_phase, uint32_t n_hub_phases,
                                                                 uint32_t n_classes, uint32_t range_shift, uint32_t* __restrict__ cnt) {
        const uint32_t lane    = threadIdx.x & 31u;
        const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
        for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
            const uint32_t k0 = Ap[row], k1 = Ap[row + 1];
            uint32_t       mine = 0;
            for (uint32_t kb = k0; kb < k1; kb += 32) {
                const uint32_t k = kb + lane;
                uint32_t       p = 0xffu;
                if (k < k1) {
                    const uint32_t col = Aj[k];
                    p                  = phase_of(slot[col], col, slots_per_phase, n_hub_phases, range_shift);
                }
                for (uint32_t q = 0; q < n_classes; ++q) {
                    const uint32_t m = __ballot_sync(0xffffffffu, p == q);
                    if (lane == q) mine += __popc(m);
                }
            }
            if (lane < n_classes) cnt[(size_t) lane * (n_rows + 1) + row] = mine;
        }
    }
    // a warp per row: stable partition of the row's entries into the class arrays (column order is kept inside a class)
    __global__ void __launch_bounds__(kBlock) phase_scatter_kernel(const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj,
                                                                   const uint32_t* __restrict__ Ax, uint32_t n_rows, const uint32_t* __restrict__ slot,
                                                                   uint32_t slots_per_phase, uint32_t n_hub_phases, uint32_t n_classes,
                                                                   uint32_t range_shift, PhasePtrs out, int seg, const uint32_t* __restrict__ row_slot) {
        const uint32_t lane    = threadIdx.x & 31u;
        const uint32_t lt      = (1u << lane) - 1u;
        const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
        for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
            const uint32_t k0 = Ap[row], k1 = Ap[row + 1];
            const bool     row_class = row_slot && row_slot[row] != 0xffffffffu;// its tail entries live in a row class (mxv_scat.cu)
            uint32_t       off = lane < n_classes ? out.Ap[lane][row] : 0u;// lane q tracks the write position of class q
            for (uint32_t kb = k0; kb < k1; kb += 32) {
                const uint32_t k  = kb + lane;
                const bool     ok = k < k1;
                uint32_t       col = 0, sl = 0xffffffffu, val = 0;
                if (ok) {
                    col = Aj[k];
                    sl  = slot[col];
                    val = Ax[k];
                }
                const uint32_t p = (ok && !(row_class && sl == 0xffffffffu)) ? phase_of(sl, col, slots_per_phase, n_hub_phases, range_shift) : 0xffu;
                for (uint32_t q = 0; q < n_classes; ++q) {
                    const uint32_t m    = __ballot_sync(0xffffffffu, p == q);
                    const uint32_t base = __shfl_sync(0xffffffffu, off, q);
                    if (p == q) {
                        const uint32_t dst = base + __popc(m & lt);// position in the row order of the class
                        if (q < n_hub_phases) reinterpret_cast<uint16_t*>(out.Aj[q])[seg ? seg_pos16(dst) : dst] = (uint16_t) (sl - q * slots_per_phase);
                        else reinterpret_cast<uint32_t*>(out.Aj[q])[seg ? seg_pos32(dst) : dst] = col;
                        out.Ax[q][seg ? seg_pos32(dst) : dst] = val;
                    }
                    if (lane == q) off += __popc(m);
                }
            }
        }
    }

    // rows ranked by their number of tail entries (all tail windows): sort keys, most entries first
    __global__ void __launch_bounds__(kBlock) tail_keys_kernel(const uint32_t* __restrict__ cnt, size_t stride, uint32_t first_tail, uint32_t n_classes,
                                                               uint32_t n_rows, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
        const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
        if (row >= n_rows) return;
        uint32_t c = 0;
        for (uint32_t w = first_tail; w < n_classes; ++w) c += cnt[w * stride + row];
        keys[row] = ~c;
        ids[row]  = row;
    }
    // the rows of the row classes leave the tail classes
    __global__ void __launch_bounds__(kBlock) tail_clear_kernel(uint32_t* __restrict__ cnt, size_t stride, uint32_t first_tail, uint32_t n_classes,
                                                                const uint32_t* __restrict__ rows, uint32_t n_hub_rows) {
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k >= n_hub_rows) return;
        const uint32_t row = rows[k];
        for (uint32_t w = first_tail; w < n_classes; ++w) cnt[w * stride + row] = 0u;
    }

    static void free_phases(Csr* M) {
        scat_free(M);
        for (int p = 0; p < M->n_phases; ++p) {
            CsrPhase& ph = M->phase[p];
            cudaFree(ph.Ap); cudaFree(ph.Aj); cudaFree(ph.Ax); cudaFree(ph.tile_rows); cudaFree(ph.carry);
            cudaFree(ph.flags); cudaFree(ph.seg_base); cudaFree(ph.seg_row); cudaFree(ph.chain); cudaFree(ph.chain_row); cudaFree(ph.head); cudaFree(ph.tail);
            ph = CsrPhase();
        }
        M->n_phases = 0;
    }

    // slot[col] = rank of the column among the n_hub most referenced ones (else 0xffffffff), hub_cols = those columns
    static int build_phases(Csr* M, const uint32_t* d_slot, uint32_t n_hub, uint32_t slots_per_phase, cudaStream_t s) {
        const uint32_t n_hub_phases = (n_hub + slots_per_phase - 1) / slots_per_phase;
        // tail classes: one per window of 2^range_shift columns; widen the windows until the classes fit the handle
        uint32_t range_shift = (uint32_t) get_option(OPT_MXV_TAIL_RANGE_LOG2);
        if (range_shift < 10 || range_shift > 31) range_shift = 31;
        while (range_shift < 31 && n_hub_phases + (((uint64_t) M->n_cols - 1) >> range_shift) + 1 > (uint64_t) kMaxHubPhases + 1) ++range_shift;
        const uint32_t n_classes = n_hub_phases + ((M->n_cols - 1) >> range_shift) + 1;
        const size_t   stride       = (size_t) M->n_rows + 1;
        uint32_t*      cnt          = nullptr;
        void*          tmp          = nullptr;
        int            rc           = 0;
        uint32_t *     rkeys = nullptr, *rids = nullptr, *rkeys_out = nullptr, *rows_sorted = nullptr, *row_slot = nullptr;
        void*          sort_tmp     = nullptr;
        uint32_t       n_hub_rows   = 0;
#define PH_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);          \
            goto done;                                                        \
        }                                                                     \
    } while (0)
        {
            PhasePtrs ptrs;
            memset(&ptrs, 0, sizeof(ptrs));
            size_t   tmp_bytes = 0;
            uint32_t totals[kMaxHubPhases + 1];
            const int seg = get_option(OPT_MXV_SEG) ? 1 : 0;
            M->n_phases = (int) n_classes;
            PH_CUDA(cudaMalloc(&cnt, n_classes * stride * 4));
            PH_CUDA(cudaMemsetAsync(cnt, 0, n_classes * stride * 4, s));
            phase_count_kernel<<<grid_for((size_t) M->n_rows * 32, kBlock, 8), kBlock, 0, s>>>(M->Ap, M->Aj, M->n_rows, d_slot, slots_per_phase, n_hub_phases, n_classes, range_shift, cnt);
            // row classes of the tail (mxv_scat.cu): the rows with the most tail entries hand those entries to a column-ordered copy
            const uint32_t max_row_classes = (uint32_t) get_option(OPT_MXV_ROW_CLASSES) < (uint32_t) kMaxScat ? (uint32_t) get_option(OPT_MXV_ROW_CLASSES) : (uint32_t) kMaxScat;
            if (seg && max_row_classes && n_hub_phases) {
                const uint32_t nr = M->n_rows;
                size_t         sort_bytes = 0;
                PH_CUDA(cudaMalloc(&rkeys, (size_t) nr * 4));
                PH_CUDA(cudaMalloc(&rids, (size_t) nr * 4));
                PH_CUDA(cudaMalloc(&rkeys_out, (size_t) nr * 4));
                PH_CUDA(cudaMalloc(&rows_sorted, (size_t) nr * 4));
                PH_CUDA(cudaMalloc(&row_slot, (size_t) nr * 4));
                tail_keys_kernel<<<(nr + kBlock - 1) / kBlock, kBlock, 0, s>>>(cnt, stride, n_hub_phases, n_classes, nr, rkeys, rids);
                PH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, rkeys, rkeys_out, rids, rows_sorted, (int) nr, 0, 32, s));
                PH_CUDA(cudaMalloc(&sort_tmp, sort_bytes));
                PH_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, rkeys, rkeys_out, rids, rows_sorted, (int) nr, 0, 32, s));
                uint32_t cap = max_row_classes * slots_per_phase;
                if (cap > nr) cap = nr;
                std::vector<uint32_t> h_keys(cap);
                PH_CUDA(cudaMemcpyAsync(h_keys.data(), rkeys_out, (size_t) cap * 4, cudaMemcpyDeviceToHost, s));
                PH_CUDA(cudaStreamSynchronize(s));
                const uint32_t min_count = (uint32_t) get_option(OPT_MXV_ROW_MIN_COUNT) ? (uint32_t) get_option(OPT_MXV_ROW_MIN_COUNT) : 1u;
                while (n_hub_rows < cap && ~h_keys[n_hub_rows] >= min_count) ++n_hub_rows;
                // a row class has fixed costs (table init, one partial table per CTA, the merge: ~45 us on a B200) and saves ~2 ps per
                // entry against the tail pass: below ~24 M entries it loses (a rank of an 8-GPU run on RMAT-24: 92 us for 17.9 M entries
                // against 57 us more in the tail pass), so the classes that would hold fewer entries are not built
                {
                    const uint64_t min_nnz = (uint64_t) get_option(OPT_MXV_ROW_MIN_NNZ);
                    uint32_t       keep    = 0;
                    for (uint32_t q0 = 0; q0 < n_hub_rows; q0 += slots_per_phase) {
                        uint64_t       sum = 0;
                        const uint32_t q1  = q0 + slots_per_phase < n_hub_rows ? q0 + slots_per_phase : n_hub_rows;
                        for (uint32_t q = q0; q < q1; ++q) sum += ~h_keys[q];
                        if (sum < min_nnz) break;
                        keep = q1;
                    }
                    n_hub_rows = keep;
                }
                count_launch(5);
                if (n_hub_rows) {
                    PH_CUDA(cudaMemsetAsync(row_slot, 0xff, (size_t) nr * 4, s));
                    hub_slots_kernel<<<(n_hub_rows + kBlock - 1) / kBlock, kBlock, 0, s>>>(rows_sorted, n_hub_rows, row_slot);
                    tail_clear_kernel<<<(n_hub_rows + kBlock - 1) / kBlock, kBlock, 0, s>>>(cnt, stride, n_hub_phases, n_classes, rows_sorted, n_hub_rows);
                    count_launch(2);
                }
            }
            PH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, cnt, (int) stride, s));
            PH_CUDA(cudaMalloc(&tmp, tmp_bytes));
            for (uint32_t p = 0; p < n_classes; ++p) {
                CsrPhase& ph = M->phase[p];
                PH_CUDA(cudaMalloc(&ph.Ap, stride * 4));
                PH_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt + p * stride, ph.Ap, (int) stride, s));
                PH_CUDA(cudaMemcpyAsync(&totals[p], ph.Ap + M->n_rows, 4, cudaMemcpyDeviceToHost, s));
            }
            PH_CUDA(cudaStreamSynchronize(s));
            count_launch(1 + 2 * (int) n_classes);
            for (uint32_t p = 0; p < n_classes; ++p) {
                CsrPhase& ph  = M->phase[p];
                ph.nnz        = totals[p];
                ph.idx16      = p < n_hub_phases;
                ph.slot_base  = ph.idx16 ? p * slots_per_phase : 0u;
                ph.n_slots    = ph.idx16 ? (n_hub - ph.slot_base < slots_per_phase ? n_hub - ph.slot_base : slots_per_phase) : 0u;
                ph.n_tiles    = (uint32_t) (((uint64_t) ph.nnz + kMxvTile - 1) / kMxvTile);
                const size_t padded = (size_t) (ph.n_tiles ? ph.n_tiles : 1) * kMxvTile;
                PH_CUDA(cudaMalloc(&ph.Aj, padded * (ph.idx16 ? 2 : 4)));
                PH_CUDA(cudaMalloc(&ph.Ax, padded * 4));
                if (seg) {// the padding of the last tile is read by the kernel
                    PH_CUDA(cudaMemsetAsync(ph.Aj, 0, padded * (ph.idx16 ? 2 : 4), s));
                    PH_CUDA(cudaMemsetAsync(ph.Ax, 0, padded * 4, s));
                }
                ptrs.Ap[p] = ph.Ap, ptrs.Aj[p] = ph.Aj, ptrs.Ax[p] = ph.Ax;
            }
            phase_scatter_kernel<<<grid_for((size_t) M->n_rows * 32, kBlock, 8), kBlock, 0, s>>>(M->Ap, M->Aj, M->Ax, M->n_rows, d_slot, slots_per_phase, n_hub_phases, n_classes, range_shift, ptrs, seg, n_hub_rows ? row_slot : nullptr);
            count_launch(1);
            for (uint32_t p = 0; p < n_classes; ++p) {
                CsrPhase& ph = M->phase[p];
                if (ph.n_tiles == 0) continue;
                if (seg) {
                    if ((rc = seg_build(M, ph, cnt + p * stride, s))) goto done;
                    continue;
                }
                PH_CUDA(cudaMalloc(&ph.tile_rows, (size_t) ph.n_tiles * sizeof(uint2)));
                PH_CUDA(cudaMalloc(&ph.carry, (size_t) ph.n_tiles * 2 * sizeof(uint32_t)));
                tile_rows_kernel<<<(ph.n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(ph.Ap, M->n_rows, kMxvTile, ph.n_tiles, ph.tile_rows);
                count_launch(1);
            }
            PH_CUDA(cudaStreamSynchronize(s));
            PH_CUDA(cudaGetLastError());
            if (n_hub_rows && (rc = scat_build(M, d_slot, rows_sorted, n_hub_rows, slots_per_phase, s))) goto done;
            if (seg && (rc = seg_build_fixlist(M, s))) goto done;
        }
    done:
#undef PH_CUDA
        cudaFree(cnt);
        cudaFree(tmp);
        cudaFree(rkeys); cudaFree(rids); cudaFree(rkeys_out); cudaFree(rows_sorted); cudaFree(row_slot); cudaFree(sort_tmp);
        if (rc) free_phases(M);
        return rc;
    }

    static int build_hub(Csr* M, cudaStream_t s) {
        const int mode = (int) get_option(OPT_MXV_HUB);// 0 off, 1 auto (phases), 2 force the single-pass hub cache, 3 force phases
        if (mode == 0 || !M->vec_ok || M->n_cols >= kHubFlag) return 0;
        if (mode == 1 && (M->nnz < (1u << 22) || M->n_cols < 4 * kHubCap)) return 0;// v already fits on chip / too little work
        const bool     phases = mode == 1 || mode == 3;
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
        // how many of the most referenced columns are referenced often enough to pay for their slot
        uint32_t slots_per_phase = (uint32_t) get_option(OPT_MXV_PHASE_SLOTS) & ~3u;
        if (slots_per_phase > kPhaseCap) slots_per_phase = kPhaseCap;
        if (slots_per_phase < 4) slots_per_phase = 4;
        uint32_t max_phases = (uint32_t) get_option(OPT_MXV_PHASES);
        if (max_phases > (uint32_t) kMaxHubPhases) max_phases = kMaxHubPhases;
        if (max_phases < 1) max_phases = 1;
        uint32_t cap = phases ? max_phases * slots_per_phase : (uint32_t) get_option(OPT_MXV_HUB_TOTAL);
        if (cap > n) cap = n;
        if (cap < 4) cap = 4;
        const uint32_t min_count = (uint32_t) get_option(OPT_MXV_HUB_MIN_COUNT);
        std::vector<uint32_t> h_keys(cap);
        HUB_CUDA(cudaMemcpyAsync(h_keys.data(), keys_out, (size_t) cap * 4, cudaMemcpyDeviceToHost, s));
        HUB_CUDA(cudaStreamSynchronize(s));
        uint32_t n_hub = 0;
        while (n_hub < cap && ~h_keys[n_hub] >= min_count) ++n_hub;
        if (n_hub >= 64 || mode >= 2) {
            if (n_hub == 0) n_hub = cap < 4 ? cap : 4;
            HUB_CUDA(cudaMalloc(&M->hub_cols, (size_t) n_hub * 4));
            HUB_CUDA(cudaMalloc(&M->hub_vals, ((size_t) n_hub + 4) * 4));
            HUB_CUDA(cudaMemsetAsync(M->hub_vals, 0, ((size_t) n_hub + 4) * 4, s));
            HUB_CUDA(cudaMemcpyAsync(M->hub_cols, ids_out, (size_t) n_hub * 4, cudaMemcpyDeviceToDevice, s));
            HUB_CUDA(cudaMemsetAsync(count, 0xff, (size_t) n * 4, s));// reuse as the slot map
            hub_slots_kernel<<<(n_hub + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->hub_cols, n_hub, count);
            count_launch(1);
            M->n_hub = n_hub;
            if (phases) {
                const int rc = build_phases(M, count, n_hub, slots_per_phase, s);
                if (rc) {
                    cleanup();
                    return rc;
                }
                HUB_CUDA(cudaMalloc(&M->sel_count, 16));// [0] the count, [1] running total, [2] ticket of the count pass (both return to 0)
                HUB_CUDA(cudaMemsetAsync(M->sel_count, 0, 16, s));
                HUB_CUDA(cudaMalloc(&M->sel_bits, ((size_t) M->n_rows + 31) / 32 * 4 + 4));
            } else {
                HUB_CUDA(cudaMalloc(&M->Aj_hub, (size_t) M->nnz * 4));
                hub_encode_kernel<<<grid_for(M->nnz, kBlock, 8), kBlock, 0, s>>>(M->Aj, M->nnz, count, M->Aj_hub);
                count_launch(1);
                M->n_hub_smem = (uint32_t) get_option(OPT_MXV_HUB_SMEM) & ~3u;
                if (M->n_hub_smem > kHubCap) M->n_hub_smem = kHubCap;
                if (M->n_hub_smem > n_hub) M->n_hub_smem = n_hub & ~3u;
            }
        }
        HUB_CUDA(cudaStreamSynchronize(s));
#undef HUB_CUDA
        cleanup();
        return 0;
    }

    // every stored value equal to the first one? (adjacency matrices: the push product can then run structure-only)
    __global__ void __launch_bounds__(kBlock) ax_uniform_kernel(const uint32_t* __restrict__ Ax, uint32_t nnz, uint32_t* __restrict__ differs) {
        const uint32_t first  = Ax[0];
        const uint32_t stride = gridDim.x * blockDim.x;
        bool           d      = false;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) d |= Ax[k] != first;
        if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31u) == 0u) *differs = 1u;
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
        {
            // carry[] is per-call scratch: its first word doubles as the flag of this one-off check
            uint32_t h[2] = {0u, 0u};
            SPLACU_CUDA(cudaMemsetAsync(M->carry, 0, 4, s));
            ax_uniform_kernel<<<grid_for(M->nnz, kBlock, 8), kBlock, 0, s>>>(M->Ax, M->nnz, M->carry);
            SPLACU_LAUNCH_CHECK();
            SPLACU_CUDA(cudaMemcpyAsync(&h[0], M->carry, 4, cudaMemcpyDeviceToHost, s));
            SPLACU_CUDA(cudaMemcpyAsync(&h[1], M->Ax, 4, cudaMemcpyDeviceToHost, s));
            SPLACU_CUDA(cudaStreamSynchronize(s));
            M->ax_uniform = h[0] == 0u;
            M->ax_value   = h[1];
        }
        return build_hub(M, s);
    }

    // ---- the streaming kernel ------------------------------------------------------------------------
    // MODE_PLAIN  : Aj = 32-bit column ids, every gather goes to v (L1 / L2)
    // MODE_HUB    : Aj = Aj_hub (hub columns replaced by slots); slots below n_hub_smem are served from shared memory
    // MODE_SMEM16 : one hub column class: Aj = 16-bit slots into a table of n_hub_smem values of v held in shared memory,
    //               no gather ever leaves the SM
    // accum != 0  : a later column-class phase: r[row] = add(r[row], sum of the class) for the rows that have entries in it
    // registers of one tile in flight: its Aj / Ax slices (4 x 128 bit each per lane) and its row range
    template<typename T> struct TileRegs {
        uint4    j[kGroups], a[kGroups];
        uint2    rows;
        bool     streamed;
    };

    template<typename T, typename S, bool MASKED, int MODE, int WARPS>
    __global__ void __launch_bounds__(WARPS * 32, 1)
            mxv_wtile_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj, const T* __restrict__ Ax,
                             const T* __restrict__ v, const T* __restrict__ mask, T* r, T init, uint32_t nnz, uint32_t n_tiles,
                             const uint2* __restrict__ tile_rows, T* __restrict__ carry, int vec_ok, const uint32_t* __restrict__ hub_vals,
                             uint32_t n_hub_smem, int accum, const uint32_t* __restrict__ gate, uint32_t gate_max) {
        extern __shared__ __align__(16) uint32_t smem[];
        if (gate && *gate >= gate_max) return;// dense mask: the column-class passes (mxv_seg.cu) run instead
        // DEPTH tiles of Aj / Ax in flight per warp, RPL rows per lane and row-loop step. Both stay at 1: unrolling the tile body
        // (DEPTH 2, RPL 4: 5.5 K instructions) made the kernel instruction-fetch bound (ncu: stalled_no_instruction 9.6 per issue),
        // and 16 warps with two tiles in flight by register rotation were slower than 24 warps with one.
        constexpr int      DEPTH    = 1;
        constexpr int      RPL      = 1;
        // a hub class spans many rows per tile (few entries of the class per row): its row loops request the extents of the
        // next rows before they handle the current ones
        constexpr bool     ROWPIPE  = MODE == MODE_SMEM16;
        constexpr uint32_t kThreads = WARPS * 32;
        const uint32_t     tid      = threadIdx.x;
        const uint32_t     lane     = tid & 31u;
        const uint32_t     warp     = tid >> 5;
        T*                 s_prod   = reinterpret_cast<T*>(smem) + warp * kSliceWords;// this warp's product slice ...
        uint32_t*          s_flag   = smem + warp * kSliceWords + kMxvTile;           // ... and its row-end flags
        const T*           s_hub    = reinterpret_cast<const T*>(smem) + WARPS * kSliceWords;
        const uint16_t*    Aj16     = reinterpret_cast<const uint16_t*>(Aj);

        if (MODE != MODE_PLAIN) {// hub values of v -> shared memory, 128-bit coalesced
            uint4*       dst = reinterpret_cast<uint4*>(smem + WARPS * kSliceWords);
            const uint4* src = reinterpret_cast<const uint4*>(hub_vals);
            for (uint32_t i = tid; i < (n_hub_smem + 3u) / 4u; i += kThreads) dst[i] = __ldg(src + i);
            __syncthreads();
        }

        const bool     all = !MASKED && (sel.classes != 0u);// ALWAYS (NEVER never gets here)
        const uint64_t pol = policy_evict_first();
        // MODE_HUB: hub slots below n_hub_smem live in shared memory, the rest of the (dense, packed) hub table is served by L1.
        // Branch-free: predicated loads.
        const uint32_t s_hub_addr = (uint32_t) __cvta_generic_to_shared(s_hub);
        auto gather = [&](uint32_t j) -> T {
            if (MODE == MODE_SMEM16) return s_hub[j];
            if (MODE == MODE_HUB) return from_bits<T>(gather_hub(j, reinterpret_cast<const uint32_t*>(v), hub_vals, s_hub_addr, n_hub_smem));
            return from_bits<T>(ld_gather_plain(reinterpret_cast<const uint32_t*>(v) + j));
        };
        // the 4 column ids of group c of the tile at entry lo: MODE_SMEM16 keeps them packed (2 x 16 bit in .x and .y)
        auto load_idx = [&](uint32_t lo, int c) -> uint4 {
            if (MODE == MODE_SMEM16) {
                const uint2 q = ld_stream_u2(reinterpret_cast<const uint2*>(Aj16 + lo) + c * 32 + lane, pol);
                return make_uint4(q.x, q.y, 0u, 0u);
            }
            return ld_stream_u4(reinterpret_cast<const uint4*>(Aj + lo) + c * 32 + lane, pol);
        };

        // Software pipeline: everything a tile needs from HBM (its row range and, unless the mask is sparse, its Aj / Ax slices)
        // is requested DEPTH tiles ahead, right after the registers of the tile DEPTH back are consumed, so that on entry a
        // warp only waits for its gathers.
        const uint32_t n_warps = gridDim.x * WARPS;
        bool           dense   = !MASKED;// masked variant: stream ahead only while the mask keeps selecting most of a tile
        auto           prefetch = [&](TileRegs<T>& tr, uint32_t t) {
            tr.streamed = false;
            if (t >= n_tiles) return;
            tr.rows = __ldg(tile_rows + t);
            if (!dense || !vec_ok || nnz - t * (uint32_t) kMxvTile < (uint32_t) kMxvTile) return;
#pragma unroll
            for (int c = 0; c < kGroups; ++c) tr.j[c] = load_idx(t * (uint32_t) kMxvTile, c);
#pragma unroll
            for (int c = 0; c < kGroups; ++c) tr.a[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Ax + t * (uint32_t) kMxvTile) + c * 32 + lane, pol);
            tr.streamed = true;
        };
        TileRegs<T> buf[DEPTH];
        buf[0].rows = make_uint2(0u, 0u);
        prefetch(buf[0], blockIdx.x * WARPS + warp);

        for (uint32_t tile = blockIdx.x * WARPS + warp; tile < n_tiles; tile += n_warps) {
            {
                TileRegs<T>&   tr        = buf[0];
                const uint32_t lo        = tile * (uint32_t) kMxvTile;
                const uint32_t hi        = (nnz - lo > (uint32_t) kMxvTile) ? lo + (uint32_t) kMxvTile : nnz;
                const uint32_t row_first = tr.rows.x;// first row with entries in this tile (the head row if one reaches in)
                const uint32_t row_hi    = tr.rows.y;// one past the last row that starts in this tile
                const bool     streamed  = tr.streamed;

                // row extents of the first 32 rows of the tile: requested now, consumed after the gathers
                uint32_t a0[RPL], b0[RPL];
                bool     take0[RPL];
                T        old0[RPL];
#pragma unroll
                for (int u = 0; u < RPL; ++u) {
                    const uint32_t row = row_first + u * 32 + lane;
                    a0[u] = b0[u] = 0u;
                    take0[u]      = false;
                    old0[u]       = init;
                    if (row < row_hi) {
                        a0[u]    = __ldg(Ap + row);
                        b0[u]    = __ldg(Ap + row + 1);
                        take0[u] = all ? true : (MASKED ? sel.test(mask[row]) : false);
                        if (accum) old0[u] = r[row];
                    }
                }

                // row-end flags of the tile (bit e: entry e is the last entry of its row inside the tile), set by the lanes that own
                // the rows, consumed by the segmented scan of phase B
                if (lane < 16) s_flag[lane] = 0u;
                __syncwarp();
                auto mark_end = [&](uint32_t s, uint32_t e) {// [s, e) tile-relative entry range of one row
                    if (e > s) atomicOr(&s_flag[(e - 1u) >> 5], 1u << ((e - 1u) & 31u));
                };
                if (MODE == MODE_SMEM16) {// its tiles span hundreds of rows: pull their extents (and r) into L2 ahead of the row loops
                    for (uint32_t row = row_first + 32 * RPL + lane * 32; row <= row_hi; row += 32 * 32) prefetch_l2(Ap + row);
                    if (accum)
                        for (uint32_t row = row_first + lane * 32; row < row_hi; row += 32 * 32) prefetch_l2(r + row);
                }

                uint32_t need[kGroups];
#pragma unroll
                for (int c = 0; c < kGroups; ++c) need[c] = 0xffffffffu;
                // a hub class whose slices are already in registers skips this pass (its gathers are on chip and cost nothing):
                // products of unselected rows are simply never folded; `dense` is then re-estimated by the row loop
                const bool need_pass = MASKED && !(MODE == MODE_SMEM16 && streamed);
                bool       idle      = false;// the mask selects no row of the tile: nothing to stream, gather or sum
                if (need_pass) {
                    // selected rows mark the 4-entry groups they need: the mask is tested before Aj / Ax / v are touched
#pragma unroll
                    for (int c = 0; c < kGroups; ++c) need[c] = 0u;
                    for (uint32_t row = row_first + lane; row < row_hi; row += 32) {
                        uint32_t ra, rb;
                        bool     tk;
                        if (row == row_first + lane) {
                            ra = a0[0], rb = b0[0], tk = take0[0];
                        } else {
                            tk = sel.test(mask[row]);
                            ra = __ldg(Ap + row), rb = __ldg(Ap + row + 1);
                        }
                        const uint32_t s = max(ra, lo), e = min(rb, hi);
                        mark_end(s - lo, e > s ? e - lo : s - lo);
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
                    idle  = n_need == 0u;
                }

                // ---- phase A: stream the tile, gather, multiply, park products in shared memory ----
                if (idle) {
                } else if (vec_ok && hi - lo == (uint32_t) kMxvTile) {
                    T x[kGroups][4];
                    if (!streamed) {
#pragma unroll
                        for (int c = 0; c < kGroups; ++c)
                            if ((need[c] >> lane) & 1u) tr.j[c] = load_idx(lo, c);
#pragma unroll
                        for (int c = 0; c < kGroups; ++c)
                            if ((need[c] >> lane) & 1u) tr.a[c] = ld_stream_u4(reinterpret_cast<const uint4*>(Ax + lo) + c * 32 + lane, pol);
                    }
#pragma unroll
                    for (int c = 0; c < kGroups; ++c)
                        if ((need[c] >> lane) & 1u) {
                            if (MODE == MODE_SMEM16) {
                                x[c][0] = gather(tr.j[c].x & 0xffffu);
                                x[c][1] = gather(tr.j[c].x >> 16);
                                x[c][2] = gather(tr.j[c].y & 0xffffu);
                                x[c][3] = gather(tr.j[c].y >> 16);
                            } else {
                                x[c][0] = gather(tr.j[c].x);
                                x[c][1] = gather(tr.j[c].y);
                                x[c][2] = gather(tr.j[c].z);
                                x[c][3] = gather(tr.j[c].w);
                            }
                        }
#pragma unroll
                    for (int c = 0; c < kGroups; ++c)
                        if ((need[c] >> lane) & 1u) {
                            uint4 p;
                            p.x = to_bits(sr.mult(from_bits<T>(tr.a[c].x), x[c][0]));
                            p.y = to_bits(sr.mult(from_bits<T>(tr.a[c].y), x[c][1]));
                            p.z = to_bits(sr.mult(from_bits<T>(tr.a[c].z), x[c][2]));
                            p.w = to_bits(sr.mult(from_bits<T>(tr.a[c].w), x[c][3]));
                            reinterpret_cast<uint4*>(s_prod)[swz_group(c * 32 + lane)] = p;
                        }
                } else {
                    for (uint32_t k = lo + lane; k < hi; k += 32) {
                        const uint32_t g  = (k - lo) >> 2;
                        bool           nd = true;
#pragma unroll
                        for (int c = 0; c < kGroups; ++c)
                            if ((g >> 5) == (uint32_t) c) nd = ((need[c] >> (g & 31u)) & 1u) != 0u;
                        if (nd) {
                            const uint32_t col = (MODE == MODE_SMEM16) ? ld_stream_u16(Aj16 + k, pol) : ld_stream_u32(Aj + k, pol);
                            s_prod[swz_entry(k - lo)] = sr.mult(from_bits<T>(ld_stream_u32(reinterpret_cast<const uint32_t*>(Ax) + k, pol)), gather(col));
                        }
                    }
                }
                prefetch(tr, tile + n_warps);
                __syncwarp();

                // ---- phase B: segmented sums of the 512 products, one segment per row ----
                // (1) the lanes that own the rows flag the last entry of every row (done by the need pass when that ran)
                uint32_t ra[RPL], rb[RPL];
                bool     tk[RPL];
                T        old[RPL];
                auto     load_rows = [&](uint32_t base, bool emit, uint32_t(&xa)[RPL], uint32_t(&xb)[RPL], bool(&xt)[RPL], T(&xo)[RPL]) {
#pragma unroll
                    for (int u = 0; u < RPL; ++u) {
                        const uint32_t row = base + u * 32 + lane;
                        xa[u] = xb[u] = 0u;
                        xt[u]         = false;
                        xo[u]         = init;
                        if (row < row_hi) {
                            xa[u] = __ldg(Ap + row), xb[u] = __ldg(Ap + row + 1);
                            if (emit) {// the marking pass needs the extents only
                                xt[u] = all ? true : (MASKED ? sel.test(mask[row]) : false);
                                if (accum) xo[u] = r[row];
                            }
                        }
                    }
                };
                if (!need_pass) {
#pragma unroll
                    for (int u = 0; u < RPL; ++u) ra[u] = a0[u], rb[u] = b0[u];
                    for (uint32_t base = row_first; base < row_hi; base += 32 * RPL) {
                        uint32_t ran[RPL], rbn[RPL];
                        bool     tkn[RPL];
                        T        oldn[RPL];
                        if (ROWPIPE) load_rows(base + 32 * RPL, false, ran, rbn, tkn, oldn);
                        else if (base != row_first) load_rows(base, false, ra, rb, tk, old);
#pragma unroll
                        for (int u = 0; u < RPL; ++u)
                            if (base + u * 32 + lane < row_hi) {
                                const uint32_t s = max(ra[u], lo), e = min(rb[u], hi);
                                mark_end(s - lo, e > s ? e - lo : s - lo);
                            }
                        if (ROWPIPE) {
#pragma unroll
                            for (int u = 0; u < RPL; ++u) ra[u] = ran[u], rb[u] = rbn[u];
                        }
                    }
                }
                __syncwarp();
                // (2) every lane scans its 16 consecutive products: the sum of a segment lands on its flagged entry. Segments that
                //     span lanes are joined by a warp-level segmented scan of the lanes' open tails (fixed order: deterministic).
                if (!idle) {
                    const uint32_t fl = (s_flag[lane >> 1] >> ((lane & 1u) * 16u)) & 0xffffu;
                    // pass 1: the lane's open tail = sum of the entries after its last flag (all 16 when it has none); the products
                    // are read 4 at a time, and again in pass 2, rather than held in 16 registers across the warp scan
                    T open = sr.identity();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 w = reinterpret_cast<const uint4*>(s_prod)[swz_group(4 * lane + q)];
                        open          = ((fl >> (4 * q + 0)) & 1u) ? sr.identity() : sr.add(open, from_bits<T>(w.x));
                        open          = ((fl >> (4 * q + 1)) & 1u) ? sr.identity() : sr.add(open, from_bits<T>(w.y));
                        open          = ((fl >> (4 * q + 2)) & 1u) ? sr.identity() : sr.add(open, from_bits<T>(w.z));
                        open          = ((fl >> (4 * q + 3)) & 1u) ? sr.identity() : sr.add(open, from_bits<T>(w.w));
                    }
                    T    v = open;
                    bool f = fl != 0u;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const T    vv = __shfl_up_sync(0xffffffffu, v, d);
                        const bool ff = __shfl_up_sync(0xffffffffu, (int) f, d) != 0;
                        if ((int) lane >= d) {
                            if (!f) v = sr.add(vv, v);
                            f = f || ff;
                        }
                    }
                    T acc = __shfl_up_sync(0xffffffffu, v, 1);// carry-in: open tails of the lanes before, back to the nearest flag
                    if (lane == 0) acc = sr.identity();
                    // pass 2: the sum of a segment lands on its flagged (last) entry
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4*      slot = reinterpret_cast<uint4*>(s_prod) + swz_group(4 * lane + q);
                        const uint4 w    = *slot;
                        uint4       o    = w;
                        acc              = sr.add(acc, from_bits<T>(w.x));
                        if ((fl >> (4 * q + 0)) & 1u) o.x = to_bits(acc), acc = sr.identity();
                        acc = sr.add(acc, from_bits<T>(w.y));
                        if ((fl >> (4 * q + 1)) & 1u) o.y = to_bits(acc), acc = sr.identity();
                        acc = sr.add(acc, from_bits<T>(w.z));
                        if ((fl >> (4 * q + 2)) & 1u) o.z = to_bits(acc), acc = sr.identity();
                        acc = sr.add(acc, from_bits<T>(w.w));
                        if ((fl >> (4 * q + 3)) & 1u) o.w = to_bits(acc), acc = sr.identity();
                        if ((fl >> (4 * q)) & 15u) *slot = o;
                    }
                }
                __syncwarp();
                // (3) the lanes that own the rows pick the sums up
                uint32_t n_sel = 0;// entries of selected rows (re-estimates `dense` when the need pass was skipped)
#pragma unroll
                for (int u = 0; u < RPL; ++u) ra[u] = a0[u], rb[u] = b0[u], tk[u] = take0[u], old[u] = old0[u];
                for (uint32_t base = row_first; base < row_hi; base += 32 * RPL) {
                    uint32_t ran[RPL], rbn[RPL];
                    bool     tkn[RPL];
                    T        oldn[RPL];
                    if (ROWPIPE) load_rows(base + 32 * RPL, true, ran, rbn, tkn, oldn);
                    else if (base != row_first) load_rows(base, true, ra, rb, tk, old);
#pragma unroll
                    for (int u = 0; u < RPL; ++u) {
                        const uint32_t row = base + u * 32 + lane;
                        if (row < row_hi) {
                            const bool     head = ra[u] < lo, tail = rb[u] > hi;
                            const uint32_t s = max(ra[u], lo) - lo, e = min(rb[u], hi) - lo;
                            if (tk[u]) {
                                if (MASKED && !need_pass) n_sel += e - s;
                                if (e > s) {
                                    const T sum = s_prod[swz_entry(e - 1u)];
                                    if (head) carry[2 * tile] = sum;
                                    else if (tail) carry[2 * tile + 1] = sum;
                                    else r[row] = sr.add(accum ? old[u] : init, sum);
                                } else if (!accum) {
                                    r[row] = init;
                                }
                            } else if (!head && !tail && !accum) {
                                // partial segments of unselected rows are never read by the fix-up; later phases leave unselected rows alone
                                r[row] = init;
                            }
                        }
                    }
                    if (ROWPIPE) {
#pragma unroll
                        for (int u = 0; u < RPL; ++u) ra[u] = ran[u], rb[u] = rbn[u], tk[u] = tkn[u], old[u] = oldn[u];
                    }
                }
                if (MASKED && !need_pass) dense = __reduce_add_sync(0xffffffffu, n_sel) * 2u >= hi - lo;
                __syncwarp();// the product slice is reused by this warp's next tile
            }
        }
    }

    // r[row] of rows that start in tile t and end in a later tile: tail(t) + head(t+1) + ... chained left to right.
    // One thread per tile; chains longer than 4 tiles (hub rows) are summed by the whole warp in a fixed order.
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_fixup_kernel(S sr, Select sel, const uint32_t* __restrict__ Ap, const T* __restrict__ mask,
                                                               T* r, T init, uint32_t n_tiles, const uint2* __restrict__ tile_rows,
                                                               const T* __restrict__ carry, int accum, const uint32_t* __restrict__ gate,
                                                               uint32_t gate_max) {
        if (gate && *gate >= gate_max) return;
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
                    else if (!accum) r[row] = init;
                }
            }
        }
        if (chain > 0 && chain <= 4) {
            T acc = carry[2 * t + 1];
            for (uint32_t u = 1; u <= chain; ++u) acc = sr.add(acc, carry[2 * (t + u)]);
            r[row] = sr.add(accum ? r[row] : init, acc);
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
            if ((int) lane == src) r[row] = sr.add(accum ? r[row] : init, sr.add(carry[2 * t + 1], acc));
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

    // one pass of the streaming kernel + its border fix-up over a CSR (the whole matrix or one column class)
    struct TileJob {
        const uint32_t* Ap;
        const void*     Aj;
        const uint32_t* Ax;
        uint32_t        nnz, n_tiles;
        const uint2*    tile_rows;
        uint32_t*       carry;
        int             vec_ok;
        const uint32_t* hub_vals;// table base for MODE_HUB / MODE_SMEM16
        uint32_t        n_smem;  // slots staged in shared memory
        int             accum;
        const uint32_t* gate;    // device counter of mask-selected rows: the pass runs only while *gate < gate_max (null: always)
        uint32_t        gate_max;
    };

    // rows selected by the mask of this call -> *out (zeroed before): lets the device choose between the column-class passes
    // (dense masks) and the CSR kernel that tests the mask before any gather (sparse masks) without a host round trip
    // (the same pass pre-fills r with init for the class passes, which accumulate onto it, and leaves select(mask[i]) as a bitmap:
    //  2 MB that stay in L2 instead of 64 MB read again by every class pass)
    // The blocks [0, n_main) count / fill; the blocks behind them pack the hub values of v (hub_vals[s] = v[hub_cols[s]]) when the call
    // is not split -- one launch less in front of the class passes. The counter needs no memset: the blocks add onto out[1], the last
    // one to finish (ticket out[2]) publishes the total in out[0] and returns out[1] / out[2] to zero.
    template<typename T>
    __global__ void __launch_bounds__(kBlock) mask_count_fill_kernel(Select sel, const T* __restrict__ mask, uint32_t n, uint32_t* __restrict__ out,
                                                                     uint32_t* __restrict__ sel_bits, T* __restrict__ r, T init, uint32_t n_main,
                                                                     const uint32_t* __restrict__ hub_cols, uint32_t n_hub, const uint32_t* __restrict__ v,
                                                                     uint32_t* __restrict__ hub_vals) {
        if (blockIdx.x >= n_main) {
            const uint32_t q = (blockIdx.x - n_main) * blockDim.x + threadIdx.x;
            if (q < n_hub) hub_vals[q] = __ldg(v + hub_cols[q]);
            return;
        }
        __shared__ uint32_t s_count;
        if (threadIdx.x == 0) s_count = 0u;
        __syncthreads();
        uint32_t       c      = 0;
        const uint32_t stride = n_main * blockDim.x;// a multiple of 32: a warp always covers 32 consecutive rows
        const uint32_t n_pad  = (n + 31u) & ~31u;
        constexpr int  U      = 4;// independent 32-row groups per warp and step
        for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_pad; i0 += stride * U) {
            bool p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + u * stride;
                p[u]             = i < n && sel.test(mask[i]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + u * stride;
                if (i >= n_pad) break;// warp-uniform: n_pad and the strides are multiples of 32
                const uint32_t m = __ballot_sync(0xffffffffu, p[u]);
                if (i < n) r[i] = init;
                if ((threadIdx.x & 31u) == 0u) {
                    sel_bits[i >> 5] = m;
                    c += __popc(m);
                }
            }
        }
        if ((threadIdx.x & 31u) == 0u && c) atomicAdd(&s_count, c);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_count) atomicAdd(out + 1, s_count);
            __threadfence();
            if (atomicAdd(out + 2, 1u) == n_main - 1u) {// the last block
                __threadfence();
                out[0] = atomicExch(out + 1, 0u);
                out[2] = 0u;
            }
        }
    }

    // select = ALWAYS (no mask pass): r = init and the hub pack in one launch (the PageRank step of spla::pr, src/algorithm.cpp:312)
    template<typename T>
    __global__ void __launch_bounds__(kBlock) fill_pack_kernel(T* __restrict__ r, T init, uint32_t n, uint32_t n_main, const uint32_t* __restrict__ hub_cols,
                                                               uint32_t n_hub, const uint32_t* __restrict__ v, uint32_t* __restrict__ hub_vals) {
        if (blockIdx.x >= n_main) {
            const uint32_t q = (blockIdx.x - n_main) * blockDim.x + threadIdx.x;
            if (q < n_hub) hub_vals[q] = __ldg(v + hub_cols[q]);
            return;
        }
        const uint32_t stride = n_main * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) r[i] = init;
    }

    template<typename T, typename S, bool MASKED, int MODE, int WARPS>
    static int launch_wtile(S sr, Select sel, const TileJob& job, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        auto           kern = mxv_wtile_kernel<T, S, MASKED, MODE, WARPS>;
        const uint32_t smem = (uint32_t) WARPS * kSliceWords * 4u + (MODE != MODE_PLAIN ? ((job.n_smem + 3u) & ~3u) * 4u : 0u);
        static uint64_t attr_done = 0;// per instantiation, one bit per device: a function attribute belongs to the device it was set on
        const int       dev_bit   = current_device() & 63;
        if (!((attr_done >> dev_bit) & 1u)) {
            SPLACU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemLimit));
            attr_done |= (uint64_t) 1 << dev_bit;
        }
        const uint32_t want = (job.n_tiles + WARPS - 1) / WARPS;
        const int      grid = (int) (want < (uint32_t) sm_count() ? want : (uint32_t) sm_count());
        kern<<<grid, WARPS * 32, smem, s>>>(sr, sel, job.Ap, reinterpret_cast<const uint32_t*>(job.Aj), reinterpret_cast<const T*>(job.Ax), v, mask, r, init,
                                            job.nnz, job.n_tiles, job.tile_rows, reinterpret_cast<T*>(job.carry), job.vec_ok, job.hub_vals, job.n_smem,
                                            job.accum, job.gate, job.gate_max);
        SPLACU_LAUNCH_CHECK();
        mxv_fixup_kernel<T, S><<<(job.n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, sel, job.Ap, mask, r, init, job.n_tiles, job.tile_rows,
                                                                                     reinterpret_cast<const T*>(job.carry), job.accum, job.gate, job.gate_max);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    template<typename T, typename S, int MODE, int WARPS>
    static int launch_job(S sr, Select sel, const TileJob& job, const T* v, const T* mask, T* r, T init, cudaStream_t s) {
        return sel.reads_mask ? launch_wtile<T, S, true, MODE, WARPS>(sr, sel, job, v, mask, r, init, s)
                              : launch_wtile<T, S, false, MODE, WARPS>(sr, sel, job, v, mask, r, init, s);
    }

    // keep v resident in L2 while the CSR arrays stream through it (access-policy window on the launching stream)
    static int set_persisting_window(const void* base, size_t bytes, cudaStream_t s) {
        // (opt-in, off by default: measured no gain) -- nothing is cached across calls: the limits belong to the current device
        int dev = 0, v1 = 0, v2 = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v1, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        cudaDeviceGetAttribute(&v2, cudaDevAttrMaxPersistingL2CacheSize, dev);
        const size_t max_win = (size_t) v1, max_persist = (size_t) v2;
        if (!max_win || !max_persist) return 0;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist);
        cudaGetLastError();
        if (bytes > max_win) bytes = max_win;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr  = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio  = bytes <= max_persist ? 1.0f : (float) max_persist / (float) bytes;
        attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
        SPLACU_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
        return 0;
    }

    template<typename T, typename S>
    // parts: 4 = prologue (r = init / the mask pass), 1 = hub classes, 2 = everything that reads v + the fix-ups; 7 = the whole product
    static int launch_tiles(S sr, Select sel, const Csr* M, const T* v, const T* mask, T* r, T init, cudaStream_t s, int parts = 7,
                            const void* d_hub_vals = nullptr) {
        int rc;
        const bool classes = M->n_phases && M->phase[0].seg;
        if (parts != 7 && !classes) {// no column classes: nothing can start before v is complete, the last part is the whole product
            if (!(parts & 2)) return 0;
            parts = 7;
        }
        if (get_option(OPT_MXV_L2_PERSIST) && (size_t) M->nnz * 8 > (size_t) 64 << 20)
            if ((rc = set_persisting_window(v, (size_t) M->n_cols * 4, s))) return rc;
        const bool seg_classes = M->n_phases && M->phase[0].seg;
        const bool fuse_pack   = seg_classes && M->n_hub && parts == 7 && !d_hub_vals && sel.reads_mask && M->sel_count;// packed by the mask pass below
        const bool fuse_fill   = seg_classes && M->n_hub && parts == 7 && !d_hub_vals && !(sel.reads_mask && M->sel_count);// packed by the fill of r
        if (M->n_hub && (parts & 1) && d_hub_vals) {// the caller brings the hub values (gathered from their owners)
            SPLACU_CUDA(cudaMemcpyAsync(M->hub_vals, d_hub_vals, (size_t) M->n_hub * 4, cudaMemcpyDeviceToDevice, s));
        } else if (M->n_hub && (parts & 1) && !fuse_pack && !fuse_fill) {
            SPLACU_PROFILE("splacu/mxv/hub_pack", s);
            mxv_hub_pack_kernel<<<(M->n_hub + kBlock - 1) / kBlock, kBlock, 0, s>>>(M->hub_cols, M->n_hub, reinterpret_cast<const uint32_t*>(v), M->hub_vals);
            SPLACU_LAUNCH_CHECK();
        }
        if (M->n_phases && M->phase[0].seg) {
            // dense masks (and ALWAYS): the column-class passes. Sparse masks: the CSR kernel below tests the mask before it
            // touches Aj / Ax / v and wins once more than half of the rows are unselected. The device decides.
            const uint32_t* gate     = nullptr;
            uint32_t        gate_min = 0;
            if (sel.reads_mask && M->sel_count) {
                gate     = M->sel_count;
                gate_min = (uint32_t) ((uint64_t) M->n_rows * (uint64_t) get_option(OPT_MXV_SEG_MIN_DENSITY) / 100u);
            }
            if (gate && (parts & 4)) {
                SPLACU_PROFILE("splacu/mxv/mask_count_fill", s);
                const uint32_t n_main = (uint32_t) grid_for(M->n_rows, kBlock, 8);
                const uint32_t n_pack = fuse_pack ? (M->n_hub + kBlock - 1) / kBlock : 0u;
                mask_count_fill_kernel<T><<<n_main + n_pack, kBlock, 0, s>>>(sel, mask, M->n_rows, M->sel_count, M->sel_bits, r, init, n_main, M->hub_cols,
                                                                           fuse_pack ? M->n_hub : 0u, reinterpret_cast<const uint32_t*>(v), M->hub_vals);
                SPLACU_LAUNCH_CHECK();
            }
            // The mask-first CSR pass of the same call (it returns at once unless the mask is sparse) runs on the handle's SIDE stream,
            // forked here and joined at the end: for a dense mask its two idle launches (~12 us in a row on the main stream) disappear
            // behind the class passes; for a sparse mask it is the class passes that return at once.
            cudaStream_t side = nullptr;
            if (gate && (parts & 2)) {
                if (!M->side) {
                    SPLACU_CUDA(cudaStreamCreateWithFlags(&M->side, cudaStreamNonBlocking));
                    SPLACU_CUDA(cudaEventCreateWithFlags(&M->ev_fork, cudaEventDisableTiming));
                    SPLACU_CUDA(cudaEventCreateWithFlags(&M->ev_join, cudaEventDisableTiming));
                }
                side = M->side;
                SPLACU_CUDA(cudaEventRecord(M->ev_fork, s));
                SPLACU_CUDA(cudaStreamWaitEvent(side, M->ev_fork, 0));
                {
                    SPLACU_PROFILE("splacu/mxv/csr_pass_gated", side);
                    const TileJob job = {M->Ap, M->Aj, M->Ax, M->nnz, M->n_tiles, M->tile_rows, M->carry, (int) M->vec_ok, nullptr, 0u, 0, gate, gate_min};
                    rc = launch_job<T, S, MODE_PLAIN, kWarps>(sr, sel, job, v, mask, r, init, side);
                }
                SPLACU_CUDA(cudaEventRecord(M->ev_join, side));
                if (rc) {
                    cudaStreamWaitEvent(s, M->ev_join, 0);
                    return rc;
                }
            }
            if (fuse_fill) {
                SPLACU_PROFILE("splacu/mxv/fill_pack", s);
                const uint32_t n_main = (uint32_t) grid_for(M->n_rows, kBlock, 8);
                fill_pack_kernel<T><<<n_main + (M->n_hub + kBlock - 1) / kBlock, kBlock, 0, s>>>(r, init, M->n_rows, n_main, M->hub_cols, M->n_hub,
                                                                                               reinterpret_cast<const uint32_t*>(v), M->hub_vals);
                SPLACU_LAUNCH_CHECK();
                parts = 3;// the prologue is done
            }
            rc = seg_mxv(M, TypeCode<T>::value, sr.mult_op(), sr.add_op(), sel, v, mask, r, to_bits(init), gate, gate_min, s, parts);
            if (side) SPLACU_CUDA(cudaStreamWaitEvent(s, M->ev_join, 0));
            return rc;
        }
        if (M->n_phases) {
            // column classes, hottest first: the hub classes gather from shared memory only, the tail class from v.
            // The first non-empty class writes every row, the later ones accumulate onto it in this fixed order.
            int       accum = 0;
            const int only  = (int) get_option(OPT_MXV_PHASE_ONLY);
            for (int p = 0; p < M->n_phases; ++p) {
                const CsrPhase& ph = M->phase[p];
                if (ph.nnz == 0 || (only && only != p + 1)) continue;
                if (only) accum = p > 0;
                const TileJob job = {ph.Ap, ph.Aj, ph.Ax, ph.nnz, ph.n_tiles, ph.tile_rows, ph.carry, 1, M->hub_vals + ph.slot_base, ph.n_slots, accum, nullptr, 0u};
                rc = ph.idx16 ? launch_job<T, S, MODE_SMEM16, kWarpsHub>(sr, sel, job, v, mask, r, init, s)
                              : launch_job<T, S, MODE_PLAIN, kWarps>(sr, sel, job, v, mask, r, init, s);
                if (rc) return rc;
                accum = 1;
            }
            return 0;
        }
        const TileJob job = {M->Ap, M->n_hub ? M->Aj_hub : M->Aj, M->Ax, M->nnz, M->n_tiles, M->tile_rows, M->carry, (int) M->vec_ok, M->hub_vals, M->n_hub_smem, 0, nullptr, 0u};
        return M->n_hub ? launch_job<T, S, MODE_HUB, kWarps>(sr, sel, job, v, mask, r, init, s)
                        : launch_job<T, S, MODE_PLAIN, kWarps>(sr, sel, job, v, mask, r, init, s);
    }

}// namespace splacu

using namespace splacu;

extern "C" int splacu_csr_hub_cols(splacu_csr handle, uint32_t* n_hub, const uint32_t** d_cols) {
    SPLACU_REQUIRE(handle, "null matrix handle");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    const bool classes = M->n_phases && M->phase[0].seg;
    if (n_hub) *n_hub = classes ? M->n_hub : 0u;
    if (d_cols) *d_cols = classes ? M->hub_cols : nullptr;
    return SPLACU_OK;
}

extern "C" int splacu_mxv_masked_part(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select, const void* d_v, const void* d_hub_vals,
                                      const void* d_mask, void* d_r, uint32_t init_bits, int part, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE((part & 2) ? "splacu/mxv_masked_part2" : ((part & 1) ? "splacu/mxv_masked_part1" : "splacu/mxv_masked_part0"), resolve_stream(stream));
    SPLACU_REQUIRE(handle, "null matrix handle");
    SPLACU_REQUIRE(part == SPLACU_PART_PROLOGUE || part == SPLACU_PART_HUB || part == (SPLACU_PART_PROLOGUE | SPLACU_PART_HUB) || part == SPLACU_PART_REST,
                   "part must be PROLOGUE, HUB, PROLOGUE | HUB or REST");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    SPLACU_REQUIRE(op_valid_for(dtype, op_mult), "op_mult not defined for dtype");
    SPLACU_REQUIRE(op_valid_for(dtype, op_add), "op_add not defined for dtype");
    SPLACU_REQUIRE(is_assoc_commutative(op_add), "the two-part product needs an associative + commutative op_add");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (M->n_rows == 0) return SPLACU_OK;
    const Select sel = make_select(op_select);
    SPLACU_REQUIRE(d_r, "null result pointer");
    SPLACU_REQUIRE(d_mask || !sel.reads_mask, "null mask pointer");
    SPLACU_REQUIRE((part & 2) ? (d_v || M->nnz == 0) : (!(part & 1) || d_hub_vals || d_v || M->n_hub == 0), "null vector pointer");
    cudaStream_t s = resolve_stream(stream);
    if (M->nnz == 0 || (!sel.reads_mask && sel.classes == 0u)) return (part & 2) ? splacu_fill(d_r, init_bits, M->n_rows, stream) : SPLACU_OK;
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        return dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
            using S = decltype(sr);
            return launch_tiles<T, S>(sr, sel, M, static_cast<const T*>(d_v), static_cast<const T*>(d_mask), static_cast<T*>(d_r), from_bits<T>(init_bits), s,
                                      part, d_hub_vals);
        });
    });
}

extern "C" int splacu_mxv_masked(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select,
                                 const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits, int early_exit, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/mxv_masked", resolve_stream(stream));
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
