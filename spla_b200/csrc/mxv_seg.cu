// mxv_seg.cu -- the column classes of the pull product in the SEGMENTED-TILE format (see mxv_pull.cu for the classes).
//
// Semantics: reference src/cpu/cpu_mxv.hpp:88-103 for associative + commutative op_add without early exit:
//            r[i] = select(mask[i]) ? add(init, sum of the products of row i) : init.
//
// Why a second format: in the CSR tile kernel a class that holds 2 .. 40 % of the entries still walks all n_rows row
// extents per pass and parks / re-reads every product in shared memory (~700 instructions per 512-entry tile: the hub
// classes were issue bound, ncu). Here the entries of a class are stored per 512-entry tile in LANE-BLOCKED order: a
// coalesced 128-bit load hands every lane 16 CONSECUTIVE entries of the row order, so the products stay in registers.
// Row structure is not Ap but, per tile,
//   flags    16 bits per lane: entry i of the lane is the last entry of its row
//   seg_row  the row of every flagged entry, in order (= the non-empty rows of the class; empty rows cost nothing)
//   seg_base number of flags before the tile
//   chain    does the tile start inside a row / how many tiles back that row began
// A lane folds its 16 products serially, cutting at the flags; a warp-level segmented scan (fixed order: deterministic)
// joins the pieces of rows that span lanes; the sums are handed through a 2 KB shared-memory slice to the lanes that own
// the segments, which add them onto r (pre-filled with init): coalesced seg_row reads, near-coalesced r updates.
// Rows that span tiles leave head / tail partials; mxv_seg_fixup_kernel chains them left to right.
#include "common.cuh"
#include "ops.cuh"
#include "profile.cuh"

#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace splacu {

    namespace {
        constexpr int      kBlock    = 256;
        constexpr int      kSegWarps = 20;// warps per persistent CTA of a hub class (shared-memory gathers)
#ifndef SPLACU_SEG_TAIL_WARPS
#define SPLACU_SEG_TAIL_WARPS 20
#endif
        constexpr int      kSegTailWarps = SPLACU_SEG_TAIL_WARPS;// ... of the tail class: 102 registers, no spills with 16 gathers + a tile in flight
        constexpr uint32_t kSmemMax  = 227u * 1024u;

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
        __device__ __forceinline__ uint32_t ld_gather(const uint32_t* p) {
            uint32_t r;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
            return r;
        }

        // ---- build ---------------------------------------------------------------------------------
        __global__ void __launch_bounds__(kBlock) seg_flags_kernel(const uint32_t* __restrict__ Ap, uint32_t n_rows, uint32_t* __restrict__ flags) {
            const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
            if (row >= n_rows) return;
            const uint32_t a = Ap[row], b = Ap[row + 1];
            if (b == a) return;
            const uint32_t e = b - 1u, tile = e >> 9, lane = (e & 511u) >> 4, i = e & 15u;
            atomicOr(&flags[tile * 16u + (lane >> 1)], 1u << (i + 16u * (lane & 1u)));
        }
        __global__ void __launch_bounds__(kBlock) seg_tile_count_kernel(const uint32_t* __restrict__ flags, uint32_t n_tiles, uint32_t* __restrict__ count) {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t > n_tiles) return;
            uint32_t c = 0;
            if (t < n_tiles)
                for (int w = 0; w < 16; ++w) c += __popc(flags[t * 16u + w]);
            count[t] = c;
        }
        __global__ void __launch_bounds__(kBlock) seg_chain_kernel(const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ flags,
                                                                   const uint32_t* __restrict__ seg_base, const uint32_t* __restrict__ seg_row, uint32_t n_tiles,
                                                                   uint32_t* __restrict__ chain, uint32_t* __restrict__ chain_row) {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= n_tiles) return;
            uint32_t w = 0, row = 0;
            if (t > 0) {
                const bool prev_ends = (flags[(t - 1u) * 16u + 15u] >> 31) & 1u;// entry 511 of tile t - 1 = (lane 31, i 15)
                if (!prev_ends) {
                    w = 0x80000000u;
                    if (seg_base[t + 1] > seg_base[t]) {// the row that reaches into the tile ends here
                        row = seg_row[seg_base[t]];
                        w |= t - (Ap[row] >> 9);
                    }
                }
            }
            chain[t]     = w;
            chain_row[t] = row;
        }
        // Build-time reordering of a hub class against shared-memory bank conflicts. The kernel's 16 table gathers of a tile are 16 warp
        // instructions: step i reads the i-th entry of every lane, 32 random slots of a 45 K-word table -> ~3.4 wavefronts per
        // instruction instead of 1 (the largest share of the LSU pipe's work in a hub pass). The ORDER of a row's entries is free (the add
        // is associative + commutative, and the fold order is fixed either way), so inside every lane the entries of one row run are
        // permuted such that at each step the lanes hit banks that are as distinct as possible: steps in order, lanes in order, each lane
        // takes from the rest of its current run the entry whose bank is least loaded at this step. One warp per tile, lane-serial
        // (512 decisions per tile, once per matrix). Flags, segment lists and chains are untouched (run lengths do not change).
        __global__ void __launch_bounds__(kBlock) seg_bank_order_kernel(uint32_t* __restrict__ idx, uint32_t* __restrict__ vals, const uint32_t* __restrict__ flags,
                                                                        uint32_t n_tiles) {
            __shared__ uint16_t s_slot[kBlock / 32][32][17];// [warp][lane][entry] (+1: no bank conflicts on the lane-serial walk)
            __shared__ uint32_t s_val[kBlock / 32][32][17];
            __shared__ uint8_t  s_load[kBlock / 32][32];
            const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
            const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
            for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += n_warps) {
                uint4*         i4 = reinterpret_cast<uint4*>(idx) + (size_t) t * 64;
                uint4*         v4 = reinterpret_cast<uint4*>(vals) + (size_t) t * 128;
                const uint32_t fl = (flags[t * 16u + (lane >> 1)] >> ((lane & 1u) * 16u)) & 0xffffu;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4    x    = i4[h * 32 + lane];
                    const uint32_t q[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        s_slot[w][lane][8 * h + 2 * k]     = (uint16_t) (q[k] & 0xffffu);
                        s_slot[w][lane][8 * h + 2 * k + 1] = (uint16_t) (q[k] >> 16);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint4 x = v4[g * 32 + lane];
                    s_val[w][lane][4 * g + 0] = x.x, s_val[w][lane][4 * g + 1] = x.y, s_val[w][lane][4 * g + 2] = x.z, s_val[w][lane][4 * g + 3] = x.w;
                }
                __syncwarp();
                for (int i = 0; i < 16; ++i) {
                    s_load[w][lane] = 0;
                    __syncwarp();
                    // end of the run that holds position i: the first flagged position >= i (15 when the lane's tail stays open)
                    const uint32_t rest = fl >> i;
                    const int      end  = rest ? i + __ffs(rest) - 1 : 15;
                    for (uint32_t L = 0; L < 32; ++L) {
                        if (lane == L) {
                            int      best = i;
                            uint32_t bl   = s_load[w][s_slot[w][lane][i] & 31u];
                            for (int c = i + 1; c <= end && bl; ++c) {
                                const uint32_t l = s_load[w][s_slot[w][lane][c] & 31u];
                                if (l < bl) bl = l, best = c;
                            }
                            if (best != i) {
                                const uint16_t a = s_slot[w][lane][i];
                                const uint32_t b = s_val[w][lane][i];
                                s_slot[w][lane][i] = s_slot[w][lane][best], s_val[w][lane][i] = s_val[w][lane][best];
                                s_slot[w][lane][best] = a, s_val[w][lane][best] = b;
                            }
                            s_load[w][s_slot[w][lane][i] & 31u] = (uint8_t) (bl + 1u);
                        }
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t q[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) q[k] = (uint32_t) s_slot[w][lane][8 * h + 2 * k] | ((uint32_t) s_slot[w][lane][8 * h + 2 * k + 1] << 16);
                    i4[h * 32 + lane] = make_uint4(q[0], q[1], q[2], q[3]);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    v4[g * 32 + lane] = make_uint4(s_val[w][lane][4 * g + 0], s_val[w][lane][4 * g + 1], s_val[w][lane][4 * g + 2], s_val[w][lane][4 * g + 3]);
                __syncwarp();
            }
        }
    }// namespace

    // flags / seg_base / segment list of a lane-blocked tile array from the extents of its units (rows of a column class, columns
    // of a row class): unit u owns entries [d_ext[u], d_ext[u + 1]), d_count[u] = its length. *seg_unit gets n_units + 32 words.
    int seg_structure(const uint32_t* d_ext, const uint32_t* d_count, uint32_t n_units, uint32_t nt, uint32_t** flags, uint32_t** seg_base,
                      uint32_t** seg_unit, uint32_t* n_segs, cudaStream_t s) {
        void*     tmp   = nullptr;
        uint32_t* count = nullptr;
        uint32_t* d_num = nullptr;
        int       rc    = 0;
#define SEG_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);          \
            goto done;                                                        \
        }                                                                     \
    } while (0)
        {
            size_t                              b1 = 0, b2 = 0;
            thrust::counting_iterator<uint32_t> units(0u);
            SEG_CUDA(cudaMalloc(flags, (size_t) nt * 16 * 4));
            SEG_CUDA(cudaMemsetAsync(*flags, 0, (size_t) nt * 16 * 4, s));
            SEG_CUDA(cudaMalloc(seg_base, ((size_t) nt + 1) * 4));
            SEG_CUDA(cudaMalloc(&count, ((size_t) nt + 1) * 4));
            SEG_CUDA(cudaMalloc(&d_num, 4));
            // every non-empty unit is one segment; + 32: the kernels read the segment list of a tile 32 at a time
            SEG_CUDA(cudaMalloc(seg_unit, ((size_t) n_units + 32) * 4));
            SEG_CUDA(cudaMemsetAsync(*seg_unit, 0, ((size_t) n_units + 32) * 4, s));
            SEG_CUDA(cub::DeviceSelect::Flagged(nullptr, b1, units, d_count, *seg_unit, d_num, (int) n_units, s));
            SEG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, count, *seg_base, (int) nt + 1, s));
            SEG_CUDA(cudaMalloc(&tmp, b1 > b2 ? b1 : b2));
            SEG_CUDA(cub::DeviceSelect::Flagged(tmp, b1, units, d_count, *seg_unit, d_num, (int) n_units, s));
            seg_flags_kernel<<<(n_units + kBlock - 1) / kBlock, kBlock, 0, s>>>(d_ext, n_units, *flags);
            seg_tile_count_kernel<<<(nt + 1 + kBlock - 1) / kBlock, kBlock, 0, s>>>(*flags, nt, count);
            SEG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, b2, count, *seg_base, (int) nt + 1, s));
            count_launch(4);
            SEG_CUDA(cudaMemcpyAsync(n_segs, d_num, 4, cudaMemcpyDeviceToHost, s));
            SEG_CUDA(cudaStreamSynchronize(s));
            SEG_CUDA(cudaGetLastError());
        }
    done:
        cudaFree(tmp);
        cudaFree(count);
        cudaFree(d_num);
        return rc;
    }

    int seg_build(const Csr* M, CsrPhase& ph, const uint32_t* d_row_count, cudaStream_t s) {
        int rc = seg_structure(ph.Ap, d_row_count, M->n_rows, ph.n_tiles, &ph.flags, &ph.seg_base, &ph.seg_row, &ph.n_segs, s);
        if (rc) return rc;
        {
            const uint32_t nt = ph.n_tiles;
            SEG_CUDA(cudaMalloc(&ph.chain, (size_t) nt * 4));
            SEG_CUDA(cudaMalloc(&ph.chain_row, (size_t) nt * 4));
            SEG_CUDA(cudaMalloc(&ph.head, (size_t) nt * 4));
            SEG_CUDA(cudaMalloc(&ph.tail, (size_t) nt * 4));
            seg_chain_kernel<<<(nt + kBlock - 1) / kBlock, kBlock, 0, s>>>(ph.Ap, ph.flags, ph.seg_base, ph.seg_row, nt, ph.chain, ph.chain_row);
            count_launch(1);
            if (ph.idx16 && get_option(OPT_MXV_BANK_ORDER)) {
                seg_bank_order_kernel<<<grid_for((size_t) nt * 32, kBlock, 8), kBlock, 0, s>>>(reinterpret_cast<uint32_t*>(ph.Aj), ph.Ax, ph.flags, nt);
                count_launch(1);
            }
            SEG_CUDA(cudaStreamSynchronize(s));
            SEG_CUDA(cudaGetLastError());
            ph.seg = true;
            cudaFree(ph.Ap);// the row extents were only needed to derive the segments
            ph.Ap = nullptr;
        }
    done:
#undef SEG_CUDA
        return rc;
    }

    // ---- the list of rows that span tiles (all classes), for the two-launch fix-up ------------------------
    namespace {
        __global__ void __launch_bounds__(kBlock) fix_collect_kernel(const uint32_t* __restrict__ chain, const uint32_t* __restrict__ chain_row, uint32_t n_tiles,
                                                                     uint32_t cls, uint32_t* __restrict__ counter, uint64_t* __restrict__ keys,
                                                                     uint32_t* __restrict__ tiles) {
            const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
            if (t >= n_tiles || !(chain[t] & 0x7fffffffu)) return;
            const uint32_t pos = atomicAdd(counter, 1u);
            keys[pos]  = ((uint64_t) chain_row[t] << 8) | cls;// (row, class) is unique: a row ends at most once per class
            tiles[pos] = t;
        }
    }// namespace
    int seg_build_fixlist(Csr* M, cudaStream_t s) {
        uint64_t  total = 0;
        uint32_t *counter = nullptr, *tiles = nullptr;
        uint64_t* keys  = nullptr;
        void*     tmp   = nullptr;
        int       rc    = 0;
        for (int p = 0; p < M->n_phases; ++p)
            if (M->phase[p].seg) total += M->phase[p].n_tiles;
        if (total == 0) return 0;
#define FX_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__);          \
            goto done;                                                        \
        }                                                                     \
    } while (0)
        {
            uint32_t n_fix = 0;
            FX_CUDA(cudaMalloc(&counter, 4));
            FX_CUDA(cudaMemsetAsync(counter, 0, 4, s));
            FX_CUDA(cudaMalloc(&keys, total * 8));
            FX_CUDA(cudaMalloc(&tiles, total * 4));
            for (int p = 0; p < M->n_phases; ++p) {
                const CsrPhase& ph = M->phase[p];
                if (!ph.seg || ph.n_tiles == 0) continue;
                fix_collect_kernel<<<(ph.n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(ph.chain, ph.chain_row, ph.n_tiles, (uint32_t) p, counter, keys, tiles);
                count_launch(1);
            }
            FX_CUDA(cudaMemcpyAsync(&n_fix, counter, 4, cudaMemcpyDeviceToHost, s));
            FX_CUDA(cudaStreamSynchronize(s));
            M->n_fix = n_fix;
            FX_CUDA(cudaMalloc(&M->fix_key, ((size_t) n_fix + 1) * 8));// allocated even when empty: marks the list as built
            FX_CUDA(cudaMalloc(&M->fix_tile, ((size_t) n_fix + 1) * 4));
            if (n_fix) {
                size_t tmp_bytes = 0;
                FX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, M->fix_key, tiles, M->fix_tile, (int) n_fix, 0, 40, s));
                FX_CUDA(cudaMalloc(&tmp, tmp_bytes));
                FX_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, M->fix_key, tiles, M->fix_tile, (int) n_fix, 0, 40, s));
                FX_CUDA(cudaStreamSynchronize(s));
            }
        }
    done:
#undef FX_CUDA
        cudaFree(counter);
        cudaFree(keys);
        cudaFree(tiles);
        cudaFree(tmp);
        if (rc) {
            cudaFree(M->fix_key);
            cudaFree(M->fix_tile);
            M->fix_key = nullptr, M->fix_tile = nullptr, M->n_fix = 0;
        }
        return rc;
    }

    // ---- the kernel ----------------------------------------------------------------------------------
    // RED: the add is PLUS and r[row] += sum is issued as a reduction at the L2 (red.global.add, no value returned) instead of a
    // load + add + store: one request and no load latency per segment. The result is the same two-operand rounding, and the order of
    // the adds onto a row stays fixed (one add per row and launch; launches are ordered), so it remains deterministic. Caveat
    // (PTX ISA, atom / red .f32 on global memory): subnormal operands and results are flushed to zero.
    template<typename T, typename S, bool MASKED, bool IDX16, int WARPS, bool RED>
    __global__ void __launch_bounds__(WARPS * 32, 1)
            mxv_seg_kernel(S sr, Select sel, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ flags,
                           const uint32_t* __restrict__ seg_base, const uint32_t* __restrict__ seg_row, const uint32_t* __restrict__ chain,
                           uint32_t* __restrict__ head, uint32_t* __restrict__ tail, const T* __restrict__ v, const uint32_t* __restrict__ sel_bits, T* r,
                           uint32_t n_tiles, const uint32_t* __restrict__ hub_vals, uint32_t n_slots, const uint32_t* __restrict__ gate,
                           uint32_t gate_min) {
        extern __shared__ __align__(16) uint32_t smem[];
        if (MASKED && gate && *gate < gate_min) return;// sparse mask: the CSR kernel (mask tested before any gather) runs instead
        constexpr int  NI   = IDX16 ? 2 : 4;// 128-bit index loads per lane and tile
        const uint32_t tid  = threadIdx.x;
        const uint32_t lane = tid & 31u;
        const uint32_t warp = tid >> 5;
        T*             s_out = reinterpret_cast<T*>(smem) + warp * 512;// segment sums of this warp's tile, in segment order
        const T*       s_hub = reinterpret_cast<const T*>(smem) + WARPS * 512;
        if (IDX16) {
            uint4*       dst = reinterpret_cast<uint4*>(smem + WARPS * 512);
            const uint4* src = reinterpret_cast<const uint4*>(hub_vals);
            for (uint32_t i = tid; i < (n_slots + 3u) / 4u; i += WARPS * 32) dst[i] = __ldg(src + i);
            __syncthreads();
        }
        const uint64_t pol     = policy_evict_first();
        const uint32_t n_warps = gridDim.x * WARPS;
        const uint32_t first   = blockIdx.x * WARPS + warp;
        const uint4*   idx4    = reinterpret_cast<const uint4*>(idx);
        const uint4*   val4    = reinterpret_cast<const uint4*>(vals);

        // ---- software pipeline: slices + flags + first 32 segment rows one tile ahead, seg_base / chain two tiles ahead ----
        uint4    xv[4], xi[NI];
        uint32_t fw = 0, sb0 = 0, sb1 = 0, ch = 0, srow = 0;
        uint32_t q0 = 0, q1 = 0, qc = 0;
        auto     load_meta = [&](uint32_t t) {
            if (t < n_tiles) {
                q0 = __ldg(seg_base + t);
                q1 = __ldg(seg_base + t + 1);
                qc = __ldg(chain + t);
            }
        };
        auto prefetch = [&](uint32_t t) {
            if (t >= n_tiles) return;
            sb0 = q0, sb1 = q1, ch = qc;
            load_meta(t + n_warps);
#pragma unroll
            for (int h = 0; h < NI; ++h) xi[h] = ld_stream_u4(idx4 + (size_t) t * (NI * 32) + h * 32 + lane, pol);
#pragma unroll
            for (int q = 0; q < 4; ++q) xv[q] = ld_stream_u4(val4 + (size_t) t * 128 + q * 32 + lane, pol);
            fw   = __ldg(flags + t * 16u + (lane >> 1));
            srow = __ldg(seg_row + sb0 + lane);// padded by 32 rows
        };
        load_meta(first);
        prefetch(first);
        // Programmatic dependent launch (option mxv_pdl): this CTA may have been scheduled while the last CTAs of the previous class pass
        // were still running -- everything above read only the matrix and the packed hub values. r, head / tail and the selection bitmap
        // are touched below: wait until the previous grid has completed and its writes are visible (a no-op in a plain launch).
        asm volatile("griddepcontrol.wait;" ::: "memory");

        for (uint32_t tile = first; tile < n_tiles; tile += n_warps) {
            const uint32_t base = sb0, nfl = sb1 - sb0;
            const bool     cont = (ch >> 31) != 0u;
            const uint32_t row0 = srow;
            const uint32_t fl   = (fw >> ((lane & 1u) * 16u)) & 0xffffu;
            // (prefetch.global.L2 hints for the seg_row / r lines of the later hand-over rounds were measured: no gain)
            // r / mask of the first 32 segments: requested now, used by the hand-over at the end of the tile
            T    old0  = sr.identity();
            bool take0 = false;
            if (lane < nfl) {
                take0 = MASKED ? ((sel_bits[row0 >> 5] >> (row0 & 31u)) & 1u) != 0u : true;
                if (!RED) old0 = r[row0];
            }

            // ---- products of the lane's 16 consecutive entries ----
            T p[16];
            if (IDX16) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w[4] = {xi[h].x, xi[h].y, xi[h].z, xi[h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        p[8 * h + 2 * k]     = s_hub[w[k] & 0xffffu];
                        p[8 * h + 2 * k + 1] = s_hub[w[k] >> 16];
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    p[4 * q + 0] = from_bits<T>(ld_gather(reinterpret_cast<const uint32_t*>(v) + xi[q].x));
                    p[4 * q + 1] = from_bits<T>(ld_gather(reinterpret_cast<const uint32_t*>(v) + xi[q].y));
                    p[4 * q + 2] = from_bits<T>(ld_gather(reinterpret_cast<const uint32_t*>(v) + xi[q].z));
                    p[4 * q + 3] = from_bits<T>(ld_gather(reinterpret_cast<const uint32_t*>(v) + xi[q].w));
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                p[4 * q + 0] = sr.mult(from_bits<T>(xv[q].x), p[4 * q + 0]);
                p[4 * q + 1] = sr.mult(from_bits<T>(xv[q].y), p[4 * q + 1]);
                p[4 * q + 2] = sr.mult(from_bits<T>(xv[q].z), p[4 * q + 2]);
                p[4 * q + 3] = sr.mult(from_bits<T>(xv[q].w), p[4 * q + 3]);
            }
            prefetch(tile + n_warps);// the slice registers are free again

            // ---- position of the lane's first segment among the segments of the tile ----
            const uint32_t cnt  = __popc(fl);
            uint32_t       incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int) lane >= d) incl += t;
            }
            uint32_t k = incl - cnt;

            // ---- pass 1: the lane's open tail (entries after its last flag); warp segmented scan -> carry-in ----
            T open = sr.identity();
#pragma unroll
            for (int i = 0; i < 16; ++i) open = ((fl >> i) & 1u) ? sr.identity() : sr.add(open, p[i]);
            T    sv = open;
            bool sf = fl != 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T    vv = __shfl_up_sync(0xffffffffu, sv, d);
                const bool ff = __shfl_up_sync(0xffffffffu, (int) sf, d) != 0;
                if ((int) lane >= d) {
                    if (!sf) sv = sr.add(vv, sv);
                    sf = sf || ff;
                }
            }
            T acc = __shfl_up_sync(0xffffffffu, sv, 1);
            if (lane == 0) acc = sr.identity();
            // ---- pass 2: segment sums -> shared memory in segment order ----
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acc = sr.add(acc, p[i]);
                if ((fl >> i) & 1u) {
                    s_out[k] = acc;
                    ++k;
                    acc = sr.identity();
                }
            }
            if (lane == 31) tail[tile] = to_bits(acc);// what follows the tile's last flag (the whole tile when it has none)
            __syncwarp();

            // ---- hand-over: the lane that owns segment o adds its sum onto r. Two-stage pipeline over the rounds of 32 segments:
            //      the rows of round i + 2 and the r / selection values of round i + 1 are requested before round i is written
            //      (needs the registers of 20-warp CTAs; at 24 warps / 80 registers it spilled and lost) ----
            {
                uint32_t rowA = row0, rowB = (nfl > 32u + lane) ? __ldg(seg_row + base + 32u + lane) : 0u;
                bool     takeA = take0;
                T        oldA  = old0;
                for (uint32_t ob = 0; ob < nfl; ob += 32) {
                    const uint32_t o = ob + lane;
                    uint32_t       rowC = 0;
                    bool           takeB = false;
                    T              oldB  = sr.identity();
                    if (o + 64 < nfl) rowC = __ldg(seg_row + base + o + 64);
                    if (o + 32 < nfl) {
                        takeB = MASKED ? ((sel_bits[rowB >> 5] >> (rowB & 31u)) & 1u) != 0u : true;
                        if (!RED) oldB = r[rowB];
                    }
                    if (o < nfl) {
                        const T sum = s_out[o];
                        if (o == 0 && cont) head[tile] = to_bits(sum);// the row began in an earlier tile: the fix-up adds the chain
                        else if (takeA) {
                            if constexpr (RED) atomicAdd(&r[rowA], sum);// result unused: compiles to RED
                            else r[rowA] = sr.add(oldA, sum);
                        }
                    }
                    rowA = rowB, takeA = takeB, oldA = oldB, rowB = rowC;
                }
            }
            __syncwarp();// s_out is reused by the next tile
        }
    }

    // rows that span tiles: r[row] += tail(t0) + tail(t0 + 1) + ... + tail(t - 1) + head(t), left to right; one thread per end
    // tile, the whole warp for chains longer than 4 tiles (hub rows)
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_seg_fixup_kernel(S sr, Select sel, const uint32_t* __restrict__ chain, const uint32_t* __restrict__ chain_row,
                                                                   const uint32_t* __restrict__ head,
                                                                   const uint32_t* __restrict__ tail, const uint32_t* __restrict__ sel_bits, T* r,
                                                                   uint32_t n_tiles,
                                                                   const uint32_t* __restrict__ gate, uint32_t gate_min) {
        if (gate && *gate < gate_min) return;
        const uint32_t t    = blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t       len = 0, row = 0;
        if (t < n_tiles) {
            len = chain[t] & 0x7fffffffu;
            if (len) {
                row = chain_row[t];
                if (sel_bits && !((sel_bits[row >> 5] >> (row & 31u)) & 1u)) len = 0;
            }
        }
        if (len > 0 && len <= 4) {
            T acc = from_bits<T>(tail[t - len]);
            for (uint32_t u = t - len + 1; u < t; ++u) acc = sr.add(acc, from_bits<T>(tail[u]));
            acc    = sr.add(acc, from_bits<T>(head[t]));
            r[row] = sr.add(r[row], acc);
        }
        uint32_t long_mask = __ballot_sync(0xffffffffu, len > 4);
        while (long_mask) {
            const int      src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            const uint32_t t1 = __shfl_sync(0xffffffffu, t, src), L = __shfl_sync(0xffffffffu, len, src);
            T              acc = sr.identity();
            for (uint32_t u = lane; u < L; u += 32) acc = sr.add(acc, from_bits<T>(tail[t1 - L + u]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if ((int) lane == src) r[row] = sr.add(r[row], sr.add(acc, from_bits<T>(head[t])));
        }
    }

    static thread_local bool g_split_call = false;// splacu_mxv_masked_part: the merged fix-up closes the second part

    // The fix-ups of ALL classes in one cooperative launch: class by class with a grid barrier in between, so that the additions onto a
    // row that spans tiles in several classes keep their fixed order (class passes first, then the chains in class order). One launch
    // and n - 1 grid barriers instead of n dependent launches of ~10 us each between the class passes (in the stream: 60 us of a
    // 1.33 ms step on RMAT-24, and the same 60 us of the 0.31 ms a rank of an 8-GPU run spends in its kernels).
    struct FixArgs {
        const uint32_t* chain;
        const uint32_t* chain_row;
        const uint32_t* head;
        const uint32_t* tail;
        uint32_t        n_tiles;
    };
    struct FixAll {
        FixArgs c[kMaxHubPhases + 1];
        int     n;
    };
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_seg_fixup_all_kernel(S sr, FixAll a, const uint32_t* __restrict__ sel_bits, T* r,
                                                                       const uint32_t* __restrict__ gate, uint32_t gate_min) {
        if (gate && *gate < gate_min) return;// uniform over the grid: nobody reaches a barrier
        cooperative_groups::grid_group grid = cooperative_groups::this_grid();
        const uint32_t                 lane = threadIdx.x & 31u;
        for (int p = 0; p < a.n; ++p) {
            const uint32_t* __restrict__ chain     = a.c[p].chain;
            const uint32_t* __restrict__ chain_row = a.c[p].chain_row;
            const uint32_t* __restrict__ head      = a.c[p].head;
            const uint32_t* __restrict__ tail      = a.c[p].tail;
            const uint32_t n_tiles = a.c[p].n_tiles;
            const uint32_t padded  = (n_tiles + 31u) & ~31u;// whole warps iterate together (the long chains are folded by the warp)
            for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < padded; t += gridDim.x * blockDim.x) {
                uint32_t len = 0, row = 0;
                if (t < n_tiles) {
                    len = chain[t] & 0x7fffffffu;
                    if (len) {
                        row = chain_row[t];
                        if (sel_bits && !((sel_bits[row >> 5] >> (row & 31u)) & 1u)) len = 0;
                    }
                }
                if (len > 0 && len <= 4) {
                    T acc = from_bits<T>(tail[t - len]);
                    for (uint32_t u = t - len + 1; u < t; ++u) acc = sr.add(acc, from_bits<T>(tail[u]));
                    acc    = sr.add(acc, from_bits<T>(head[t]));
                    r[row] = sr.add(r[row], acc);
                }
                uint32_t long_mask = __ballot_sync(0xffffffffu, len > 4);
                while (long_mask) {
                    const int src = __ffs(long_mask) - 1;
                    long_mask &= long_mask - 1;
                    const uint32_t t1 = __shfl_sync(0xffffffffu, t, src), L = __shfl_sync(0xffffffffu, len, src);
                    T              acc = sr.identity();
                    for (uint32_t u = lane; u < L; u += 32) acc = sr.add(acc, from_bits<T>(tail[t1 - L + u]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
                    if ((int) lane == src) r[row] = sr.add(r[row], sr.add(acc, from_bits<T>(head[t])));
                }
            }
            if (p + 1 < a.n) grid.sync();
        }
    }
    template<typename T, typename S>
    static int launch_fixup_all(S sr, const Csr* M, const int* classes, int n, const uint32_t* sel_bits, T* r, const uint32_t* gate, uint32_t gate_min,
                                cudaStream_t s) {
        if (n == 0) return 0;
        FixAll   a;
        uint32_t most = 0;
        a.n = n;
        for (int i = 0; i < n; ++i) {
            const CsrPhase& ph = M->phase[classes[i]];
            a.c[i]             = {ph.chain, ph.chain_row, ph.head, ph.tail, ph.n_tiles};
            if (ph.n_tiles > most) most = ph.n_tiles;
        }
        auto       kern = mxv_seg_fixup_all_kernel<T, S>;
        static int per_sm[64] = {0};// co-resident CTAs per SM, per device
        const int  dev = current_device() & 63;
        if (!per_sm[dev]) {
            int nb = 0;
            SPLACU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kBlock, 0));
            per_sm[dev] = nb > 0 ? nb : 1;
        }
        uint32_t grid = (most + kBlock - 1) / kBlock;
        const uint32_t cap = (uint32_t) (per_sm[dev] * sm_count());
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        SPLACU_PROFILE("splacu/mxv/fixup_all", s);
        void* args[] = {&sr, &a, (void*) &sel_bits, &r, (void*) &gate, &gate_min};
        SPLACU_CUDA(cudaLaunchCooperativeKernel((const void*) kern, dim3(grid), dim3(kBlock), args, 0, s));
        count_launch(1);
        return 0;
    }

    // The same fix-ups in TWO plain launches, no grid barriers (fixup_all: 43 us in the stream, most of it barrier and dependent-load
    // latency of five small passes in a row):
    //   (1) mxv_seg_chain_sums_kernel, over the tiles of ALL classes at once: the sum of every chain, tail(t0) + ... + tail(t - 1) +
    //       head(t) left to right, written back into head[t] (only the thread of tile t reads head[t]);
    //   (2) mxv_seg_fix_rows_kernel, one thread per row that spans tiles in any class: r[row] = add(... add(add(r[row], sum in its
    //       first class), sum in its next class) ...) in class order, from the list the handle keeps of (row, class, end tile) sorted by
    //       row and class (built once, seg_build_fixlist).
    // Same additions in the same order as the ordered cooperative launch: bit-identical results, run-to-run deterministic.
    // Option mxv_fixup_merge = 2 (default); 1 keeps the cooperative launch, 0 one launch per class.
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_seg_chain_sums_kernel(S sr, FixAll a, const uint32_t* __restrict__ sel_bits,
                                                                        const uint32_t* __restrict__ gate, uint32_t gate_min) {
        if (gate && *gate < gate_min) return;
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t       t    = blockIdx.x * blockDim.x + threadIdx.x;// position in the concatenation of the classes, each padded to whole warps
        int            p    = 0;
        for (; p < a.n; ++p) {
            const uint32_t padded = (a.c[p].n_tiles + 31u) & ~31u;
            if (t < padded) break;
            t -= padded;
        }
        if (p == a.n) return;// whole warps leave together (the classes are padded to warps)
        const uint32_t* __restrict__ chain     = a.c[p].chain;
        const uint32_t* __restrict__ chain_row = a.c[p].chain_row;
        uint32_t*                    head      = const_cast<uint32_t*>(a.c[p].head);
        const uint32_t* __restrict__ tail      = a.c[p].tail;
        uint32_t len = 0;
        if (t < a.c[p].n_tiles) {
            len = chain[t] & 0x7fffffffu;
            if (len && sel_bits) {
                const uint32_t row = chain_row[t];
                if (!((sel_bits[row >> 5] >> (row & 31u)) & 1u)) len = 0;
            }
        }
        if (len > 0 && len <= 4) {
            T acc = from_bits<T>(tail[t - len]);
            for (uint32_t u = t - len + 1; u < t; ++u) acc = sr.add(acc, from_bits<T>(tail[u]));
            head[t] = to_bits(sr.add(acc, from_bits<T>(head[t])));
        }
        uint32_t long_mask = __ballot_sync(0xffffffffu, len > 4);
        while (long_mask) {
            const int src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            const uint32_t t1 = __shfl_sync(0xffffffffu, t, src), L = __shfl_sync(0xffffffffu, len, src);
            T              acc = sr.identity();
            for (uint32_t u = lane; u < L; u += 32) acc = sr.add(acc, from_bits<T>(tail[t1 - L + u]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = sr.add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if ((int) lane == src) head[t] = to_bits(sr.add(acc, from_bits<T>(head[t])));
        }
    }
    // a.c[] is indexed by CLASS here (n_tiles == 0: the class did not run in this call)
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) mxv_seg_fix_rows_kernel(S sr, FixAll a, const uint64_t* __restrict__ fix_key, const uint32_t* __restrict__ fix_tile,
                                                                      uint32_t n_fix, const uint32_t* __restrict__ sel_bits, T* r,
                                                                      const uint32_t* __restrict__ gate, uint32_t gate_min) {
        if (gate && *gate < gate_min) return;
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n_fix) return;
        uint64_t       key = fix_key[i];
        const uint32_t row = (uint32_t) (key >> 8);
        if (i > 0 && (uint32_t) (fix_key[i - 1] >> 8) == row) return;// the first entry of a row does the row
        if (sel_bits && !((sel_bits[row >> 5] >> (row & 31u)) & 1u)) return;
        T    acc = r[row];
        bool any = false;
        for (uint32_t j = i;;) {
            const uint32_t p = (uint32_t) key & 0xffu;
            if (a.c[p].n_tiles) {
                acc = sr.add(acc, from_bits<T>(a.c[p].head[fix_tile[j]]));
                any = true;
            }
            if (++j >= n_fix) break;
            key = fix_key[j];
            if ((uint32_t) (key >> 8) != row) break;
        }
        if (any) r[row] = acc;
    }
    template<typename T, typename S>
    static int launch_fixup_rows(S sr, const Csr* M, const int* classes, int n, const uint32_t* sel_bits, T* r, const uint32_t* gate, uint32_t gate_min,
                                 cudaStream_t s) {
        if (n == 0 || M->n_fix == 0) return 0;
        FixAll   a, by_class;
        uint64_t total = 0;
        a.n = n;
        by_class.n = kMaxHubPhases + 1;
        for (int p = 0; p <= kMaxHubPhases; ++p) by_class.c[p] = {nullptr, nullptr, nullptr, nullptr, 0u};
        for (int i = 0; i < n; ++i) {
            const CsrPhase& ph     = M->phase[classes[i]];
            a.c[i]                 = {ph.chain, ph.chain_row, ph.head, ph.tail, ph.n_tiles};
            by_class.c[classes[i]] = a.c[i];
            total += (ph.n_tiles + 31u) & ~31u;
        }
        if (total == 0) return 0;
        SPLACU_PROFILE("splacu/mxv/fixup_rows", s);
        mxv_seg_chain_sums_kernel<T, S><<<(unsigned) ((total + kBlock - 1) / kBlock), kBlock, 0, s>>>(sr, a, sel_bits, gate, gate_min);
        SPLACU_LAUNCH_CHECK();
        mxv_seg_fix_rows_kernel<T, S><<<(M->n_fix + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, by_class, M->fix_key, M->fix_tile, M->n_fix, sel_bits, r, gate,
                                                                                         gate_min);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    template<typename T, typename S, bool MASKED, bool IDX16, bool RED = false>
    static int launch_seg(S sr, Select sel, const Csr* M, const CsrPhase& ph, const T* v, const uint32_t* sel_bits, T* r, const uint32_t* gate,
                          uint32_t gate_min, cudaStream_t s) {
        // (hub classes only: in the tail class, bound by the L1 -> L2 request port, every reduction is one more request while the
        //  near-coalesced load + store of r costs a few requests per 32 rows; in the stream the tail pass takes 472 us either way)
        if constexpr (!RED && S::is_static && IDX16) {
            if (sr.add_op() == SPLACU_PLUS && get_option(OPT_MXV_RED)) return launch_seg<T, S, MASKED, IDX16, true>(sr, sel, M, ph, v, sel_bits, r, gate, gate_min, s);
        }
        constexpr int  kW   = IDX16 ? kSegWarps : kSegTailWarps;
        auto           kern = mxv_seg_kernel<T, S, MASKED, IDX16, kW, RED>;
        const uint32_t smem = kW * 512u * 4u + (IDX16 ? ((ph.n_slots + 3u) & ~3u) * 4u : 0u);
        static uint64_t attr_done = 0;// per instantiation, one bit per device: a function attribute belongs to the device it was set on
        const int       dev_bit   = current_device() & 63;
        if (!((attr_done >> dev_bit) & 1u)) {
            SPLACU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemMax));
            attr_done |= (uint64_t) 1 << dev_bit;
        }
        if (smem > kSmemMax) {
            set_error("mxv: hub class of %u slots does not fit in shared memory", ph.n_slots);
            return SPLACU_E_INVALID;
        }
        const uint32_t want = (ph.n_tiles + kW - 1) / kW;
        const uint32_t cap  = persistent_grid_cap();
        const int      grid = (int) (want < cap ? want : cap);
        // per-launch scopes (splacu_profile_enable): the in-stream time of every class pass -- warm caches, no serialisation, which ncu's
        // per-kernel replay does not show (the tail pass: 402 us under ncu, 472 us in the stream)
        static const char* const kLabels[2][kMaxHubPhases + 1] = {
                {"splacu/mxv/class00", "splacu/mxv/class01", "splacu/mxv/class02", "splacu/mxv/class03", "splacu/mxv/class04", "splacu/mxv/class05", "splacu/mxv/class06", "splacu/mxv/class07", "splacu/mxv/class08", "splacu/mxv/class09", "splacu/mxv/class10", "splacu/mxv/class11", "splacu/mxv/class12", "splacu/mxv/class13", "splacu/mxv/class14", "splacu/mxv/class15", "splacu/mxv/class16"},
                {"splacu/mxv/fixup00", "splacu/mxv/fixup01", "splacu/mxv/fixup02", "splacu/mxv/fixup03", "splacu/mxv/fixup04", "splacu/mxv/fixup05", "splacu/mxv/fixup06", "splacu/mxv/fixup07", "splacu/mxv/fixup08", "splacu/mxv/fixup09", "splacu/mxv/fixup10", "splacu/mxv/fixup11", "splacu/mxv/fixup12", "splacu/mxv/fixup13", "splacu/mxv/fixup14", "splacu/mxv/fixup15", "splacu/mxv/fixup16"}};
        const int p = (int) (&ph - M->phase);
        {
            SPLACU_PROFILE(kLabels[0][p], s);
            // (not in the parts of a split product: those are captured into the CUDA graphs of the multi-GPU step, whose capture was
            //  measured and verified without programmatic edges)
            if (get_option(OPT_MXV_PDL) && !g_split_call) {
                cudaLaunchConfig_t  cfg = {};
                cudaLaunchAttribute attr[1];
                cfg.gridDim          = dim3((unsigned) grid);
                cfg.blockDim         = dim3(kW * 32);
                cfg.dynamicSmemBytes = smem;
                cfg.stream           = s;
                attr[0].id           = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs    = attr;
                cfg.numAttrs = 1;
                SPLACU_CUDA(cudaLaunchKernelEx(&cfg, kern, sr, sel, reinterpret_cast<const uint32_t*>(ph.Aj), static_cast<const uint32_t*>(ph.Ax),
                                               static_cast<const uint32_t*>(ph.flags), static_cast<const uint32_t*>(ph.seg_base),
                                               static_cast<const uint32_t*>(ph.seg_row), static_cast<const uint32_t*>(ph.chain), ph.head, ph.tail, v, sel_bits, r,
                                               ph.n_tiles, static_cast<const uint32_t*>(M->hub_vals + ph.slot_base), ph.n_slots, gate, gate_min));
                count_launch(1);
            } else {
                kern<<<grid, kW * 32, smem, s>>>(sr, sel, reinterpret_cast<const uint32_t*>(ph.Aj), ph.Ax, ph.flags, ph.seg_base, ph.seg_row, ph.chain, ph.head,
                                                        ph.tail, v, sel_bits, r, ph.n_tiles, M->hub_vals + ph.slot_base, ph.n_slots, gate, gate_min);
                SPLACU_LAUNCH_CHECK();
            }
        }
        if (!get_option(OPT_MXV_FIXUP_MERGE) && !g_split_call) {
            SPLACU_PROFILE(kLabels[1][p], s);
            mxv_seg_fixup_kernel<T, S><<<(ph.n_tiles + kBlock - 1) / kBlock, kBlock, 0, s>>>(sr, sel, ph.chain, ph.chain_row, ph.head, ph.tail, sel_bits, r,
                                                                                            ph.n_tiles, gate, gate_min);
            SPLACU_LAUNCH_CHECK();
        }
        return 0;
    }

    int seg_mxv(const Csr* M, int dtype, int op_mult, int op_add, const Select& sel, const void* d_v, const void* d_mask, void* d_r, uint32_t init_bits,
                const uint32_t* gate, uint32_t gate_min, cudaStream_t s, int parts) {
        // every class accumulates onto r, which starts as init everywhere (unselected and empty rows keep it); with a gate the
        // caller's mask-count pass has filled it already
        if (!gate && (parts & 4)) {
            const int rc = splacu_fill(d_r, init_bits, M->n_rows, s);
            if (rc) return rc;
        }
        if (sel.reads_mask && !(gate && M->sel_bits)) {
            set_error("mxv: the class passes need the selection bitmap of the mask-count pass");
            return SPLACU_E_INVALID;
        }
        (void) d_mask;
        g_split_call   = (parts & 3) != 3;
        const int only = (int) get_option(OPT_MXV_PHASE_ONLY);
        // the row classes of the tail first (mxv_scat.cu): their merge kernels are long done when the tail class needs the SMs
        if ((parts & 2) && (!only || only > M->n_phases)) {
            const int rc = scat_mxv(M, dtype, op_mult, op_add, sel, d_v, d_r, gate, gate_min, s);
            if (rc) return rc;
        }
        return dispatch_dtype(dtype, [&](auto tag) {
            using T       = decltype(tag);
            const T* v    = static_cast<const T*>(d_v);
            T*       r    = static_cast<T*>(d_r);
            return dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
                using S = decltype(sr);
                int ran[kMaxHubPhases + 1], n_ran = 0;
                for (int p = 0; p < M->n_phases; ++p) {
                    const CsrPhase& ph = M->phase[p];
                    if (ph.nnz == 0 || (only && only != p + 1)) continue;
                    ran[n_ran++] = p;
                    if (!(parts & (ph.idx16 ? 1 : 2))) continue;
                    int e;
                    if (sel.reads_mask && gate) e = ph.idx16 ? launch_seg<T, S, true, true>(sr, sel, M, ph, v, M->sel_bits, r, gate, gate_min, s) : launch_seg<T, S, true, false>(sr, sel, M, ph, v, M->sel_bits, r, gate, gate_min, s);
                    else e = ph.idx16 ? launch_seg<T, S, false, true>(sr, sel, M, ph, v, nullptr, r, nullptr, 0u, s) : launch_seg<T, S, false, false>(sr, sel, M, ph, v, nullptr, r, nullptr, 0u, s);
                    if (e) return e;
                }
                if (!(parts & 2)) return 0;// the fix-ups close the product: with the part that runs last
                if (get_option(OPT_MXV_FIXUP_MERGE) >= 2 && M->fix_key)
                    return launch_fixup_rows<T, S>(sr, M, ran, n_ran, (sel.reads_mask && gate) ? M->sel_bits : nullptr, r, gate, gate_min, s);
                if (get_option(OPT_MXV_FIXUP_MERGE) || (parts & 3) != 3)
                    return launch_fixup_all<T, S>(sr, M, ran, n_ran, (sel.reads_mask && gate) ? M->sel_bits : nullptr, r, gate, gate_min, s);
                return 0;
            });
        });
    }

}// namespace splacu
