// vector_ops.cu -- vector format glue and the element-wise neighbours of the hot path.
//   fill / coo<->dense        replaces reference src/opencl/kernels/fill.cl, vector_formats.cl
//   bitmap count / ordered emit  replaces the atomic (unordered) compaction of vector_formats.cl:42-57 and the
//                             radix-sort + reduce-by-key tail of cl_vxm.hpp:157-173 with an order-preserving scan
//   v_assign / v_count_mf / v_eadd(_fdb) / v_reduce   reference src/cpu/cpu_v_*.hpp semantics
// All kernels are HBM-bound streaming passes: 128-bit accesses where alignment allows, grid sized in
// multiples of the SM count, grid-stride loops.
#include "common.cuh"
#include "profile.cuh"
#include "ops.cuh"

#include <cub/block/block_scan.cuh>

namespace splacu {

    static constexpr int kBlock = 256;

    // ------------------------------------------------------------------------------------------
    // fill
    __global__ void __launch_bounds__(kBlock) fill_kernel(uint32_t* __restrict__ dst, uint32_t value, size_t n) {
        const size_t tid    = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
        const size_t stride = (size_t) gridDim.x * blockDim.x;
        // 16-byte aligned body
        const size_t head = min(n, (size_t) ((16 - ((uintptr_t) dst & 15)) & 15) / 4);
        if (tid < head) dst[tid] = value;
        uint4*       d4 = reinterpret_cast<uint4*>(dst + head);
        const size_t n4 = (n - head) / 4;
        const uint4  v4 = make_uint4(value, value, value, value);
        for (size_t i = tid; i < n4; i += stride) d4[i] = v4;
        const size_t tail = head + n4 * 4;
        if (tid < n - tail) dst[tail + tid] = value;
    }

    // ------------------------------------------------------------------------------------------
    // exclusive scan (uint32): block-local scan -> scan of block sums (single CTA) -> add back.
    static constexpr int kScanItems = 4;
    static constexpr int kScanTile  = kBlock * kScanItems;

    __device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        return x;
    }

    // exclusive block scan of one value per thread; returns exclusive prefix, total via smem broadcast
    __device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t* s_warp /*[33]*/, uint32_t& total) {
        const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const uint32_t incl = warp_incl_scan(x);
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            const int nw = blockDim.x >> 5;
            uint32_t  v  = lane < nw ? s_warp[lane] : 0u;
            uint32_t  iv = warp_incl_scan(v);
            s_warp[lane] = iv - v;
            if (lane == 31) s_warp[32] = iv;
        }
        __syncthreads();
        total            = s_warp[32];
        const uint32_t r = s_warp[w] + incl - x;
        __syncthreads();
        return r;
    }

    __global__ void __launch_bounds__(kBlock) scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                               uint32_t n, uint32_t* __restrict__ block_sums) {
        __shared__ uint32_t s_warp[33];
        const uint32_t      base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
        uint32_t            v[kScanItems];
        uint32_t            sum = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (base + k < n) ? in[base + k] : 0u;
            sum += v[k];
        }
        uint32_t total;
        uint32_t prefix = block_excl_scan(sum, s_warp, total);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n) out[base + k] = prefix;
            prefix += v[k];
        }
        if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
    }

    // single CTA: exclusive scan of m block sums in place, total -> *d_total
    __global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ sums, uint32_t m, uint32_t* __restrict__ d_total) {
        __shared__ uint32_t s_warp[33];
        uint32_t            carry = 0;
        for (uint32_t base = 0; base < m; base += blockDim.x) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t x = i < m ? sums[i] : 0u;
            uint32_t       total;
            const uint32_t p = block_excl_scan(x, s_warp, total);
            if (i < m) sums[i] = carry + p;
            carry += total;
        }
        if (threadIdx.x == 0 && d_total) *d_total = carry;
    }

    __global__ void __launch_bounds__(kBlock) scan_add_kernel(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ block_sums) {
        const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
        const uint32_t add  = block_sums[blockIdx.x];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (base + k < n) out[base + k] += add;
    }

    int scan_exclusive_u32(Workspace* ws, const uint32_t* d_in, uint32_t* d_out, uint32_t n, uint32_t* d_total, cudaStream_t s) {
        if (n == 0) {
            if (d_total) SPLACU_CUDA(cudaMemsetAsync(d_total, 0, 4, s));
            return 0;
        }
        const uint32_t nb = (n + kScanTile - 1) / kScanTile;
        int            rc = ws_reserve_blocks(ws, nb);
        if (rc) return rc;
        scan_tile_kernel<<<nb, kBlock, 0, s>>>(d_in, d_out, n, ws->block_sums);
        SPLACU_LAUNCH_CHECK();
        scan_sums_kernel<<<1, 1024, 0, s>>>(ws->block_sums, nb, d_total);
        SPLACU_LAUNCH_CHECK();
        if (nb > 1) {
            scan_add_kernel<<<nb, kBlock, 0, s>>>(d_out, n, ws->block_sums);
            SPLACU_LAUNCH_CHECK();
        }
        return 0;
    }

    // ------------------------------------------------------------------------------------------
    // bitmap count + ordered emit. One thread owns one 32-bit bitmap word (32 vector entries); a CTA
    // owns kBlock words. Output order = ascending bit index, i.e. sorted COO like the CPU converters.
    __global__ void __launch_bounds__(kBlock) bitmap_count_kernel(const uint32_t* __restrict__ bitmap, uint32_t n_words,
                                                                  uint32_t tail_mask, uint32_t* __restrict__ block_sums) {
        __shared__ uint32_t s_warp[33];
        const uint32_t      w = blockIdx.x * kBlock + threadIdx.x;
        uint32_t            x = 0;
        if (w < n_words) {
            x = bitmap[w];
            if (w == n_words - 1) x &= tail_mask;
        }
        uint32_t total;
        (void) block_excl_scan((uint32_t) __popc(x), s_warp, total);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
    }

    template<int MODE>
    __global__ void __launch_bounds__(kBlock) bitmap_emit_kernel(uint32_t* __restrict__ bitmap, uint32_t n_words, uint32_t tail_mask,
                                                                 const uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ src,
                                                                 const uint32_t* __restrict__ vi, uint32_t identity,
                                                                 uint32_t* __restrict__ ri, uint32_t* __restrict__ rx) {
        __shared__ uint32_t s_warp[33];
        const uint32_t      w = blockIdx.x * kBlock + threadIdx.x;
        uint32_t            x = 0;
        if (w < n_words) {
            x = bitmap[w];
            if (w == n_words - 1) x &= tail_mask;
        }
        uint32_t total;
        uint32_t pos = block_offsets[blockIdx.x] + block_excl_scan((uint32_t) __popc(x), s_warp, total);
        if (x) {
            if (MODE != EMIT_DENSE) bitmap[w] = 0u;
            const uint32_t base = w << 5;
            while (x) {
                const uint32_t b = __ffs(x) - 1;
                x &= x - 1;
                const uint32_t k = base + b;
                if (MODE == EMIT_INDIRECT) {
                    const uint32_t i = vi[k];
                    ri[pos]          = i;
                    rx[pos]          = src[i];
                } else if (MODE == EMIT_CONST) {
                    ri[pos] = k;
                    rx[pos] = identity;
                } else {
                    ri[pos] = k;
                    rx[pos] = src[k];
                    if (MODE == EMIT_ACC_RESET) src[k] = identity;
                }
                ++pos;
            }
        }
    }

    int bitmap_count(Workspace* ws, const uint32_t* d_bitmap, uint32_t n, cudaStream_t s) {
        if (n == 0) {
            SPLACU_CUDA(cudaMemsetAsync(ws->d_scalars, 0, 4, s));
            return 0;
        }
        const uint32_t n_words   = (n + 31) / 32;
        const uint32_t tail_mask = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
        const uint32_t nb        = (n_words + kBlock - 1) / kBlock;
        int            rc        = ws_reserve_blocks(ws, nb);
        if (rc) return rc;
        bitmap_count_kernel<<<nb, kBlock, 0, s>>>(d_bitmap, n_words, tail_mask, ws->block_sums);
        SPLACU_LAUNCH_CHECK();
        scan_sums_kernel<<<1, 1024, 0, s>>>(ws->block_sums, nb, ws->d_scalars);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    int bitmap_emit(Workspace* ws, uint32_t* d_bitmap, uint32_t n, int mode, uint32_t* d_src, const uint32_t* d_vi,
                    uint32_t identity, uint32_t* d_ri, uint32_t* d_rx, cudaStream_t s) {
        if (n == 0) return 0;
        const uint32_t n_words   = (n + 31) / 32;
        const uint32_t tail_mask = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
        const uint32_t nb        = (n_words + kBlock - 1) / kBlock;
        if (mode == EMIT_ACC_RESET)
            bitmap_emit_kernel<EMIT_ACC_RESET><<<nb, kBlock, 0, s>>>(d_bitmap, n_words, tail_mask, ws->block_sums, d_src, d_vi, identity, d_ri, d_rx);
        else if (mode == EMIT_DENSE)
            bitmap_emit_kernel<EMIT_DENSE><<<nb, kBlock, 0, s>>>(d_bitmap, n_words, tail_mask, ws->block_sums, d_src, d_vi, identity, d_ri, d_rx);
        else if (mode == EMIT_CONST)
            bitmap_emit_kernel<EMIT_CONST><<<nb, kBlock, 0, s>>>(d_bitmap, n_words, tail_mask, ws->block_sums, d_src, d_vi, identity, d_ri, d_rx);
        else
            bitmap_emit_kernel<EMIT_INDIRECT><<<nb, kBlock, 0, s>>>(d_bitmap, n_words, tail_mask, ws->block_sums, d_src, d_vi, identity, d_ri, d_rx);
        SPLACU_LAUNCH_CHECK();
        return 0;
    }

    // ------------------------------------------------------------------------------------------
    // coo -> dense scatter (dense pre-filled with the fill value)
    __global__ void __launch_bounds__(kBlock) scatter_kernel(uint32_t nv, const uint32_t* __restrict__ vi, const uint32_t* __restrict__ vx,
                                                             uint32_t* __restrict__ dense, uint32_t n) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += stride) {
            const uint32_t i = vi[k];
            if (i < n) dense[i] = vx[k];
        }
    }

    __global__ void __launch_bounds__(kBlock) gather_kernel(uint32_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src,
                                                            uint32_t* __restrict__ dst) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) dst[k] = src[idx[k]];
    }

    // element k of segment q (seg_off[q] <= k < seg_off[q + 1]): peer_dst[q][dst_idx[k]] = src[src_idx[k]] -- a gather from the local
    // vector stored straight into the peers' buffers (NVLink peer stores when peer_dst[q] is peer-mapped memory)
    __global__ void __launch_bounds__(kBlock) push_peers_kernel(uint32_t n, const uint32_t* __restrict__ src_idx, const uint32_t* __restrict__ dst_idx,
                                                                const uint32_t* __restrict__ seg_off, uint32_t n_peers, uint32_t* const* __restrict__ peer_dst,
                                                                const uint32_t* __restrict__ src) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
            uint32_t q = 0;
            while (q + 1 < n_peers && k >= seg_off[q + 1]) ++q;
            peer_dst[q][dst_idx[k]] = src[src_idx[k]];
        }
    }

    // dense -> bitmap of entries != fill (value comparison in T). One warp ballot = one bitmap word.
    template<typename T>
    __global__ void __launch_bounds__(kBlock) mark_nonfill_kernel(const T* __restrict__ dense, uint32_t n, T fill, uint32_t* __restrict__ bitmap) {
        const uint32_t n_pad  = (n + 31) & ~31u;
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
            const bool     p = (i < n) && value_neq(dense[i], fill);
            const uint32_t m = __ballot_sync(0xffffffffu, p);
            if ((threadIdx.x & 31) == 0) bitmap[i >> 5] = m;
        }
    }

    // ------------------------------------------------------------------------------------------
    // neighbours
    template<typename T>
    __global__ void __launch_bounds__(kBlock) assign_dense_kernel(int op, Select sel, uint32_t n, T* __restrict__ r, const T* __restrict__ mask, T value) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const bool take = sel.reads_mask ? sel.test(mask[i]) : (sel.classes != 0u);
            if (take) r[i] = bin_dynamic<T>(op, r[i], value);
        }
    }

    template<typename T>
    __global__ void __launch_bounds__(kBlock) assign_sparse_kernel(int op, Select sel, T* __restrict__ r, uint32_t nm, const uint32_t* __restrict__ mi,
                                                                   const T* __restrict__ mx, T value) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nm; k += stride) {
            const bool take = sel.reads_mask ? sel.test(mx[k]) : (sel.classes != 0u);
            if (take) {
                const uint32_t i = mi[k];
                r[i]             = bin_dynamic<T>(op, r[i], value);
            }
        }
    }

    template<typename T>
    __global__ void __launch_bounds__(kBlock) count_mf_kernel(const T* __restrict__ v, uint32_t n, T fill, uint32_t* __restrict__ d_count) {
        __shared__ uint32_t s_warp[33];
        uint32_t            c      = 0;
        const uint32_t      stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += value_neq(v[i], fill) ? 1u : 0u;
        uint32_t total;
        (void) block_excl_scan(c, s_warp, total);
        if (threadIdx.x == 0 && total) atomicAdd(d_count, total);
    }

    template<typename T>
    __global__ void __launch_bounds__(kBlock) eadd_fdb_dense_kernel(int op, uint32_t n, T* __restrict__ r, const T* __restrict__ v,
                                                                    T* __restrict__ fdb, T fdb_fill) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const T prev = r[i];
            const T next = bin_dynamic<T>(op, prev, v[i]);
            r[i]         = next;
            fdb[i]       = value_neq(prev, next) ? next : fdb_fill;
        }
    }

    // the same for nv <= kSmallFront in ONE launch of one CTA: update, flag, scan in shared memory, write the changed (index,
    // new value) pairs in input order to the small-front scratch and their count to *d_count
    template<typename T>
    __global__ void __launch_bounds__(1024) eadd_fdb_sparse_small_kernel(int op, T* __restrict__ r, uint32_t nv, const uint32_t* __restrict__ vi,
                                                                         const T* __restrict__ vx, uint32_t* __restrict__ out_i,
                                                                         uint32_t* __restrict__ out_x, uint32_t* __restrict__ d_count) {
        using BlockScan = cub::BlockScan<uint32_t, 1024>;
        __shared__ typename BlockScan::TempStorage tmp;
        constexpr int  kItems = kSmallFront / 1024;
        uint32_t       flag[kItems], idx[kItems], val[kItems], pos[kItems];
        const uint32_t base = threadIdx.x * kItems;
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t t = base + k;
            flag[k] = 0u, idx[k] = 0u, val[k] = 0u;
            if (t < nv) {
                const uint32_t i    = vi[t];
                const T        prev = r[i];
                const T        next = bin_dynamic<T>(op, prev, vx[t]);
                r[i]                = next;
                flag[k]             = value_neq(prev, next) ? 1u : 0u;
                idx[k]              = i;
                val[k]              = to_bits(next);
            }
        }
        uint32_t total;
        BlockScan(tmp).ExclusiveSum(flag, pos, total);
#pragma unroll
        for (int k = 0; k < kItems; ++k)
            if (flag[k]) {
                out_i[pos[k]] = idx[k];
                out_x[pos[k]] = val[k];
            }
        if (threadIdx.x == 0) *d_count = total;
    }

    // sparse v (unique indices): r[vi[k]] = op(r[vi[k]], vx[k]); bit k of the bitmap = changed
    template<typename T>
    __global__ void __launch_bounds__(kBlock) eadd_fdb_sparse_kernel(int op, T* __restrict__ r, uint32_t nv, const uint32_t* __restrict__ vi,
                                                                     const T* __restrict__ vx, uint32_t* __restrict__ bitmap) {
        const uint32_t n_pad  = (nv + 31) & ~31u;
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += stride) {
            bool changed = false;
            if (k < nv) {
                const uint32_t i    = vi[k];
                const T        prev = r[i];
                const T        next = bin_dynamic<T>(op, prev, vx[k]);
                r[i]                = next;
                changed             = value_neq(prev, next);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, changed);
            if ((threadIdx.x & 31) == 0) bitmap[k >> 5] = m;
        }
    }

    template<typename T>
    __global__ void __launch_bounds__(kBlock) eadd_dense_kernel(int op, uint32_t n, T* __restrict__ r, const T* __restrict__ u, const T* __restrict__ v) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) r[i] = bin_dynamic<T>(op, u[i], v[i]);
    }

    // two-stage reduction; stage 2 runs in the last CTA to finish (threadfence + ticket)
    template<typename T>
    __global__ void __launch_bounds__(kBlock) reduce_kernel(int op, const T* __restrict__ v, uint32_t n, T identity, T init,
                                                            T* __restrict__ partials, uint32_t* __restrict__ ticket, T* __restrict__ result) {
        __shared__ T    s_part[kBlock / 32];
        __shared__ bool s_last;
        T               acc    = identity;
        const uint32_t  stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc = bin_dynamic<T>(op, acc, v[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = bin_dynamic<T>(op, acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            T a = s_part[0];
            for (int w = 1; w < kBlock / 32; ++w) a = bin_dynamic<T>(op, a, s_part[w]);
            partials[blockIdx.x] = a;
            __threadfence();
            s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            T a = identity;
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x) a = bin_dynamic<T>(op, a, *((volatile T*) &partials[b]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a = bin_dynamic<T>(op, a, __shfl_xor_sync(0xffffffffu, a, o));
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = a;
            __syncthreads();
            if (threadIdx.x == 0) {
                T t = s_part[0];
                for (int w = 1; w < kBlock / 32; ++w) t = bin_dynamic<T>(op, t, s_part[w]);
                *result = bin_dynamic<T>(op, init, t);
                *ticket = 0u;
            }
        }
    }

}// namespace splacu

using namespace splacu;

namespace splacu {
    // Multi-GPU pull: the all-gather of the result windows as ONE kernel of peer stores. Every rank writes its own window into
    // the copy of the vector that each peer holds (symmetric allocation, same offset everywhere) with 128-bit stores over NVLink /
    // NVSwitch; the caller closes the step with a cross-device barrier. 8 peers x 8 MB at 8 GPUs: a collective library call costs
    // more in launch + protocol than the bytes do.
    struct PeerPtrs {
        uint32_t* p[SPLACU_MAX_PEERS];
    };
    __global__ void __launch_bounds__(kBlock) publish_window_kernel(PeerPtrs peers, int n_peers, int self, const uint32_t* __restrict__ src, size_t offset,
                                                                    size_t count) {
        // offset and count are multiples of 4 elements and the bases 16-byte aligned (checked on the host)
        const size_t n4     = count / 4;
        const size_t stride = (size_t) gridDim.x * blockDim.x;
        const uint4* s4     = reinterpret_cast<const uint4*>(src + offset);
        constexpr int U     = 4;// 64 bytes per thread and peer in flight
        for (size_t i0 = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
            uint4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < n4) x[u] = s4[i0 + u * stride];
            for (int q = 0; q < n_peers; ++q) {
                if (q == self) continue;
                uint4* d4 = reinterpret_cast<uint4*>(peers.p[q] + offset);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u * stride < n4) d4[i0 + u * stride] = x[u];
            }
        }
    }
}// namespace splacu

extern "C" {

int splacu_publish_window(void* const* peer_bases, int n_peers, int self, size_t offset, size_t count, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(peer_bases && n_peers >= 1 && n_peers <= SPLACU_MAX_PEERS, "bad peer list");
    SPLACU_REQUIRE(self >= 0 && self < n_peers, "bad own index");
    SPLACU_REQUIRE((offset % 4) == 0 && (count % 4) == 0, "window offset and length must be multiples of 4 elements");
    if (count == 0 || n_peers == 1) return SPLACU_OK;
    PeerPtrs pp;
    for (int q = 0; q < SPLACU_MAX_PEERS; ++q) pp.p[q] = q < n_peers ? static_cast<uint32_t*>(peer_bases[q]) : nullptr;
    for (int q = 0; q < n_peers; ++q) SPLACU_REQUIRE(pp.p[q] && (reinterpret_cast<uintptr_t>(pp.p[q]) & 15u) == 0, "peer base must be 16-byte aligned");
    cudaStream_t s = resolve_stream(stream);
    publish_window_kernel<<<grid_for((count / 4 + 3) / 4, kBlock, 8), kBlock, 0, s>>>(pp, n_peers, self, pp.p[self], offset, count);
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

int splacu_fill(void* d_dst, uint32_t value_bits, size_t n, void* stream) {
    SPLACU_CHECK_INIT();
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_dst, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    fill_kernel<<<grid_for((n + 3) / 4, kBlock, 8), kBlock, 0, s>>>(static_cast<uint32_t*>(d_dst), value_bits, n);
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

int splacu_coo_to_dense(uint32_t n, uint32_t fill_bits, uint32_t nv, const uint32_t* d_vi, const void* d_vx, void* d_dense, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/coo_to_dense", resolve_stream(stream));
    SPLACU_REQUIRE(d_dense || n == 0, "null dense pointer");
    int rc = splacu_fill(d_dense, fill_bits, n, stream);
    if (rc) return rc;
    if (nv == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_vi && d_vx, "null coo pointers");
    cudaStream_t s = resolve_stream(stream);
    scatter_kernel<<<grid_for(nv, kBlock, 8), kBlock, 0, s>>>(nv, d_vi, static_cast<const uint32_t*>(d_vx), static_cast<uint32_t*>(d_dense), n);
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

int splacu_v_gather(uint32_t n, const uint32_t* d_idx, const void* d_src, void* d_dst, void* stream) {
    SPLACU_CHECK_INIT();
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_idx && d_src && d_dst, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    gather_kernel<<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(n, d_idx, static_cast<const uint32_t*>(d_src), static_cast<uint32_t*>(d_dst));
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

int splacu_v_push_peers(uint32_t n, const uint32_t* d_src_idx, const uint32_t* d_dst_idx, const uint32_t* d_seg_off, uint32_t n_peers,
                        void* const* d_peer_dst, const void* d_src, void* stream) {
    SPLACU_CHECK_INIT();
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_src_idx && d_dst_idx && d_seg_off && d_peer_dst && d_src && n_peers > 0, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    push_peers_kernel<<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(n, d_src_idx, d_dst_idx, d_seg_off, n_peers, reinterpret_cast<uint32_t* const*>(d_peer_dst),
                                                               static_cast<const uint32_t*>(d_src));
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

int splacu_v_scatter(uint32_t n, const uint32_t* d_idx, const void* d_src, void* d_dst, uint32_t n_dst, void* stream) {
    SPLACU_CHECK_INIT();
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_idx && d_src && d_dst, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    scatter_kernel<<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(n, d_idx, static_cast<const uint32_t*>(d_src), static_cast<uint32_t*>(d_dst), n_dst);
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

static int read_scalar0(Workspace* ws, uint32_t* h_out, cudaStream_t s) {
    SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars, ws->d_scalars, 4, cudaMemcpyDeviceToHost, s));
    SPLACU_CUDA(cudaStreamSynchronize(s));
    *h_out = ws->h_scalars[0];
    return 0;
}

int splacu_dense_to_coo_count(int dtype, uint32_t n, uint32_t fill_bits, const void* d_dense, splacu_workspace handle, uint32_t* h_nr, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dense_to_coo_count", resolve_stream(stream));
    SPLACU_REQUIRE(handle && h_nr, "null pointer");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    *h_nr           = 0;
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_dense, "null dense pointer");
    SPLACU_REQUIRE(ws->pending == 0, "workspace has a pending emit");
    int rc = ws_reserve_vector(ws, n, s);
    if (rc) return rc;
    rc = dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        mark_nonfill_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(static_cast<const T*>(d_dense), n, from_bits<T>(fill_bits), ws->bitmap);
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
    if (rc) return rc;
    // the scratch bitmap now holds marks: acc[] content untouched, but the bitmap must be cleared again by emit
    rc = bitmap_count(ws, ws->bitmap, n, s);
    if (rc) return rc;
    rc = read_scalar0(ws, h_nr, s);
    if (rc) return rc;
    ws->pending    = 3;
    ws->pend_n     = n;
    ws->pend_count = *h_nr;
    return SPLACU_OK;
}

int splacu_dense_to_coo_emit(int dtype, uint32_t n, uint32_t fill_bits, const void* d_dense, splacu_workspace handle, uint32_t* d_ri, void* d_rx, void* stream) {
    (void) dtype;
    (void) fill_bits;
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dense_to_coo_emit", resolve_stream(stream));
    SPLACU_REQUIRE(handle, "null workspace");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(ws->pending == 3 && ws->pend_n == n, "dense_to_coo_emit without matching dense_to_coo_count");
    ws->pending = 0;
    const uint32_t n_words = (n + 31) / 32;
    if (ws->pend_count) {
        SPLACU_REQUIRE(d_ri && d_rx, "null output pointers");
        int rc = bitmap_emit(ws, ws->bitmap, n, EMIT_DENSE, const_cast<uint32_t*>(static_cast<const uint32_t*>(d_dense)), nullptr, 0u,
                             d_ri, static_cast<uint32_t*>(d_rx), s);
        if (rc) return rc;
    }
    SPLACU_CUDA(cudaMemsetAsync(ws->bitmap, 0, (size_t) n_words * 4, s));
    return SPLACU_OK;
}

int splacu_v_assign_masked_dense(int dtype, int op_assign, int op_select, uint32_t n, void* d_r, const void* d_mask, uint32_t value_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_assign_masked", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op_assign), "op_assign not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (n == 0) return SPLACU_OK;
    Select sel = make_select(op_select);
    SPLACU_REQUIRE(d_r && (d_mask || !sel.reads_mask), "null pointer");
    cudaStream_t s = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        assign_dense_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(op_assign, sel, n, static_cast<T*>(d_r), static_cast<const T*>(d_mask), from_bits<T>(value_bits));
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
}

int splacu_v_assign_masked_sparse(int dtype, int op_assign, int op_select, void* d_r, uint32_t nm, const uint32_t* d_mi, const void* d_mx,
                                  uint32_t value_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_assign_masked", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op_assign), "op_assign not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (nm == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_mi && d_mx, "null pointer");
    Select       sel = make_select(op_select);
    cudaStream_t s   = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        assign_sparse_kernel<T><<<grid_for(nm, kBlock, 8), kBlock, 0, s>>>(op_assign, sel, static_cast<T*>(d_r), nm, d_mi, static_cast<const T*>(d_mx), from_bits<T>(value_bits));
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
}

int splacu_v_count_mf_dense(int dtype, uint32_t n, const void* d_v, uint32_t fill_bits, splacu_workspace handle, uint32_t* h_count, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_count_mf", resolve_stream(stream));
    SPLACU_REQUIRE(handle && h_count, "null pointer");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    *h_count        = 0;
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_v, "null pointer");
    SPLACU_CUDA(cudaMemsetAsync(ws->d_scalars, 0, 4, s));
    int rc = dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        count_mf_kernel<T><<<grid_for(n, kBlock, 4), kBlock, 0, s>>>(static_cast<const T*>(d_v), n, from_bits<T>(fill_bits), ws->d_scalars);
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
    if (rc) return rc;
    return read_scalar0(ws, h_count, s);
}

int splacu_v_eadd_fdb_dense(int dtype, int op, uint32_t n, void* d_r, const void* d_v, void* d_fdb, uint32_t fdb_fill_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_eadd_fdb", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op), "op not defined for dtype");
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_v && d_fdb, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        eadd_fdb_dense_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(op, n, static_cast<T*>(d_r), static_cast<const T*>(d_v), static_cast<T*>(d_fdb), from_bits<T>(fdb_fill_bits));
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
}

int splacu_v_eadd_fdb_sparse_begin(int dtype, int op, void* d_r, uint32_t nv, const uint32_t* d_vi, const void* d_vx,
                                   splacu_workspace handle, uint32_t* h_nf, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_eadd_fdb", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op), "op not defined for dtype");
    SPLACU_REQUIRE(handle && h_nf, "null pointer");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    *h_nf           = 0;
    SPLACU_REQUIRE(ws->pending == 0, "workspace has a pending emit");
    if (nv == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_vi && d_vx, "null pointer");
    int rc;
    ws->pend_small = nv <= kSmallFront && get_option(OPT_SMALL_FRONT);
    if (ws->pend_small) {
        rc = dispatch_dtype(dtype, [&](auto tag) {
            using T = decltype(tag);
            eadd_fdb_sparse_small_kernel<T><<<1, 1024, 0, s>>>(op, static_cast<T*>(d_r), nv, d_vi, static_cast<const T*>(d_vx), ws->small + kSmallList,
                                                              ws->small + kSmallList + kSmallFront, ws->d_scalars);
            SPLACU_LAUNCH_CHECK();
            return 0;
        });
        if (rc) return rc;
        rc = read_scalar0(ws, h_nf, s);
        if (rc) return rc;
        ws->pending    = 2;
        ws->pend_n     = nv;
        ws->pend_count = *h_nf;
        return SPLACU_OK;
    }
    rc = ws_reserve_vector(ws, nv, s);
    if (rc) return rc;
    rc = dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        eadd_fdb_sparse_kernel<T><<<grid_for(nv, kBlock, 8), kBlock, 0, s>>>(op, static_cast<T*>(d_r), nv, d_vi, static_cast<const T*>(d_vx), ws->bitmap);
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
    if (rc) return rc;
    rc = bitmap_count(ws, ws->bitmap, nv, s);
    if (rc) return rc;
    rc = read_scalar0(ws, h_nf, s);
    if (rc) return rc;
    ws->pending    = 2;
    ws->pend_n     = nv;
    ws->pend_count = *h_nf;
    ws->pend_vi    = d_vi;
    ws->pend_src   = static_cast<const uint32_t*>(d_r);
    return SPLACU_OK;
}

int splacu_v_eadd_fdb_sparse_emit(splacu_workspace handle, uint32_t* d_fi, void* d_fx, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_eadd_fdb_emit", resolve_stream(stream));
    SPLACU_REQUIRE(handle, "null workspace");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    if (ws->pending == 0) return SPLACU_OK;// nv == 0
    SPLACU_REQUIRE(ws->pending == 2, "eadd_fdb_sparse_emit without matching begin");
    ws->pending = 0;
    SPLACU_REQUIRE(ws->pend_count == 0 || (d_fi && d_fx), "null output pointers");
    if (ws->pend_small) {
        if (ws->pend_count) {
            SPLACU_CUDA(cudaMemcpyAsync(d_fi, ws->small + kSmallList, (size_t) ws->pend_count * 4, cudaMemcpyDeviceToDevice, s));
            SPLACU_CUDA(cudaMemcpyAsync(d_fx, ws->small + kSmallList + kSmallFront, (size_t) ws->pend_count * 4, cudaMemcpyDeviceToDevice, s));
        }
        return SPLACU_OK;
    }
    // the emit kernel also clears the words it visits, so the scratch bitmap is all-zero again afterwards
    return bitmap_emit(ws, ws->bitmap, ws->pend_n, EMIT_INDIRECT, const_cast<uint32_t*>(ws->pend_src), ws->pend_vi, 0u, d_fi,
                       static_cast<uint32_t*>(d_fx), s);
}

int splacu_v_eadd_dense(int dtype, int op, uint32_t n, void* d_r, const void* d_u, const void* d_v, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_eadd", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op), "op not defined for dtype");
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_u && d_v, "null pointer");
    cudaStream_t s = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        eadd_dense_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(op, n, static_cast<T*>(d_r), static_cast<const T*>(d_u), static_cast<const T*>(d_v));
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
}

int splacu_v_reduce_dense(int dtype, int op, uint32_t n, const void* d_v, uint32_t init_bits, splacu_workspace handle, uint32_t* h_result_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/v_reduce", resolve_stream(stream));
    SPLACU_REQUIRE(op_valid_for(dtype, op), "op not defined for dtype");
    SPLACU_REQUIRE(is_assoc_commutative(op), "v_reduce needs an associative and commutative op on the device");
    SPLACU_REQUIRE(handle && h_result_bits, "null pointer");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    *h_result_bits  = init_bits;
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_v, "null pointer");
    const int grid = grid_for(n, kBlock, 4);
    int       rc   = ws_reserve_blocks(ws, (uint32_t) grid);
    if (rc) return rc;
    rc = dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        reduce_kernel<T><<<grid, kBlock, 0, s>>>(op, static_cast<const T*>(d_v), n, add_identity<T>(op), from_bits<T>(init_bits),
                                                 reinterpret_cast<T*>(ws->block_sums), ws->d_scalars + 1, reinterpret_cast<T*>(ws->d_scalars));
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
    if (rc) return rc;
    return read_scalar0(ws, h_result_bits, s);
}

}// extern "C"
