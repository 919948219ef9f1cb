// vxm_push.cu -- masked semiring push product r = v x M over a sparse frontier (spla exec_vxm_masked, SpMSpV).
//
// Semantics: reference src/cpu/cpu_vxm.hpp:92-125 (see include/splacu.h). Replaces the OpenCL pipeline
// count -> blocking read -> collect -> radix sort (6-8 passes) -> reduce-by-key -> blocking read
// (reference src/opencl/cl_vxm.hpp:107-173) with
//   1. degree gather + exclusive scan over the frontier            (load-balancing offsets, no host sync)
//   2. load-balanced expand: every thread owns a contiguous run of edge slots of the expanded frontier, finds its
//      frontier entry by binary search in the offsets, tests the mask BEFORE touching Ax, and folds the product into a
//      dense per-vector accumulator with an op-specialised atomic; touched columns are recorded in a bitmap
//   3. popcount + scan of the bitmap (n/8 bytes) -> result count    (the single 4-byte device->host sync)
//   4. ordered emit of (j, acc[j]) and reset of exactly the touched scratch
// Step 2 is exact for every associative + commutative add whose identity e satisfies add(e, x) == x
// (PLUS, MULT, MIN, MAX, BOR, BAND, BXOR; LOR/LAND when the products are already 0/1), because the reference
// stores the first product of a column directly. Every other op pair takes the exact ordered path:
// expand (column, product) pairs at their deterministic frontier-order positions, stable radix sort by column,
// strict left-to-right fold per column.
#include "common.cuh"
#include "profile.cuh"
#include "jit.cuh"
#include "ops.cuh"

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace splacu {

    static constexpr int kBlock = 256;
#ifndef SPLACU_VXM_EPT
#define SPLACU_VXM_EPT 4
#endif
    static constexpr int kEpt   = SPLACU_VXM_EPT;// edge slots per thread per chunk (4: measured against 2 / 8 / 16, profiles/r02_exp_notes.txt)

    // deg[t] = length of row vi[t]; optionally also rowstart[t] = Ap[vi[t]] (so that the expand reads it sequentially instead of
    // chasing vi -> Ap) and *differs |= (vx[t] != vx[0]) (bit patterns): the structure-only push needs one frontier value
    __global__ void __launch_bounds__(kBlock) vxm_degrees_kernel(uint32_t nv, const uint32_t* __restrict__ vi, const uint32_t* __restrict__ Ap,
                                                                 uint32_t* __restrict__ deg, uint32_t* __restrict__ rowstart,
                                                                 const uint32_t* __restrict__ vx, uint32_t* __restrict__ differs) {
        const uint32_t stride = gridDim.x * blockDim.x;
        const uint32_t x0     = vx ? vx[0] : 0u;
        bool           d      = false;
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nv; t += stride) {
            const uint32_t i = vi[t];
            const uint32_t a = Ap[i];
            deg[t]           = Ap[i + 1] - a;
            if (rowstart) rowstart[t] = a;
            if (vx) d |= vx[t] != x0;
        }
        if (vx && __any_sync(__activemask(), d) && d) *differs = 1u;
    }

    // Small fronts (nv <= kSmallFront) are launch-latency bound: ONE CTA gathers the degrees, scans them in shared memory and
    // clears the touched-column counter of the expand (instead of degree kernel + 3 scan launches).
    __global__ void __launch_bounds__(1024) vxm_offsets_small_kernel(uint32_t nv, const uint32_t* __restrict__ vi, const uint32_t* __restrict__ Ap,
                                                                     uint32_t* __restrict__ off /*[nv + 1]*/, uint32_t* __restrict__ rowstart,
                                                                     uint32_t* __restrict__ counter) {
        using BlockScan = cub::BlockScan<uint32_t, 1024>;
        __shared__ typename BlockScan::TempStorage tmp;
        constexpr int  kItems = kSmallFront / 1024;
        uint32_t       d[kItems];
        const uint32_t base = threadIdx.x * kItems;
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t t = base + k;
            uint32_t       x = 0;
            if (t < nv) {
                const uint32_t i = vi[t];
                const uint32_t a = Ap[i];
                x                = Ap[i + 1] - a;
                rowstart[t]      = a;
            }
            d[k] = x;
        }
        uint32_t total;
        BlockScan(tmp).ExclusiveSum(d, d, total);
#pragma unroll
        for (int k = 0; k < kItems; ++k)
            if (base + k < nv) off[base + k] = d[k];
        if (threadIdx.x == 0) {
            off[nv]  = total;
            *counter = 0u;
        }
    }

    // ... and when the expand touched <= kSmallList columns, one CTA sorts the list it left, emits (j, acc[j]) in ascending
    // order and resets exactly that scratch (instead of two passes over the n-bit bitmap plus a scan).
    __global__ void __launch_bounds__(1024) vxm_emit_small_kernel(const uint32_t* __restrict__ list, uint32_t nr, uint32_t* __restrict__ acc,
                                                                  uint32_t* __restrict__ bitmap, uint32_t identity, int end_bit,
                                                                  uint32_t* __restrict__ ri, uint32_t* __restrict__ rx) {
        constexpr int kItems = kSmallList / 1024;
        using BlockSort     = cub::BlockRadixSort<uint32_t, 1024, kItems>;
        __shared__ typename BlockSort::TempStorage tmp;
        uint32_t key[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t q = threadIdx.x * kItems + k;
            key[k]           = q < nr ? list[q] : 0xffffffffu;
        }
        BlockSort(tmp).Sort(key, 0, end_bit);// only the bits a column id has (the padding sorts last: stable, all-ones digits); blocked arrangement: thread t holds ranks t * kItems .. t * kItems + kItems - 1
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const uint32_t q = threadIdx.x * kItems + k;
            if (q < nr) {
                const uint32_t j = key[k];
                ri[q]            = j;
                rx[q]            = acc[j];
                acc[j]           = identity;
                bitmap[j >> 5]   = 0u;// every set bit of the word is in the list, so all of them are being reset
            }
        }
    }

    // last t in [0, nv) with off[t] <= e   (off is non-decreasing, off[0] == 0 <= e)
    __device__ __forceinline__ uint32_t find_entry(const uint32_t* __restrict__ off, uint32_t nv, uint32_t e) {
        uint32_t lo = 0, hi = nv;// invariant: off[lo] <= e, (hi == nv or off[hi] > e)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off[mid] <= e) lo = mid;
            else hi = mid;
        }
        return lo;
    }

    // last t in [lo, hi) with off[t] <= e, given off[lo] <= e
    __device__ __forceinline__ uint32_t find_entry_in(const uint32_t* __restrict__ off, uint32_t lo, uint32_t hi, uint32_t e) {
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off[mid] <= e) lo = mid;
            else hi = mid;
        }
        return lo;
    }
    // coarse index of the expanded frontier: first[c] = frontier entry that owns edge slot c * (kBlock * kEpt), first[n_chunks] = nv - 1.
    // A thread of the expand then searches only between the entries of its chunk's borders -- a handful of offsets that all
    // threads of the CTA share (L1 hits) -- instead of walking a 20-level binary search whose lower levels miss (ncu round 2: the
    // searches were ~0.6 of the 2.3 L2 sectors per edge of the structure-only expand).
    __global__ void __launch_bounds__(kBlock) vxm_chunk_index_kernel(const uint32_t* __restrict__ off, uint32_t nv, uint32_t* __restrict__ first) {
        const uint32_t total    = off[nv];
        const uint32_t chunk    = kBlock * kEpt;
        const uint32_t n_chunks = (total + chunk - 1) / chunk;
        const uint32_t stride   = gridDim.x * blockDim.x;
        for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c <= n_chunks; c += stride)
            first[c] = c < n_chunks ? find_entry(off, nv, c * chunk) : nv - 1u;
    }

    // MODE 0: atomic accumulate into acc + bitmap (fast path). MODE 1: write (key, product) pairs (exact path).
    // Every thread owns kEpt consecutive edge slots of the expanded frontier: no CTA barrier, 64 fully independent warps per SM.
    // The slots are handled in PHASES, each a batch of kEpt independent loads: (1) walk the frontier entries, load the column ids
    // and values (8 consecutive words per thread: one or two sectors when the loads are in flight together -- issued one edge at
    // a time, with the random accesses of the previous edge in between, every one of them was its own L2 request, ncu round 2:
    // 5.9 L2 sectors per edge), (2) the selection bits / mask values, (3) the accumulator words, (4) the atomics.
    // rowstart[t] = Ap[vi[t]] comes from the degree kernel, so the walk reads three sequential arrays and never chases vi -> Ap.
    // (Measured alternative, dropped: a CTA-tiled variant that stages the frontier entries of 2048 slots in shared memory and
    // assigns slots round-robin for coalesced Aj loads was 2-3x SLOWER on the 400 M-edge level of an RMAT-24 BFS -- its barriers
    // serialise the dependent random accesses of a tile.)
    template<typename T, typename S, int MODE>
    __global__ void __launch_bounds__(kBlock) vxm_expand_kernel(S sr, Select sel, const uint32_t* __restrict__ Aj, const T* __restrict__ Ax, uint32_t nv,
                                                                    const T* __restrict__ vx, const T* __restrict__ mask,
                                                                    const uint32_t* __restrict__ sel_bits, const uint32_t* __restrict__ off /*[nv+1]*/,
                                                                    const uint32_t* __restrict__ rowstart, const uint32_t* __restrict__ first,
                                                                    T* __restrict__ acc, uint32_t* __restrict__ bitmap, uint32_t* __restrict__ keys,
                                                                    T* __restrict__ vals, uint32_t invalid_key, uint32_t identity_bits,
                                                                    uint32_t* __restrict__ counter, uint32_t* __restrict__ list,
                                                                    const uint32_t* __restrict__ run_if /*null: always; else only when *run_if != 0*/) {
        if (run_if && *run_if == 0u) return;// the structure-only kernel took this call
        const uint32_t total = off[nv];
        const uint32_t chunk = kBlock * kEpt;
        for (uint64_t base = (uint64_t) blockIdx.x * chunk; base < total; base += (uint64_t) gridDim.x * chunk) {
            if (base + threadIdx.x * kEpt >= total) continue;
            const uint32_t e0    = (uint32_t) base + threadIdx.x * kEpt;
            const uint32_t n_e   = min(total - e0, (uint32_t) kEpt);
            const uint32_t c     = (uint32_t) (base / chunk);
            uint32_t       t     = first ? find_entry_in(off, first[c], first[c + 1] + 1u, e0) : find_entry(off, nv, e0);
            uint32_t       t_end = off[t + 1];
            uint32_t       row0  = rowstart[t] - off[t];// k = row0 + e
            T              x     = vx[t];
            uint32_t       j[kEpt];
            T              p[kEpt];
            // (1) column ids and products
#pragma unroll
            for (int q = 0; q < kEpt; ++q) {
                j[q] = 0u;
                p[q] = T(0);
                if ((uint32_t) q < n_e) {
                    while (e0 + q >= t_end) {// next frontier entry (skips entries with empty rows)
                        ++t;
                        t_end = off[t + 1];
                        row0  = rowstart[t] - off[t];
                        x     = vx[t];
                    }
                    const uint32_t k = row0 + e0 + q;
                    j[q]             = Aj[k];
                    p[q]             = sr.mult(x, Ax[k]);
                }
            }
            // (2) select(mask[j])
            bool take[kEpt];
            if (sel_bits) {
                uint32_t w[kEpt];
#pragma unroll
                for (int q = 0; q < kEpt; ++q) w[q] = sel_bits[j[q] >> 5];
#pragma unroll
                for (int q = 0; q < kEpt; ++q) take[q] = (uint32_t) q < n_e && ((w[q] >> (j[q] & 31u)) & 1u) != 0u;
            } else if (sel.reads_mask) {
                T m[kEpt];
#pragma unroll
                for (int q = 0; q < kEpt; ++q) m[q] = mask[j[q]];
#pragma unroll
                for (int q = 0; q < kEpt; ++q) take[q] = (uint32_t) q < n_e && sel.test(m[q]);
            } else {
#pragma unroll
                for (int q = 0; q < kEpt; ++q) take[q] = (uint32_t) q < n_e && sel.classes != 0u;
            }
            if (MODE == 0) {
                // (3) current accumulator words (possibly stale: they only serve to skip atomics that cannot change the value)
                uint32_t cur[kEpt];
#pragma unroll
                for (int q = 0; q < kEpt; ++q) cur[q] = take[q] ? *reinterpret_cast<volatile uint32_t*>(&acc[j[q]]) : 0u;
                // (4) fold
#pragma unroll
                for (int q = 0; q < kEpt; ++q) {
                    if (!take[q]) continue;
                    atomic_combine<T>(sr.add_op(), &acc[j[q]], p[q], from_bits<T>(cur[q]));
                    if (cur[q] == identity_bits) {
                        const uint32_t bit = 1u << (j[q] & 31u);
                        const uint32_t old = atomicOr(&bitmap[j[q] >> 5], bit);
                        if (counter && !(old & bit)) {// first touch of column j: count it, remember it while the list has room
                            const uint32_t pos = atomicAdd(counter, 1u);
                            if (pos < kSmallList) list[pos] = j[q];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < kEpt; ++q)
                    if ((uint32_t) q < n_e) {
                        keys[e0 + q] = take[q] ? j[q] : invalid_key;
                        vals[e0 + q] = take[q] ? p[q] : T(0);
                    }
            }
        }
    }

    // Structure-only push. When every product mult(x, a) is provably ONE value p (uniform frontier values x uniform matrix values,
    // or a mult that ignores the varying side) and the add is idempotent on it (add(p, p) == p: MIN, MAX, BOR, BAND, and LOR / LAND
    // over 0/1 products), the result is (j, p) for every column j reached through a selected edge -- the accumulator, Ax and vx are
    // never read. `cand` starts as the select(mask) bitmap; an edge costs ONE L2 request (its word of cand, read through L2 so that
    // a cleared bit is seen by every SM) instead of two (selection bit + accumulator word); the first edge to reach j clears its bit
    // (atomicAnd) and, if it really was the first, records it in the touched bitmap the ordered emit reads. This is what a BFS level
    // is (reference src/algorithm.cpp:97-99: BAND / BOR over a frontier of ones and an adjacency matrix of ones).
    // `differs` != 0 (the frontier values are not uniform after all): the general kernel runs instead, this one returns at once.
    __device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
        uint32_t r;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
        return r;
    }
    template<typename T, typename S>
    __global__ void __launch_bounds__(kBlock) vxm_expand_struct_kernel(S sr, const uint32_t* __restrict__ Aj, uint32_t nv, const T* __restrict__ vx,
                                                                           uint32_t ax_value, const uint32_t* __restrict__ off /*[nv+1]*/,
                                                                           const uint32_t* __restrict__ rowstart, const uint32_t* __restrict__ first,
                                                                           uint32_t* __restrict__ cand,
                                                                           uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ differs,
                                                                           uint32_t* __restrict__ value_out) {
        if (differs && *differs) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) *value_out = to_bits(sr.mult(vx[0], from_bits<T>(ax_value)));
        const uint32_t total = off[nv];
        const uint32_t chunk = kBlock * kEpt;
        for (uint64_t base = (uint64_t) blockIdx.x * chunk; base < total; base += (uint64_t) gridDim.x * chunk) {
            if (base + threadIdx.x * kEpt >= total) continue;
            uint32_t       e     = (uint32_t) base + threadIdx.x * kEpt;
            const uint32_t e_end = min(total, e + kEpt);
            const uint32_t c     = (uint32_t) (base / chunk);
            uint32_t       t     = first ? find_entry_in(off, first[c], first[c + 1] + 1u, e) : find_entry(off, nv, e);
            uint32_t       t_end = off[t + 1];
            uint32_t       row0  = rowstart[t] - off[t];// k = row0 + e
            uint32_t       j[kEpt];
            uint32_t       w[kEpt];
#pragma unroll
            for (int q = 0; q < kEpt; ++q) {
                j[q] = 0xffffffffu;
                if (e + q < e_end) {
                    while (e + q >= t_end) {// next frontier entry (skips entries with empty rows)
                        ++t;
                        t_end = off[t + 1];
                        row0  = rowstart[t] - off[t];
                    }
                    j[q] = Aj[row0 + e + q];
                }
            }
#pragma unroll
            for (int q = 0; q < kEpt; ++q) w[q] = j[q] != 0xffffffffu ? ld_cg_u32(cand + (j[q] >> 5)) : 0u;// all requests in flight together
#pragma unroll
            for (int q = 0; q < kEpt; ++q) {
                const uint32_t bit = 1u << (j[q] & 31u);
                if (w[q] & bit) {
                    const uint32_t old = atomicAnd(&cand[j[q] >> 5], ~bit);
                    if (old & bit) atomicOr(&bitmap[j[q] >> 5], bit);
                }
            }
        }
    }

    // bit j of sel_bits = select(mask[j]): a 2 MB L2-resident stand-in for the 64 MB mask when a large frontier is expanded
    template<typename T>
    __global__ void __launch_bounds__(kBlock) select_bits_kernel(Select sel, const T* __restrict__ mask, uint32_t n, uint32_t* __restrict__ sel_bits) {
        const uint32_t n_pad  = (n + 31) & ~31u;
        const uint32_t stride = gridDim.x * blockDim.x;
        constexpr int  U      = 4;// independent 32-element groups per warp and step
        for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_pad; i0 += stride * U) {
            bool p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + u * stride;
                p[u]             = (i < n) && sel.test(mask[i]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + u * stride;
                if (i >= n_pad) break;// warp-uniform
                const uint32_t m = __ballot_sync(0xffffffffu, p[u]);
                if ((threadIdx.x & 31) == 0) sel_bits[i >> 5] = m;
            }
        }
    }

    __global__ void __launch_bounds__(kBlock) unpack_bits_kernel(const uint32_t* __restrict__ bits, uint32_t n, uint32_t one, uint32_t zero,
                                                                uint32_t* __restrict__ out) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = ((bits[i >> 5] >> (i & 31u)) & 1u) ? one : zero;
    }

    // exact path: after the stable sort, the head of every key run folds its run left to right
    template<typename T>
    __global__ void __launch_bounds__(kBlock) vxm_fold_runs_kernel(int op_add, uint32_t n_pairs, const uint32_t* __restrict__ keys,
                                                                   const T* __restrict__ vals, uint32_t invalid_key, T* __restrict__ acc,
                                                                   uint32_t* __restrict__ bitmap) {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += stride) {
            const uint32_t key = keys[q];
            if (key == invalid_key) continue;
            if (q > 0 && keys[q - 1] == key) continue;
            T a = vals[q];
            for (uint32_t p = q + 1; p < n_pairs && keys[p] == key; ++p) a = bin_dynamic<T>(op_add, a, vals[p]);
            acc[key] = a;
            atomicOr(&bitmap[key >> 5], 1u << (key & 31u));
        }
    }

    static bool fast_path_ok(int op_mult, int op_add) {
        switch (op_add) {
            case SPLACU_PLUS: case SPLACU_MULT: case SPLACU_MIN: case SPLACU_MAX:
            case SPLACU_BOR: case SPLACU_BAND: case SPLACU_BXOR: return true;
            case SPLACU_LOR: case SPLACU_LAND:
                // the reference returns the RAW product for a column with a single contribution; the accumulator
                // normalises to 0/1, so the two agree only when products are 0/1 already
                return op_mult == SPLACU_LOR || op_mult == SPLACU_LAND || op_mult == SPLACU_BONE;
            default: return false;
        }
    }

    int vxm_finish(Workspace* ws, uint32_t* h_nr, cudaStream_t s);

    template<typename T>
    static int vxm_begin_typed(const Csr* M, int op_mult, int op_add, const Select& sel, uint32_t nv, const uint32_t* d_vi, const T* d_vx,
                               const T* d_mask, Workspace* ws, uint32_t* h_nr, cudaStream_t s, bool defer_finish) {
        const uint32_t n = M->n_cols;
        int            rc;
        if ((rc = ws_reserve_vector(ws, n, s))) return rc;
        if ((rc = ws_reserve_pairs(ws, 0, (size_t) nv + 1))) return rc;

        // 1. load-balancing offsets
        const bool fast  = fast_path_ok(op_mult, op_add);
        const bool small = fast && nv <= kSmallFront && get_option(OPT_SMALL_FRONT);
        uint32_t*  counter = small ? ws->d_scalars : nullptr;// the result count of a small front comes from the expand itself
        // structure-only candidate: big frontier, idempotent add, products that cannot vary (decided here except for the
        // uniformity of the frontier values, which the degree kernel checks on the device)
        const bool idem      = op_add == SPLACU_MIN || op_add == SPLACU_MAX || op_add == SPLACU_BOR || op_add == SPLACU_BAND || op_add == SPLACU_LOR || op_add == SPLACU_LAND;
        const bool ax_free   = op_mult == SPLACU_FIRST || op_mult == SPLACU_BONE; // vxm: mult(v, a)
        const bool vx_free   = op_mult == SPLACU_SECOND || op_mult == SPLACU_BONE;
        const bool strct     = fast && !small && idem && get_option(OPT_VXM_STRUCT) && (uint64_t) nv * 64 >= n && (M->ax_uniform || ax_free);
        uint32_t*  differs   = strct ? ws->d_scalars + 2 : nullptr;
        uint32_t*  rowstart  = ws->offsets + ws->cap_offsets;
        ws->last_struct      = false;
        if (small) {
            vxm_offsets_small_kernel<<<1, 1024, 0, s>>>(nv, d_vi, M->Ap, ws->offsets, rowstart, counter);
            SPLACU_LAUNCH_CHECK();
        } else {
            if (strct) SPLACU_CUDA(cudaMemsetAsync(differs, 0, 8, s));// [2] differs, [3] the product
            vxm_degrees_kernel<<<grid_for(nv, kBlock, 8), kBlock, 0, s>>>(nv, d_vi, M->Ap, ws->offsets, rowstart,
                                                                         strct && !vx_free ? reinterpret_cast<const uint32_t*>(d_vx) : nullptr, differs);
            SPLACU_LAUNCH_CHECK();
            if ((rc = scan_exclusive_u32(ws, ws->offsets, ws->offsets, nv, ws->offsets + nv, s))) return rc;
        }

        // coarse index of the expanded frontier (one entry per 2048 edge slots; sized for the whole matrix)
        const uint32_t* chunk_first = nullptr;// small fronts are launch bound: they keep the full search
        if (!small) {
            if ((rc = ws_reserve_chunks(ws, (size_t) M->nnz / (kBlock * kEpt) + 2))) return rc;
            vxm_chunk_index_kernel<<<grid_for((size_t) M->nnz / (kBlock * kEpt) + 2, kBlock, 8), kBlock, 0, s>>>(ws->offsets, nv, ws->chunk_first);
            SPLACU_LAUNCH_CHECK();
            chunk_first = ws->chunk_first;
        }
        const T    identity = fast ? add_identity<T>(op_add) : from_bits<T>(ws->acc_identity);
        T*         acc      = reinterpret_cast<T*>(ws->acc);
        const int  grid     = sm_count() * 8;

        if (fast) {
            if (!ws->acc_clean || ws->acc_identity != to_bits(identity)) {
                if ((rc = splacu_fill(ws->acc, to_bits(identity), ws->cap_n, (void*) s))) return rc;
                ws->acc_identity = to_bits(identity);
                ws->acc_clean    = true;
            }
            const uint32_t* sel_bits = nullptr;
            if (strct || (sel.reads_mask && get_option(OPT_VXM_SELBITS) && (uint64_t) nv * 64 >= n)) {
                // (nv * 64 >= n: a frontier this large expands to at least ~n edges on the graphs this path sees)
                if ((rc = ws_reserve_selbits(ws, n))) return rc;
                if (sel.reads_mask) {
                    select_bits_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(sel, d_mask, n, ws->sel_bits);
                    SPLACU_LAUNCH_CHECK();
                    sel_bits = ws->sel_bits;
                } else {// ALWAYS, structure-only: every column is a candidate
                    SPLACU_CUDA(cudaMemsetAsync(ws->sel_bits, 0xff, ((size_t) n + 31) / 32 * 4, s));
                }
            }
            rc = dispatch_semiring<T>(op_mult, op_add, [&](auto sr) {
                using S = decltype(sr);
                if (strct) {
                    vxm_expand_struct_kernel<T, S><<<grid, kBlock, 0, s>>>(sr, M->Aj, nv, d_vx, M->ax_value, ws->offsets, rowstart, chunk_first, ws->sel_bits,
                                                                          ws->bitmap, differs, ws->d_scalars + 3);
                    SPLACU_LAUNCH_CHECK();
                }
                vxm_expand_kernel<T, S, 0><<<grid, kBlock, 0, s>>>(sr, sel, M->Aj, reinterpret_cast<const T*>(M->Ax), nv, d_vx, d_mask,
                                                                   sel_bits, ws->offsets, rowstart, chunk_first, acc, ws->bitmap, nullptr, nullptr, 0u, to_bits(identity), counter, ws->small,
                                                                   strct ? differs : nullptr);
                SPLACU_LAUNCH_CHECK();
                return 0;
            });
            if (rc) return rc;
        } else {
            // exact ordered path: needs the pair count on the host to size the sort buffers
            SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars + 1, ws->offsets + nv, 4, cudaMemcpyDeviceToHost, s));
            SPLACU_CUDA(cudaStreamSynchronize(s));
            const uint32_t n_pairs = ws->h_scalars[1];
            if (n_pairs) {
                if ((rc = ws_reserve_pairs(ws, n_pairs, (size_t) nv + 1))) return rc;
                SemiringDynamic<T> sr;
                sr.mul   = op_mult;
                sr.ad    = op_add;
                sr.ident = T(0);
                vxm_expand_kernel<T, SemiringDynamic<T>, 1><<<grid, kBlock, 0, s>>>(sr, sel, M->Aj, reinterpret_cast<const T*>(M->Ax), nv, d_vx,
                                                                                   d_mask, nullptr, ws->offsets, rowstart, chunk_first, nullptr, nullptr, ws->keys_a,
                                                                                   reinterpret_cast<T*>(ws->vals_a), n, 0u, nullptr, nullptr, nullptr);
                SPLACU_LAUNCH_CHECK();
                int end_bit = 1;
                while (end_bit < 32 && (n >> end_bit) != 0u) ++end_bit;// keys are in [0, n]
                size_t tmp_bytes = 0;
                SPLACU_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ws->keys_a, ws->keys_b, ws->vals_a, ws->vals_b, (int) n_pairs, 0, end_bit, s));
                if (tmp_bytes > ws->cap_sort_tmp) {
                    SPLACU_CUDA(cudaStreamSynchronize(s));
                    cudaFree(ws->sort_tmp);
                    ws->sort_tmp = nullptr;
                    SPLACU_CUDA(cudaMalloc(&ws->sort_tmp, tmp_bytes + tmp_bytes / 4));
                    ws->cap_sort_tmp = tmp_bytes + tmp_bytes / 4;
                }
                SPLACU_CUDA(cub::DeviceRadixSort::SortPairs(ws->sort_tmp, tmp_bytes, ws->keys_a, ws->keys_b, ws->vals_a, ws->vals_b, (int) n_pairs, 0, end_bit, s));
                count_launch(8);
                vxm_fold_runs_kernel<T><<<grid_for(n_pairs, kBlock, 8), kBlock, 0, s>>>(op_add, n_pairs, ws->keys_b, reinterpret_cast<const T*>(ws->vals_b), n, acc, ws->bitmap);
                SPLACU_LAUNCH_CHECK();
            }
        }

        // 3. count: enqueued here, read by vxm_finish (the one host synchronisation of the call)
        if (!small && (rc = bitmap_count(ws, ws->bitmap, n, s))) return rc;
        SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars, ws->d_scalars, strct ? 16 : 4, cudaMemcpyDeviceToHost, s));
        ws->fin_strct    = strct;
        ws->fin_small    = small;
        ws->fin_n        = n;
        ws->fin_identity = to_bits(identity);
        ws->pending      = 4;// enqueued, not yet finished
        if (defer_finish) return 0;
        return vxm_finish(ws, h_nr, s);
    }

    int vxm_finish(Workspace* ws, uint32_t* h_nr, cudaStream_t s) {
        int rc;
        SPLACU_CUDA(cudaStreamSynchronize(s));
        const bool strct = ws->fin_strct, small = ws->fin_small;
        *h_nr             = ws->h_scalars[0];
        ws->pend_const    = strct && ws->h_scalars[2] == 0u;// the frontier values were uniform: the structure-only kernel ran
        ws->pend_value    = ws->h_scalars[3];
        ws->last_struct   = ws->pend_const;
        ws->pend_small    = small && *h_nr <= kSmallList;
        if (small && !ws->pend_small && (rc = bitmap_count(ws, ws->bitmap, ws->fin_n, s))) return rc;// block offsets for the bitmap emit
        ws->pending       = 1;
        ws->pend_n        = ws->fin_n;
        ws->pend_count    = *h_nr;
        ws->pend_identity = ws->fin_identity;
        return 0;
    }

    // User-defined ops (jit.cu): the exact ordered path with the user's mult / add / select inside -- offsets, (column, product) pairs
    // at their frontier-order positions, stable radix sort by column, left-to-right fold per column (reference src/cpu/cpu_vxm.hpp:92-125).
    int vxm_begin_jit(const Csr* M, const jit::Module* jm, uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask, Workspace* ws,
                      uint32_t* h_nr, cudaStream_t s) {
        uint32_t n = M->n_cols;
        int      rc;
        if ((rc = ws_reserve_vector(ws, n, s))) return rc;
        if ((rc = ws_reserve_pairs(ws, 0, (size_t) nv + 1))) return rc;
        uint32_t* rowstart = ws->offsets + ws->cap_offsets;
        ws->last_struct    = false;
        vxm_degrees_kernel<<<grid_for(nv, kBlock, 8), kBlock, 0, s>>>(nv, d_vi, M->Ap, ws->offsets, rowstart, nullptr, nullptr);
        SPLACU_LAUNCH_CHECK();
        if ((rc = scan_exclusive_u32(ws, ws->offsets, ws->offsets, nv, ws->offsets + nv, s))) return rc;
        SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars + 1, ws->offsets + nv, 4, cudaMemcpyDeviceToHost, s));
        SPLACU_CUDA(cudaStreamSynchronize(s));
        uint32_t n_pairs = ws->h_scalars[1];
        if (n_pairs) {
            if ((rc = ws_reserve_pairs(ws, n_pairs, (size_t) nv + 1))) return rc;
            rowstart              = ws->offsets + ws->cap_offsets;
            const uint32_t* aj    = M->Aj;
            const uint32_t* ax    = M->Ax;
            const uint32_t* off   = ws->offsets;
            uint32_t*       keys  = ws->keys_a;
            uint32_t*       vals  = ws->vals_a;
            uint32_t        inval = n;
            void* a1[] = {&aj, &ax, &nv, &d_vx, &d_mask, &off, &rowstart, &keys, &vals, &inval};
            if ((rc = jit::launch(jm, jit::K_VXM_PAIRS, n_pairs, a1, s))) return rc;
            int end_bit = 1;
            while (end_bit < 32 && (n >> end_bit) != 0u) ++end_bit;// keys are in [0, n]
            size_t tmp_bytes = 0;
            SPLACU_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ws->keys_a, ws->keys_b, ws->vals_a, ws->vals_b, (int) n_pairs, 0, end_bit, s));
            if (tmp_bytes > ws->cap_sort_tmp) {
                SPLACU_CUDA(cudaStreamSynchronize(s));
                cudaFree(ws->sort_tmp);
                ws->sort_tmp = nullptr;
                SPLACU_CUDA(cudaMalloc(&ws->sort_tmp, tmp_bytes + tmp_bytes / 4));
                ws->cap_sort_tmp = tmp_bytes + tmp_bytes / 4;
            }
            SPLACU_CUDA(cub::DeviceRadixSort::SortPairs(ws->sort_tmp, tmp_bytes, ws->keys_a, ws->keys_b, ws->vals_a, ws->vals_b, (int) n_pairs, 0, end_bit, s));
            count_launch(8);
            const uint32_t* skeys = ws->keys_b;
            const uint32_t* svals = ws->vals_b;
            uint32_t*       acc   = ws->acc;
            uint32_t*       bm    = ws->bitmap;
            void* a2[] = {&n_pairs, &skeys, &svals, &inval, &acc, &bm};
            if ((rc = jit::launch(jm, jit::K_VXM_FOLD, n_pairs, a2, s))) return rc;
        }
        if ((rc = bitmap_count(ws, ws->bitmap, n, s))) return rc;
        SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars, ws->d_scalars, 4, cudaMemcpyDeviceToHost, s));
        SPLACU_CUDA(cudaStreamSynchronize(s));
        *h_nr             = ws->h_scalars[0];
        ws->pend_small    = false;
        ws->pend_const    = false;
        ws->pending       = 1;
        ws->pend_n        = n;
        ws->pend_count    = *h_nr;
        ws->pend_identity = ws->acc_identity;// emit puts back what the accumulator was filled with
        return 0;
    }

}// namespace splacu

using namespace splacu;

extern "C" {

int splacu_v_pack_bits(int dtype, int op_select, uint32_t n, const void* d_v, uint32_t* d_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_v && d_bits, "null pointer");
    const Select sel = make_select(op_select);
    cudaStream_t s   = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        select_bits_kernel<T><<<grid_for(n, kBlock, 8), kBlock, 0, s>>>(sel, static_cast<const T*>(d_v), n, d_bits);
        SPLACU_LAUNCH_CHECK();
        return 0;
    });
}

int splacu_v_unpack_bits(uint32_t n, const uint32_t* d_bits, uint32_t one_bits, uint32_t zero_bits, void* d_out, void* stream) {
    SPLACU_CHECK_INIT();
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_bits && d_out, "null pointer");
    unpack_bits_kernel<<<grid_for(n, kBlock, 8), kBlock, 0, resolve_stream(stream)>>>(d_bits, n, one_bits, zero_bits, static_cast<uint32_t*>(d_out));
    SPLACU_LAUNCH_CHECK();
    return SPLACU_OK;
}

static int vxm_begin_common(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select, uint32_t nv, const uint32_t* d_vi, const void* d_vx,
                            const void* d_mask, splacu_workspace wsh, uint32_t* h_nr, void* stream, bool defer_finish) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/vxm_masked_begin", resolve_stream(stream));
    SPLACU_REQUIRE(handle && wsh && (h_nr || defer_finish), "null handle");
    const Csr* M  = reinterpret_cast<const Csr*>(handle);
    Workspace* ws = reinterpret_cast<Workspace*>(wsh);
    SPLACU_REQUIRE(op_valid_for(dtype, op_mult), "op_mult not defined for dtype");
    SPLACU_REQUIRE(op_valid_for(dtype, op_add), "op_add not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    SPLACU_REQUIRE(ws->pending == 0, "workspace has a pending emit");
    const Select sel = make_select(op_select);
    if (h_nr) *h_nr = 0;
    if (nv == 0 || M->n_cols == 0 || M->nnz == 0 || sel.classes == 0u) return SPLACU_OK;// nothing can be touched
    SPLACU_REQUIRE(d_vi && d_vx, "null frontier pointers");
    SPLACU_REQUIRE(d_mask || !sel.reads_mask, "null mask pointer");
    cudaStream_t s = resolve_stream(stream);
    return dispatch_dtype(dtype, [&](auto tag) {
        using T = decltype(tag);
        return vxm_begin_typed<T>(M, op_mult, op_add, sel, nv, d_vi, static_cast<const T*>(d_vx), static_cast<const T*>(d_mask), ws, h_nr, s, defer_finish);
    });
}

int splacu_vxm_masked_begin(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select,
                            uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                            splacu_workspace wsh, uint32_t* h_nr, void* stream) {
    return vxm_begin_common(handle, dtype, op_mult, op_add, op_select, nv, d_vi, d_vx, d_mask, wsh, h_nr, stream, false);
}

int splacu_vxm_masked_begin_async(splacu_csr handle, int dtype, int op_mult, int op_add, int op_select,
                                  uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                                  splacu_workspace wsh, void* stream) {
    return vxm_begin_common(handle, dtype, op_mult, op_add, op_select, nv, d_vi, d_vx, d_mask, wsh, nullptr, stream, true);
}

int splacu_vxm_masked_begin_finish(splacu_workspace wsh, uint32_t* h_nr, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/vxm_masked_begin_finish", resolve_stream(stream));
    SPLACU_REQUIRE(wsh && h_nr, "null pointer");
    Workspace* ws = reinterpret_cast<Workspace*>(wsh);
    *h_nr         = 0;
    if (ws->pending == 0) return SPLACU_OK;// the async begin short-circuited on an empty product
    SPLACU_REQUIRE(ws->pending == 4, "begin_finish without a matching begin_async");
    return vxm_finish(ws, h_nr, resolve_stream(stream));
}

int splacu_vxm_masked_emit(splacu_workspace wsh, uint32_t* d_ri, void* d_rx, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/vxm_masked_emit", resolve_stream(stream));
    SPLACU_REQUIRE(wsh, "null workspace");
    Workspace* ws = reinterpret_cast<Workspace*>(wsh);
    if (ws->pending == 0) return SPLACU_OK;// begin() short-circuited on an empty product
    SPLACU_REQUIRE(ws->pending == 1, "vxm_masked_emit without matching begin");
    ws->pending = 0;
    if (ws->pend_count == 0) return SPLACU_OK;// nothing touched: scratch still clean
    SPLACU_REQUIRE(d_ri && d_rx, "null output pointers");
    if (ws->pend_small) {
        int end_bit = 1;
        while (end_bit < 32 && ((ws->pend_n - 1u) >> end_bit) != 0u) ++end_bit;
        vxm_emit_small_kernel<<<1, 1024, 0, resolve_stream(stream)>>>(ws->small, ws->pend_count, ws->acc, ws->bitmap, ws->pend_identity, end_bit, d_ri,
                                                                      static_cast<uint32_t*>(d_rx));
        SPLACU_LAUNCH_CHECK();
        return SPLACU_OK;
    }
    if (ws->pend_const)// structure-only: one value for every touched column, the accumulator was never written
        return bitmap_emit(ws, ws->bitmap, ws->pend_n, EMIT_CONST, nullptr, nullptr, ws->pend_value, d_ri, static_cast<uint32_t*>(d_rx), resolve_stream(stream));
    return bitmap_emit(ws, ws->bitmap, ws->pend_n, EMIT_ACC_RESET, ws->acc, nullptr, ws->pend_identity, d_ri, static_cast<uint32_t*>(d_rx),
                       resolve_stream(stream));
}

int splacu_vxm_masked(splacu_csr M, int dtype, int op_mult, int op_add, int op_select,
                      uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask,
                      uint32_t* d_ri, void* d_rx, uint32_t capacity, uint32_t* h_nr, splacu_workspace ws, void* stream) {
    int rc = splacu_vxm_masked_begin(M, dtype, op_mult, op_add, op_select, nv, d_vi, d_vx, d_mask, ws, h_nr, stream);
    if (rc) return rc;
    if (*h_nr > capacity) {
        // still drain the scratch so that the workspace stays usable: emit needs room, so report and reset by fill
        Workspace* w = reinterpret_cast<Workspace*>(ws);
        w->pending   = 0;
        w->acc_clean = false;
        cudaMemsetAsync(w->bitmap, 0, ((size_t) w->pend_n + 31) / 32 * 4, resolve_stream(stream));
        set_error("splacu_vxm_masked: result has %u entries, capacity is %u", *h_nr, capacity);
        return SPLACU_E_CAPACITY;
    }
    return splacu_vxm_masked_emit(ws, d_ri, d_rx, stream);
}

}// extern "C"
