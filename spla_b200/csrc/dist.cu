// dist.cu -- the hot path sharded over the GPUs of one box, behind the C ABI (net-new relative to the single-device
// reference: its accelerator interface only reserves the knob, reference src/core/accelerator.hpp:58-69 set_queues_count).
//
// One process, one host thread, N shards (normally one per device; several shards may share a device, which is how the logic is
// tested on a single GPU). The caller's vectors live on the HOME device (shard 0's device), exactly where the single-device entry
// points expect them, so everything around the two products (v_assign, v_eadd, v_reduce, format conversions) is unchanged:
//
//   pull  rows are cut into contiguous blocks of ~nnz / N entries (nnz-balanced, not n-balanced); shard p holds M[rows_p, :] with its
//         own handle (column classes included). A product = broadcast of v to the shards (ncclBroadcast over NVLink when the shards
//         sit on distinct devices, a plain copy otherwise), the mask / result windows by peer copies, the N local products
//         concurrently on the shards' streams, the home stream waits for all of them. Nothing synchronises with the host.
//   push  column-sharded: shard p holds M[:, cols_p] as a CSR over all rows with column ids rebased to its window (column windows
//         nnz-balanced). Every shard expands the WHOLE frontier against its slice under its window of the mask: the results are
//         disjoint and already ordered by shard, so the output is their concatenation (local ids shifted by the window start). All
//         shards are enqueued before the host waits for any count.
//
// Streams: shard p runs on the backend stream of its device (shards that share a device share its stream). Events order the home
// stream before and after the shards.
#include "common.cuh"
#include "profile.cuh"
#include "ops.cuh"

#include <dlfcn.h>

#include <vector>

namespace splacu {

    namespace {
        constexpr int kBlock = 256;

        // ---- NCCL, bound lazily (single process, ncclCommInitAll) ----------------------------------------
        struct Nccl {
            bool  tried = false, ok = false;
            int (*CommInitAll)(void**, int, const int*)                                           = nullptr;
            int (*CommDestroy)(void*)                                                             = nullptr;
            int (*GroupStart)()                                                                   = nullptr;
            int (*GroupEnd)()                                                                     = nullptr;
            int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t)           = nullptr;
            const char* (*GetErrorString)(int)                                                    = nullptr;
        } g_nccl;
        template<typename F> bool bind(void* lib, const char* name, F& fn) {
            fn = reinterpret_cast<F>(dlsym(lib, name));
            return fn != nullptr;
        }
        const Nccl& nccl() {
            if (!g_nccl.tried) {
                g_nccl.tried = true;
                void* lib    = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
                if (lib)
                    g_nccl.ok = bind(lib, "ncclCommInitAll", g_nccl.CommInitAll) && bind(lib, "ncclCommDestroy", g_nccl.CommDestroy) &&
                                bind(lib, "ncclGroupStart", g_nccl.GroupStart) && bind(lib, "ncclGroupEnd", g_nccl.GroupEnd) &&
                                bind(lib, "ncclBroadcast", g_nccl.Broadcast) && bind(lib, "ncclGetErrorString", g_nccl.GetErrorString);
            }
            return g_nccl;
        }
        constexpr int kNcclUint32 = 3;// ncclUint32 / ncclUint (nccl.h: ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3)

        // device buffer that only grows
        struct Buf {
            void*  p   = nullptr;
            size_t cap = 0;
            int    reserve(size_t bytes) {
                if (bytes <= cap && p) return 0;
                if (p) {
                    SPLACU_CUDA(cudaDeviceSynchronize());
                    cudaFree(p);
                    p = nullptr;
                }
                const size_t want = bytes + bytes / 4 + 1024;
                SPLACU_CUDA(cudaMalloc(&p, want));
                cap = want;
                return 0;
            }
            void release() {
                if (p) cudaFree(p);
                p   = nullptr;
                cap = 0;
            }
        };

        struct DevGuard {
            int prev;
            DevGuard() { cudaGetDevice(&prev); }
            ~DevGuard() { cudaSetDevice(prev); }
        };

        // ---- slice builders (run on the home device) -----------------------------------------------------
        __global__ void __launch_bounds__(kBlock) col_hist_kernel(const uint32_t* __restrict__ Aj, uint32_t nnz, uint32_t* __restrict__ count) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&count[Aj[k]], 1u);
        }
        // a warp per row: entries of the row with c0 <= column < c1
        __global__ void __launch_bounds__(kBlock) slice_count_kernel(const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj, uint32_t n_rows,
                                                                     uint32_t c0, uint32_t c1, uint32_t* __restrict__ cnt) {
            const uint32_t lane = threadIdx.x & 31u, n_warps = (gridDim.x * blockDim.x) >> 5;
            for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
                const uint32_t k1 = Ap[row + 1];
                uint32_t       c  = 0;
                for (uint32_t k = Ap[row] + lane; k < k1; k += 32) {
                    const uint32_t j = Aj[k];
                    c += (j >= c0 && j < c1) ? 1u : 0u;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) cnt[row] = c;
            }
        }
        // a warp per row: stable copy of those entries (column order kept), column ids rebased to c0
        __global__ void __launch_bounds__(kBlock) slice_scatter_kernel(const uint32_t* __restrict__ Ap, const uint32_t* __restrict__ Aj,
                                                                       const uint32_t* __restrict__ Ax, uint32_t n_rows, uint32_t c0, uint32_t c1,
                                                                       const uint32_t* __restrict__ out_Ap, uint32_t* __restrict__ out_Aj,
                                                                       uint32_t* __restrict__ out_Ax) {
            const uint32_t lane = threadIdx.x & 31u, n_warps = (gridDim.x * blockDim.x) >> 5, lt = (1u << lane) - 1u;
            for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
                const uint32_t k0 = Ap[row], k1 = Ap[row + 1];
                uint32_t       dst = out_Ap[row];
                for (uint32_t kb = k0; kb < k1; kb += 32) {
                    const uint32_t k    = kb + lane;
                    uint32_t       j    = 0, a = 0;
                    bool           keep = false;
                    if (k < k1) {
                        j    = Aj[k];
                        keep = j >= c0 && j < c1;
                        if (keep) a = Ax[k];
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const uint32_t q = dst + __popc(m & lt);
                        out_Aj[q]        = j - c0;
                        out_Ax[q]        = a;
                    }
                    dst += __popc(m);
                }
            }
        }
        __global__ void __launch_bounds__(kBlock) rebase_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t base, uint32_t* __restrict__ out) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] - base;
        }
        __global__ void __launch_bounds__(kBlock) shift_kernel(uint32_t* __restrict__ x, uint32_t n, uint32_t add) {
            const uint32_t stride = gridDim.x * blockDim.x;
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] += add;
        }
    }// namespace

    struct Shard {
        int          device = 0;
        cudaStream_t stream = nullptr;
        Workspace*   ws     = nullptr;
        cudaEvent_t  done   = nullptr;
        void*        comm   = nullptr;// ncclComm_t
        // per-call staging (on the shard's device; shard 0 works on the caller's buffers directly)
        Buf v_full, mask_win, r_win, f_vi, f_vx, o_ri, o_rx;
    };

    struct Group {
        int                n = 0;
        std::vector<Shard> shard;
        bool               use_nccl = false;
        cudaEvent_t        ready    = nullptr;// on the home device
        int                home     = 0;
    };

    struct DSlice {
        uint32_t  r0 = 0, r1 = 0, c0 = 0, c1 = 0;
        uint32_t *rAp = nullptr, *rAj = nullptr, *rAx = nullptr;// M[r0:r1, :]   (shard 0: views of the caller's arrays)
        uint32_t *cAp = nullptr, *cAj = nullptr, *cAx = nullptr;// M[:, c0:c1], column ids rebased to c0
        bool      own_rows = false;
        Csr*      rows = nullptr;
        Csr*      cols = nullptr;
        uint32_t  nr   = 0;// result count of the pending push
    };

    struct DCsr {
        Group*              g = nullptr;
        uint32_t            n_rows = 0, n_cols = 0, nnz = 0;
        std::vector<DSlice> s;
        bool                push_built = false;
        const uint32_t *    Ap = nullptr, *Aj = nullptr, *Ax = nullptr;// the caller's arrays on the home device
        int                 pending = 0;
    };

    int csr_build_metadata(Csr* M, cudaStream_t s);

    static int copy_between(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, cudaStream_t s) {
        if (bytes == 0) return 0;
        if (dst_dev == src_dev) SPLACU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
        else SPLACU_CUDA(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, s));
        return 0;
    }

    static int make_local_csr(Csr** out, uint32_t n_rows, uint32_t n_cols, uint32_t nnz, const uint32_t* Ap, const uint32_t* Aj, const uint32_t* Ax,
                              cudaStream_t s) {
        Csr* M    = new Csr();
        M->n_rows = n_rows, M->n_cols = n_cols, M->nnz = nnz;
        M->Ap = Ap, M->Aj = Aj, M->Ax = Ax;
        const int rc = csr_build_metadata(M, s);
        if (rc) {
            splacu_csr_destroy(reinterpret_cast<splacu_csr>(M));
            return rc;
        }
        *out = M;
        return 0;
    }

    // boundaries b[0..n] over a prefix array P[0..m] (P[m] = total) such that every part holds ~total / n
    static void balanced(const std::vector<uint32_t>& P, int n, std::vector<uint32_t>& b) {
        const size_t   m     = P.size() - 1;
        const uint64_t total = P[m];
        b.assign(n + 1, 0);
        b[n] = (uint32_t) m;
        for (int p = 1; p < n; ++p) {
            const uint64_t target = total * (uint64_t) p / (uint64_t) n;
            size_t         lo = 0, hi = m;
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                if (P[mid] < target) lo = mid + 1;
                else hi = mid;
            }
            b[p] = (uint32_t) lo;
            if (b[p] < b[p - 1]) b[p] = b[p - 1];
        }
    }

}// namespace splacu

using namespace splacu;

extern "C" {

int splacu_dist_create(splacu_dist* out, int n_shards, const int* device_ids) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(out && n_shards >= 1 && n_shards <= SPLACU_MAX_PEERS && device_ids, "bad shard list");
    int count = 0;
    SPLACU_CUDA(cudaGetDeviceCount(&count));
    for (int p = 0; p < n_shards; ++p) SPLACU_REQUIRE(device_ids[p] >= 0 && device_ids[p] < count, "device id out of range");
    SPLACU_REQUIRE(device_ids[0] == current_device(), "shard 0 must sit on the home device (splacu_init)");
    DevGuard guard;
    Group*   g = new Group();
    g->n       = n_shards;
    g->home    = device_ids[0];
    g->shard.resize(n_shards);
    bool distinct = n_shards > 1;
    for (int p = 0; p < n_shards; ++p)
        for (int q = 0; q < p; ++q)
            if (device_ids[p] == device_ids[q]) distinct = false;
    int rc = 0;
    for (int p = 0; p < n_shards && !rc; ++p) {
        Shard& sh = g->shard[p];
        sh.device = device_ids[p];
        if ((rc = ensure_device(sh.device))) break;
        cudaSetDevice(sh.device);
        sh.stream = resolve_stream(nullptr);
        for (int q = 0; q < n_shards; ++q)
            if (device_ids[q] != sh.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, sh.device, device_ids[q]);
                if (can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[q], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                    cudaGetLastError();
                }
            }
        splacu_workspace w = nullptr;
        if (!rc) rc = splacu_workspace_create(&w);
        sh.ws = reinterpret_cast<Workspace*>(w);
        if (!rc && cudaEventCreateWithFlags(&sh.done, cudaEventDisableTiming) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
    }
    cudaSetDevice(g->home);
    if (!rc && cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
    if (!rc && distinct && nccl().ok) {
        std::vector<void*> comms(n_shards, nullptr);
        const int          e = nccl().CommInitAll(comms.data(), n_shards, device_ids);
        if (e == 0) {
            for (int p = 0; p < n_shards; ++p) g->shard[p].comm = comms[p];
            g->use_nccl = true;
        }// else: peer copies
        cudaSetDevice(g->home);
    }
    if (rc) {
        splacu_dist_destroy(reinterpret_cast<splacu_dist>(g));
        return rc;
    }
    *out = reinterpret_cast<splacu_dist>(g);
    return SPLACU_OK;
}

int splacu_dist_destroy(splacu_dist handle) {
    if (!handle) return SPLACU_OK;
    Group*   g = reinterpret_cast<Group*>(handle);
    DevGuard guard;
    for (Shard& sh : g->shard) {
        cudaSetDevice(sh.device);
        if (sh.stream) cudaStreamSynchronize(sh.stream);
        if (sh.comm && nccl().ok) nccl().CommDestroy(sh.comm);
        if (sh.ws) splacu_workspace_destroy(reinterpret_cast<splacu_workspace>(sh.ws));
        if (sh.done) cudaEventDestroy(sh.done);
        for (Buf* b : {&sh.v_full, &sh.mask_win, &sh.r_win, &sh.f_vi, &sh.f_vx, &sh.o_ri, &sh.o_rx}) b->release();
    }
    cudaSetDevice(g->home);
    if (g->ready) cudaEventDestroy(g->ready);
    cudaGetLastError();
    delete g;
    return SPLACU_OK;
}

int splacu_dist_info(splacu_dist handle, int* n_shards, int* uses_nccl) {
    SPLACU_REQUIRE(handle, "null group");
    Group* g = reinterpret_cast<Group*>(handle);
    if (n_shards) *n_shards = g->n;
    if (uses_nccl) *uses_nccl = g->use_nccl ? 1 : 0;
    return SPLACU_OK;
}

int splacu_dcsr_destroy(splacu_dcsr handle) {
    if (!handle) return SPLACU_OK;
    DCsr*    M = reinterpret_cast<DCsr*>(handle);
    DevGuard guard;
    for (int p = 0; p < (int) M->s.size(); ++p) {
        DSlice& sl = M->s[p];
        cudaSetDevice(M->g->shard[p].device);
        cudaStreamSynchronize(M->g->shard[p].stream);
        if (sl.rows) splacu_csr_destroy(reinterpret_cast<splacu_csr>(sl.rows));
        if (sl.cols) splacu_csr_destroy(reinterpret_cast<splacu_csr>(sl.cols));
        if (sl.own_rows) {
            cudaFree(sl.rAp);
            cudaFree(sl.rAj);
            cudaFree(sl.rAx);
        }
        cudaFree(sl.cAp);
        cudaFree(sl.cAj);
        cudaFree(sl.cAx);
    }
    cudaGetLastError();
    delete M;
    return SPLACU_OK;
}

int splacu_dcsr_create(splacu_dcsr* out, splacu_dist group, uint32_t n_rows, uint32_t n_cols, uint32_t nnz, const uint32_t* d_Ap, const uint32_t* d_Aj,
                       const void* d_Ax, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dcsr_create", resolve_stream(stream));
    SPLACU_REQUIRE(out && group && d_Ap, "null pointer");
    SPLACU_REQUIRE(nnz == 0 || (d_Aj && d_Ax), "null Aj/Ax");
    Group* g = reinterpret_cast<Group*>(group);
    SPLACU_REQUIRE(current_device() == g->home, "the home device of the group must be current");
    DevGuard     guard;
    cudaStream_t s0 = resolve_stream(stream);
    DCsr*        M  = new DCsr();
    M->g = g, M->n_rows = n_rows, M->n_cols = n_cols, M->nnz = nnz;
    M->Ap = d_Ap, M->Aj = d_Aj, M->Ax = static_cast<const uint32_t*>(d_Ax);
    M->s.resize(g->n);
    int rc = 0;
#define D_CUDA(expr)                                                          \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess && !rc) rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
    // ---- row windows, nnz-balanced on Ap ----
    std::vector<uint32_t> hAp((size_t) n_rows + 1), rb;
    D_CUDA(cudaMemcpyAsync(hAp.data(), d_Ap, ((size_t) n_rows + 1) * 4, cudaMemcpyDeviceToHost, s0));
    D_CUDA(cudaStreamSynchronize(s0));
    if (!rc) balanced(hAp, g->n, rb);
    for (int p = 0; p < g->n && !rc; ++p) {
        DSlice&  sl = M->s[p];
        Shard&   sh = g->shard[p];
        sl.r0 = rb[p], sl.r1 = rb[p + 1];
        const uint32_t rows = sl.r1 - sl.r0, k0 = hAp[sl.r0], k1 = hAp[sl.r1], cnt = k1 - k0;
        if (p == 0 && k0 == 0) {// shard 0: its rows are a prefix of the caller's arrays -- no copy
            sl.rAp = const_cast<uint32_t*>(d_Ap), sl.rAj = const_cast<uint32_t*>(d_Aj), sl.rAx = const_cast<uint32_t*>(M->Ax);
        } else {
            // rebase the row extents on the home device, then move the slice to its shard
            uint32_t* tmp = nullptr;
            D_CUDA(cudaMalloc(&tmp, ((size_t) rows + 1) * 4));
            if (!rc) {
                rebase_kernel<<<grid_for((size_t) rows + 1, kBlock, 8), kBlock, 0, s0>>>(d_Ap + sl.r0, rows + 1, k0, tmp);
                count_launch();
            }
            D_CUDA(cudaStreamSynchronize(s0));
            cudaSetDevice(sh.device);
            sl.own_rows = true;
            D_CUDA(cudaMalloc(&sl.rAp, ((size_t) rows + 1) * 4));
            D_CUDA(cudaMalloc(&sl.rAj, ((size_t) cnt + 4) * 4));
            D_CUDA(cudaMalloc(&sl.rAx, ((size_t) cnt + 4) * 4));
            if (!rc) rc = copy_between(sl.rAp, sh.device, tmp, g->home, ((size_t) rows + 1) * 4, sh.stream);
            if (!rc) rc = copy_between(sl.rAj, sh.device, d_Aj + k0, g->home, (size_t) cnt * 4, sh.stream);
            if (!rc) rc = copy_between(sl.rAx, sh.device, M->Ax + k0, g->home, (size_t) cnt * 4, sh.stream);
            D_CUDA(cudaStreamSynchronize(sh.stream));
            cudaSetDevice(g->home);
            cudaFree(tmp);
        }
        if (!rc) {
            cudaSetDevice(sh.device);
            rc = make_local_csr(&sl.rows, rows, n_cols, cnt, sl.rAp, sl.rAj, sl.rAx, sh.stream);
            cudaSetDevice(g->home);
        }
    }
#undef D_CUDA
    if (rc) {
        splacu_dcsr_destroy(reinterpret_cast<splacu_dcsr>(M));
        return rc;
    }
    *out = reinterpret_cast<splacu_dcsr>(M);
    return SPLACU_OK;
}

// column slices for the push direction, built at the first push (a pull-only user never pays for them)
static int build_push_slices(DCsr* M, cudaStream_t s0) {
    Group* g = M->g;
    int    rc = 0;
#define D_CUDA(expr)                                                          \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess && !rc) rc = ::splacu::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
    // column windows, nnz-balanced on the column histogram
    uint32_t* d_hist = nullptr;
    D_CUDA(cudaMalloc(&d_hist, ((size_t) M->n_cols + 1) * 4));
    D_CUDA(cudaMemsetAsync(d_hist, 0, ((size_t) M->n_cols + 1) * 4, s0));
    if (!rc && M->nnz) {
        col_hist_kernel<<<grid_for(M->nnz, kBlock, 8), kBlock, 0, s0>>>(M->Aj, M->nnz, d_hist);
        count_launch();
    }
    std::vector<uint32_t> hist((size_t) M->n_cols + 1), cb;
    D_CUDA(cudaMemcpyAsync(hist.data(), d_hist, ((size_t) M->n_cols + 1) * 4, cudaMemcpyDeviceToHost, s0));
    D_CUDA(cudaStreamSynchronize(s0));
    cudaFree(d_hist);
    if (rc) return rc;
    uint32_t run = 0;
    for (size_t j = 0; j <= M->n_cols; ++j) {// exclusive prefix in place
        const uint32_t c = hist[j];
        hist[j]          = run;
        run += c;
    }
    balanced(hist, g->n, cb);
    uint32_t* cnt = nullptr;
    D_CUDA(cudaMalloc(&cnt, ((size_t) M->n_rows + 1) * 4));
    Workspace* ws0 = g->shard[0].ws;
    for (int p = 0; p < g->n && !rc; ++p) {
        DSlice& sl = M->s[p];
        Shard&  sh = g->shard[p];
        sl.c0 = cb[p], sl.c1 = cb[p + 1];
        const uint32_t nnz_p = hist[sl.c1] - hist[sl.c0];
        // count -> scan -> scatter on the home device (it reads the whole matrix once per shard), then move the slice
        uint32_t *tAp = nullptr, *tAj = nullptr, *tAx = nullptr;
        D_CUDA(cudaMalloc(&tAp, ((size_t) M->n_rows + 1) * 4));
        D_CUDA(cudaMalloc(&tAj, ((size_t) nnz_p + 4) * 4));
        D_CUDA(cudaMalloc(&tAx, ((size_t) nnz_p + 4) * 4));
        D_CUDA(cudaMemsetAsync(cnt, 0, ((size_t) M->n_rows + 1) * 4, s0));
        if (!rc) {
            slice_count_kernel<<<grid_for((size_t) M->n_rows * 32, kBlock, 8), kBlock, 0, s0>>>(M->Ap, M->Aj, M->n_rows, sl.c0, sl.c1, cnt);
            count_launch();
            rc = scan_exclusive_u32(ws0, cnt, tAp, M->n_rows + 1, nullptr, s0);
        }
        if (!rc) {
            slice_scatter_kernel<<<grid_for((size_t) M->n_rows * 32, kBlock, 8), kBlock, 0, s0>>>(M->Ap, M->Aj, M->Ax, M->n_rows, sl.c0, sl.c1, tAp, tAj, tAx);
            count_launch();
        }
        D_CUDA(cudaStreamSynchronize(s0));
        if (sh.device == g->home) {
            sl.cAp = tAp, sl.cAj = tAj, sl.cAx = tAx;
        } else {
            cudaSetDevice(sh.device);
            D_CUDA(cudaMalloc(&sl.cAp, ((size_t) M->n_rows + 1) * 4));
            D_CUDA(cudaMalloc(&sl.cAj, ((size_t) nnz_p + 4) * 4));
            D_CUDA(cudaMalloc(&sl.cAx, ((size_t) nnz_p + 4) * 4));
            if (!rc) rc = copy_between(sl.cAp, sh.device, tAp, g->home, ((size_t) M->n_rows + 1) * 4, sh.stream);
            if (!rc) rc = copy_between(sl.cAj, sh.device, tAj, g->home, (size_t) nnz_p * 4, sh.stream);
            if (!rc) rc = copy_between(sl.cAx, sh.device, tAx, g->home, (size_t) nnz_p * 4, sh.stream);
            D_CUDA(cudaStreamSynchronize(sh.stream));
            cudaSetDevice(g->home);
            cudaFree(tAp);
            cudaFree(tAj);
            cudaFree(tAx);
        }
        if (!rc) {
            cudaSetDevice(sh.device);
            // the push never runs the pull's column classes: build the slice handle without them
            const int64_t hub = get_option(OPT_MXV_HUB);
            splacu_set_option("mxv_hub", 0);
            rc = make_local_csr(&sl.cols, M->n_rows, sl.c1 - sl.c0, nnz_p, sl.cAp, sl.cAj, sl.cAx, sh.stream);
            splacu_set_option("mxv_hub", hub);
            cudaSetDevice(g->home);
        }
    }
    cudaFree(cnt);
#undef D_CUDA
    if (!rc) M->push_built = true;
    return rc;
}

int splacu_dcsr_bounds(splacu_dcsr handle, int* n_shards, uint32_t* row_bounds, uint32_t* col_bounds) {
    SPLACU_REQUIRE(handle, "null handle");
    DCsr* M = reinterpret_cast<DCsr*>(handle);
    if (n_shards) *n_shards = M->g->n;
    for (int p = 0; p < M->g->n; ++p) {
        if (row_bounds) row_bounds[p] = M->s[p].r0, row_bounds[p + 1] = M->s[p].r1;
        if (col_bounds) col_bounds[p] = M->s[p].c0, col_bounds[p + 1] = M->s[p].c1;
    }
    return SPLACU_OK;
}

int splacu_dist_mxv_masked(splacu_dcsr handle, int dtype, int op_mult, int op_add, int op_select, const void* d_v, const void* d_mask, void* d_r,
                           uint32_t init_bits, int early_exit, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dist_mxv_masked", resolve_stream(stream));
    SPLACU_REQUIRE(handle, "null handle");
    DCsr*  M = reinterpret_cast<DCsr*>(handle);
    Group* g = M->g;
    SPLACU_REQUIRE(op_valid_for(dtype, op_mult) && op_valid_for(dtype, op_add), "op not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    if (M->n_rows == 0) return SPLACU_OK;
    const Select sel = make_select(op_select);
    SPLACU_REQUIRE(d_r && (d_mask || !sel.reads_mask) && (d_v || M->nnz == 0), "null pointer");
    SPLACU_REQUIRE(current_device() == g->home, "the home device of the group must be current");
    DevGuard     guard;
    cudaStream_t s0 = resolve_stream(stream);
    const size_t vb = (size_t) M->n_cols * 4;
    int          rc = 0;
    SPLACU_CUDA(cudaEventRecord(g->ready, s0));
    // 1. the input vector on every shard
    for (int p = 0; p < g->n; ++p) {
        Shard& sh = g->shard[p];
        cudaSetDevice(sh.device);
        if (sh.stream != s0) SPLACU_CUDA(cudaStreamWaitEvent(sh.stream, g->ready, 0));
        if (sh.device != g->home && (rc = sh.v_full.reserve(vb))) return rc;
    }
    if (g->use_nccl && M->nnz) {
        nccl().GroupStart();
        for (int p = 0; p < g->n; ++p) {
            Shard& sh = g->shard[p];
            void*  dst = sh.device == g->home ? const_cast<void*>(d_v) : sh.v_full.p;// in place at the root
            const int e = nccl().Broadcast(d_v, dst, M->n_cols, kNcclUint32, 0, sh.comm, sh.stream);
            if (e) {
                nccl().GroupEnd();
                set_error("ncclBroadcast failed: %s", nccl().GetErrorString(e));
                return SPLACU_E_INVALID;
            }
        }
        const int e = nccl().GroupEnd();
        if (e) {
            set_error("ncclGroupEnd failed: %s", nccl().GetErrorString(e));
            return SPLACU_E_INVALID;
        }
    } else if (M->nnz) {
        for (int p = 0; p < g->n; ++p) {
            Shard& sh = g->shard[p];
            if (sh.device == g->home) continue;
            bool seen = false;// shards that share a device share the copy
            for (int q = 0; q < p; ++q) seen |= g->shard[q].device == sh.device;
            if (seen) continue;
            cudaSetDevice(sh.device);
            if ((rc = copy_between(sh.v_full.p, sh.device, d_v, g->home, vb, sh.stream))) return rc;
        }
    }
    // 2. the N local products, each on its shard's stream
    for (int p = 0; p < g->n; ++p) {
        Shard&         sh   = g->shard[p];
        DSlice&        sl   = M->s[p];
        const uint32_t rows = sl.r1 - sl.r0;
        if (rows == 0) continue;
        cudaSetDevice(sh.device);
        const bool  local = sh.device == g->home;
        const void* v_p   = d_v;
        if (!local) {
            v_p = sh.v_full.p;
            for (int q = 0; q < p; ++q)
                if (g->shard[q].device == sh.device) v_p = g->shard[q].v_full.p ? g->shard[q].v_full.p : v_p;
        }
        const void* m_p = d_mask ? static_cast<const uint32_t*>(d_mask) + sl.r0 : nullptr;
        void*       r_p = static_cast<uint32_t*>(d_r) + sl.r0;
        if (!local) {
            if ((rc = sh.r_win.reserve((size_t) rows * 4))) return rc;
            r_p = sh.r_win.p;
            if (sel.reads_mask) {
                if ((rc = sh.mask_win.reserve((size_t) rows * 4))) return rc;
                if ((rc = copy_between(sh.mask_win.p, sh.device, m_p, g->home, (size_t) rows * 4, sh.stream))) return rc;
                m_p = sh.mask_win.p;
            }
        }
        if ((rc = splacu_mxv_masked(reinterpret_cast<splacu_csr>(sl.rows), dtype, op_mult, op_add, op_select, v_p, m_p, r_p, init_bits, early_exit, sh.stream)))
            return rc;
        if (!local && (rc = copy_between(static_cast<uint32_t*>(d_r) + sl.r0, g->home, r_p, sh.device, (size_t) rows * 4, sh.stream))) return rc;
        if (sh.stream != s0) SPLACU_CUDA(cudaEventRecord(sh.done, sh.stream));
    }
    // 3. the home stream continues when every shard is done
    cudaSetDevice(g->home);
    for (int p = 0; p < g->n; ++p) {
        Shard& sh = g->shard[p];
        if (sh.stream != s0 && M->s[p].r1 > M->s[p].r0) SPLACU_CUDA(cudaStreamWaitEvent(s0, sh.done, 0));
    }
    return SPLACU_OK;
}

int splacu_dist_vxm_masked_begin(splacu_dcsr handle, int dtype, int op_mult, int op_add, int op_select, uint32_t nv, const uint32_t* d_vi, const void* d_vx,
                                 const void* d_mask, uint32_t* h_nr, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dist_vxm_masked_begin", resolve_stream(stream));
    SPLACU_REQUIRE(handle && h_nr, "null handle");
    DCsr*  M = reinterpret_cast<DCsr*>(handle);
    Group* g = M->g;
    SPLACU_REQUIRE(op_valid_for(dtype, op_mult) && op_valid_for(dtype, op_add), "op not defined for dtype");
    SPLACU_REQUIRE(op_select >= 0 && op_select < SPLACU_SELOP_COUNT, "unknown op_select");
    SPLACU_REQUIRE(M->pending == 0, "the matrix has a pending emit");
    const Select sel = make_select(op_select);
    *h_nr            = 0;
    for (DSlice& sl : M->s) sl.nr = 0;
    if (nv == 0 || M->n_cols == 0 || M->nnz == 0 || sel.classes == 0u) return SPLACU_OK;
    SPLACU_REQUIRE(d_vi && d_vx && (d_mask || !sel.reads_mask), "null pointer");
    SPLACU_REQUIRE(current_device() == g->home, "the home device of the group must be current");
    DevGuard     guard;
    cudaStream_t s0 = resolve_stream(stream);
    int          rc = 0;
    if (!M->push_built && (rc = build_push_slices(M, s0))) return rc;
    SPLACU_CUDA(cudaEventRecord(g->ready, s0));
    // enqueue on every shard: frontier + mask window over, expand, count
    for (int p = 0; p < g->n; ++p) {
        Shard&         sh = g->shard[p];
        DSlice&        sl = M->s[p];
        const uint32_t w  = sl.c1 - sl.c0;
        if (w == 0 || sl.cols->nnz == 0) continue;
        cudaSetDevice(sh.device);
        if (sh.stream != s0) SPLACU_CUDA(cudaStreamWaitEvent(sh.stream, g->ready, 0));
        const bool      local = sh.device == g->home;
        const uint32_t* vi_p  = d_vi;
        const void*     vx_p  = d_vx;
        const void*     m_p   = d_mask ? static_cast<const uint32_t*>(d_mask) + sl.c0 : nullptr;
        if (!local) {
            if ((rc = sh.f_vi.reserve((size_t) nv * 4)) || (rc = sh.f_vx.reserve((size_t) nv * 4))) return rc;
            if ((rc = copy_between(sh.f_vi.p, sh.device, d_vi, g->home, (size_t) nv * 4, sh.stream))) return rc;
            if ((rc = copy_between(sh.f_vx.p, sh.device, d_vx, g->home, (size_t) nv * 4, sh.stream))) return rc;
            vi_p = static_cast<const uint32_t*>(sh.f_vi.p), vx_p = sh.f_vx.p;
            if (sel.reads_mask) {
                if ((rc = sh.mask_win.reserve((size_t) w * 4))) return rc;
                if ((rc = copy_between(sh.mask_win.p, sh.device, m_p, g->home, (size_t) w * 4, sh.stream))) return rc;
                m_p = sh.mask_win.p;
            }
        }
        // shards that share a device share its workspace-bearing stream but not the workspace: each shard has its own
        if ((rc = splacu_vxm_masked_begin_async(reinterpret_cast<splacu_csr>(sl.cols), dtype, op_mult, op_add, op_select, nv, vi_p, vx_p, m_p,
                                                reinterpret_cast<splacu_workspace>(sh.ws), sh.stream)))
            return rc;
    }
    // only now wait: one count per shard
    uint32_t total = 0;
    for (int p = 0; p < g->n; ++p) {
        Shard& sh = g->shard[p];
        cudaSetDevice(sh.device);
        uint32_t c = 0;
        if ((rc = splacu_vxm_masked_begin_finish(reinterpret_cast<splacu_workspace>(sh.ws), &c, sh.stream))) return rc;
        M->s[p].nr = c;
        total += c;
    }
    *h_nr      = total;
    M->pending = 1;
    return SPLACU_OK;
}

int splacu_dist_vxm_masked_emit(splacu_dcsr handle, uint32_t* d_ri, void* d_rx, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/dist_vxm_masked_emit", resolve_stream(stream));
    SPLACU_REQUIRE(handle, "null handle");
    DCsr*  M = reinterpret_cast<DCsr*>(handle);
    Group* g = M->g;
    if (M->pending == 0) return SPLACU_OK;
    M->pending = 0;
    SPLACU_REQUIRE(current_device() == g->home, "the home device of the group must be current");
    DevGuard     guard;
    cudaStream_t s0     = resolve_stream(stream);
    uint32_t     offset = 0;
    int          rc     = 0;
    for (int p = 0; p < g->n; ++p) {
        Shard&         sh = g->shard[p];
        DSlice&        sl = M->s[p];
        const uint32_t c  = sl.nr;
        cudaSetDevice(sh.device);
        if (c == 0) {// nothing touched on this shard: its scratch is clean, but a begun call must still be closed
            if ((rc = splacu_vxm_masked_emit(reinterpret_cast<splacu_workspace>(sh.ws), nullptr, nullptr, sh.stream))) return rc;
            continue;
        }
        SPLACU_REQUIRE(d_ri && d_rx, "null output pointers");
        const bool local = sh.device == g->home;
        uint32_t*  ri_p  = d_ri + offset;
        void*      rx_p  = static_cast<uint32_t*>(d_rx) + offset;
        if (!local) {
            if ((rc = sh.o_ri.reserve((size_t) c * 4)) || (rc = sh.o_rx.reserve((size_t) c * 4))) return rc;
            ri_p = static_cast<uint32_t*>(sh.o_ri.p), rx_p = sh.o_rx.p;
        }
        if ((rc = splacu_vxm_masked_emit(reinterpret_cast<splacu_workspace>(sh.ws), ri_p, rx_p, sh.stream))) return rc;
        if (sl.c0) {
            shift_kernel<<<grid_for(c, kBlock, 8), kBlock, 0, sh.stream>>>(ri_p, c, sl.c0);// local column ids -> global
            SPLACU_LAUNCH_CHECK();
        }
        if (!local) {
            if ((rc = copy_between(d_ri + offset, g->home, ri_p, sh.device, (size_t) c * 4, sh.stream))) return rc;
            if ((rc = copy_between(static_cast<uint32_t*>(d_rx) + offset, g->home, rx_p, sh.device, (size_t) c * 4, sh.stream))) return rc;
        }
        if (sh.stream != s0) SPLACU_CUDA(cudaEventRecord(sh.done, sh.stream));
        offset += c;
    }
    cudaSetDevice(g->home);
    for (int p = 0; p < g->n; ++p) {
        Shard& sh = g->shard[p];
        if (sh.stream != s0 && M->s[p].nr) SPLACU_CUDA(cudaStreamWaitEvent(s0, sh.done, 0));
    }
    return SPLACU_OK;
}

}// extern "C"
