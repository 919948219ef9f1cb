// runtime.cu -- device runtime of the splacu backend: init, stream, memory, error reporting,
// device CSR handle and workspace. Replaces the OpenCL runtime of the reference
// (src/opencl/cl_accelerator.cpp:84-201, cl_alloc_*.cpp, cl_counter.cpp) for the hot path.
#include "common.cuh"
#include "profile.cuh"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>

namespace splacu {

    // One backend stream per device. g_device is the HOME device (splacu_init): where the caller's vectors live and where every
    // single-device entry point runs; a multi-GPU group (dist.cu) initialises further devices through ensure_device().
    static constexpr int  kMaxDevices   = 64;
    bool                  g_initialised = false;
    static int            g_device      = -1;
    static int            g_sm_count    = 148;
    static cudaStream_t   g_streams[kMaxDevices] = {};
#define g_stream g_streams[g_device >= 0 ? g_device : 0]
    static char           g_device_name[256] = "none";
    static std::atomic<uint64_t> g_launches{0};
    static thread_local char     g_error[1024] = "";

    void set_error(const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_error, sizeof(g_error), fmt, ap);
        va_end(ap);
    }

    int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
        set_error("CUDA error %d (%s) at %s:%d: %s", (int) e, cudaGetErrorName(e), file, line, what);
        cudaGetLastError();// clear sticky launch-config errors
        return (int) e;
    }

    static int64_t g_options[OPT_COUNT] = {/*mxv_hub: 0 off, 1 auto (column-class phases), 2 force the single-pass hub cache, 3 force phases*/ 1, /*mxv_hub_min_count*/ 16, /*mxv_hub_total*/ 1048576, /*mxv_hub_smem*/ 16384, /*mxv_l2_persist*/ 0, /*vxm_selbits*/ 1, /*mxv_phases*/ 4, /*mxv_phase_slots*/ 45056, /*mxv_phase_only: profiling aid, p + 1 runs class p alone (wrong result)*/ 0, /*mxv_seg: classes in the segmented-tile format*/ 1, /*mxv_seg_min_density: percent of rows a mask must select for the class passes*/ 45, /*mxv_tail_range_log2: the tail class is split into windows of 2^k columns of v (one pass each, L2-resident gathers)*/ 24, /*small_front: single-CTA offset / emit / filter kernels for fronts of <= 8192 entries (launch-latency paths)*/ 1, /*vxm_struct: structure-only push when every product is provably the same value and the add is idempotent*/ 1, /*mxv_red: hub class passes of a PLUS semiring add their segment sums onto r with L2 reductions (red.add) instead of load + add + store; FLOAT: red.global.add.f32 flushes subnormal sums to zero (PTX ISA), 0 keeps them*/ 1, /*mxv_row_classes: row classes of the tail (the tail entries of the rows with the most of them, scattered into a shared-memory table of partial results while v streams)*/ 1, /*mxv_row_min_count: tail entries a row needs to get a slot*/ 64, /*mxv_fixup_merge: 0 one fix-up launch per class, 1 the fix-ups of all classes in one cooperative launch after the last class pass (grid barrier between classes, fixed order), 2 two plain launches: all chain sums at once, then one thread per row adds them in class order (same order as 1: bit-identical)*/ 2, /*mxv_row_min_nnz: entries a row class must hold to be built (its fixed costs against ~2 ps saved per entry)*/ 25165824, /*mxv_bank_order: at handle build, permute the entries of every row run inside a lane of a hub class against shared-memory bank conflicts of the table gathers*/ 1, /*mxv_reserve_sms: SMs the persistent class kernels of the pull product leave free (for the kernels of a collective that runs beside them in a multi-GPU step)*/ 0, /*mxv_pdl: the class passes are launched with programmatic stream serialization: the CTAs of the next pass load their hub table and first tile while the last CTAs of the previous pass finish, and wait (griddepcontrol.wait) before they touch r*/ 1};
    static const char* const g_option_names[OPT_COUNT] = {"mxv_hub", "mxv_hub_min_count", "mxv_hub_total", "mxv_hub_smem", "mxv_l2_persist", "vxm_selbits", "mxv_phases", "mxv_phase_slots", "mxv_phase_only", "mxv_seg", "mxv_seg_min_density", "mxv_tail_range_log2", "small_front", "vxm_struct", "mxv_red", "mxv_row_classes", "mxv_row_min_count", "mxv_fixup_merge", "mxv_row_min_nnz", "mxv_bank_order", "mxv_reserve_sms", "mxv_pdl"};
    int64_t get_option(int opt) { return g_options[opt]; }

    void count_launch(int n) { g_launches.fetch_add((uint64_t) n, std::memory_order_relaxed); }
    static std::atomic<uint64_t> g_jit_compiles{0};
    void count_jit_compile() { g_jit_compiles.fetch_add(1, std::memory_order_relaxed); }
    uint64_t jit_compiles() { return g_jit_compiles.load(); }

    int current_device() {
        int dev = 0;
        cudaGetDevice(&dev);
        return dev;
    }
    cudaStream_t resolve_stream(void* stream) {
        if (stream) return (cudaStream_t) stream;
        const int dev = current_device();
        return (dev >= 0 && dev < kMaxDevices && g_streams[dev]) ? g_streams[dev] : g_stream;
    }
    int sm_count() { return g_sm_count; }
    int ensure_device(int device) {
        SPLACU_REQUIRE(device >= 0 && device < kMaxDevices, "device index out of range");
        if (g_streams[device]) return 0;
        const int prev = current_device();
        SPLACU_CUDA(cudaSetDevice(device));
        cudaError_t e = cudaStreamCreateWithFlags(&g_streams[device], cudaStreamNonBlocking);
        cudaSetDevice(prev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
        return 0;
    }

    int ws_reserve_vector(Workspace* ws, uint32_t n, cudaStream_t s) {
        if (n <= ws->cap_n && ws->acc) return 0;
        SPLACU_CUDA(cudaStreamSynchronize(s));
        if (ws->acc) cudaFree(ws->acc);
        if (ws->bitmap) cudaFree(ws->bitmap);
        ws->acc = ws->bitmap = nullptr;
        size_t cap   = (size_t) n + (n >> 2) + 1024;// head-room so that growing vectors do not thrash
        if (cap > 0xffffffffull) cap = 0xffffffffull;
        size_t words = (cap + 31) / 32 + 1;
        SPLACU_CUDA(cudaMalloc(&ws->acc, cap * sizeof(uint32_t)));
        SPLACU_CUDA(cudaMalloc(&ws->bitmap, words * sizeof(uint32_t)));
        SPLACU_CUDA(cudaMemsetAsync(ws->bitmap, 0, words * sizeof(uint32_t), s));
        ws->cap_n     = (uint32_t) cap;
        ws->acc_clean = false;
        return 0;
    }

    int ws_reserve_selbits(Workspace* ws, uint32_t n) {
        if (n <= ws->cap_sel && ws->sel_bits) return 0;
        if (ws->sel_bits) {
            SPLACU_CUDA(cudaDeviceSynchronize());
            cudaFree(ws->sel_bits);
            ws->sel_bits = nullptr;
        }
        const size_t cap = (size_t) n + (n >> 2) + 1024;
        SPLACU_CUDA(cudaMalloc(&ws->sel_bits, ((cap + 31) / 32 + 1) * sizeof(uint32_t)));
        ws->cap_sel = (uint32_t) (cap > 0xffffffffull ? 0xffffffffull : cap);
        return 0;
    }

    int ws_reserve_blocks(Workspace* ws, uint32_t n_blocks) {
        if (n_blocks <= ws->cap_blocks && ws->block_sums) return 0;
        if (ws->block_sums) {
            SPLACU_CUDA(cudaDeviceSynchronize());
            cudaFree(ws->block_sums);
            ws->block_sums = nullptr;
        }
        size_t cap = (size_t) n_blocks * 2 + 1024;
        SPLACU_CUDA(cudaMalloc(&ws->block_sums, cap * sizeof(uint32_t)));
        ws->cap_blocks = (uint32_t) cap;
        return 0;
    }

    int ws_reserve_pairs(Workspace* ws, size_t n_pairs, size_t n_offsets) {
        if (n_pairs > ws->cap_pairs) {
            SPLACU_CUDA(cudaDeviceSynchronize());
            cudaFree(ws->keys_a); cudaFree(ws->keys_b); cudaFree(ws->vals_a); cudaFree(ws->vals_b);
            ws->keys_a = ws->keys_b = ws->vals_a = ws->vals_b = nullptr;
            size_t cap = n_pairs + n_pairs / 4 + 1024;
            SPLACU_CUDA(cudaMalloc(&ws->keys_a, cap * 4));
            SPLACU_CUDA(cudaMalloc(&ws->keys_b, cap * 4));
            SPLACU_CUDA(cudaMalloc(&ws->vals_a, cap * 4));
            SPLACU_CUDA(cudaMalloc(&ws->vals_b, cap * 4));
            ws->cap_pairs = cap;
        }
        if (n_offsets > ws->cap_offsets) {
            SPLACU_CUDA(cudaDeviceSynchronize());
            cudaFree(ws->offsets);
            ws->offsets = nullptr;
            size_t cap  = n_offsets + n_offsets / 4 + 1024;
            SPLACU_CUDA(cudaMalloc(&ws->offsets, cap * 2 * 4));// [0, cap): frontier offsets, [cap, 2 cap): row starts Ap[vi[t]]
            ws->cap_offsets = cap;
        }
        return 0;
    }

    int ws_reserve_chunks(Workspace* ws, size_t n_chunks) {
        if (n_chunks <= ws->cap_chunks && ws->chunk_first) return 0;
        if (ws->chunk_first) {
            SPLACU_CUDA(cudaDeviceSynchronize());
            cudaFree(ws->chunk_first);
            ws->chunk_first = nullptr;
        }
        const size_t cap = n_chunks + n_chunks / 4 + 64;
        SPLACU_CUDA(cudaMalloc(&ws->chunk_first, cap * 4));
        ws->cap_chunks = cap;
        return 0;
    }

    // defined in mxv_pull.cu
    int csr_build_metadata(Csr* M, cudaStream_t s);

}// namespace splacu

using namespace splacu;

extern "C" {

int splacu_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("splacu_init: no CUDA device available (%s); this backend has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return SPLACU_E_NOT_INIT;
    }
    if (device < 0 || device >= count) {
        set_error("splacu_init: device %d out of range [0, %d)", device, count);
        return SPLACU_E_INVALID;
    }
    SPLACU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SPLACU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (g_initialised && g_device == device) return SPLACU_OK;
    // switching the home device: the stream of the old one stays alive (handles and workspaces created there keep working
    // as long as their device is current when they are used); per-device state -- function attributes, JIT modules -- is keyed
    // by device
    g_device = device;
    if (!g_stream) SPLACU_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_sm_count = prop.multiProcessorCount;
    snprintf(g_device_name, sizeof(g_device_name), "%s (sm_%d%d, %d SMs, %.0f GB)", prop.name, prop.major, prop.minor,
             prop.multiProcessorCount, (double) prop.totalGlobalMem / 1e9);
    g_initialised = true;
    // tuning knobs from the environment, e.g. SPLACU_OPTIONS="mxv_red=1,mxv_phases=3" (the same names splacu_set_option takes)
    static bool env_done = false;
    if (!env_done) {
        env_done = true;
        if (const char* env = getenv("SPLACU_OPTIONS")) {
            std::string all(env);
            size_t      pos = 0;
            while (pos < all.size()) {
                size_t end = all.find(',', pos);
                if (end == std::string::npos) end = all.size();
                const std::string item = all.substr(pos, end - pos);
                const size_t      eq   = item.find('=');
                if (eq != std::string::npos && splacu_set_option(item.substr(0, eq).c_str(), atoll(item.c_str() + eq + 1)) != 0)
                    fprintf(stderr, "splacu: SPLACU_OPTIONS: unknown option in '%s'\n", item.c_str());
                pos = end + 1;
            }
        }
    }
    return SPLACU_OK;
}

int splacu_finalize(void) {
    const int prev = current_device();
    for (int d = 0; d < kMaxDevices; ++d)
        if (g_streams[d]) {
            cudaSetDevice(d);
            cudaStreamSynchronize(g_streams[d]);
            cudaStreamDestroy(g_streams[d]);
            g_streams[d] = nullptr;
        }
    cudaSetDevice(prev);
    cudaGetLastError();
    g_initialised = false;
    g_device      = -1;
    return SPLACU_OK;
}

int splacu_device_count(int* count) {
    SPLACU_REQUIRE(count, "null pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
    }
    return SPLACU_OK;
}

int splacu_device_name(char* buffer, int length) {
    SPLACU_REQUIRE(buffer && length > 0, "bad buffer");
    snprintf(buffer, (size_t) length, "%s", g_device_name);
    return SPLACU_OK;
}

int splacu_sm_count(int* count) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(count, "null pointer");
    *count = g_sm_count;
    return SPLACU_OK;
}

void* splacu_default_stream(void) { return g_device >= 0 ? (void*) g_stream : nullptr; }

int splacu_sync(void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_CUDA(cudaStreamSynchronize(resolve_stream(stream)));
    return SPLACU_OK;
}

const char* splacu_last_error(void) { return g_error; }

int splacu_launch_count(uint64_t* count) {
    SPLACU_REQUIRE(count, "null pointer");
    *count = g_launches.load();
    return SPLACU_OK;
}

int splacu_set_option(const char* name, int64_t value) {
    SPLACU_REQUIRE(name, "null option name");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_option_names[i]) == 0) {
            g_options[i] = value;
            return SPLACU_OK;
        }
    set_error("splacu_set_option: unknown option '%s'", name);
    return SPLACU_E_INVALID;
}

int splacu_get_option(const char* name, int64_t* value) {
    SPLACU_REQUIRE(name && value, "null pointer");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_option_names[i]) == 0) {
            *value = g_options[i];
            return SPLACU_OK;
        }
    set_error("splacu_get_option: unknown option '%s'", name);
    return SPLACU_E_INVALID;
}

int splacu_csr_info(splacu_csr handle, uint32_t* n_tiles, uint32_t* n_hub) {
    SPLACU_REQUIRE(handle, "null matrix handle");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    if (n_tiles) *n_tiles = M->n_tiles;
    if (n_hub) *n_hub = M->n_hub;
    return SPLACU_OK;
}

int splacu_csr_phases(splacu_csr handle, int* n_phases, uint32_t* nnz_per_phase, int cap) {
    SPLACU_REQUIRE(handle, "null matrix handle");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    if (n_phases) *n_phases = M->n_phases;
    if (nnz_per_phase)
        for (int p = 0; p < M->n_phases && p < cap; ++p) nnz_per_phase[p] = M->phase[p].nnz;
    return SPLACU_OK;
}

int splacu_csr_row_classes(splacu_csr handle, int* n_classes, uint32_t* nnz_per_class, uint32_t* rows_per_class, int cap) {
    SPLACU_REQUIRE(handle, "null matrix handle");
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    if (n_classes) *n_classes = M->n_scat;
    for (int q = 0; q < M->n_scat && q < cap; ++q) {
        if (nnz_per_class) nnz_per_class[q] = M->scat[q].nnz;
        if (rows_per_class) rows_per_class[q] = M->scat[q].n_slots;
    }
    return SPLACU_OK;
}

int splacu_malloc(void** d_ptr, size_t bytes) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(d_ptr, "null pointer");
    *d_ptr = nullptr;
    if (bytes == 0) bytes = 4;
    SPLACU_CUDA(cudaMalloc(d_ptr, bytes));
    return SPLACU_OK;
}

int splacu_free(void* d_ptr) {
    if (!d_ptr) return SPLACU_OK;
    // safe after splacu_finalize (reference Library::finalize destroys the accelerator while
    // decorations may still hold device buffers, src/library.cpp:97-104): the primary context survives
    cudaError_t e = cudaFree(d_ptr);
    if (e != cudaSuccess && e != cudaErrorCudartUnloading) return cuda_fail(e, "cudaFree", __FILE__, __LINE__);
    return SPLACU_OK;
}

int splacu_malloc_host(void** h_ptr, size_t bytes) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(h_ptr, "null pointer");
    if (bytes == 0) bytes = 4;
    SPLACU_CUDA(cudaMallocHost(h_ptr, bytes));
    return SPLACU_OK;
}

int splacu_free_host(void* h_ptr) {
    if (!h_ptr) return SPLACU_OK;
    cudaError_t e = cudaFreeHost(h_ptr);
    if (e != cudaSuccess && e != cudaErrorCudartUnloading) return cuda_fail(e, "cudaFreeHost", __FILE__, __LINE__);
    return SPLACU_OK;
}

int splacu_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    SPLACU_CHECK_INIT();
    if (bytes == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_dst && h_src, "null pointer");
    SPLACU_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, resolve_stream(stream)));
    return SPLACU_OK;
}

int splacu_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    SPLACU_CHECK_INIT();
    if (bytes == 0) return SPLACU_OK;
    SPLACU_REQUIRE(h_dst && d_src, "null pointer");
    SPLACU_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, resolve_stream(stream)));
    return SPLACU_OK;
}

int splacu_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream) {
    SPLACU_CHECK_INIT();
    if (bytes == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_dst && d_src, "null pointer");
    SPLACU_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, resolve_stream(stream)));
    return SPLACU_OK;
}

int splacu_csr_create(splacu_csr* out, uint32_t n_rows, uint32_t n_cols, uint32_t nnz,
                      const uint32_t* d_Ap, const uint32_t* d_Aj, const void* d_Ax, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/csr_create", resolve_stream(stream));
    SPLACU_REQUIRE(out, "null handle pointer");
    SPLACU_REQUIRE(d_Ap, "null Ap");
    SPLACU_REQUIRE(nnz == 0 || (d_Aj && d_Ax), "null Aj/Ax");
    Csr* M    = new Csr();
    M->n_rows = n_rows;
    M->n_cols = n_cols;
    M->nnz    = nnz;
    M->Ap     = d_Ap;
    M->Aj     = d_Aj;
    M->Ax     = static_cast<const uint32_t*>(d_Ax);
    int rc    = csr_build_metadata(M, resolve_stream(stream));
    if (rc != 0) {
        splacu_csr_destroy(reinterpret_cast<splacu_csr>(M));// frees whatever metadata was allocated before the failure
        return rc;
    }
    *out = reinterpret_cast<splacu_csr>(M);
    return SPLACU_OK;
}

int splacu_csr_destroy(splacu_csr handle) {
    if (!handle) return SPLACU_OK;
    Csr* M = reinterpret_cast<Csr*>(handle);
    if (M->tile_rows) cudaFree(M->tile_rows);
    if (M->carry) cudaFree(M->carry);
    if (M->hub_cols) cudaFree(M->hub_cols);
    if (M->hub_vals) cudaFree(M->hub_vals);
    if (M->Aj_hub) cudaFree(M->Aj_hub);
    if (M->sel_count) cudaFree(M->sel_count);
    if (M->side) {
        cudaStreamSynchronize(M->side);
        cudaStreamDestroy(M->side);
        cudaEventDestroy(M->ev_fork);
        cudaEventDestroy(M->ev_join);
    }
    if (M->sel_bits) cudaFree(M->sel_bits);
    if (M->fix_key) cudaFree(M->fix_key);
    if (M->fix_tile) cudaFree(M->fix_tile);
    for (int p = 0; p < M->n_phases; ++p) {
        CsrPhase& ph = M->phase[p];
        cudaFree(ph.Ap); cudaFree(ph.Aj); cudaFree(ph.Ax); cudaFree(ph.tile_rows); cudaFree(ph.carry);
        cudaFree(ph.flags); cudaFree(ph.seg_base); cudaFree(ph.seg_row); cudaFree(ph.chain); cudaFree(ph.chain_row); cudaFree(ph.head); cudaFree(ph.tail);
    }
    scat_free(M);
    cudaGetLastError();
    delete M;
    return SPLACU_OK;
}

int splacu_workspace_create(splacu_workspace* out) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(out, "null handle pointer");
    Workspace* ws = new Workspace();
    cudaError_t e = cudaMalloc(&ws->d_scalars, 64 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ws->small, (size_t) (kSmallList + 2 * kSmallFront) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(&ws->h_scalars, 64 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ws->d_scalars, 0, 64 * sizeof(uint32_t));
    if (e != cudaSuccess) {
        delete ws;
        return cuda_fail(e, "workspace alloc", __FILE__, __LINE__);
    }
    *out = reinterpret_cast<splacu_workspace>(ws);
    return SPLACU_OK;
}

int splacu_workspace_reset(splacu_workspace handle, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(handle, "null workspace");
    Workspace* ws = reinterpret_cast<Workspace*>(handle);
    ws->pending    = 0;
    ws->pend_small = ws->pend_const = false;
    ws->acc_clean  = false;// the accumulator is re-filled by the next push
    if (ws->bitmap) SPLACU_CUDA(cudaMemsetAsync(ws->bitmap, 0, ((size_t) ws->cap_n + 31) / 32 * 4 + 4, resolve_stream(stream)));
    return SPLACU_OK;
}

int splacu_workspace_info(splacu_workspace handle, int* struct_only) {
    SPLACU_REQUIRE(handle, "null workspace");
    if (struct_only) *struct_only = reinterpret_cast<Workspace*>(handle)->last_struct ? 1 : 0;
    return SPLACU_OK;
}

int splacu_workspace_destroy(splacu_workspace handle) {
    if (!handle) return SPLACU_OK;
    Workspace* ws = reinterpret_cast<Workspace*>(handle);
    cudaFree(ws->acc); cudaFree(ws->bitmap); cudaFree(ws->sel_bits); cudaFree(ws->block_sums); cudaFree(ws->d_scalars); cudaFree(ws->small);
    cudaFreeHost(ws->h_scalars);
    cudaFree(ws->keys_a); cudaFree(ws->keys_b); cudaFree(ws->vals_a); cudaFree(ws->vals_b); cudaFree(ws->offsets); cudaFree(ws->chunk_first);
    cudaFree(ws->sort_tmp);
    cudaGetLastError();
    delete ws;
    return SPLACU_OK;
}

}// extern "C"
