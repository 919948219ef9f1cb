// jit.cu -- user-defined ops on the device.
//
// The reference hands every kernel the op's SOURCE TEXT, "(T a, T b) { ... }" (OpBinary::make_*, reference src/op.cpp:294-342;
// tests/test_vector.cpp:299-302), and compiles it into the OpenCL program at first use (src/opencl/cl_program_builder.cpp:65-120).
// Here the built-in ops are specialised ahead of time (ops.cuh); an op that is NOT a built-in is compiled at first use with NVRTC
// into a small module of generic kernels for (dtype, op_mult, op_add, op_select), loaded through the driver API and cached by key.
// Nothing is known about a user op (associativity, commutativity, identity), so the generic kernels keep the reference CPU
// backend's SEQUENTIAL semantics exactly:
//   mxv   one thread per row, strict left-to-right fold, early exit as src/cpu/cpu_mxv.hpp:88-103
//   vxm   (column, product) pairs at their frontier-order positions -> stable sort by column -> left-to-right fold per column
//         (src/cpu/cpu_vxm.hpp:92-125), i.e. the "exact ordered path" of vxm_push.cu with the user's mult / add inside
//   v_assign_masked / v_eadd / v_eadd_fdb  elementwise (src/cpu/cpu_v_assign.hpp, cpu_v_eadd.hpp, cpu_v_eadd_fdb.hpp)
// Compiled with --fmad=false: the host calls mult and add as separate std::functions, a fused multiply-add would round differently.
// libnvrtc and libcuda are opened lazily (dlopen): without them a user op yields SPLACU_E_NOT_IMPLEMENTED, built-ins are unaffected.
#include "common.cuh"
#include "jit.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace splacu { namespace jit {

    namespace {
        // ---- lazily bound entry points -----------------------------------------------------------
        struct Api {
            bool ok = false;
            std::string why;
            nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
            nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*)                                              = nullptr;
            nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*)                                                                = nullptr;
            nvrtcResult (*GetCUBIN)(nvrtcProgram, char*)                                                                      = nullptr;
            nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*)                                                           = nullptr;
            nvrtcResult (*GetProgramLog)(nvrtcProgram, char*)                                                                 = nullptr;
            nvrtcResult (*DestroyProgram)(nvrtcProgram*)                                                                      = nullptr;
            CUresult (*ModuleLoadData)(CUmodule*, const void*)                                                                = nullptr;
            CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*)                                                 = nullptr;
            CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
        };
        Api g_api;
        bool g_api_tried = false;

        template<typename F> bool bind(void* lib, const char* name, F& fn) {
            fn = reinterpret_cast<F>(dlsym(lib, name));
            return fn != nullptr;
        }
        const Api& api(bool need_driver) {
            if (!g_api_tried) {
                g_api_tried = true;
                void* rtc   = nullptr;
                for (const char* p : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
                    rtc = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
                    if (rtc) break;
                }
                if (!rtc) {
                    g_api.why = "libnvrtc.so.12 not found";
                    return g_api;
                }
                bool ok = bind(rtc, "nvrtcCreateProgram", g_api.CreateProgram) && bind(rtc, "nvrtcCompileProgram", g_api.CompileProgram) &&
                          bind(rtc, "nvrtcGetCUBINSize", g_api.GetCUBINSize) && bind(rtc, "nvrtcGetCUBIN", g_api.GetCUBIN) &&
                          bind(rtc, "nvrtcGetProgramLogSize", g_api.GetProgramLogSize) && bind(rtc, "nvrtcGetProgramLog", g_api.GetProgramLog) &&
                          bind(rtc, "nvrtcDestroyProgram", g_api.DestroyProgram);
                if (!ok) {
                    g_api.why = "libnvrtc lacks a required symbol";
                    return g_api;
                }
                g_api.ok = true;
                void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
                if (drv) {
                    bind(drv, "cuModuleLoadData", g_api.ModuleLoadData);
                    bind(drv, "cuModuleGetFunction", g_api.ModuleGetFunction);
                    bind(drv, "cuLaunchKernel", g_api.LaunchKernel);
                }
            }
            (void) need_driver;
            return g_api;
        }

        // ---- source generation -----------------------------------------------------------------------
        const char* type_name(int dtype) { return dtype == SPLACU_INT ? "int" : (dtype == SPLACU_UINT ? "uint" : "float"); }

        // body "(T a, T b) { ... }" of a built-in binary op, bit-compatible with ops.cuh / reference src/op.cpp:194-241
        std::string builtin_binop(int dtype, int op) {
            const bool f = dtype == SPLACU_FLOAT, s = dtype == SPLACU_INT;
            const std::string T = type_name(dtype);
            const std::string h = "(" + T + " a, " + T + " b) ";
            auto wrap = [&](const char* expr) {// integer arithmetic wraps: computed in unsigned
                return f ? h + "{ return " + expr + "; }" : h + "{ uint x = (uint) a, y = (uint) b; (void) x; (void) y; return (" + T + ") (" + expr + "); }";
            };
            switch (op) {
                case SPLACU_PLUS: return wrap(f ? "a + b" : "x + y");
                case SPLACU_MINUS: return wrap(f ? "a - b" : "x - y");
                case SPLACU_MULT: return wrap(f ? "a * b" : "x * y");
                case SPLACU_DIV:
                    if (f) return h + "{ return a / b; }";
                    if (s) return h + "{ return b == 0 ? 0 : ((a == (-2147483647 - 1) && b == -1) ? (-2147483647 - 1) : a / b); }";
                    return h + "{ return b ? a / b : 0u; }";
                case SPLACU_MINUS_POW2: return f ? h + "{ float d = a - b; return d * d; }" : wrap("(x - y) * (x - y)");
                case SPLACU_FIRST: return h + "{ return a; }";
                case SPLACU_SECOND: return h + "{ return b; }";
                case SPLACU_BONE: return h + "{ return (" + T + ") 1; }";
                case SPLACU_MIN: return h + "{ return (b < a) ? b : a; }";
                case SPLACU_MAX: return h + "{ return (a < b) ? b : a; }";
                case SPLACU_LOR: return h + "{ return (a != 0 || b != 0) ? (" + T + ") 1 : (" + T + ") 0; }";
                case SPLACU_LAND: return h + "{ return (a != 0 && b != 0) ? (" + T + ") 1 : (" + T + ") 0; }";
                case SPLACU_BOR: return f ? "" : h + "{ return a | b; }";
                case SPLACU_BAND: return f ? "" : h + "{ return a & b; }";
                case SPLACU_BXOR: return f ? "" : h + "{ return a ^ b; }";
                default: return "";
            }
        }
        std::string builtin_selop(int dtype, int op) {
            const std::string h = std::string("(") + type_name(dtype) + " a) ";
            switch (op) {
                case SPLACU_EQZERO: return h + "{ return a == 0; }";
                case SPLACU_NQZERO: return h + "{ return a != 0; }";
                case SPLACU_GTZERO: return h + "{ return a > 0; }";
                case SPLACU_GEZERO: return h + "{ return a >= 0; }";
                case SPLACU_LTZERO: return h + "{ return a < 0; }";
                case SPLACU_LEZERO: return h + "{ return a <= 0; }";
                case SPLACU_ALWAYS: return h + "{ return true; }";
                case SPLACU_NEVER: return h + "{ return false; }";
                default: return "";
            }
        }

        // the generic kernels; T, op_mult, op_add, op_select are spliced in front
        const char* const kKernels = R"JIT(
__device__ __forceinline__ T    as_t(uint b) { union { uint u; T t; } c; c.u = b; return c.t; }
__device__ __forceinline__ uint as_u(T t)    { union { uint u; T t; } c; c.t = t; return c.u; }

// reference src/cpu/cpu_mxv.hpp:88-103: r[i] = select(mask[i]) ? fold(add, init, mult(a, v[j]) in stored order) : init
extern "C" __global__ void __launch_bounds__(256) jit_mxv_seq(uint n_rows, const uint* __restrict__ Ap, const uint* __restrict__ Aj,
        const T* __restrict__ Ax, const T* __restrict__ v, const T* __restrict__ mask, T* __restrict__ r, uint init_bits, int early_exit) {
    const T    init   = as_t(init_bits);
    const uint stride = gridDim.x * blockDim.x;
    for (uint row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
        T sum = init;
        if (op_select(mask ? mask[row] : (T) 0)) {
            const uint k1 = Ap[row + 1];
            for (uint k = Ap[row]; k < k1; ++k) {
                sum = op_add(sum, op_mult(Ax[k], v[Aj[k]]));
                if (early_exit && sum != init) break;
            }
        }
        r[row] = sum;
    }
}

// last t in [0, nv) with off[t] <= e
__device__ __forceinline__ uint find_entry(const uint* __restrict__ off, uint nv, uint e) {
    uint lo = 0, hi = nv;
    while (hi - lo > 1) {
        const uint mid = (lo + hi) >> 1;
        if (off[mid] <= e) lo = mid; else hi = mid;
    }
    return lo;
}

// reference src/cpu/cpu_vxm.hpp:96-110: products mult(x, a) of the selected edges, at their frontier-order positions
extern "C" __global__ void __launch_bounds__(256) jit_vxm_pairs(const uint* __restrict__ Aj, const T* __restrict__ Ax, uint nv, const T* __restrict__ vx,
        const T* __restrict__ mask, const uint* __restrict__ off, const uint* __restrict__ rowstart, uint* __restrict__ keys, T* __restrict__ vals,
        uint invalid_key) {
    const uint total  = off[nv];
    const uint stride = gridDim.x * blockDim.x;
    for (uint e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint t    = find_entry(off, nv, e);
        const uint k    = rowstart[t] + (e - off[t]);
        const uint j    = Aj[k];
        const bool take = op_select(mask ? mask[j] : (T) 0);
        keys[e]         = take ? j : invalid_key;
        vals[e]         = take ? op_mult(vx[t], Ax[k]) : (T) 0;
    }
}

// after the stable sort by column: the head of every run folds it left to right (first product stored as is, cpu_vxm.hpp:104-107)
extern "C" __global__ void __launch_bounds__(256) jit_vxm_fold(uint n_pairs, const uint* __restrict__ keys, const T* __restrict__ vals, uint invalid_key,
        T* __restrict__ acc, uint* __restrict__ bitmap) {
    const uint stride = gridDim.x * blockDim.x;
    for (uint q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += stride) {
        const uint key = keys[q];
        if (key == invalid_key) continue;
        if (q > 0 && keys[q - 1] == key) continue;
        T a = vals[q];
        for (uint p = q + 1; p < n_pairs && keys[p] == key; ++p) a = op_add(a, vals[p]);
        acc[key] = a;
        atomicOr(&bitmap[key >> 5], 1u << (key & 31u));
    }
}

// reference src/cpu/cpu_v_assign.hpp:95-127 / :66-93 (op_add = the assign op)
extern "C" __global__ void __launch_bounds__(256) jit_assign_dense(uint n, T* __restrict__ r, const T* __restrict__ mask, uint value_bits) {
    const T    value  = as_t(value_bits);
    const uint stride = gridDim.x * blockDim.x;
    for (uint i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (op_select(mask ? mask[i] : (T) 0)) r[i] = op_add(r[i], value);
}
extern "C" __global__ void __launch_bounds__(256) jit_assign_sparse(T* __restrict__ r, uint nm, const uint* __restrict__ mi, const T* __restrict__ mx,
        uint value_bits) {
    const T    value  = as_t(value_bits);
    const uint stride = gridDim.x * blockDim.x;
    for (uint k = blockIdx.x * blockDim.x + threadIdx.x; k < nm; k += stride)
        if (op_select(mx[k])) { const uint i = mi[k]; r[i] = op_add(r[i], value); }
}

// reference src/cpu/cpu_v_eadd.hpp:128-152 (op_add = the op)
extern "C" __global__ void __launch_bounds__(256) jit_eadd_dense(uint n, T* __restrict__ r, const T* __restrict__ u, const T* __restrict__ v) {
    const uint stride = gridDim.x * blockDim.x;
    for (uint i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) r[i] = op_add(u[i], v[i]);
}

// reference src/cpu/cpu_v_eadd_fdb.hpp:104-137 / :70-102 (op_add = the op)
extern "C" __global__ void __launch_bounds__(256) jit_eadd_fdb_dense(uint n, T* __restrict__ r, const T* __restrict__ v, T* __restrict__ fdb, uint fill_bits) {
    const T    fill   = as_t(fill_bits);
    const uint stride = gridDim.x * blockDim.x;
    for (uint i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T prev = r[i];
        const T next = op_add(prev, v[i]);
        r[i]         = next;
        fdb[i]       = (prev != next) ? next : fill;
    }
}
extern "C" __global__ void __launch_bounds__(256) jit_eadd_fdb_sparse(T* __restrict__ r, uint nv, const uint* __restrict__ vi, const T* __restrict__ vx,
        uint* __restrict__ bitmap) {
    const uint n_pad  = (nv + 31u) & ~31u;
    const uint stride = gridDim.x * blockDim.x;
    for (uint k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += stride) {
        bool changed = false;
        if (k < nv) {
            const uint i    = vi[k];
            const T    prev = r[i];
            const T    next = op_add(prev, vx[k]);
            r[i]            = next;
            changed         = prev != next;
        }
        const uint m = __ballot_sync(0xffffffffu, changed);
        if ((threadIdx.x & 31u) == 0u) bitmap[k >> 5] = m;
    }
}
)JIT";

        const char* const kKernelNames[K_COUNT] = {"jit_mxv_seq", "jit_vxm_pairs", "jit_vxm_fold", "jit_assign_dense", "jit_assign_sparse",
                                                   "jit_eadd_dense", "jit_eadd_fdb_dense", "jit_eadd_fdb_sparse"};

        // "(T a, T b) {...}" of one op slot, or "" when it cannot be expressed
        std::string slot_source(int dtype, const splacu_op* op, bool select, int fallback_id) {
            if (!op) return select ? builtin_selop(dtype, fallback_id) : builtin_binop(dtype, fallback_id);
            if (op->id >= 0) return select ? builtin_selop(dtype, op->id) : builtin_binop(dtype, op->id);
            return op->source ? std::string(op->source) : std::string();
        }
        std::string slot_key(const splacu_op* op, int fallback_id) {
            if (!op) return "#" + std::to_string(fallback_id);
            if (op->id >= 0) return "#" + std::to_string(op->id);
            return std::string(op->name ? op->name : "?") + ":" + (op->source ? op->source : "");
        }

        std::mutex                                   g_mutex;
        std::map<std::string, std::vector<char>>     g_cubins; // key -> compiled image (device independent)
        std::map<std::string, Module*>               g_modules;// key@device -> loaded module

        int compile(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel, const std::string& key, std::vector<char>& image) {
            const Api& a = api(false);
            if (!a.ok) {
                set_error("user-defined op: NVRTC is unavailable (%s); no device code can be generated", a.why.c_str());
                return SPLACU_E_NOT_IMPLEMENTED;
            }
            const std::string T  = type_name(dtype);
            const std::string sm = slot_source(dtype, mult, false, SPLACU_FIRST), sa = slot_source(dtype, add, false, SPLACU_SECOND),
                              ss = slot_source(dtype, sel, true, SPLACU_ALWAYS);
            if (sm.empty() || sa.empty() || ss.empty()) {
                set_error("user-defined op: an op has no source text / is not defined for this type");
                return SPLACU_E_INVALID;
            }
            std::string src = "typedef unsigned int uint;\ntypedef " + T + " T;\n";
            src += "__device__ __forceinline__ " + T + " op_mult" + sm + "\n";
            src += "__device__ __forceinline__ " + T + " op_add" + sa + "\n";
            src += "__device__ __forceinline__ bool op_select" + ss + "\n";
            src += kKernels;
            nvrtcProgram prog = nullptr;
            if (a.CreateProgram(&prog, src.c_str(), "splacu_user_ops.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
                set_error("user-defined op: nvrtcCreateProgram failed");
                return SPLACU_E_COMPILE;
            }
            const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo"};
            const nvrtcResult rc = a.CompileProgram(prog, 4, opts);
            if (rc != NVRTC_SUCCESS) {
                size_t n = 0;
                a.GetProgramLogSize(prog, &n);
                std::string log(n, '\0');
                if (n) a.GetProgramLog(prog, &log[0]);
                if (log.size() > 700) log.resize(700);
                set_error("user-defined op failed to compile (key %s): %s", key.substr(0, 120).c_str(), log.c_str());
                a.DestroyProgram(&prog);
                return SPLACU_E_COMPILE;
            }
            size_t bytes = 0;
            a.GetCUBINSize(prog, &bytes);
            image.resize(bytes);
            a.GetCUBIN(prog, image.data());
            a.DestroyProgram(&prog);
            return SPLACU_OK;
        }
    }// namespace

    std::string make_key(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel) {
        return std::to_string(dtype) + "|" + slot_key(mult, SPLACU_FIRST) + "|" + slot_key(add, SPLACU_SECOND) + "|" + slot_key(sel, SPLACU_ALWAYS);
    }

    int compile_only(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel, size_t* image_bytes) {
        const std::string           key = make_key(dtype, mult, add, sel);
        std::lock_guard<std::mutex> lock(g_mutex);
        auto                        it = g_cubins.find(key);
        if (it == g_cubins.end()) {
            std::vector<char> image;
            const int         rc = compile(dtype, mult, add, sel, key, image);
            if (rc) return rc;
            it = g_cubins.emplace(key, std::move(image)).first;
            count_jit_compile();
        }
        if (image_bytes) *image_bytes = it->second.size();
        return SPLACU_OK;
    }

    int get_module(int dtype, const splacu_op* mult, const splacu_op* add, const splacu_op* sel, const Module** out) {
        *out = nullptr;
        int rc = compile_only(dtype, mult, add, sel, nullptr);
        if (rc) return rc;
        int dev = 0;
        SPLACU_CUDA(cudaGetDevice(&dev));
        const std::string           key  = make_key(dtype, mult, add, sel);
        const std::string           dkey = key + "@" + std::to_string(dev);
        std::lock_guard<std::mutex> lock(g_mutex);
        auto                        it = g_modules.find(dkey);
        if (it == g_modules.end()) {
            const Api& a = api(true);
            if (!a.ModuleLoadData || !a.ModuleGetFunction || !a.LaunchKernel) {
                set_error("user-defined op: libcuda.so.1 (driver API) is unavailable");
                return SPLACU_E_NOT_IMPLEMENTED;
            }
            SPLACU_CUDA(cudaFree(nullptr));// make sure the primary context of the device is current
            Module*  m   = new Module();
            CUmodule mod = nullptr;
            CUresult cr  = a.ModuleLoadData(&mod, g_cubins[key].data());
            if (cr != CUDA_SUCCESS) {
                delete m;
                set_error("user-defined op: cuModuleLoadData failed (%d)", (int) cr);
                return SPLACU_E_COMPILE;
            }
            m->module = mod;
            for (int k = 0; k < K_COUNT; ++k) {
                CUfunction fn = nullptr;
                cr            = a.ModuleGetFunction(&fn, mod, kKernelNames[k]);
                if (cr != CUDA_SUCCESS) {
                    delete m;
                    set_error("user-defined op: kernel %s missing from the module (%d)", kKernelNames[k], (int) cr);
                    return SPLACU_E_COMPILE;
                }
                m->fn[k] = fn;
            }
            it = g_modules.emplace(dkey, m).first;
        }
        *out = it->second;
        return SPLACU_OK;
    }

    int launch(const Module* m, int kernel, size_t work_items, void** args, cudaStream_t s) {
        const Api& a    = api(true);
        const int  grid = grid_for(work_items, 256, 8);
        CUresult   cr   = a.LaunchKernel(reinterpret_cast<CUfunction>(m->fn[kernel]), (unsigned) grid, 1, 1, 256, 1, 1, 0, reinterpret_cast<CUstream>(s), args, nullptr);
        if (cr != CUDA_SUCCESS) {
            set_error("user-defined op: cuLaunchKernel(%s) failed (%d)", kKernelNames[kernel], (int) cr);
            return SPLACU_E_INVALID;
        }
        count_launch();
        return SPLACU_OK;
    }

}}// namespace splacu::jit
