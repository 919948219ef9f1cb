// user_ops.cu -- the op-descriptor entry points of include/splacu.h: a call whose ops are all built-ins forwards to the
// ahead-of-time specialised entry point; a call with a user-defined op runs the NVRTC-compiled generic kernels of jit.cu.
// Replaces the per-kernel JIT of the reference's OpenCL backend (src/opencl/cl_program_builder.cpp:65-120) for the hot path and
// its neighbour tasks.
#include "common.cuh"
#include "profile.cuh"
#include "jit.cuh"
#include "ops.cuh"

namespace splacu {
    int      vxm_begin_jit(const Csr* M, const jit::Module* jm, uint32_t nv, const uint32_t* d_vi, const void* d_vx, const void* d_mask, Workspace* ws,
                           uint32_t* h_nr, cudaStream_t s);
    uint64_t jit_compiles();

    static bool select_reads_mask(const splacu_op* sel) { return !sel || sel->id < 0 || make_select(sel->id).reads_mask; }
    static bool bad_builtin(int dtype, const splacu_op* op, bool select) {
        if (!op || op->id < 0) return false;
        return select ? op->id >= SPLACU_SELOP_COUNT : !op_valid_for(dtype, op->id);
    }
}// namespace splacu

using namespace splacu;

#define SPLACU_REQUIRE_OPS(dtype, m, a, s)                                                                       \
    SPLACU_REQUIRE((dtype) == SPLACU_INT || (dtype) == SPLACU_UINT || (dtype) == SPLACU_FLOAT, "unknown dtype"); \
    SPLACU_REQUIRE(!bad_builtin(dtype, m, false) && !bad_builtin(dtype, a, false) && !bad_builtin(dtype, s, true), "built-in op not defined for dtype")

extern "C" {

int splacu_jit_compile(int dtype, const splacu_op* m, const splacu_op* a, const splacu_op* s, size_t* image_bytes) {
    SPLACU_REQUIRE_OPS(dtype, m, a, s);
    return jit::compile_only(dtype, m, a, s, image_bytes);
}

int splacu_jit_compile_count(uint64_t* count) {
    SPLACU_REQUIRE(count, "null pointer");
    *count = jit_compiles();
    return SPLACU_OK;
}

int splacu_mxv_masked_ops(splacu_csr handle, int dtype, const splacu_op* m, const splacu_op* a, const splacu_op* sl, const void* d_v,
                          const void* d_mask, void* d_r, uint32_t init_bits, int early_exit, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/mxv_masked_ops", resolve_stream(stream));
    SPLACU_REQUIRE(handle && m && a && sl, "null handle / op");
    if (!jit::is_user(m) && !jit::is_user(a) && !jit::is_user(sl))
        return splacu_mxv_masked(handle, dtype, m->id, a->id, sl->id, d_v, d_mask, d_r, init_bits, early_exit, stream);
    SPLACU_REQUIRE_OPS(dtype, m, a, sl);
    const Csr* M = reinterpret_cast<const Csr*>(handle);
    if (M->n_rows == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r, "null result pointer");
    SPLACU_REQUIRE(d_mask || !select_reads_mask(sl), "null mask pointer");
    SPLACU_REQUIRE(d_v || M->nnz == 0, "null vector pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, m, a, sl, &jm);
    if (rc) return rc;
    uint32_t        n_rows = M->n_rows;
    const uint32_t *ap = M->Ap, *aj = M->Aj, *ax = M->Ax;
    void*           args[] = {&n_rows, &ap, &aj, &ax, &d_v, &d_mask, &d_r, &init_bits, &early_exit};
    return jit::launch(jm, jit::K_MXV_SEQ, n_rows, args, resolve_stream(stream));
}

int splacu_vxm_masked_begin_ops(splacu_csr handle, int dtype, const splacu_op* m, const splacu_op* a, const splacu_op* sl, uint32_t nv,
                                const uint32_t* d_vi, const void* d_vx, const void* d_mask, splacu_workspace wsh, uint32_t* h_nr, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_PROFILE("splacu/vxm_masked_begin_ops", resolve_stream(stream));
    SPLACU_REQUIRE(handle && wsh && h_nr && m && a && sl, "null handle / op");
    if (!jit::is_user(m) && !jit::is_user(a) && !jit::is_user(sl))
        return splacu_vxm_masked_begin(handle, dtype, m->id, a->id, sl->id, nv, d_vi, d_vx, d_mask, wsh, h_nr, stream);
    SPLACU_REQUIRE_OPS(dtype, m, a, sl);
    const Csr* M  = reinterpret_cast<const Csr*>(handle);
    Workspace* ws = reinterpret_cast<Workspace*>(wsh);
    SPLACU_REQUIRE(ws->pending == 0, "workspace has a pending emit");
    *h_nr = 0;
    if (nv == 0 || M->n_cols == 0 || M->nnz == 0) return SPLACU_OK;
    if (!jit::is_user(sl) && make_select(sl->id).classes == 0u) return SPLACU_OK;// NEVER
    SPLACU_REQUIRE(d_vi && d_vx, "null frontier pointers");
    SPLACU_REQUIRE(d_mask || !select_reads_mask(sl), "null mask pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, m, a, sl, &jm);
    if (rc) return rc;
    return vxm_begin_jit(M, jm, nv, d_vi, d_vx, d_mask, ws, h_nr, resolve_stream(stream));
}

int splacu_v_assign_masked_dense_ops(int dtype, const splacu_op* as, const splacu_op* sl, uint32_t n, void* d_r, const void* d_mask, uint32_t value_bits,
                                     void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(as && sl, "null op");
    if (!jit::is_user(as) && !jit::is_user(sl)) return splacu_v_assign_masked_dense(dtype, as->id, sl->id, n, d_r, d_mask, value_bits, stream);
    SPLACU_REQUIRE_OPS(dtype, (const splacu_op*) nullptr, as, sl);
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && (d_mask || !select_reads_mask(sl)), "null pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, nullptr, as, sl, &jm);
    if (rc) return rc;
    void* args[] = {&n, &d_r, &d_mask, &value_bits};
    return jit::launch(jm, jit::K_ASSIGN_DENSE, n, args, resolve_stream(stream));
}

int splacu_v_assign_masked_sparse_ops(int dtype, const splacu_op* as, const splacu_op* sl, void* d_r, uint32_t nm, const uint32_t* d_mi, const void* d_mx,
                                      uint32_t value_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(as && sl, "null op");
    if (!jit::is_user(as) && !jit::is_user(sl)) return splacu_v_assign_masked_sparse(dtype, as->id, sl->id, d_r, nm, d_mi, d_mx, value_bits, stream);
    SPLACU_REQUIRE_OPS(dtype, (const splacu_op*) nullptr, as, sl);
    if (nm == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_mi && d_mx, "null pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, nullptr, as, sl, &jm);
    if (rc) return rc;
    void* args[] = {&d_r, &nm, &d_mi, &d_mx, &value_bits};
    return jit::launch(jm, jit::K_ASSIGN_SPARSE, nm, args, resolve_stream(stream));
}

int splacu_v_eadd_dense_op(int dtype, const splacu_op* op, uint32_t n, void* d_r, const void* d_u, const void* d_v, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(op, "null op");
    if (!jit::is_user(op)) return splacu_v_eadd_dense(dtype, op->id, n, d_r, d_u, d_v, stream);
    SPLACU_REQUIRE_OPS(dtype, (const splacu_op*) nullptr, op, (const splacu_op*) nullptr);
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_u && d_v, "null pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, nullptr, op, nullptr, &jm);
    if (rc) return rc;
    void* args[] = {&n, &d_r, &d_u, &d_v};
    return jit::launch(jm, jit::K_EADD_DENSE, n, args, resolve_stream(stream));
}

int splacu_v_eadd_fdb_dense_op(int dtype, const splacu_op* op, uint32_t n, void* d_r, const void* d_v, void* d_fdb, uint32_t fdb_fill_bits, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(op, "null op");
    if (!jit::is_user(op)) return splacu_v_eadd_fdb_dense(dtype, op->id, n, d_r, d_v, d_fdb, fdb_fill_bits, stream);
    SPLACU_REQUIRE_OPS(dtype, (const splacu_op*) nullptr, op, (const splacu_op*) nullptr);
    if (n == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_v && d_fdb, "null pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, nullptr, op, nullptr, &jm);
    if (rc) return rc;
    void* args[] = {&n, &d_r, &d_v, &d_fdb, &fdb_fill_bits};
    return jit::launch(jm, jit::K_EADD_FDB_DENSE, n, args, resolve_stream(stream));
}

int splacu_v_eadd_fdb_sparse_begin_op(int dtype, const splacu_op* op, void* d_r, uint32_t nv, const uint32_t* d_vi, const void* d_vx,
                                      splacu_workspace handle, uint32_t* h_nf, void* stream) {
    SPLACU_CHECK_INIT();
    SPLACU_REQUIRE(op, "null op");
    if (!jit::is_user(op)) return splacu_v_eadd_fdb_sparse_begin(dtype, op->id, d_r, nv, d_vi, d_vx, handle, h_nf, stream);
    SPLACU_REQUIRE_OPS(dtype, (const splacu_op*) nullptr, op, (const splacu_op*) nullptr);
    SPLACU_REQUIRE(handle && h_nf, "null pointer");
    Workspace*   ws = reinterpret_cast<Workspace*>(handle);
    cudaStream_t s  = resolve_stream(stream);
    *h_nf           = 0;
    SPLACU_REQUIRE(ws->pending == 0, "workspace has a pending emit");
    if (nv == 0) return SPLACU_OK;
    SPLACU_REQUIRE(d_r && d_vi && d_vx, "null pointer");
    const jit::Module* jm = nullptr;
    int                rc = jit::get_module(dtype, nullptr, op, nullptr, &jm);
    if (rc) return rc;
    if ((rc = ws_reserve_vector(ws, nv, s))) return rc;
    uint32_t* bm     = ws->bitmap;
    void*     args[] = {&d_r, &nv, &d_vi, &d_vx, &bm};
    if ((rc = jit::launch(jm, jit::K_EADD_FDB_SPARSE, ((size_t) nv + 31) & ~(size_t) 31, args, s))) return rc;
    if ((rc = bitmap_count(ws, ws->bitmap, nv, s))) return rc;
    SPLACU_CUDA(cudaMemcpyAsync(ws->h_scalars, ws->d_scalars, 4, cudaMemcpyDeviceToHost, s));
    SPLACU_CUDA(cudaStreamSynchronize(s));
    *h_nf          = ws->h_scalars[0];
    ws->pend_small = false;
    ws->pending    = 2;
    ws->pend_n     = nv;
    ws->pend_count = *h_nf;
    ws->pend_vi    = d_vi;
    ws->pend_src   = static_cast<const uint32_t*>(d_r);
    return SPLACU_OK;
}

}// extern "C"
