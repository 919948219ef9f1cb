// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A flat C entry layer over the UNMODIFIED reference library (oracle/_ref/libspla_ref.so, built by
// oracle/Makefile straight from /root/reference). It lets tests/, tests/golden/make_golden.py and
// bench.py's cpu_baseline / --impl reference leg drive the reference's own CPU backend
// (src/cpu/cpu_mxv.hpp, src/cpu/cpu_vxm.hpp, src/algorithm.cpp) through its public C++ API from
// ctypes. Nothing in the product path links to this file.
//
// Dense vectors are fed with Vector::fill_with + set_* (src/core/tvector.hpp:159-168,260-275) so the
// CpuDense decoration holds exactly the caller's values; results are read from the decorations the
// CPU algorithms write (CpuDenseVec::Ax for mxv, CpuCooVec::Ai/Ax for vxm).

#include <spla.hpp>

#include <core/tmatrix.hpp>
#include <core/tvector.hpp>
#include <cpu/cpu_formats.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

using namespace spla;

namespace {

    ref_ptr<Type> type_of(int dtype) { return dtype == 0 ? INT : (dtype == 1 ? UINT : FLOAT); }

    ref_ptr<OpBinary> bin_of(int dtype, int op) {
        static ref_ptr<OpBinary>* table[15][3] = {
                {&PLUS_INT, &PLUS_UINT, &PLUS_FLOAT},
                {&MINUS_INT, &MINUS_UINT, &MINUS_FLOAT},
                {&MULT_INT, &MULT_UINT, &MULT_FLOAT},
                {&DIV_INT, &DIV_UINT, &DIV_FLOAT},
                {&MINUS_POW2_INT, &MINUS_POW2_UINT, &MINUS_POW2_FLOAT},
                {&FIRST_INT, &FIRST_UINT, &FIRST_FLOAT},
                {&SECOND_INT, &SECOND_UINT, &SECOND_FLOAT},
                {&BONE_INT, &BONE_UINT, &BONE_FLOAT},
                {&MIN_INT, &MIN_UINT, &MIN_FLOAT},
                {&MAX_INT, &MAX_UINT, &MAX_FLOAT},
                {&LOR_INT, &LOR_UINT, &LOR_FLOAT},
                {&LAND_INT, &LAND_UINT, &LAND_FLOAT},
                {&BOR_INT, &BOR_UINT, nullptr},
                {&BAND_INT, &BAND_UINT, nullptr},
                {&BXOR_INT, &BXOR_UINT, nullptr}};
        if (op < 0 || op >= 15 || !table[op][dtype]) return ref_ptr<OpBinary>();
        return *table[op][dtype];
    }

    ref_ptr<OpSelect> sel_of(int dtype, int op) {
        static ref_ptr<OpSelect>* table[8][3] = {
                {&EQZERO_INT, &EQZERO_UINT, &EQZERO_FLOAT},
                {&NQZERO_INT, &NQZERO_UINT, &NQZERO_FLOAT},
                {&GTZERO_INT, &GTZERO_UINT, &GTZERO_FLOAT},
                {&GEZERO_INT, &GEZERO_UINT, &GEZERO_FLOAT},
                {&LTZERO_INT, &LTZERO_UINT, &LTZERO_FLOAT},
                {&LEZERO_INT, &LEZERO_UINT, &LEZERO_FLOAT},
                {&ALWAYS_INT, &ALWAYS_UINT, &ALWAYS_FLOAT},
                {&NEVER_INT, &NEVER_UINT, &NEVER_FLOAT}};
        if (op < 0 || op >= 8) return ref_ptr<OpSelect>();
        return *table[op][dtype];
    }

    ref_ptr<Scalar> scalar_of(int dtype, uint32_t bits) {
        if (dtype == 0) { int32_t x; std::memcpy(&x, &bits, 4); return Scalar::make_int(x); }
        if (dtype == 1) return Scalar::make_uint(bits);
        float f; std::memcpy(&f, &bits, 4);
        return Scalar::make_float(f);
    }

    template<typename T>
    ref_ptr<Vector> dense_vector_t(uint n, const uint32_t* bits, uint32_t fill_bits, int dtype) {
        auto v = Vector::make(n, type_of(dtype));
        v->set_fill_value(scalar_of(dtype, fill_bits));
        v->fill_with(scalar_of(dtype, fill_bits));
        auto* tv = dynamic_cast<TVector<T>*>(v.get());
        auto* d  = tv->template get<CpuDenseVec<T>>();
        std::memcpy(d->Ax.data(), bits, sizeof(T) * n);
        return v;
    }
    ref_ptr<Vector> dense_vector(int dtype, uint n, const uint32_t* bits, uint32_t fill_bits) {
        if (dtype == 0) return dense_vector_t<T_INT>(n, bits, fill_bits, dtype);
        if (dtype == 1) return dense_vector_t<T_UINT>(n, bits, fill_bits, dtype);
        return dense_vector_t<T_FLOAT>(n, bits, fill_bits, dtype);
    }

    ref_ptr<Vector> sparse_vector(int dtype, uint n, uint nv, const uint32_t* vi, const uint32_t* vx, uint32_t fill_bits) {
        auto v = Vector::make(n, type_of(dtype));
        v->set_fill_value(scalar_of(dtype, fill_bits));
        // Vector::build copies keys/values into the CpuCoo decoration (src/core/tvector.hpp:277-303)
        static uint32_t dummy = 0;
        auto keys   = MemView::make(nv ? (void*) vi : (void*) &dummy, sizeof(uint32_t) * nv, false);
        auto values = MemView::make(nv ? (void*) vx : (void*) &dummy, sizeof(uint32_t) * nv, false);
        v->build(keys, values);
        return v;
    }

    template<typename T>
    void read_dense_t(const ref_ptr<Vector>& v, uint32_t* out) {
        auto* tv = dynamic_cast<TVector<T>*>(v.get());
        tv->validate_rw(FormatVector::CpuDense);
        auto* d = tv->template get<CpuDenseVec<T>>();
        std::memcpy(out, d->Ax.data(), sizeof(T) * d->Ax.size());
    }
    void read_dense(int dtype, const ref_ptr<Vector>& v, uint32_t* out) {
        if (dtype == 0) read_dense_t<T_INT>(v, out);
        else if (dtype == 1) read_dense_t<T_UINT>(v, out);
        else read_dense_t<T_FLOAT>(v, out);
    }

    template<typename T>
    uint32_t read_coo_t(const ref_ptr<Vector>& v, uint32_t* ri, uint32_t* rx, uint32_t cap) {
        auto* tv = dynamic_cast<TVector<T>*>(v.get());
        tv->validate_rw(FormatVector::CpuCoo);
        auto*    c = tv->template get<CpuCooVec<T>>();
        uint32_t n = c->values;
        if (ri && rx) {
            uint32_t m = n < cap ? n : cap;
            std::memcpy(ri, c->Ai.data(), sizeof(uint32_t) * m);
            std::memcpy(rx, c->Ax.data(), sizeof(T) * m);
        }
        return n;
    }
    uint32_t read_coo(int dtype, const ref_ptr<Vector>& v, uint32_t* ri, uint32_t* rx, uint32_t cap) {
        if (dtype == 0) return read_coo_t<T_INT>(v, ri, rx, cap);
        if (dtype == 1) return read_coo_t<T_UINT>(v, ri, rx, cap);
        return read_coo_t<T_FLOAT>(v, ri, rx, cap);
    }

    struct RefMatrix {
        ref_ptr<Matrix> M;
        int             dtype;
        uint            n_rows, n_cols;
    };

    double now_s() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

}// namespace

extern "C" {

// Build a reference Matrix from row-sorted COO triples via Matrix::build (src/core/tmatrix.hpp:220-253),
// the same entry the examples use after MtxLoader. Stored row order = order given.
void* refshim_matrix_create(int dtype, uint32_t n_rows, uint32_t n_cols, uint64_t nnz,
                            const uint32_t* Ai, const uint32_t* Aj, const uint32_t* Ax) {
    Library::get()->set_force_no_acceleration(true);
    auto* h   = new RefMatrix();
    h->dtype  = dtype;
    h->n_rows = n_rows;
    h->n_cols = n_cols;
    h->M      = Matrix::make(n_rows, n_cols, type_of(dtype));
    static uint32_t dummy = 0;
    auto k1 = MemView::make(nnz ? (void*) Ai : (void*) &dummy, sizeof(uint32_t) * nnz, false);
    auto k2 = MemView::make(nnz ? (void*) Aj : (void*) &dummy, sizeof(uint32_t) * nnz, false);
    auto vs = MemView::make(nnz ? (void*) Ax : (void*) &dummy, sizeof(uint32_t) * nnz, false);
    if (h->M->build(k1, k2, vs) != Status::Ok) { delete h; return nullptr; }
    h->M->set_format(FormatMatrix::CpuLil);
    return h;
}

void refshim_matrix_free(void* handle) { delete static_cast<RefMatrix*>(handle); }

// exec_mxv_masked on the reference CPU backend. `seconds` (optional) receives the mean wall time of
// `repeats` timed calls after one untimed call (formats already converted), as spla::Timer would.
int refshim_mxv_masked(void* handle, int op_mult, int op_add, int op_select,
                       const uint32_t* v, const uint32_t* mask, uint32_t init_bits, int early_exit,
                       uint32_t* r, int repeats, double* seconds) {
    auto* h = static_cast<RefMatrix*>(handle);
    const int dt = h->dtype;
    auto vv = dense_vector(dt, h->n_cols, v, 0);
    auto mm = dense_vector(dt, h->n_rows, mask, 0);
    auto rr = Vector::make(h->n_rows, type_of(dt));
    auto desc = Descriptor::make();
    desc->set_early_exit(early_exit != 0);
    auto om = bin_of(dt, op_mult); auto oa = bin_of(dt, op_add); auto os = sel_of(dt, op_select);
    if (!om || !oa || !os) return 6;
    auto init = scalar_of(dt, init_bits);
    Status st = exec_mxv_masked(rr, mm, h->M, vv, om, oa, os, init, desc);
    if (st != Status::Ok) return (int) st;
    if (repeats > 0 && seconds) {
        double t0 = now_s();
        for (int k = 0; k < repeats; ++k) exec_mxv_masked(rr, mm, h->M, vv, om, oa, os, init, desc);
        *seconds = (now_s() - t0) / repeats;
    }
    if (r) read_dense(dt, rr, r);
    return 0;
}

// exec_vxm_masked on the reference CPU backend. Returns the entry count in *nr; ri/rx get min(cap, *nr).
int refshim_vxm_masked(void* handle, int op_mult, int op_add, int op_select,
                       uint32_t nv, const uint32_t* vi, const uint32_t* vx, const uint32_t* mask, uint32_t init_bits,
                       uint32_t* ri, uint32_t* rx, uint32_t cap, uint32_t* nr, int repeats, double* seconds) {
    auto* h = static_cast<RefMatrix*>(handle);
    const int dt = h->dtype;
    auto vv = sparse_vector(dt, h->n_rows, nv, vi, vx, 0);
    auto mm = dense_vector(dt, h->n_cols, mask, 0);
    auto rr = Vector::make(h->n_cols, type_of(dt));
    auto om = bin_of(dt, op_mult); auto oa = bin_of(dt, op_add); auto os = sel_of(dt, op_select);
    if (!om || !oa || !os) return 6;
    auto init = scalar_of(dt, init_bits);
    Status st = exec_vxm_masked(rr, mm, vv, h->M, om, oa, os, init);
    if (st != Status::Ok) return (int) st;
    if (repeats > 0 && seconds) {
        double t0 = now_s();
        for (int k = 0; k < repeats; ++k) exec_vxm_masked(rr, mm, vv, h->M, om, oa, os, init);
        *seconds = (now_s() - t0) / repeats;
    }
    *nr = read_coo(dt, rr, ri, rx, cap);
    return 0;
}

// spla::bfs (src/algorithm.cpp:45-120). mode: 0 push, 1 pull, 2 push-pull. depth_out: n dense INT (0 = unreached).
int refshim_bfs(void* handle, uint32_t source, int mode, float front_factor, int32_t* depth_out, double* seconds) {
    auto* h = static_cast<RefMatrix*>(handle);
    auto  v = Vector::make(h->n_rows, INT);
    auto  d = Descriptor::make();
    d->set_traversal_mode(mode == 0 ? Descriptor::TraversalMode::Push : (mode == 1 ? Descriptor::TraversalMode::Pull : Descriptor::TraversalMode::PushPull));
    d->set_front_factor(front_factor);
    double t0 = now_s();
    Status st = bfs(v, h->M, source, d);
    if (seconds) *seconds = now_s() - t0;
    if (st != Status::Ok) return (int) st;
    read_dense(0, v, reinterpret_cast<uint32_t*>(depth_out));
    return 0;
}

// spla::sssp (src/algorithm.cpp:158-229). dist_out: n dense FLOAT (FLT_MAX = unreached).
int refshim_sssp(void* handle, uint32_t source, int mode, float front_factor, float* dist_out, double* seconds) {
    auto* h = static_cast<RefMatrix*>(handle);
    auto  v = Vector::make(h->n_rows, FLOAT);
    auto  d = Descriptor::make();
    d->set_traversal_mode(mode == 0 ? Descriptor::TraversalMode::Push : (mode == 1 ? Descriptor::TraversalMode::Pull : Descriptor::TraversalMode::PushPull));
    d->set_front_factor(front_factor);
    double t0 = now_s();
    Status st = sssp(v, h->M, source, d);
    if (seconds) *seconds = now_s() - t0;
    if (st != Status::Ok) return (int) st;
    read_dense(2, v, reinterpret_cast<uint32_t*>(dist_out));
    return 0;
}

// spla::pr (src/algorithm.cpp:278-335). rank_out: n dense FLOAT.
int refshim_pr(void* handle, float alpha, float eps, float* rank_out, double* seconds) {
    auto* h = static_cast<RefMatrix*>(handle);
    ref_ptr<Vector> p = Vector::make(h->n_rows, FLOAT);
    auto            d = Descriptor::make();
    double t0 = now_s();
    Status st = pr(p, h->M, alpha, eps, d);
    if (seconds) *seconds = now_s() - t0;
    if (st != Status::Ok) return (int) st;
    read_dense(2, p, reinterpret_cast<uint32_t*>(rank_out));
    return 0;
}

}// extern "C"
