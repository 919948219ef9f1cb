/*
 * oracle/spla_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see spla_oracle.h).
 *
 * Sequential restatement of spla's CPU backend for the masked mxv / vxm path and its
 * neighbours. Written from the semantics of the reference, in plain C over raw 32-bit
 * patterns. Parity: pinned against the reference's own known-answer tests and against
 * outputs of the unmodified reference CPU backend (tests/test_oracle.py, tests/golden/).
 */
#include "spla_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline float    as_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_as(float f)    { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int32_t  as_i(uint32_t u) { int32_t i; memcpy(&i, &u, 4); return i; }

/* Built-in binary ops, reference src/op.cpp:194-241.
 * INT arithmetic wraps (two's complement) as the compiled reference does in practice.
 * MIN/MAX follow std::min/std::max: min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a  (op.cpp:150-154).
 * LOR/LAND return the C++ bool converted to T, i.e. exactly 0 or 1 (op.cpp:229-234). */
uint32_t orc_binary(int dtype, int op, uint32_t a, uint32_t b) {
    if (dtype == ORC_FLOAT) {
        const float x = as_f(a), y = as_f(b);
        switch (op) {
            case ORC_PLUS:       return f_as(x + y);
            case ORC_MINUS:      return f_as(x - y);
            case ORC_MULT:       return f_as(x * y);
            case ORC_DIV:        return f_as(x / y);
            case ORC_MINUS_POW2: { volatile float d = x - y; return f_as(d * d); }
            case ORC_FIRST:      return a;
            case ORC_SECOND:     return b;
            case ORC_BONE:       return f_as(1.0f);
            case ORC_MIN:        return (y < x) ? b : a;
            case ORC_MAX:        return (x < y) ? b : a;
            case ORC_LOR:        return f_as((x != 0.0f || y != 0.0f) ? 1.0f : 0.0f);
            case ORC_LAND:       return f_as((x != 0.0f && y != 0.0f) ? 1.0f : 0.0f);
            default:             return 0; /* BOR/BAND/BXOR do not exist for FLOAT (op.cpp:236-241) */
        }
    }
    if (dtype == ORC_INT) {
        const int32_t x = as_i(a), y = as_i(b);
        switch (op) {
            case ORC_PLUS:       return a + b;
            case ORC_MINUS:      return a - b;
            case ORC_MULT:       return a * b;
            case ORC_DIV:        return (uint32_t) (x / y);
            case ORC_MINUS_POW2: return (a - b) * (a - b);
            case ORC_FIRST:      return a;
            case ORC_SECOND:     return b;
            case ORC_BONE:       return 1u;
            case ORC_MIN:        return (y < x) ? b : a;
            case ORC_MAX:        return (x < y) ? b : a;
            case ORC_LOR:        return (x || y) ? 1u : 0u;
            case ORC_LAND:       return (x && y) ? 1u : 0u;
            case ORC_BOR:        return a | b;
            case ORC_BAND:       return a & b;
            case ORC_BXOR:       return a ^ b;
            default:             return 0;
        }
    }
    switch (op) { /* ORC_UINT */
        case ORC_PLUS:       return a + b;
        case ORC_MINUS:      return a - b;
        case ORC_MULT:       return a * b;
        case ORC_DIV:        return a / b;
        case ORC_MINUS_POW2: return (a - b) * (a - b);
        case ORC_FIRST:      return a;
        case ORC_SECOND:     return b;
        case ORC_BONE:       return 1u;
        case ORC_MIN:        return (b < a) ? b : a;
        case ORC_MAX:        return (a < b) ? b : a;
        case ORC_LOR:        return (a || b) ? 1u : 0u;
        case ORC_LAND:       return (a && b) ? 1u : 0u;
        case ORC_BOR:        return a | b;
        case ORC_BAND:       return a & b;
        case ORC_BXOR:       return a ^ b;
        default:             return 0;
    }
}

/* Built-in select ops, reference src/op.cpp:243-266 */
int orc_select(int dtype, int op, uint32_t a) {
    if (op == ORC_ALWAYS) return 1;
    if (op == ORC_NEVER) return 0;
    if (dtype == ORC_FLOAT) {
        const float x = as_f(a);
        switch (op) {
            case ORC_EQZERO: return x == 0;
            case ORC_NQZERO: return x != 0;
            case ORC_GTZERO: return x > 0;
            case ORC_GEZERO: return x >= 0;
            case ORC_LTZERO: return x < 0;
            case ORC_LEZERO: return x <= 0;
        }
    } else if (dtype == ORC_INT) {
        const int32_t x = as_i(a);
        switch (op) {
            case ORC_EQZERO: return x == 0;
            case ORC_NQZERO: return x != 0;
            case ORC_GTZERO: return x > 0;
            case ORC_GEZERO: return x >= 0;
            case ORC_LTZERO: return x < 0;
            case ORC_LEZERO: return x <= 0;
        }
    } else {
        switch (op) {
            case ORC_EQZERO: return a == 0;
            case ORC_NQZERO: return a != 0;
            case ORC_GTZERO: return a > 0;
            case ORC_GEZERO: return 1;
            case ORC_LTZERO: return 0;
            case ORC_LEZERO: return a == 0;
        }
    }
    return 0;
}

/* `x != y` in the value type (reference compares T values, so float NaN != NaN and -0 == +0) */
static inline int neq(int dtype, uint32_t a, uint32_t b) {
    return dtype == ORC_FLOAT ? (as_f(a) != as_f(b)) : (a != b);
}

/* reference src/cpu/cpu_mxv.hpp:88-103: every r[i] is written; rows failing the select (and empty
 * rows) get init; the fold is left to right in stored order starting FROM init; mult(a_ij, v_j);
 * early_exit stops at the first position where sum != init. */
int orc_mxv_masked(int dtype, int op_mult, int op_add, int op_select,
                   uint32_t n_rows, const uint32_t* Ap, const uint32_t* Aj, const uint32_t* Ax,
                   const uint32_t* v, const uint32_t* mask, uint32_t init, int early_exit,
                   uint32_t* r) {
    for (uint32_t i = 0; i < n_rows; ++i) {
        uint32_t sum = init;
        if (orc_select(dtype, op_select, mask[i])) {
            for (uint32_t k = Ap[i]; k < Ap[i + 1]; ++k) {
                sum = orc_binary(dtype, op_add, sum, orc_binary(dtype, op_mult, Ax[k], v[Aj[k]]));
                if (early_exit && neq(dtype, sum, init)) break;
            }
        }
        r[i] = sum;
    }
    return 0;
}

/* reference src/cpu/cpu_vxm.hpp:92-125: frontier entries in stored order, row entries in stored
 * order, mask tested per target column; the first product for a column is stored as is (init is
 * never used), later ones folded with add(acc, mult(v_i, a_ij)); output sorted by column. The
 * hash map + sort of the reference is restated as a dense accumulator + touched flags. */
int64_t orc_vxm_masked(int dtype, int op_mult, int op_add, int op_select,
                       uint32_t n_cols, const uint32_t* Ap, const uint32_t* Aj, const uint32_t* Ax,
                       uint32_t nv, const uint32_t* vi, const uint32_t* vx, const uint32_t* mask,
                       uint32_t* ri, uint32_t* rx) {
    uint32_t* acc     = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) (n_cols ? n_cols : 1));
    uint8_t*  touched = (uint8_t*) calloc(n_cols ? n_cols : 1, 1);
    if (!acc || !touched) { free(acc); free(touched); return -1; }

    for (uint32_t idx = 0; idx < nv; ++idx) {
        const uint32_t i = vi[idx];
        const uint32_t x = vx[idx];
        for (uint32_t k = Ap[i]; k < Ap[i + 1]; ++k) {
            const uint32_t j = Aj[k];
            if (!orc_select(dtype, op_select, mask[j])) continue;
            const uint32_t p = orc_binary(dtype, op_mult, x, Ax[k]);
            if (touched[j]) {
                acc[j] = orc_binary(dtype, op_add, acc[j], p);
            } else {
                acc[j]     = p;
                touched[j] = 1;
            }
        }
    }

    int64_t nr = 0;
    for (uint32_t j = 0; j < n_cols; ++j) {
        if (touched[j]) {
            ri[nr] = j;
            rx[nr] = acc[j];
            nr++;
        }
    }
    free(acc);
    free(touched);
    return nr;
}

/* reference src/cpu/cpu_v_assign.hpp:113-121 */
void orc_v_assign_masked_dense(int dtype, int op_assign, int op_select, uint32_t n,
                               uint32_t* r, const uint32_t* mask, uint32_t value) {
    for (uint32_t i = 0; i < n; ++i)
        if (orc_select(dtype, op_select, mask[i])) r[i] = orc_binary(dtype, op_assign, r[i], value);
}

/* reference src/cpu/cpu_v_assign.hpp:84-92 */
void orc_v_assign_masked_sparse(int dtype, int op_assign, int op_select,
                                uint32_t* r, uint32_t nm, const uint32_t* mi, const uint32_t* mx, uint32_t value) {
    for (uint32_t k = 0; k < nm; ++k)
        if (orc_select(dtype, op_select, mx[k])) r[mi[k]] = orc_binary(dtype, op_assign, r[mi[k]], value);
}

/* reference src/cpu/cpu_v_count_mf.hpp:91-107 */
uint32_t orc_v_count_mf_dense(uint32_t n, const uint32_t* v, uint32_t fill, int dtype) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; ++i) c += neq(dtype, v[i], fill) ? 1u : 0u;
    return c;
}

/* reference src/cpu/cpu_v_eadd_fdb.hpp:90-101 */
uint32_t orc_v_eadd_fdb_sparse(int dtype, int op, uint32_t* r, uint32_t nv, const uint32_t* vi, const uint32_t* vx,
                               uint32_t* fi, uint32_t* fx) {
    uint32_t nf = 0;
    for (uint32_t k = 0; k < nv; ++k) {
        const uint32_t i    = vi[k];
        const uint32_t prev = r[i];
        r[i]                = orc_binary(dtype, op, prev, vx[k]);
        if (neq(dtype, prev, r[i])) {
            fi[nf] = i;
            fx[nf] = r[i];
            nf++;
        }
    }
    return nf;
}

/* reference src/cpu/cpu_v_eadd_fdb.hpp:124-134 */
void orc_v_eadd_fdb_dense(int dtype, int op, uint32_t n, uint32_t* r, const uint32_t* v,
                          uint32_t* fdb, uint32_t fdb_fill) {
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t prev = r[i];
        r[i]                = orc_binary(dtype, op, prev, v[i]);
        fdb[i]              = neq(dtype, prev, r[i]) ? r[i] : fdb_fill;
    }
}

/* reference src/cpu/cpu_v_eadd.hpp:146-150 */
void orc_v_eadd_dense(int dtype, int op, uint32_t n, uint32_t* r, const uint32_t* u, const uint32_t* v) {
    for (uint32_t i = 0; i < n; ++i) r[i] = orc_binary(dtype, op, u[i], v[i]);
}

/* reference src/cpu/cpu_v_reduce.hpp:100-112 */
uint32_t orc_v_reduce_dense(int dtype, int op, uint32_t n, const uint32_t* v, uint32_t init) {
    uint32_t s = init;
    for (uint32_t i = 0; i < n; ++i) s = orc_binary(dtype, op, s, v[i]);
    return s;
}
