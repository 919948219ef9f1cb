/*
 * oracle/spla_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of spla's sequential CPU backend for the masked mxv / vxm hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product path (spla_b200/csrc, spla_b200/src/cuda) never does.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle.py) against
 *   - every known-answer test the reference holds for the path
 *     (tests/test_mxv.cpp:33-130, tests/test_vxm.cpp:33-187, python/pyspla/matrix.py:942-963,
 *      python/pyspla/vector.py:479-500), and
 *   - outputs of the unmodified reference CPU backend itself (oracle/_ref/libspla_ref.so driven by
 *     oracle/ref_shim.cpp), committed as tests/golden/ fixtures by tests/golden/make_golden.py.
 *
 * All values are 4-byte and travel as raw uint32_t bit patterns; `dtype` says how to read them.
 */
#ifndef SPLA_ORACLE_H
#define SPLA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* spla value types, reference src/type.cpp:32-35 (codes "I", "U", "F") */
enum { ORC_INT = 0, ORC_UINT = 1, ORC_FLOAT = 2 };

/* built-in binary ops in the order of reference src/op.cpp:194-241 */
enum {
    ORC_PLUS = 0, ORC_MINUS, ORC_MULT, ORC_DIV, ORC_MINUS_POW2, ORC_FIRST, ORC_SECOND, ORC_BONE,
    ORC_MIN, ORC_MAX, ORC_LOR, ORC_LAND, ORC_BOR, ORC_BAND, ORC_BXOR, ORC_BIN_COUNT
};

/* built-in select ops in the order of reference src/op.cpp:243-266 */
enum {
    ORC_EQZERO = 0, ORC_NQZERO, ORC_GTZERO, ORC_GEZERO, ORC_LTZERO, ORC_LEZERO, ORC_ALWAYS, ORC_NEVER,
    ORC_SEL_COUNT
};

uint32_t orc_binary(int dtype, int op, uint32_t a, uint32_t b);
int      orc_select(int dtype, int op, uint32_t a);

/* reference src/cpu/cpu_mxv.hpp:56-106. CSR rows hold the LIL rows in stored order. */
int orc_mxv_masked(int dtype, int op_mult, int op_add, int op_select,
                   uint32_t n_rows, const uint32_t* Ap, const uint32_t* Aj, const uint32_t* Ax,
                   const uint32_t* v, const uint32_t* mask, uint32_t init, int early_exit,
                   uint32_t* r);

/* reference src/cpu/cpu_vxm.hpp:58-128. Returns the number of result entries (ascending ri),
 * ri/rx must have room for min(n_cols, products) entries. */
int64_t orc_vxm_masked(int dtype, int op_mult, int op_add, int op_select,
                       uint32_t n_cols, const uint32_t* Ap, const uint32_t* Aj, const uint32_t* Ax,
                       uint32_t nv, const uint32_t* vi, const uint32_t* vx, const uint32_t* mask,
                       uint32_t* ri, uint32_t* rx);

/* neighbours of the hot path inside bfs/sssp/pr loops */

/* reference src/cpu/cpu_v_assign.hpp:55-133, dense branch: r[i] = select(mask[i]) ? assign(r[i], value) : r[i] */
void orc_v_assign_masked_dense(int dtype, int op_assign, int op_select, uint32_t n,
                               uint32_t* r, const uint32_t* mask, uint32_t value);
/* sparse-mask branch: for every stored mask entry (index, value) that passes select */
void orc_v_assign_masked_sparse(int dtype, int op_assign, int op_select,
                                uint32_t* r, uint32_t nm, const uint32_t* mi, const uint32_t* mx, uint32_t value);
/* reference src/cpu/cpu_v_count_mf.hpp:55-111, dense branch: count of entries != fill */
uint32_t orc_v_count_mf_dense(uint32_t n, const uint32_t* v, uint32_t fill, int dtype);
/* reference src/cpu/cpu_v_eadd_fdb.hpp:55-139 sparse->dense branch:
 * for every stored (i,x) of v: prev=r[i]; r[i]=op(prev,x); if (prev != r[i]) emit (i, r[i]) to fdb */
uint32_t orc_v_eadd_fdb_sparse(int dtype, int op, uint32_t* r, uint32_t nv, const uint32_t* vi, const uint32_t* vx,
                               uint32_t* fi, uint32_t* fx);
/* dense->dense branch: fdb dense, fdb[i] = changed ? r[i] : fdb_fill */
void orc_v_eadd_fdb_dense(int dtype, int op, uint32_t n, uint32_t* r, const uint32_t* v,
                          uint32_t* fdb, uint32_t fdb_fill);
/* reference src/cpu/cpu_v_eadd.hpp dense branch: r[i] = op(u[i], v[i]) */
void orc_v_eadd_dense(int dtype, int op, uint32_t n, uint32_t* r, const uint32_t* u, const uint32_t* v);
/* reference src/cpu/cpu_v_reduce.hpp dense branch: s = fold(op, init, v[0..n)) left to right */
uint32_t orc_v_reduce_dense(int dtype, int op, uint32_t n, const uint32_t* v, uint32_t init);

#ifdef __cplusplus
}
#endif
#endif
