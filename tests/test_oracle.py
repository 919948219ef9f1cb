"""CPU tests: pin the plain-C oracle (oracle/spla_oracle.c) to the reference.

1. every known-answer vector the reference's own tests / docstrings hold for the path (tests/known_answers.py)
2. outputs of the unmodified reference CPU backend stored in tests/golden/ (made by tests/golden/make_golden.py)
3. when oracle/_ref is present (build container): a live differential run against the reference itself
"""
import os

import numpy as np
import pytest

import cases
import known_answers
from cases import FLOAT, INT, UINT

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", known_answers.KNOWN, ids=[f'{c["kind"]}:{c["source"]}' for c in known_answers.KNOWN])
def test_known_answers(oracle, case):
    Ap, Aj, Ax, dt = known_answers.as_arrays(case)
    om, oa, osel = case["ops"]
    mask = np.asarray(case["mask"], dtype=dt)
    if case["kind"] == "mxv":
        r = oracle.mxv_masked(case["dtype"], om, oa, osel, Ap, Aj, Ax, np.asarray(case["v"], dtype=dt), mask, case["init"])
        np.testing.assert_array_equal(r, np.asarray(case["expect"], dtype=dt))
    else:
        ri, rx = oracle.vxm_masked(case["dtype"], om, oa, osel, Ap, Aj, Ax, case["n_cols"],
                                   np.asarray(case["vi"], dtype=np.uint32), np.asarray(case["vx"], dtype=dt), mask)
        dense = np.zeros(case["n_cols"], dtype=dt)  # reading back untouched entries gives the fill value 0
        dense[ri] = rx
        np.testing.assert_array_equal(dense, np.asarray(case["expect"], dtype=dt))
        assert np.all(np.diff(ri.astype(np.int64)) > 0)  # strictly ascending (reference src/cpu/cpu_vxm.hpp:117)


def _golden_cases():
    z = np.load(os.path.join(GOLDEN, "mxv_vxm_reference.npz"))
    for line in z["meta"]:
        cid, dtype, om, oa, osel, n_rows, n_cols, ee = str(line).split(",")
        yield z, int(cid), int(dtype), om, oa, osel, int(n_rows), int(n_cols), bool(int(ee))


def test_golden_mxv_vxm_bit_exact(oracle):
    """The sequential restatement must reproduce the reference bit for bit, floats included."""
    n = 0
    for z, cid, dtype, om, oa, osel, n_rows, n_cols, ee in _golden_cases():
        p = f"c{cid}_"
        r = oracle.mxv_masked(dtype, om, oa, osel, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], z[p + "v"], z[p + "mask_r"],
                              z[p + "init"][0], ee)
        assert np.array_equal(r.view(np.uint32), z[p + "r"].view(np.uint32)), (cid, om, oa, osel, ee)
        ri, rx = oracle.vxm_masked(dtype, om, oa, osel, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], n_cols, z[p + "vi"], z[p + "vx"],
                                   z[p + "mask_c"])
        assert np.array_equal(ri, z[p + "ri"]), (cid, om, oa, osel)
        assert np.array_equal(rx.view(np.uint32), z[p + "rx"].view(np.uint32)), (cid, om, oa, osel)
        n += 1
    assert n >= 100


def test_op_tables(oracle):
    """Spot checks of the op semantics the kernels rely on (reference src/op.cpp:194-266)."""
    assert oracle.binary(INT, "LOR", 0, 6) == 1 and oracle.binary(INT, "LAND", 5, 7) == 1
    assert oracle.binary(INT, "BAND", 6, 3) == 2 and oracle.binary(UINT, "BXOR", 6, 3) == 5
    assert oracle.binary(INT, "MINUS_POW2", 2, 5) == 9 and oracle.binary(INT, "BONE", 9, 9) == 1
    assert oracle.binary(INT, "FIRST", 4, 8) == 4 and oracle.binary(INT, "SECOND", 4, 8) == 8
    assert oracle.binary(INT, "PLUS", 2**31 - 1, 1) == -2**31  # wraps
    assert oracle.binary(UINT, "MINUS", 0, 1) == 2**32 - 1
    assert oracle.binary(FLOAT, "MIN", 2.5, -1.0) == -1.0 and oracle.binary(FLOAT, "LOR", 0.0, 0.25) == 1.0


def test_edge_cases(oracle):
    # empty matrix, empty frontier, NEVER select, explicit fill-valued frontier entries (SURVEY 8a notes B, D, F)
    Ap = np.zeros(6, dtype=np.uint32)
    e = np.zeros(0, dtype=np.uint32)
    r = oracle.mxv_masked(INT, "MULT", "PLUS", "ALWAYS", Ap, e, e.view(np.int32), np.ones(3, np.int32), np.ones(5, np.int32), 7)
    np.testing.assert_array_equal(r, np.full(5, 7, np.int32))  # every r[i] = init
    Ap = np.array([0, 2, 3], dtype=np.uint32)
    Aj = np.array([0, 1, 1], dtype=np.uint32)
    Ax = np.array([0, 5, 7], dtype=np.int32)
    ri, rx = oracle.vxm_masked(INT, "MULT", "PLUS", "ALWAYS", Ap, Aj, Ax, 2, e, e.view(np.int32), np.zeros(2, np.int32))
    assert len(ri) == 0
    ri, rx = oracle.vxm_masked(INT, "MULT", "PLUS", "NEVER", Ap, Aj, Ax, 2, np.array([0], np.uint32), np.array([1], np.int32), np.zeros(2, np.int32))
    assert len(ri) == 0
    # a frontier entry whose value is 0 still produces structural entries (value 0 == fill stays in the pattern)
    ri, rx = oracle.vxm_masked(INT, "MULT", "PLUS", "ALWAYS", Ap, Aj, Ax, 2, np.array([0], np.uint32), np.array([0], np.int32), np.zeros(2, np.int32))
    np.testing.assert_array_equal(ri, [0, 1])
    np.testing.assert_array_equal(rx, [0, 0])
    # early exit: sequential first-hit semantics (note E): init 0, products 0, 3, 4 -> stops at 3
    Ap = np.array([0, 3], dtype=np.uint32)
    r = oracle.mxv_masked(INT, "MULT", "PLUS", "ALWAYS", Ap, np.array([0, 1, 2], np.uint32), np.array([1, 1, 1], np.int32),
                          np.array([0, 3, 4], np.int32), np.zeros(1, np.int32), 0, early_exit=True)
    assert r[0] == 3


def test_neighbour_ops(oracle):
    rng = np.random.default_rng(5)
    r = cases.rand_values(rng, INT, 50)
    mask = cases.rand_values(rng, INT, 50)
    out = oracle.v_assign_masked_dense(INT, "SECOND", "NQZERO", r, mask, 9)
    np.testing.assert_array_equal(out, np.where(mask != 0, 9, r))
    assert oracle.v_count_mf_dense(INT, out, 9) == int(np.sum(out != 9))
    d = np.full(20, np.float32(np.finfo(np.float32).max))
    d[3] = 0.0
    vi = np.array([3, 4, 7], np.uint32)
    vx = np.array([1.0, 2.0, 5.0], np.float32)
    r2, fi, fx = oracle.v_eadd_fdb_sparse(FLOAT, "MIN", d, vi, vx)
    np.testing.assert_array_equal(fi, [4, 7])
    np.testing.assert_array_equal(fx, [2.0, 5.0])
    assert r2[3] == 0.0
    s = oracle.v_reduce_dense(INT, "PLUS", np.arange(10, dtype=np.int32), 5)
    assert s == 50


@pytest.mark.skipif(not os.path.exists("/root/reference/src/library.cpp"), reason="reference sources only exist in the build container")
def test_live_differential_vs_reference(oracle):
    """Restatement vs the unmodified reference CPU backend on fresh random inputs, every built-in op pair."""
    from oracle.oracle import RefSpla

    ref = RefSpla()
    rng = np.random.default_rng(99)
    checked = 0
    for dtype in (INT, UINT, FLOAT):
        Ap, Aj, Ax = cases.rand_csr(rng, dtype, 48, 40, 4, skew=True)
        M = ref.matrix(dtype, 48, 40, cases.csr_to_coo_rows(Ap), Aj, Ax)
        for om in cases.BIN_OPS:
            for oa in cases.BIN_OPS:
                if not (cases.op_valid(dtype, om) and cases.op_valid(dtype, oa)):
                    continue
                if "DIV" in (om, oa) and dtype != FLOAT:
                    continue  # integer division by zero traps on the host
                osel = cases.SEL_OPS[checked % len(cases.SEL_OPS)]
                ee = bool(checked & 1)
                v = cases.rand_values(rng, dtype, 40)
                mask = cases.rand_values(rng, dtype, 48)
                init = cases.rand_values(rng, dtype, 1)[0]
                a = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, ee)
                b = ref.mxv_masked(M, om, oa, osel, v, mask, init, ee)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (dtype, om, oa, osel, ee)
                vi, vx = cases.rand_frontier(rng, dtype, 48, 17)
                maskc = cases.rand_values(rng, dtype, 40)
                ai, ax = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Ax, 40, vi, vx, maskc)
                bi, bx = ref.vxm_masked(M, om, oa, osel, vi, vx, maskc)
                assert np.array_equal(ai, bi) and np.array_equal(ax.view(np.uint32), bx.view(np.uint32)), (dtype, om, oa, osel)
                checked += 1
    assert checked > 500


def test_float_bound_helpers_against_oracle(oracle):
    """tests/gpu_util.py's escape for cancelling float sums: its fp32 restatement of the binary ops agrees with the C oracle bit for
    bit, and the sequential reference fold itself lies inside the derived any-order summation bound around the float64 sum."""
    from gpu_util import mxv_bound, np_binop_f32, np_select, vxm_bound

    rng = np.random.default_rng(5)
    a = cases.rand_values(rng, FLOAT, 400)
    b = cases.rand_values(rng, FLOAT, 400, "positive")
    for op in cases.BIN_OPS:
        if not cases.op_valid(FLOAT, op):
            continue
        got = np_binop_f32(op, a, b)
        want = np.array([oracle.binary(FLOAT, op, x, y) for x, y in zip(a, b)], dtype=np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), op
    n_rows, n_cols = 300, 200
    Ap, Aj, Ax = cases.rand_csr(rng, FLOAT, n_rows, n_cols, 12, skew=True)
    v = cases.rand_values(rng, FLOAT, n_cols)
    mask = np.ones(n_rows, np.float32)
    for om in ("MULT", "PLUS", "MINUS", "MIN"):
        r = oracle.mxv_masked(FLOAT, om, "PLUS", "ALWAYS", Ap, Aj, Ax, v, mask, np.float32(0.25), False)
        exact, ab = mxv_bound(om, "PLUS", Ap, Aj, Ax, v, np.float32(0.25))
        assert (np.abs(r.astype(np.float64) - exact) <= ab + 1e-12 * np.abs(exact)).all(), om
        vi, vx = cases.rand_frontier(rng, FLOAT, n_rows, 60)
        maskc = cases.rand_values(rng, FLOAT, n_cols)
        ri, rx = oracle.vxm_masked(FLOAT, om, "PLUS", "EQZERO", Ap, Aj, Ax, n_cols, vi, vx, maskc)
        exact, ab = vxm_bound(om, "PLUS", Ap, Aj, Ax, n_cols, vi, vx, np_select("EQZERO", maskc), ri)
        assert (np.abs(rx.astype(np.float64) - exact) <= ab + 1e-12 * np.abs(exact)).all(), om
    assert mxv_bound("MULT", "MIN", Ap, Aj, Ax, v, 0) is None
