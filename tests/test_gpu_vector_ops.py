"""GPU parity of the NEIGHBOUR ops of the hot path, op by op and randomised: every built-in binary op x every
select op x int / uint / float for v_assign_masked (dense and sparse mask), v_eadd, v_eadd_fdb (dense and sparse), v_reduce, v_count_mf
against the C oracle (oracle/spla_oracle.c restates src/cpu/cpu_v_assign.hpp:66-93, cpu_v_eadd.hpp, cpu_v_eadd_fdb.hpp:66-137,
cpu_v_reduce.hpp:60-74, cpu_v_count_mf.hpp:91-107). Elementwise ops are bit-exact for every dtype; FLOAT PLUS reductions are held to 1e-5
relative per result with the float64 escape of tests/gpu_util.py, FLOAT MULT reductions are tested on exactly representable factors. All calls go through the C ABI (spla_b200.backend)."""
import numpy as np
import pytest
import torch

import cases
from cases import FLOAT, INT, UINT
from gpu_util import assert_values, idx_dev, to_dev, to_np

pytestmark = pytest.mark.gpu

SIZES = (1, 33, 5000, 70001)


def _ops(dtype, with_div=True):
    for op in cases.BIN_OPS:
        if not cases.op_valid(dtype, op):
            continue
        if op == "DIV" and (dtype != FLOAT or not with_div):
            continue
        yield op


def _vals(rng, dtype, n, op):
    # FLOAT DIV: operands away from 0 (no NaN / inf: their ordering is not part of the contract)
    return cases.rand_values(rng, dtype, n, "positive" if op == "DIV" else "small")


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_v_assign_masked_all_ops(backend, oracle, dtype):
    """r[i] = op_assign(r[i], value) where op_select(mask[i]) -- dense mask and sparse (mi, mx) mask, every op pair."""
    rng = np.random.default_rng(900 + dtype)
    k = 0
    for op in _ops(dtype):
        for osel in cases.SEL_OPS:
            n = SIZES[k % len(SIZES)]
            k += 1
            r0 = _vals(rng, dtype, n, op)
            mask = cases.rand_values(rng, dtype, n)
            value = _vals(rng, dtype, 1, op)[0]
            want = oracle.v_assign_masked_dense(dtype, op, osel, r0, mask, value)
            r = to_dev(r0, backend)
            backend.v_assign_masked(r, to_dev(mask, backend), value, op, osel)
            backend.sync()
            assert_values(to_np(r, cases.NP[dtype]), want, True, what=f"v_assign dense {op}/{osel} n={n}")
            # sparse mask: ascending indices, stored values tested by op_select
            nm = int(rng.integers(0, n + 1))
            mi = np.sort(rng.choice(n, nm, replace=False)).astype(np.uint32)
            mx = cases.rand_values(rng, dtype, nm)
            want = oracle.v_assign_masked_sparse(dtype, op, osel, r0, mi, mx, value)
            r = to_dev(r0, backend)
            backend.v_assign_masked(r, (idx_dev(mi, backend), to_dev(mx, backend)), value, op, osel)
            backend.sync()
            assert_values(to_np(r, cases.NP[dtype]), want, True, what=f"v_assign sparse {op}/{osel} n={n} nm={nm}")
    assert k >= 96


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_v_eadd_all_ops(backend, oracle, dtype):
    """r[i] = op(u[i], v[i]) for every built-in op: elementwise, bit-exact."""
    rng = np.random.default_rng(910 + dtype)
    for k, op in enumerate(_ops(dtype)):
        n = SIZES[k % len(SIZES)]
        u, v = _vals(rng, dtype, n, op), _vals(rng, dtype, n, op)
        want = oracle.v_eadd_dense(dtype, op, u, v)
        got = backend.v_eadd(to_dev(u, backend), to_dev(v, backend), op)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, True, what=f"v_eadd {op} n={n}")


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_v_eadd_fdb_all_ops(backend, oracle, dtype):
    """r[i] = op(r[i], v[i]) with the feedback vector of the changed entries: dense v (fdb[i] = changed ? r[i] : fill) and sparse v
    ((fi, fx) = the changed entries in the order of v), every built-in op."""
    rng = np.random.default_rng(920 + dtype)
    for k, op in enumerate(_ops(dtype)):
        n = SIZES[(k + 1) % len(SIZES)]
        r0, v = _vals(rng, dtype, n, op), _vals(rng, dtype, n, op)
        fill = _vals(rng, dtype, 1, op)[0]
        want_r, want_f = oracle.v_eadd_fdb_dense(dtype, op, r0, v, fill)
        r = to_dev(r0, backend)
        fdb = backend.v_eadd_fdb_dense(r, to_dev(v, backend), op, fill)
        backend.sync()
        assert_values(to_np(r, cases.NP[dtype]), want_r, True, what=f"v_eadd_fdb dense r {op} n={n}")
        assert_values(to_np(fdb, cases.NP[dtype]), want_f, True, what=f"v_eadd_fdb dense fdb {op} n={n}")
        for nv in sorted({0, 1, min(n, 17), int(rng.integers(0, n + 1))}):
            vi = np.sort(rng.choice(n, nv, replace=False)).astype(np.uint32)
            vx = _vals(rng, dtype, nv, op)
            want_r, want_fi, want_fx = oracle.v_eadd_fdb_sparse(dtype, op, r0, vi, vx)
            r = to_dev(r0, backend)
            fi, fx = backend.v_eadd_fdb_sparse(r, idx_dev(vi, backend), to_dev(vx, backend), op)
            backend.sync()
            assert_values(to_np(r, cases.NP[dtype]), want_r, True, what=f"v_eadd_fdb sparse r {op} n={n} nv={nv}")
            assert np.array_equal(to_np(fi, np.uint32), want_fi), f"v_eadd_fdb sparse feedback pattern {op} n={n} nv={nv}"
            assert_values(to_np(fx, cases.NP[dtype]), want_fx, True, what=f"v_eadd_fdb sparse fx {op} n={n} nv={nv}")


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_v_reduce_assoc_ops(backend, oracle, dtype):
    """s = op(init, v[0], v[1], ...) for the associative + commutative ops (the device folds as a tree, the reference left to right):
    integers, MIN / MAX, logical and bitwise results and FLOAT products of exactly representable factors are bit-exact; FLOAT PLUS within
    1e-5 relative or no further from the float64 value than the reference's own sequential fold."""
    rng = np.random.default_rng(930 + dtype)
    for k, op in enumerate(o for o in cases.ASSOC if cases.op_valid(dtype, o)):
        for n in (1, 33, 5000, 300007):
            if dtype == FLOAT and op == "MULT":
                # factors whose product is exact in ANY order (powers of two, a few dozen of them != 1), compared bit for bit. Random
                # factors near 1 are no test of the kernel: a pairwise fp32 product of 3e5 factors from [0.999, 1.001] lands 2.6e-4 below
                # the float64 value in every tree order (numpy float32 in the device's order reproduces the device bit for bit) while the
                # reference's left-to-right fold stays within 2e-5 -- a property of fp32 products, and no caller on the path reduces
                # with MULT (spla::pr reduces with PLUS, src/algorithm.cpp:317)
                v = np.ones(n, dtype=np.float32)
                pick = rng.random(n) < 40.0 / max(n, 40)
                v[pick] = rng.choice(np.array([0.5, 2.0, 4.0, 0.25], dtype=np.float32), int(pick.sum()))
            elif dtype != FLOAT and op == "MULT":
                v = np.where(rng.random(n) < 4.0 / max(n, 4), 3, 1).astype(cases.NP[dtype])  # a few factors, wraps like the reference
            elif dtype == FLOAT and op == "PLUS":
                v = rng.uniform(0.0, 1.0, n).astype(np.float32)
            else:
                v = cases.rand_values(rng, dtype, n)
            init = {"MULT": 1, "MIN": 7, "MAX": -7 if dtype != UINT else 0, "LAND": 1, "BAND": 0x7fffffff if dtype != FLOAT else 1}.get(op, 0)
            want = oracle.v_reduce_dense(dtype, op, v, init)
            got = backend.v_reduce(to_dev(v, backend), op, init)
            g = np.array([got]).astype(cases.NP[dtype])
            w = np.array([want]).astype(cases.NP[dtype])
            exact = dtype != FLOAT or op != "PLUS"
            bound = None
            if not exact:
                bound = (np.array([np.float64(init) + v.astype(np.float64).sum()]), None)
            assert_values(g, w, exact, what=f"v_reduce {op} n={n}", bound=bound)


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_v_count_mf(backend, oracle, dtype):
    """number of entries that differ from the fill value (value comparison: -0.0 == +0.0)."""
    rng = np.random.default_rng(940 + dtype)
    for n in (1, 31, 32, 33, 5000, 300007):
        for fill in (0, 1, 3):
            v = cases.rand_values(rng, dtype, n)
            if dtype == FLOAT:
                v[rng.random(n) < 0.1] = -0.0
            v[rng.random(n) < 0.5] = fill
            want = oracle.v_count_mf_dense(dtype, v, fill)
            got = backend.v_count_mf(to_dev(v, backend), fill)
            assert got == want, (n, fill, got, want)
