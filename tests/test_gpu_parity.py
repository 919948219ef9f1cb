"""GPU parity tests proper: the CUDA path, called through the C ABI (include/splacu.h), against
 (1) the reference's known-answer vectors, (2) the committed outputs of the reference CPU backend
 (tests/golden), (3) the plain-C oracle on seeded random inputs over every built-in op / type / select,
 and (4) size-independent properties at full benchmark sizes.

Bar: bit-exact for INT / UINT and every order-independent op; FLOAT PLUS / MULT reductions within 1e-5
relative (north_star), the tolerance being written in gpu_util.assert_values.
"""
import os
import zlib

import numpy as np
import pytest
import torch

import cases
import known_answers
from cases import FLOAT, INT, UINT
from gpu_util import assert_values, idx_dev, make_csr, mxv_bound, np_select, to_dev, to_np, vxm_bound

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_mxv(backend, dtype, om, oa, osel, n_cols, Ap, Aj, Ax, v, mask, init, ee):
    M = make_csr(backend, len(Ap) - 1, n_cols, Ap, Aj, Ax)
    r = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init, early_exit=ee)
    backend.sync()
    return to_np(r, cases.NP[dtype])


def run_vxm(backend, dtype, om, oa, osel, n_cols, Ap, Aj, Ax, vi, vx, mask):
    M = make_csr(backend, len(Ap) - 1, n_cols, Ap, Aj, Ax)
    ri, rx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend), om, oa, osel)
    backend.sync()
    return to_np(ri, np.uint32), to_np(rx, cases.NP[dtype])


@pytest.mark.parametrize("case", known_answers.KNOWN, ids=[f'{c["kind"]}:{c["source"]}' for c in known_answers.KNOWN])
def test_known_answers(backend, case):
    Ap, Aj, Ax, dt = known_answers.as_arrays(case)
    om, oa, osel = case["ops"]
    mask = np.asarray(case["mask"], dtype=dt)
    if case["kind"] == "mxv":
        r = run_mxv(backend, case["dtype"], om, oa, osel, case["n_cols"], Ap, Aj, Ax, np.asarray(case["v"], dtype=dt), mask, case["init"], False)
        np.testing.assert_array_equal(r, np.asarray(case["expect"], dtype=dt))
    else:
        ri, rx = run_vxm(backend, case["dtype"], om, oa, osel, case["n_cols"], Ap, Aj, Ax, np.asarray(case["vi"], dtype=np.uint32),
                         np.asarray(case["vx"], dtype=dt), mask)
        dense = np.zeros(case["n_cols"], dtype=dt)
        dense[ri] = rx
        np.testing.assert_array_equal(dense, np.asarray(case["expect"], dtype=dt))
        assert np.all(np.diff(ri.astype(np.int64)) > 0)


def test_golden_reference_outputs(backend):
    """Against outputs of the unmodified reference CPU backend (tests/golden/mxv_vxm_reference.npz)."""
    z = np.load(os.path.join(GOLDEN, "mxv_vxm_reference.npz"))
    n = 0
    for line in z["meta"]:
        cid, dtype, om, oa, osel, n_rows, n_cols, ee = str(line).split(",")
        dtype, n_cols, ee = int(dtype), int(n_cols), bool(int(ee))
        p = f"c{cid}_"
        exact = cases.exact_expected(dtype, om, oa) or ee
        r = run_mxv(backend, dtype, om, oa, osel, n_cols, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], z[p + "v"], z[p + "mask_r"], z[p + "init"][0], ee)
        assert_values(r, z[p + "r"], exact, what=f"mxv case {cid} {om}/{oa}/{osel} ee={ee}",
                      bound=lambda: mxv_bound(om, oa, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], z[p + "v"], z[p + "init"][0]))
        ri, rx = run_vxm(backend, dtype, om, oa, osel, n_cols, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], z[p + "vi"], z[p + "vx"], z[p + "mask_c"])
        assert np.array_equal(ri, z[p + "ri"]), f"vxm pattern case {cid} {om}/{oa}/{osel}"
        assert_values(rx, z[p + "rx"], cases.exact_expected(dtype, om, oa), what=f"vxm case {cid} {om}/{oa}/{osel}",
                      bound=lambda: vxm_bound(om, oa, z[p + "Ap"], z[p + "Aj"], z[p + "Ax"], n_cols, z[p + "vi"], z[p + "vx"], np_select(osel, z[p + "mask_c"]), ri))
        n += 1
    assert n >= 100


@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_all_op_pairs_vs_oracle(backend, oracle, dtype):
    """Every built-in (op_mult, op_add) pair, cycling selects and early_exit, skewed rows, non-identity init."""
    rng = np.random.default_rng(1234 + dtype)
    n_rows, n_cols = 301, 257
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, 5, skew=True)
    M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    k = 0
    for om in cases.BIN_OPS:
        for oa in cases.BIN_OPS:
            if not (cases.op_valid(dtype, om) and cases.op_valid(dtype, oa)):
                continue
            if "DIV" in (om, oa) and dtype != FLOAT:
                continue
            osel = cases.SEL_OPS[k % 8]
            ee = bool((k // 8) & 1)
            k += 1
            # FLOAT DIV: keep operands away from 0 so that no NaN is produced (NaN ordering under MIN/MAX is not part of the contract)
            vk = "positive" if "DIV" in (om, oa) else "small"
            v = cases.rand_values(rng, dtype, n_cols, vk)
            mask = cases.rand_values(rng, dtype, n_rows)
            init = cases.rand_values(rng, dtype, 1, vk)[0]
            want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, ee)
            got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init, early_exit=ee)
            backend.sync()
            assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa) or ee, what=f"mxv {om}/{oa}/{osel} ee={ee}",
                          bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))

            vi, vx = cases.rand_frontier(rng, dtype, n_rows, 97, vk)
            maskc = cases.rand_values(rng, dtype, n_cols)
            wi, wx = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Ax, n_cols, vi, vx, maskc)
            gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(maskc, backend), om, oa, osel)
            backend.sync()
            assert np.array_equal(to_np(gi, np.uint32), wi), f"vxm pattern {om}/{oa}/{osel}"
            assert_values(to_np(gx, cases.NP[dtype]), wx, cases.exact_expected(dtype, om, oa), what=f"vxm {om}/{oa}/{osel}",
                          bound=lambda: vxm_bound(om, oa, Ap, Aj, Ax, n_cols, vi, vx, np_select(osel, maskc), wi))
    assert k > 100


@pytest.mark.parametrize("dtype,om,oa,osel", cases.NAMED_SEMIRINGS)
@pytest.mark.parametrize("shape", [(1, 1, 1), (5000, 3000, 2), (4096, 4096, 40), (20000, 20000, 12)])
def test_named_semirings_shapes(backend, oracle, dtype, om, oa, osel, shape):
    """The semirings bfs / sssp / pr use, on ragged, rectangular and skewed shapes (rows > 2^13 nnz included)."""
    n_rows, n_cols, avg = shape
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, n_rows)).encode()))
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, avg, skew=n_rows > 100, kind=kind)
    M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    for ee in (False, True):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        v[rng.random(n_cols) < 0.5] = 0
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else 0
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, ee)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init, early_exit=ee)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa) or ee, what=f"mxv {om}/{oa} {shape} ee={ee}",
                      bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))
    for nv in (0, 1, max(1, n_rows // 50), n_rows):
        vi, vx = cases.rand_frontier(rng, dtype, n_rows, nv, kind)
        maskc = cases.rand_values(rng, dtype, n_cols)
        wi, wx = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Ax, n_cols, vi, vx, maskc)
        gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(maskc, backend), om, oa, osel)
        backend.sync()
        assert np.array_equal(to_np(gi, np.uint32), wi), f"vxm pattern nv={nv}"
        assert_values(to_np(gx, cases.NP[dtype]), wx, cases.exact_expected(dtype, om, oa), what=f"vxm {om}/{oa} {shape} nv={nv}",
                      bound=lambda: vxm_bound(om, oa, Ap, Aj, Ax, n_cols, vi, vx, np_select(osel, maskc), wi))


@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (UINT, "BAND", "BOR", "ALWAYS"), (FLOAT, "MULT", "PLUS", "ALWAYS"),
                                               (FLOAT, "PLUS", "MIN", "NQZERO"), (INT, "LAND", "LOR", "GTZERO")])
@pytest.mark.parametrize("hub_smem,hub_total", [(0, 64), (16, 64), (256, 256), (128, 4096)])
def test_pull_hub_cache_forced(backend, oracle, dtype, om, oa, osel, hub_smem, hub_total):
    """The hub-cache variant of the streaming pull kernel (most referenced columns served from shared memory / the packed
    hub table) is normally reserved for large matrices; force it on small skewed ones and compare with the oracle."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, hub_smem, hub_total)).encode()))
    n_rows, n_cols = 3000, 2500
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, 30, skew=True, kind=kind)
    # make the column popularity skewed as well: fold most columns onto a few hubs, keep rows sorted + duplicate free
    hubs = rng.integers(0, n_cols, 40, dtype=np.int64)
    fold = rng.random(len(Aj)) < 0.6
    Aj = Aj.astype(np.int64)
    Aj[fold] = hubs[rng.integers(0, len(hubs), int(fold.sum()))]
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(Ap.astype(np.int64)))
    key, first = np.unique(rows * n_cols + Aj, return_index=True)
    Aj, Ax, rows = (key % n_cols).astype(np.uint32), Ax[first], key // n_cols
    Ap = np.zeros(n_rows + 1, dtype=np.uint32)
    Ap[1:] = np.cumsum(np.bincount(rows, minlength=n_rows))
    try:
        backend.set_option("mxv_hub", 2)
        backend.set_option("mxv_hub_smem", hub_smem)
        backend.set_option("mxv_hub_total", hub_total)
        backend.set_option("mxv_hub_min_count", 2)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        assert backend.csr_info(M)["n_hub"] > 0
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_hub_smem", 16384)
        backend.set_option("mxv_hub_total", 1 << 20)
        backend.set_option("mxv_hub_min_count", 16)
    for rep in range(2):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else 0
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, False)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"hub mxv rep {rep}",
                      bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))


def _skewed_csr(rng, dtype, n_rows, n_cols, kind, avg=30):
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, avg, skew=True, kind=kind)
    hubs = rng.integers(0, n_cols, 40, dtype=np.int64)
    fold = rng.random(len(Aj)) < 0.6
    Aj = Aj.astype(np.int64)
    Aj[fold] = hubs[rng.integers(0, len(hubs), int(fold.sum()))]
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(Ap.astype(np.int64)))
    key, first = np.unique(rows * n_cols + Aj, return_index=True)
    Aj, Ax, rows = (key % n_cols).astype(np.uint32), Ax[first], key // n_cols
    Ap = np.zeros(n_rows + 1, dtype=np.uint32)
    Ap[1:] = np.cumsum(np.bincount(rows, minlength=n_rows))
    return Ap, Aj, Ax


@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (UINT, "BAND", "BOR", "ALWAYS"), (FLOAT, "MULT", "PLUS", "ALWAYS"),
                                               (FLOAT, "PLUS", "MIN", "NQZERO"), (INT, "LAND", "LOR", "GTZERO"), (FLOAT, "MULT", "PLUS", "NQZERO")])
@pytest.mark.parametrize("slots,phases,min_count", [(4, 1, 2), (16, 3, 2), (64, 16, 1), (1024, 2, 2), (49152, 4, 1)])
def test_pull_column_class_phases_forced(backend, oracle, dtype, om, oa, osel, slots, phases, min_count):
    """The column-class phases of the streaming pull kernel (hub classes with 16-bit slots gathered from shared memory, then
    the tail class, each accumulating onto r) are normally reserved for large matrices; force them on small skewed ones --
    including classes that end up empty, rows longer than a tile and rows with entries in a single class -- and compare
    with the oracle."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, slots, phases)).encode()))
    n_rows, n_cols = 3000, 2500
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = _skewed_csr(rng, dtype, n_rows, n_cols, kind)
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_phase_slots", slots)
        backend.set_option("mxv_phases", phases)
        backend.set_option("mxv_hub_min_count", min_count)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        info = backend.csr_info(M)
        assert len(info["phase_nnz"]) >= 2 and sum(info["phase_nnz"]) + sum(info["row_class_nnz"]) == len(Aj), info
        assert info["phase_nnz"][0] > 0
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
    for rep in range(2):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else (0 if rep == 0 else 3)
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, False)
        # the three forms of the fix-up of rows that span tiles: two plain launches (default: all chain sums, then one thread per row in
        # class order), the ordered cooperative launch, one launch per class; the first two add in the same order: bit-identical
        bits = {}
        for fix in (2, 1, 0):
            try:
                backend.set_option("mxv_fixup_merge", fix)
                got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init)
                backend.sync()
            finally:
                backend.set_option("mxv_fixup_merge", 2)
            assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"phases mxv rep {rep} fix-up {fix}",
                          bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))
            bits[fix] = to_np(got, np.uint32)
        # (one launch per class interleaves the chain sums with the class passes: another order, equal only within the bar)
        d21 = np.flatnonzero(bits[2] != bits[1])
        assert d21.size == 0, f"the two-launch fix-up differs bitwise from the ordered cooperative launch at rows {d21[:8]} ({d21.size})"
    # early exit and non-associative adds keep using the original CSR of the handle
    v = cases.rand_values(rng, dtype, n_cols, kind)
    mask = cases.rand_values(rng, dtype, n_rows)
    want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, 0, True)
    got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, 0, early_exit=True)
    backend.sync()
    assert np.array_equal(to_np(got, cases.NP[dtype]), want)


def test_pull_column_class_phases_auto(backend, oracle):
    """A matrix large enough for the automatic choice (>= 4 Mi entries, >= 164 K columns): classes are built without any
    option, the result matches the oracle (INT: bit-exact, FLOAT: 1e-5)."""
    rng = np.random.default_rng(77)
    n = 1 << 18
    nnz_target = 5 << 20
    # power-law column popularity, uniform rows
    rows = rng.integers(0, n, nnz_target, dtype=np.int64)
    cols = (n * rng.random(nnz_target) ** 4).astype(np.int64)
    perm = rng.permutation(n)
    cols = perm[cols]
    key = np.unique(rows * n + cols)
    rows, cols = key // n, (key % n).astype(np.uint32)
    Ap = np.zeros(n + 1, dtype=np.uint32)
    Ap[1:] = np.cumsum(np.bincount(rows, minlength=n))
    assert len(cols) >= (1 << 22)
    for dtype, om, oa, osel in [(INT, "MULT", "PLUS", "NQZERO"), (FLOAT, "MULT", "PLUS", "ALWAYS")]:
        Ax = cases.rand_values(rng, dtype, len(cols), "unit" if dtype == FLOAT else "small")
        M = make_csr(backend, n, n, Ap, cols, Ax)
        info = backend.csr_info(M)
        assert len(info["phase_nnz"]) >= 2 and sum(info["phase_nnz"]) + sum(info["row_class_nnz"]) == len(cols), info
        v = cases.rand_values(rng, dtype, n, "unit" if dtype == FLOAT else "small")
        mask = cases.rand_values(rng, dtype, n)
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, cols, Ax, v, mask, 1, False)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, 1)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"auto phases {dtype}",
                      bound=lambda: mxv_bound(om, oa, Ap, cols, Ax, v, 1))


def test_edge_cases(backend, oracle):
    # empty matrix: every r[i] = init (SURVEY 8a note B); empty frontier; NEVER; explicit zero frontier values (note F)
    e = np.zeros(0, dtype=np.uint32)
    Ap = np.zeros(6, dtype=np.uint32)
    r = run_mxv(backend, INT, "MULT", "PLUS", "ALWAYS", 3, Ap, e, e.view(np.int32), np.ones(3, np.int32), np.ones(5, np.int32), 7, False)
    np.testing.assert_array_equal(r, np.full(5, 7, np.int32))
    Ap = np.array([0, 2, 3], dtype=np.uint32)
    Aj = np.array([0, 1, 1], dtype=np.uint32)
    Ax = np.array([0, 5, 7], dtype=np.int32)
    z2 = np.zeros(2, np.int32)
    ri, rx = run_vxm(backend, INT, "MULT", "PLUS", "ALWAYS", 2, Ap, Aj, Ax, e, e.view(np.int32), z2)
    assert len(ri) == 0
    ri, rx = run_vxm(backend, INT, "MULT", "PLUS", "NEVER", 2, Ap, Aj, Ax, np.array([0], np.uint32), np.array([1], np.int32), z2)
    assert len(ri) == 0
    ri, rx = run_vxm(backend, INT, "MULT", "PLUS", "ALWAYS", 2, Ap, Aj, Ax, np.array([0], np.uint32), np.array([0], np.int32), z2)
    np.testing.assert_array_equal(ri, [0, 1])
    np.testing.assert_array_equal(rx, [0, 0])
    # early exit stops at the first position where the running sum != init (note E)
    Ap = np.array([0, 3], dtype=np.uint32)
    r = run_mxv(backend, INT, "MULT", "PLUS", "ALWAYS", 3, Ap, np.array([0, 1, 2], np.uint32), np.array([1, 1, 1], np.int32),
                np.array([0, 3, 4], np.int32), np.zeros(1, np.int32), 0, True)
    assert r[0] == 3
    # argument order of op_mult (note A): mxv mult(a, v), vxm mult(v, a)
    Ap = np.array([0, 1], dtype=np.uint32)
    r = run_mxv(backend, INT, "MINUS", "PLUS", "ALWAYS", 1, Ap, np.array([0], np.uint32), np.array([10], np.int32), np.array([3], np.int32),
                np.zeros(1, np.int32), 0, False)
    assert r[0] == 7
    ri, rx = run_vxm(backend, INT, "MINUS", "PLUS", "ALWAYS", 1, Ap, np.array([0], np.uint32), np.array([10], np.int32), np.array([0], np.uint32),
                     np.array([3], np.int32), np.zeros(1, np.int32))
    assert rx[0] == -7


def test_repeated_vxm_reuses_clean_scratch(backend, oracle):
    """The dense accumulator / bitmap scratch is reset by emit: alternating semirings must not leak state."""
    rng = np.random.default_rng(77)
    n = 5000
    Ap, Aj, Ax = cases.rand_csr(rng, INT, n, n, 8, skew=True)
    M = make_csr(backend, n, n, Ap, Aj, Ax)
    for it in range(6):
        om, oa = [("MULT", "PLUS"), ("BAND", "BOR"), ("PLUS", "MIN"), ("MULT", "MAX"), ("FIRST", "SECOND"), ("MULT", "LOR")][it]
        vi, vx = cases.rand_frontier(rng, INT, n, 300 + 100 * it)
        mask = cases.rand_values(rng, INT, n)
        wi, wx = oracle.vxm_masked(INT, om, oa, "EQZERO", Ap, Aj, Ax, n, vi, vx, mask)
        gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend), om, oa, "EQZERO")
        backend.sync()
        assert np.array_equal(to_np(gi, np.uint32), wi) and np.array_equal(to_np(gx, np.int32), wx), (it, om, oa)


def test_format_glue(backend):
    rng = np.random.default_rng(3)
    for n in (1, 31, 32, 33, 1000, 100003):
        dense = cases.rand_values(rng, FLOAT, n)
        dense[rng.random(n) < 0.6] = 7.5
        ri, rx = backend.dense_to_coo(to_dev(dense, backend), 7.5)
        backend.sync()
        keep = np.nonzero(dense != np.float32(7.5))[0]
        np.testing.assert_array_equal(to_np(ri, np.uint32), keep.astype(np.uint32))
        np.testing.assert_array_equal(to_np(rx, np.float32), dense[keep])
        back = backend.coo_to_dense(n, 7.5, ri, rx)
        backend.sync()
        np.testing.assert_array_equal(to_np(back, np.float32), dense)


def test_properties_at_scale(backend):
    """Size-independent checks at a benchmark-like size (RMAT scale 20, ~30 M nnz) where the CPU oracle is too slow:
    linearity of MULT/PLUS, mxv <-> vxm duality on a symmetric matrix, mask monotonicity, idempotence of BOR."""
    from spla_b200 import graphs

    n, Ap64, Aj = graphs.rmat(20, seed=5, device=backend.device)
    torch.cuda.synchronize()
    nnz = Aj.numel()
    Ap = Ap64.to(torch.int32)
    ones_i = torch.ones(nnz, dtype=torch.int32, device=backend.device)
    Mi = backend.csr(n, n, Ap, Aj, ones_i)
    g = torch.Generator(device=backend.device)
    g.manual_seed(11)
    x = torch.randint(0, 5, (n,), generator=g, device=backend.device, dtype=torch.int32)
    y = torch.randint(0, 5, (n,), generator=g, device=backend.device, dtype=torch.int32)
    zeros = torch.zeros(n, dtype=torch.int32, device=backend.device)
    torch.cuda.synchronize()  # inputs were produced on torch's default stream; the backend runs on its own stream
    with torch.cuda.stream(backend.stream):
        rx_ = backend.mxv_masked(Mi, x, zeros, "MULT", "PLUS", "EQZERO", 0)
        ry_ = backend.mxv_masked(Mi, y, zeros, "MULT", "PLUS", "EQZERO", 0)
        rxy = backend.mxv_masked(Mi, x + y, zeros, "MULT", "PLUS", "EQZERO", 0)
        backend.sync()
        assert torch.equal(rxy, rx_ + ry_)  # linearity, exact in int32
        deg = (Ap64[1:] - Ap64[:-1]).to(torch.int32)
        rdeg = backend.mxv_masked(Mi, torch.ones_like(x), zeros, "MULT", "PLUS", "EQZERO", 0)
        backend.sync()
        assert torch.equal(rdeg, deg)  # A * 1 = row degrees
        # duality on the symmetric matrix: (f x A)[j] == (A x f)[j] for a sparse f (reference bfs relies on it)
        vi = torch.nonzero(x == 4).flatten().to(torch.int32)
        vx = torch.full((vi.numel(),), 3, dtype=torch.int32, device=backend.device)
        f_dense = backend.coo_to_dense(n, 0, vi, vx)
        pull = backend.mxv_masked(Mi, f_dense, zeros, "MULT", "PLUS", "EQZERO", 0)
        ri, rxv = backend.vxm_masked(Mi, vi, vx, zeros, "MULT", "PLUS", "EQZERO")
        backend.sync()
        push_dense = backend.coo_to_dense(n, 0, ri, rxv)
        backend.sync()
        assert torch.equal(push_dense, pull)
        assert bool((ri[1:] > ri[:-1]).all())  # sortedness
        # mask monotonicity: masked result == unmasked result where selected, init elsewhere
        mask = (y > 2).to(torch.int32)
        rm = backend.mxv_masked(Mi, x, mask, "MULT", "PLUS", "EQZERO", -1)
        backend.sync()
        assert torch.equal(rm, torch.where((mask == 0) & (deg > 0), rx_ - 1, torch.full_like(rx_, -1)))
        # BFS semiring: early-exit result equals the full OR-reduction (values 0/1), and is idempotent
        fb = (x == 4).to(torch.int32)
        full = backend.mxv_masked(Mi, fb, zeros, "BAND", "BOR", "EQZERO", 0)
        ee = backend.mxv_masked(Mi, fb, zeros, "BAND", "BOR", "EQZERO", 0, early_exit=True)
        backend.sync()
        assert torch.equal(full, ee)
        assert torch.equal(full, (pull > 0).to(torch.int32))


def test_publish_window_single_device(backend):
    """splacu_publish_window (the multi-GPU all-gather as one kernel of peer stores) with three 'peers' that are plain buffers on
    the same device: the window of the own copy lands in the other two, nothing else changes."""
    import ctypes as C

    import torch

    n, w0, cnt = 4096 + 40, 128, 1000  # count is rounded up to a multiple of 4 by the caller
    with torch.cuda.stream(backend.stream):
        bufs = [torch.full((n,), float(q), dtype=torch.float32, device=backend.device) for q in range(3)]
        bufs[1][:] = torch.arange(n, dtype=torch.float32, device=backend.device)
    backend.sync()
    ptrs = (C.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    rc = backend.lib.splacu_publish_window(ptrs, 3, 1, C.c_size_t(w0), C.c_size_t(cnt), backend.stream_ptr)
    assert rc == 0
    backend.sync()
    want = torch.arange(n, dtype=torch.float32)
    for q in (0, 2):
        got = bufs[q].cpu()
        assert torch.equal(got[w0:w0 + cnt], want[w0:w0 + cnt])
        assert bool((got[:w0] == float(q)).all()) and bool((got[w0 + cnt:] == float(q)).all())
    assert torch.equal(bufs[1].cpu(), want)
    # misaligned windows are refused, not silently rounded
    assert backend.lib.splacu_publish_window(ptrs, 3, 1, C.c_size_t(w0 + 1), C.c_size_t(cnt), backend.stream_ptr) != 0


@pytest.mark.parametrize("density", [0.0, 0.3, 0.44, 0.46, 0.9, 1.0])
def test_pull_mask_density_gate(backend, oracle, density):
    """With column classes the device chooses per call between the class passes (dense masks) and the mask-first CSR kernel
    (sparse masks) from the number of selected rows; both sides of the threshold, and the extremes, must match the oracle."""
    rng = np.random.default_rng(int(density * 1000) + 5)
    n_rows, n_cols = 4000, 3000
    Ap, Aj, Ax = _skewed_csr(rng, INT, n_rows, n_cols, "small")
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_phase_slots", 128)
        backend.set_option("mxv_phases", 3)
        backend.set_option("mxv_hub_min_count", 2)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        assert len(backend.csr_info(M)["phase_nnz"]) >= 2
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
    v = cases.rand_values(rng, INT, n_cols, "small")
    mask = (rng.random(n_rows) < density).astype(np.int32)
    for osel, m in (("NQZERO", mask), ("EQZERO", 1 - mask)):
        want = oracle.mxv_masked(INT, "MULT", "PLUS", osel, Ap, Aj, Ax, v, m, 7, False)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(m, backend), "MULT", "PLUS", osel, 7)
        backend.sync()
        assert np.array_equal(to_np(got, np.int32), want), f"density {density} {osel}"


@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (FLOAT, "MULT", "PLUS", "NQZERO"), (UINT, "BAND", "BOR", "ALWAYS")])
def test_pull_column_classes_csr_format(backend, oracle, dtype, om, oa, osel):
    """The column classes can also be kept as plain CSR slices run by the CSR tile kernel (option mxv_seg = 0: the format the
    segmented tiles replaced, kept as the comparison point of DESIGN.md); it must stay exact too."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, "csr-classes")).encode()))
    n_rows, n_cols = 3000, 2500
    kind = "unit" if dtype == FLOAT else "small"
    Ap, Aj, Ax = _skewed_csr(rng, dtype, n_rows, n_cols, kind)
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_seg", 0)
        backend.set_option("mxv_phase_slots", 64)
        backend.set_option("mxv_phases", 3)
        backend.set_option("mxv_hub_min_count", 2)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        assert len(backend.csr_info(M)["phase_nnz"]) >= 2
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_seg", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
    v = cases.rand_values(rng, dtype, n_cols, kind)
    mask = cases.rand_values(rng, dtype, n_rows)
    want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, 2, False)
    got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, 2)
    backend.sync()
    assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what="csr-format classes",
                  bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, 2))


@pytest.mark.parametrize("seg", [1, 0])
@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (FLOAT, "MULT", "PLUS", "ALWAYS"), (FLOAT, "PLUS", "MIN", "NQZERO"),
                                              (UINT, "BAND", "BOR", "ALWAYS")])
def test_pull_tail_column_ranges(backend, oracle, dtype, om, oa, osel, seg):
    """Vectors larger than the L2 split the tail class into windows of 2^k columns, one pass each (option mxv_tail_range_log2,
    24 by default); forced down to 1024-column windows here so that a 2500-column matrix gets 2 hub classes + 3 tail windows.
    Every class accumulates onto r in a fixed order: exact for INT / UINT / MIN, 1e-5 for FLOAT sums."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, "tail-ranges", seg)).encode()))
    n_rows, n_cols = 3000, 2500
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = _skewed_csr(rng, dtype, n_rows, n_cols, kind)
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_seg", seg)
        backend.set_option("mxv_phase_slots", 64)
        backend.set_option("mxv_phases", 2)
        backend.set_option("mxv_hub_min_count", 2)
        backend.set_option("mxv_tail_range_log2", 10)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        info = backend.csr_info(M)
        assert len(info["phase_nnz"]) == 2 + 3 and sum(info["phase_nnz"]) + sum(info["row_class_nnz"]) == len(Aj), info
        assert all(x > 0 for x in info["phase_nnz"][2:]), info
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_seg", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
        backend.set_option("mxv_tail_range_log2", 24)
    for rep in range(2):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else (0 if rep == 0 else 5)
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, False)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"tail ranges rep {rep}",
                      bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))


def test_pull_tail_ranges_widen_to_fit(backend):
    """More windows than the handle has classes for: the windows are widened (doubling) until they fit."""
    rng = np.random.default_rng(5)
    n_rows, n_cols = 2000, 60000
    Ap, Aj, Ax = cases.rand_csr(rng, INT, n_rows, n_cols, 20, skew=True, kind="small")
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_phase_slots", 64)
        backend.set_option("mxv_phases", 2)
        backend.set_option("mxv_hub_min_count", 1)
        backend.set_option("mxv_tail_range_log2", 10)  # 59 windows asked for, 17 classes available
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        info = backend.csr_info(M)
        assert 2 < len(info["phase_nnz"]) <= 17 and sum(info["phase_nnz"]) + sum(info["row_class_nnz"]) == len(Aj), info
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
        backend.set_option("mxv_tail_range_log2", 24)


@pytest.mark.parametrize("small", [1, 0])
def test_small_front_paths(backend, oracle, small):
    """Fronts of <= 8192 entries take single-CTA offset / sorted-emit / filter kernels (option small_front, on by default);
    the general kernels must give the same, bit for bit: push products whose result fits the 4096-entry list, overflows it,
    and fronts above the limit; the sparse eadd_fdb below and above the limit; scratch reuse across calls."""
    rng = np.random.default_rng(2024 + small)
    n = 30000
    Ap, Aj, Ax = cases.rand_csr(rng, FLOAT, n, n, 6, skew=False, kind="positive")
    try:
        backend.set_option("small_front", small)
        M = make_csr(backend, n, n, Ap, Aj, Ax)
        dist = np.full(n, np.float32(3.0e38), dtype=np.float32)
        d_dev = to_dev(dist, backend)
        for nv in (1, 40, 500, 3000, 8192, 8193, 20000, 7):
            vi, vx = cases.rand_frontier(rng, FLOAT, n, nv, "positive")
            mask = cases.rand_values(rng, FLOAT, n)
            for om, oa, osel in (("MULT", "PLUS", "EQZERO"), ("PLUS", "MIN", "ALWAYS")):  # the last one (exact) feeds the eadd_fdb below
                wi, wx = oracle.vxm_masked(FLOAT, om, oa, osel, Ap, Aj, Ax, n, vi, vx, mask)
                gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend), om, oa, osel)
                backend.sync()
                assert np.array_equal(to_np(gi, np.uint32), wi), f"vxm pattern nv={nv} {om}/{oa} small_front={small}"
                assert_values(to_np(gx, np.float32), wx, cases.exact_expected(FLOAT, om, oa), what=f"vxm nv={nv} {om}/{oa}")
            # the SSSP step: distances <- min(distances, candidates); feedback = the entries that improved, in index order
            want_r, want_fi, want_fx = oracle.v_eadd_fdb_sparse(FLOAT, "MIN", dist, wi, wx)
            fi, fx = backend.v_eadd_fdb_sparse(d_dev, gi, gx, "MIN")
            backend.sync()
            assert np.array_equal(to_np(fi, np.uint32), want_fi) and np.array_equal(to_np(fx, np.float32), want_fx), f"eadd_fdb feedback nv={nv}"
            assert np.array_equal(to_np(d_dev, np.float32), want_r), f"eadd_fdb distances nv={nv}"
            dist = want_r
    finally:
        backend.set_option("small_front", 1)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 70001])
def test_pack_unpack_bits(backend, n):
    """Structure-only exchange form of a dense vector: bit i = select(v[i]), trailing bits of the last word 0, and back."""
    rng = np.random.default_rng(n)
    for dtype, np_t in ((INT, np.int32), (FLOAT, np.float32)):
        v = (rng.integers(-1, 2, n)).astype(np_t)
        for osel, pred in (("NQZERO", v != 0), ("GTZERO", v > 0), ("ALWAYS", np.ones(n, bool))):
            bits = backend.pack_bits(to_dev(v, backend), osel)
            backend.sync()
            got = bits.cpu().numpy().view(np.uint8)
            want = np.packbits(pred.astype(np.uint8), bitorder="little")
            want = np.concatenate([want, np.zeros((-len(want)) % 4, dtype=np.uint8)])
            assert np.array_equal(got, want), (n, dtype, osel)
            out = to_dev(np.full(n, 9, dtype=np_t), backend)
            backend.unpack_bits(bits, n, 1, 0, out)
            backend.sync()
            assert np.array_equal(to_np(out, np_t), pred.astype(np_t)), (n, dtype, osel)


@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "BAND", "BOR", "EQZERO"), (INT, "LAND", "LOR", "EQZERO"), (UINT, "BAND", "BOR", "ALWAYS"),
                                               (FLOAT, "PLUS", "MIN", "NQZERO"), (INT, "MULT", "MAX", "GEZERO"), (FLOAT, "LAND", "LOR", "EQZERO"),
                                               (INT, "FIRST", "MIN", "EQZERO"), (INT, "SECOND", "BAND", "EQZERO"), (INT, "BONE", "LOR", "ALWAYS")])
def test_push_structure_only_path(backend, oracle, dtype, om, oa, osel):
    """Large frontiers whose products are provably ONE value under an idempotent add run structure-only (no Ax / vx / accumulator
    reads): uniform frontier values x uniform matrix values, or a mult that ignores the varying side. The path must be taken when
    it applies, must not be taken when the values vary (decided on the device), and both must match the oracle bit for bit --
    including explicit zero products (note F: touched columns stay) and a second call on the same workspace."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, "struct")).encode()))
    n = 72000  # n / 8 frontier entries: above the small-front limit (8192), which keeps its own single-CTA path
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n, n, 5, skew=False)
    np_t = cases.NP[dtype]
    for a_val, x_val in ((3, 5), (0, 2), (1, 1)):
        Au = np.full(len(Aj), a_val, dtype=np_t)
        M = make_csr(backend, n, n, Ap, Aj, Au)
        vi = np.sort(rng.choice(n, size=n // 8, replace=False)).astype(np.uint32)  # nv * 64 >= n: a "large" frontier
        mask = cases.rand_values(rng, dtype, n)
        for uniform in (True, False, True):
            vx = np.full(len(vi), x_val, dtype=np_t)
            if not uniform:
                vx[len(vx) // 2] = x_val + 1
            wi, wx = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Au, n, vi, vx, mask)
            gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend), om, oa, osel)
            backend.sync()
            expect_struct = uniform or om in ("SECOND", "BONE")
            assert backend.vxm_info()["struct_only"] == expect_struct, (om, oa, uniform)
            assert np.array_equal(to_np(gi, np.uint32), wi), f"pattern {om}/{oa} uniform={uniform}"
            assert_values(to_np(gx, np_t), wx, True, what=f"struct vxm {om}/{oa} a={a_val} x={x_val} uniform={uniform}")
    # matrix values vary: never structure-only unless the mult ignores them
    M = make_csr(backend, n, n, Ap, Aj, Ax)
    vx = np.full(len(vi), 2, dtype=np_t)
    wi, wx = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Ax, n, vi, vx, mask)
    gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend), om, oa, osel)
    backend.sync()
    assert backend.vxm_info()["struct_only"] == (om in ("FIRST", "BONE"))
    assert np.array_equal(to_np(gi, np.uint32), wi)
    assert_values(to_np(gx, np_t), wx, True, what=f"struct vxm {om}/{oa} varying matrix")
    # a sum is not idempotent: never structure-only
    Au = np.full(len(Aj), 1, dtype=np_t)
    M = make_csr(backend, n, n, Ap, Aj, Au)
    backend.vxm_masked(M, idx_dev(vi, backend), to_dev(np.full(len(vi), 1, dtype=np_t), backend), to_dev(mask, backend), "MULT", "PLUS", osel)
    backend.sync()
    assert not backend.vxm_info()["struct_only"]


def test_workspace_reset_after_abandoned_begin(backend, oracle):
    """A begin() without its emit() (an exception in the caller) leaves the workspace pending; splacu_workspace_reset makes it usable again."""
    import ctypes as C

    rng = np.random.default_rng(8)
    n = 3000
    Ap, Aj, Ax = cases.rand_csr(rng, INT, n, n, 6)
    M = make_csr(backend, n, n, Ap, Aj, Ax)
    vi, vx = cases.rand_frontier(rng, INT, n, 200)
    mask = cases.rand_values(rng, INT, n)
    d_vi, d_vx, d_m = idx_dev(vi, backend), to_dev(vx, backend), to_dev(mask, backend)
    nr = C.c_uint32(0)
    from spla_b200.backend import BIN, SEL, _ptr

    args = (M.handle, INT, BIN["MULT"], BIN["PLUS"], SEL["EQZERO"], len(vi), _ptr(d_vi), _ptr(d_vx), _ptr(d_m), backend.ws, C.byref(nr), backend.stream_ptr)
    assert backend.lib.splacu_vxm_masked_begin(*args) == 0
    assert backend.lib.splacu_vxm_masked_begin(*args) != 0  # pending emit: refused
    backend.reset_workspace()
    wi, wx = oracle.vxm_masked(INT, "MULT", "PLUS", "EQZERO", Ap, Aj, Ax, n, vi, vx, mask)
    gi, gx = backend.vxm_masked(M, d_vi, d_vx, d_m, "MULT", "PLUS", "EQZERO")
    backend.sync()
    assert np.array_equal(to_np(gi, np.uint32), wi) and np.array_equal(to_np(gx, np.int32), wx)


@pytest.mark.gpu
@pytest.mark.parametrize("n_rows,nnz,order", [(1, 0, "sorted"), (7, 0, "sorted"), (1, 5, "sorted"), (1000, 20000, "sorted"), (1000, 20000, "shuffled"),
                                              (70001, 300000, "shuffled"), (50, 4000, "reversed"), (300000, 1000, "sorted"), (300000, 1000, "shuffled")])
def test_ingest_coo_to_csr(backend, n_rows, nnz, order):
    """splacu_coo_to_csr against a stable host sort by row (reference src/cpu/cpu_format_coo.hpp:58-76: entries of a row keep their
    input order); sorted input takes the in-place path, anything else the device sort; empty leading / trailing rows included."""
    rng = np.random.default_rng(nnz + n_rows)
    lo, hi = (n_rows // 5, n_rows - n_rows // 7) if n_rows > 10 else (0, n_rows)  # leave empty rows at both ends
    Ai = rng.integers(lo, max(hi, lo + 1), nnz).astype(np.uint32)
    if order == "sorted":
        Ai.sort(kind="stable")
    elif order == "reversed":
        Ai[::-1].sort(kind="stable")
    Aj = rng.integers(0, 1 << 20, nnz).astype(np.uint32)
    Ax = rng.standard_normal(nnz).astype(np.float32)
    Ap_d, Aj_d, Ax_d, was_sorted = backend.coo_to_csr(n_rows, idx_dev(Ai, backend), idx_dev(Aj, backend), to_dev(Ax, backend))
    backend.sync()
    perm = np.argsort(Ai, kind="stable")
    want_Ap = np.zeros(n_rows + 1, dtype=np.uint32)
    np.cumsum(np.bincount(Ai, minlength=n_rows), out=want_Ap[1:])
    assert was_sorted == bool(nnz == 0 or np.all(Ai[1:] >= Ai[:-1]))
    assert np.array_equal(to_np(Ap_d, np.uint32), want_Ap)
    assert np.array_equal(to_np(Aj_d, np.uint32), Aj[perm])
    assert np.array_equal(to_np(Ax_d, np.float32).view(np.uint32), Ax[perm].view(np.uint32))
    # a row id outside the matrix is refused
    if nnz:
        bad = Ai.copy()
        bad[nnz // 2] = n_rows
        with pytest.raises(Exception):
            backend.coo_to_csr(n_rows, idx_dev(bad, backend), idx_dev(Aj, backend), to_dev(Ax, backend))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (UINT, "BAND", "BOR", "ALWAYS"), (FLOAT, "MULT", "PLUS", "ALWAYS"),
                                               (FLOAT, "PLUS", "MIN", "NQZERO"), (INT, "LAND", "LOR", "GTZERO"), (FLOAT, "MULT", "PLUS", "NQZERO"),
                                               (INT, "PLUS", "MAX", "ALWAYS"), (UINT, "MULT", "BXOR", "NQZERO"), (FLOAT, "BONE", "MULT", "ALWAYS"),
                                               (INT, "MIN", "MULT", "LEZERO")])
@pytest.mark.parametrize("slots,col_phases,row_classes,row_min", [(4, 1, 1, 1), (16, 2, 4, 2), (64, 3, 2, 1), (1024, 1, 1, 8), (45056, 1, 1, 1)])
def test_pull_row_classes_forced(backend, oracle, dtype, om, oa, osel, slots, col_phases, row_classes, row_min):
    """The row classes of the tail (mxv_scat.cu: tail-column entries of the heaviest rows, column-ordered, accumulated with
    shared-memory atomics while v streams, merged onto r under the mask) forced on small skewed matrices: every op family of the
    atomic combine (PLUS / MIN / MAX / MULT-by-CAS / logical / bitwise), ragged last tiles, column segments that span tiles (a
    heavy column), rows that live entirely in a row class, empty rows and unselected rows."""
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, slots, row_classes)).encode()))
    n_rows, n_cols = 2600, 3100
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    # skewed rows AND skewed columns; transposing the skewed generator's output makes a few rows very long
    Ap0, Aj0, Ax0 = _skewed_csr(rng, dtype, n_cols, n_rows, kind)
    rows0 = np.repeat(np.arange(n_cols), np.diff(Ap0.astype(np.int64)))
    order = np.lexsort((rows0, Aj0))
    Aj, Ax = rows0[order].astype(np.uint32), Ax0[order]
    Ap = np.zeros(n_rows + 1, dtype=np.uint32)
    Ap[1:] = np.cumsum(np.bincount(Aj0, minlength=n_rows))
    try:
        backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_phase_slots", slots)
        backend.set_option("mxv_phases", col_phases)
        backend.set_option("mxv_hub_min_count", 1)
        backend.set_option("mxv_row_classes", row_classes)
        backend.set_option("mxv_row_min_count", row_min)
        backend.set_option("mxv_row_min_nnz", 0)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
        info = backend.csr_info(M)
        assert sum(info["phase_nnz"]) + sum(info["row_class_nnz"]) == len(Aj), info
        if slots < 45056:
            assert len(info["row_class_nnz"]) >= 1 and info["row_class_nnz"][0] > 0, info
        assert len(info["row_class_nnz"]) <= row_classes
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
        backend.set_option("mxv_row_classes", 1)
        backend.set_option("mxv_row_min_count", 64)
        backend.set_option("mxv_row_min_nnz", 25165824)
    for rep in range(3):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else (0 if rep == 0 else 3)
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, False)
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init)
        backend.sync()
        assert_values(to_np(got, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"row classes mxv rep {rep}",
                      bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (FLOAT, "MULT", "PLUS", "ALWAYS"), (FLOAT, "PLUS", "MIN", "NQZERO"),
                                               (UINT, "BAND", "BOR", "NQZERO"), (FLOAT, "MULT", "PLUS", "NQZERO")])
@pytest.mark.parametrize("slots,col_phases,row_classes,given", [(16, 2, 1, True), (64, 3, 0, False), (1024, 1, 1, True), (45056, 4, 1, True)])
def test_pull_two_part_product(backend, oracle, dtype, om, oa, osel, slots, col_phases, row_classes, given):
    """splacu_mxv_masked_part: prologue + hub classes (fed with v[hub_cols] gathered by the caller or with v itself) followed by
    the rest (row classes, tail classes, fix-ups, gated CSR pass) equals the one-call product -- what the overlapped multi-GPU step relies
    on. Forced classes on a small skewed matrix; the last parameter set has no classes (part 1 is a no-op, part 2 the whole product)."""
    rng = np.random.default_rng(zlib.crc32(repr(("2part", dtype, om, oa, slots)).encode()))
    n_rows, n_cols = 2900, 3000
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = _skewed_csr(rng, dtype, n_rows, n_cols, kind)
    try:
        if slots < 45056:
            backend.set_option("mxv_hub", 3)
        backend.set_option("mxv_phase_slots", slots)
        backend.set_option("mxv_phases", col_phases)
        backend.set_option("mxv_hub_min_count", 1)
        backend.set_option("mxv_row_classes", row_classes)
        backend.set_option("mxv_row_min_count", 4)
        backend.set_option("mxv_row_min_nnz", 0)
        M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
        backend.set_option("mxv_phases", 4)
        backend.set_option("mxv_hub_min_count", 16)
        backend.set_option("mxv_row_classes", 1)
        backend.set_option("mxv_row_min_count", 64)
        backend.set_option("mxv_row_min_nnz", 25165824)
    hub_cols = backend.csr_hub_cols(M)
    assert (hub_cols.numel() > 0) == (slots < 45056)
    for rep in range(3):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        if rep == 2:
            mask[rng.random(n_rows) < 0.8] = 0  # sparse selection: the gated CSR pass of part 2 does the work
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else (0 if rep == 0 else 3)
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, False)
        dv, dm = to_dev(v, backend), to_dev(mask, backend)
        out = backend.empty(n_rows, like=dv)
        hub_vals = None
        if given and hub_cols.numel():
            hub_vals = backend.empty(hub_cols.numel(), like=dv)
            backend.v_gather(hub_cols, dv, hub_vals)
            # the scatter is the inverse on a permutation
            perm = torch.randperm(hub_cols.numel(), device=backend.device).to(torch.int32)
            tmp = backend.empty(hub_cols.numel(), like=dv)
            backend.v_gather(perm, hub_vals, tmp)
            back = backend.empty(hub_cols.numel(), like=dv)
            backend.v_scatter(perm, tmp, back)
            backend.sync()
            assert torch.equal(back.view(torch.int32), hub_vals.view(torch.int32))
        if rep == 1:  # prologue on its own, then the hub classes
            backend.mxv_masked_part(M, 4, None, None, dm, om, oa, osel, init, out)
            backend.mxv_masked_part(M, 1, dv if hub_vals is None else None, hub_vals, dm, om, oa, osel, init, out)
        else:
            backend.mxv_masked_part(M, 4 | 1, dv if hub_vals is None else None, hub_vals, dm, om, oa, osel, init, out)
        backend.mxv_masked_part(M, 2, dv, None, dm, om, oa, osel, init, out)
        backend.sync()
        assert_values(to_np(out, cases.NP[dtype]), want, cases.exact_expected(dtype, om, oa), what=f"two-part mxv rep {rep}",
                      bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))


@pytest.mark.gpu
def test_v_push_peers_local(backend):
    """splacu_v_push_peers with every 'peer' buffer on this device: segment q of the element list lands in buffer q at the given slots."""
    g = torch.Generator(device=backend.device)
    g.manual_seed(5)
    n_src, n_peers, n_dst = 5000, 3, 700
    src = torch.randint(-1000, 1000, (n_src,), generator=g, device=backend.device, dtype=torch.int32)
    counts = [300, 0, 700]
    seg_off = torch.tensor([0, 300, 300, 1000], dtype=torch.int32, device=backend.device)
    src_idx = torch.randint(0, n_src, (1000,), generator=g, device=backend.device, dtype=torch.int32)
    dst_idx = torch.cat([torch.randperm(n_dst, generator=g, device=backend.device)[:c] for c in counts]).to(torch.int32)
    bufs = [torch.full((n_dst,), -7, dtype=torch.int32, device=backend.device) for _ in range(n_peers)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=backend.device)
    backend.v_push_peers(src_idx, dst_idx, seg_off, ptrs, src)
    backend.sync()
    for q in range(n_peers):
        want = torch.full((n_dst,), -7, dtype=torch.int32, device=backend.device)
        a, b = int(seg_off[q]), int(seg_off[q + 1])
        want[dst_idx[a:b].long()] = src[src_idx[a:b].long()]
        assert torch.equal(bufs[q], want), f"buffer {q}"
