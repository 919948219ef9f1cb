"""User-defined ops: the source text "(T a, T b) { ... }" of an OpBinary::make_* / OpSelect::make_* op (reference
src/op.cpp:294-342, tests/test_vector.cpp:299-302) is compiled with NVRTC inside libsplacu (spla_b200/csrc/jit.cu).

CPU part (no GPU): NVRTC cross-compiles for sm_100a, so every generated module can be BUILT here; a broken source must come back as
SPLACU_E_COMPILE with the compiler's message. GPU part: the compiled generic kernels against the oracle / a plain sequential
restatement, bit for bit (user ops keep the CPU backend's sequential semantics)."""
import ctypes as C

import numpy as np
import pytest

import cases
from cases import FLOAT, INT, UINT

CT = {INT: "int", UINT: "uint", FLOAT: "float"}


@pytest.fixture(scope="module")
def lib():
    from spla_b200.backend import load_library

    return load_library()


def _compile(lib, dtype, mult, add, sel):
    from spla_b200.backend import BIN, SEL, make_op

    ops = [None if o is None else make_op(o, t) for o, t in ((mult, BIN), (add, BIN), (sel, SEL))]
    n = C.c_size_t(0)
    rc = lib.splacu_jit_compile(dtype, *[C.byref(o) if o is not None else None for o in ops], C.byref(n))
    return rc, n.value, lib.splacu_last_error().decode()


def test_every_builtin_op_has_device_source(lib):
    """A module may mix user ops with built-ins: the generated text of every built-in binary / select op must compile for every type."""
    for dtype in (INT, UINT, FLOAT):
        for k, op in enumerate(cases.BIN_OPS):
            if not cases.op_valid(dtype, op):
                continue
            user = ("keep_a", f"({CT[dtype]} a, {CT[dtype]} b) {{ return a; }}")
            rc, size, err = _compile(lib, dtype, op, user, cases.SEL_OPS[k % 8])
            assert rc == 0 and size > 1000, (dtype, op, err)


def test_user_sources_compile_and_are_cached(lib):
    blend = ("blend", "(float a, float b) { return 0.25f * a + 0.75f * b; }")  # reference tests/test_vector.cpp:299-302
    rc, size, err = _compile(lib, FLOAT, "MULT", blend, "ALWAYS")
    assert rc == 0 and size > 1000, err
    c0 = C.c_uint64(0)
    lib.splacu_jit_compile_count(C.byref(c0))
    rc, _, _ = _compile(lib, FLOAT, "MULT", blend, "ALWAYS")
    c1 = C.c_uint64(0)
    lib.splacu_jit_compile_count(C.byref(c1))
    assert rc == 0 and c1.value == c0.value  # second request: cache hit
    rc, _, err = _compile(lib, UINT, ("m", "(uint a, uint b) { return min(a, b) + 1u; }"), ("x", "(uint a, uint b) { return a ^ (b << 1); }"),
                          ("odd", "(uint a) { return (a & 1u) != 0u; }"))
    assert rc == 0, err
    rc, _, err = _compile(lib, INT, None, ("my_plus", "(int a, int b) { return a + b + 1; }"), None)  # neighbour tasks: one op, the other slots unused
    assert rc == 0, err


def test_broken_source_reports_compilation_error(lib):
    rc, _, err = _compile(lib, INT, "MULT", ("broken", "(int a, int b) { return a +* b; }"), "EQZERO")
    assert rc == -5 and "error" in err.lower(), (rc, err)  # SPLACU_E_COMPILE + the NVRTC log
    rc, _, err = _compile(lib, FLOAT, "BOR", ("x", "(float a, float b) { return a; }"), "EQZERO")
    assert rc == -1, (rc, err)  # a built-in that is not defined for the type is refused before compiling


# ---- GPU: the compiled kernels ------------------------------------------------------------------------------------------
def _seq_mxv(Ap, Aj, Ax, v, take, init, mult, add, early_exit):
    r = np.empty(len(Ap) - 1, dtype=Ax.dtype)
    for i in range(len(r)):
        s = init
        if take[i]:
            for k in range(Ap[i], Ap[i + 1]):
                s = add(s, mult(Ax[k], v[Aj[k]]))
                if early_exit and s != init:
                    break
        r[i] = s
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [INT, UINT, FLOAT])
def test_user_ops_that_restate_builtins_match_the_oracle(backend, oracle, dtype):
    """The generic (NVRTC) kernels against the oracle: user ops that spell built-ins out under their own names take the JIT path
    (id < 0) and must reproduce the oracle bit for bit -- mxv with and without early exit, vxm, assign, eadd, eadd_fdb."""
    from gpu_util import idx_dev, make_csr, to_dev, to_np

    t = CT[dtype]
    plus = ("u_plus", f"({t} a, {t} b) {{ return a + b; }}")
    mult = ("u_mult", f"({t} a, {t} b) {{ return a * b; }}")
    umin = ("u_min", f"({t} a, {t} b) {{ return (b < a) ? b : a; }}")
    eqz = ("u_eqzero", f"({t} a) {{ return a == 0; }}")
    rng = np.random.default_rng(100 + dtype)
    n_rows, n_cols = 700, 600
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, 7, skew=True)
    M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    np_t = cases.NP[dtype]
    for (om, oa, osel), (bm, ba, bs) in (((mult, plus, eqz), ("MULT", "PLUS", "EQZERO")), ((plus, umin, "NQZERO"), ("PLUS", "MIN", "NQZERO")),
                                          (("MULT", plus, "ALWAYS"), ("MULT", "PLUS", "ALWAYS"))):
        for ee in (False, True):
            v = cases.rand_values(rng, dtype, n_cols)
            mask = cases.rand_values(rng, dtype, n_rows)
            init = cases.rand_values(rng, dtype, 1)[0]
            want = oracle.mxv_masked(dtype, bm, ba, bs, Ap, Aj, Ax, v, mask, init, ee)
            got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init, early_exit=ee)
            backend.sync()
            assert np.array_equal(to_np(got, np_t).view(np.uint32), want.view(np.uint32)), (bm, ba, bs, ee)  # sequential fold: bit-exact, FLOAT too
        for nv in (0, 1, 150, n_rows):
            vi, vx = cases.rand_frontier(rng, dtype, n_rows, nv)
            maskc = cases.rand_values(rng, dtype, n_cols)
            wi, wx = oracle.vxm_masked(dtype, bm, ba, bs, Ap, Aj, Ax, n_cols, vi, vx, maskc)
            gi, gx = backend.vxm_masked(M, idx_dev(vi, backend), to_dev(vx, backend), to_dev(maskc, backend), om, oa, osel)
            backend.sync()
            assert np.array_equal(to_np(gi, np.uint32), wi), (bm, ba, bs, nv)
            assert np.array_equal(to_np(gx, np_t).view(np.uint32), wx.view(np.uint32)), (bm, ba, bs, nv)
    # neighbours
    r = cases.rand_values(rng, dtype, n_cols)
    mask = cases.rand_values(rng, dtype, n_cols)
    want = oracle.v_assign_masked_dense(dtype, "PLUS", "EQZERO", r, mask, 3)
    got = backend.v_assign_masked(to_dev(r, backend), to_dev(mask, backend), 3, plus, eqz)
    backend.sync()
    assert np.array_equal(to_np(got, np_t), want)
    mi, mx = cases.rand_frontier(rng, dtype, n_cols, 90)
    want = oracle.v_assign_masked_sparse(dtype, "PLUS", "EQZERO", r, mi, mx, 2)
    got = backend.v_assign_masked(to_dev(r, backend), (idx_dev(mi, backend), to_dev(mx, backend)), 2, plus, eqz)
    backend.sync()
    assert np.array_equal(to_np(got, np_t), want)
    u, w = cases.rand_values(rng, dtype, n_cols), cases.rand_values(rng, dtype, n_cols)
    got = backend.v_eadd(to_dev(u, backend), to_dev(w, backend), umin)
    backend.sync()
    assert np.array_equal(to_np(got, np_t), oracle.v_eadd_dense(dtype, "MIN", u, w))
    want_r, want_f = oracle.v_eadd_fdb_dense(dtype, "MIN", u, w, 9)
    d_r = to_dev(u, backend)
    fdb = backend.v_eadd_fdb_dense(d_r, to_dev(w, backend), umin, 9)
    backend.sync()
    assert np.array_equal(to_np(d_r, np_t), want_r) and np.array_equal(to_np(fdb, np_t), want_f)
    vi, vx = cases.rand_frontier(rng, dtype, n_cols, 200)
    want_r, want_fi, want_fx = oracle.v_eadd_fdb_sparse(dtype, "MIN", u, vi, vx)
    d_r = to_dev(u, backend)
    fi, fx = backend.v_eadd_fdb_sparse(d_r, idx_dev(vi, backend), to_dev(vx, backend), umin)
    backend.sync()
    assert np.array_equal(to_np(d_r, np_t), want_r) and np.array_equal(to_np(fi, np.uint32), want_fi) and np.array_equal(to_np(fx, np_t), want_fx)


@pytest.mark.gpu
def test_genuinely_new_ops_match_a_sequential_restatement(backend):
    """Ops no built-in provides (the reference's own example: 0.25 a + 0.75 b, tests/test_vector.cpp:299-302), against a plain
    sequential fold in numpy float32 scalars: bit-exact, which also proves that mult and add round separately (no fma)."""
    from gpu_util import make_csr, to_dev, to_np

    rng = np.random.default_rng(7)
    n_rows, n_cols = 300, 250
    Ap, Aj, Ax = cases.rand_csr(rng, FLOAT, n_rows, n_cols, 6, skew=True, kind="positive")
    v = cases.rand_values(rng, FLOAT, n_cols, "positive")
    mask = cases.rand_values(rng, FLOAT, n_rows)
    f32 = np.float32
    blend = ("blend", "(float a, float b) { return 0.25f * a + 0.75f * b; }")
    M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    for ee in (False, True):
        got = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), "MULT", blend, "GTZERO", 1.5, early_exit=ee)
        backend.sync()
        want = _seq_mxv(Ap.astype(np.int64), Aj, Ax, v, mask > 0, f32(1.5), lambda a, b: f32(a * b), lambda a, b: f32(f32(f32(0.25) * a) + f32(f32(0.75) * b)), ee)
        assert np.array_equal(to_np(got, np.float32).view(np.uint32), want.view(np.uint32)), ee
    # reference tests/test_vector.cpp:285-315 (eadd_fdb_custom), on the device
    N = 10000
    a = np.arange(N, dtype=np.float32)
    b = (np.float32(N) - a * a).astype(np.float32)
    d_r = to_dev(a, backend)
    backend.v_eadd_fdb_dense(d_r, to_dev(b, backend), blend, 0.0)
    backend.sync()
    want = (np.float32(0.25) * a).astype(np.float32) + (np.float32(0.75) * b).astype(np.float32)
    assert np.array_equal(to_np(d_r, np.float32).view(np.uint32), want.astype(np.float32).view(np.uint32))
