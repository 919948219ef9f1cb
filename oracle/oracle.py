"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two checkers built by oracle/Makefile:

* ``Oracle``   -- libspla_oracle.so, the plain-C restatement (oracle/spla_oracle.c) of spla's CPU
                  backend for masked mxv / vxm (reference src/cpu/cpu_mxv.hpp:56-106,
                  src/cpu/cpu_vxm.hpp:58-128) and the neighbour ops.
* ``RefSpla``  -- oracle/_ref/libspla_refshim.so, a flat C layer (oracle/ref_shim.cpp) over the
                  UNMODIFIED reference library compiled from /root/reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg import this
module. The product package ``spla_b200`` never does.

All arrays are numpy, values travel as 4-byte patterns: int32 / uint32 / float32 arrays are viewed as
uint32 on the way in and viewed back on the way out.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

INT, UINT, FLOAT = 0, 1, 2
NP_DTYPE = {INT: np.int32, UINT: np.uint32, FLOAT: np.float32}

BIN_OPS = ["PLUS", "MINUS", "MULT", "DIV", "MINUS_POW2", "FIRST", "SECOND", "BONE",
           "MIN", "MAX", "LOR", "LAND", "BOR", "BAND", "BXOR"]
SEL_OPS = ["EQZERO", "NQZERO", "GTZERO", "GEZERO", "LTZERO", "LEZERO", "ALWAYS", "NEVER"]
BIN = {n: i for i, n in enumerate(BIN_OPS)}
SEL = {n: i for i, n in enumerate(SEL_OPS)}

_u32p = C.POINTER(C.c_uint32)


def _p(a):
    return a.ctypes.data_as(_u32p)


def _u32(a):
    """contiguous uint32 view of a 4-byte numpy array"""
    a = np.ascontiguousarray(a)
    assert a.dtype.itemsize == 4, a.dtype
    return a.view(np.uint32)


def _bits(dtype, x):
    return int(np.array([x], dtype=NP_DTYPE[dtype]).view(np.uint32)[0])


def build(target="oracle"):
    """Compile the checkers (gcc/g++ only). `ref` is a no-op when /root/reference is absent."""
    subprocess.check_call(["make", "-s", "-C", HERE, target])


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libspla_oracle.so")
        if not os.path.exists(path):
            build("oracle")
        self.lib = lib = C.CDLL(path)
        lib.orc_binary.restype = C.c_uint32
        lib.orc_binary.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        lib.orc_select.restype = C.c_int
        lib.orc_select.argtypes = [C.c_int, C.c_int, C.c_uint32]
        lib.orc_mxv_masked.restype = C.c_int
        lib.orc_mxv_masked.argtypes = [C.c_int] * 4 + [C.c_uint32, _u32p, _u32p, _u32p, _u32p, _u32p,
                                                      C.c_uint32, C.c_int, _u32p]
        lib.orc_vxm_masked.restype = C.c_int64
        lib.orc_vxm_masked.argtypes = [C.c_int] * 4 + [C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32,
                                                      _u32p, _u32p, _u32p, _u32p, _u32p]
        lib.orc_v_assign_masked_dense.restype = None
        lib.orc_v_assign_masked_dense.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, _u32p, _u32p, C.c_uint32]
        lib.orc_v_assign_masked_sparse.restype = None
        lib.orc_v_assign_masked_sparse.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, C.c_uint32, _u32p, _u32p, C.c_uint32]
        lib.orc_v_count_mf_dense.restype = C.c_uint32
        lib.orc_v_count_mf_dense.argtypes = [C.c_uint32, _u32p, C.c_uint32, C.c_int]
        lib.orc_v_eadd_fdb_sparse.restype = C.c_uint32
        lib.orc_v_eadd_fdb_sparse.argtypes = [C.c_int, C.c_int, _u32p, C.c_uint32, _u32p, _u32p, _u32p, _u32p]
        lib.orc_v_eadd_fdb_dense.restype = None
        lib.orc_v_eadd_fdb_dense.argtypes = [C.c_int, C.c_int, C.c_uint32, _u32p, _u32p, _u32p, C.c_uint32]
        lib.orc_v_eadd_dense.restype = None
        lib.orc_v_eadd_dense.argtypes = [C.c_int, C.c_int, C.c_uint32, _u32p, _u32p, _u32p]
        lib.orc_v_reduce_dense.restype = C.c_uint32
        lib.orc_v_reduce_dense.argtypes = [C.c_int, C.c_int, C.c_uint32, _u32p, C.c_uint32]

    def binary(self, dtype, op, a, b):
        r = self.lib.orc_binary(dtype, BIN[op], _bits(dtype, a), _bits(dtype, b))
        return np.array([r], dtype=np.uint32).view(NP_DTYPE[dtype])[0]

    def mxv_masked(self, dtype, op_mult, op_add, op_select, Ap, Aj, Ax, v, mask, init, early_exit=False):
        n_rows = len(Ap) - 1
        Ap, Aj, Ax, v, mask = map(_u32, (Ap, Aj, Ax, v, mask))
        r = np.empty(n_rows, dtype=np.uint32)
        rc = self.lib.orc_mxv_masked(dtype, BIN[op_mult], BIN[op_add], SEL[op_select], n_rows,
                                     _p(Ap), _p(Aj), _p(Ax), _p(v), _p(mask), _bits(dtype, init),
                                     int(early_exit), _p(r))
        assert rc == 0
        return r.view(NP_DTYPE[dtype])

    def vxm_masked(self, dtype, op_mult, op_add, op_select, Ap, Aj, Ax, n_cols, vi, vx, mask):
        Ap, Aj, Ax, vi, vx, mask = map(_u32, (Ap, Aj, Ax, vi, vx, mask))
        assert len(mask) == n_cols
        ri = np.empty(max(n_cols, 1), dtype=np.uint32)
        rx = np.empty(max(n_cols, 1), dtype=np.uint32)
        nr = self.lib.orc_vxm_masked(dtype, BIN[op_mult], BIN[op_add], SEL[op_select], n_cols,
                                     _p(Ap), _p(Aj), _p(Ax), len(vi), _p(vi), _p(vx), _p(mask), _p(ri), _p(rx))
        assert nr >= 0
        return ri[:nr].copy(), rx[:nr].copy().view(NP_DTYPE[dtype])

    def v_assign_masked_dense(self, dtype, op_assign, op_select, r, mask, value):
        r = _u32(r).copy()
        mask = _u32(mask)
        self.lib.orc_v_assign_masked_dense(dtype, BIN[op_assign], SEL[op_select], len(r), _p(r), _p(mask), _bits(dtype, value))
        return r.view(NP_DTYPE[dtype])

    def v_assign_masked_sparse(self, dtype, op_assign, op_select, r, mi, mx, value):
        r = _u32(r).copy()
        mi, mx = _u32(mi), _u32(mx)
        self.lib.orc_v_assign_masked_sparse(dtype, BIN[op_assign], SEL[op_select], _p(r), len(mi), _p(mi), _p(mx), _bits(dtype, value))
        return r.view(NP_DTYPE[dtype])

    def v_count_mf_dense(self, dtype, v, fill):
        v = _u32(v)
        return int(self.lib.orc_v_count_mf_dense(len(v), _p(v), _bits(dtype, fill), dtype))

    def v_eadd_fdb_sparse(self, dtype, op, r, vi, vx):
        r = _u32(r).copy()
        vi, vx = _u32(vi), _u32(vx)
        fi = np.empty(max(len(vi), 1), dtype=np.uint32)
        fx = np.empty(max(len(vi), 1), dtype=np.uint32)
        nf = self.lib.orc_v_eadd_fdb_sparse(dtype, BIN[op], _p(r), len(vi), _p(vi), _p(vx), _p(fi), _p(fx))
        return r.view(NP_DTYPE[dtype]), fi[:nf].copy(), fx[:nf].copy().view(NP_DTYPE[dtype])

    def v_eadd_fdb_dense(self, dtype, op, r, v, fdb_fill):
        r = _u32(r).copy()
        v = _u32(v)
        fdb = np.empty(len(r), dtype=np.uint32)
        self.lib.orc_v_eadd_fdb_dense(dtype, BIN[op], len(r), _p(r), _p(v), _p(fdb), _bits(dtype, fdb_fill))
        return r.view(NP_DTYPE[dtype]), fdb.view(NP_DTYPE[dtype])

    def v_eadd_dense(self, dtype, op, u, v):
        u, v = _u32(u), _u32(v)
        r = np.empty(len(u), dtype=np.uint32)
        self.lib.orc_v_eadd_dense(dtype, BIN[op], len(u), _p(r), _p(u), _p(v))
        return r.view(NP_DTYPE[dtype])

    def v_reduce_dense(self, dtype, op, v, init):
        v = _u32(v)
        s = self.lib.orc_v_reduce_dense(dtype, BIN[op], len(v), _p(v), _bits(dtype, init))
        return np.array([s], dtype=np.uint32).view(NP_DTYPE[dtype])[0]


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libspla_refshim.so"))


class RefSpla:
    """The unmodified reference CPU backend (oracle/_ref), driven through oracle/ref_shim.cpp."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libspla_refshim.so")
        if not os.path.exists(path):
            build("ref")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libspla_refshim.so missing and /root/reference absent")
        self.lib = lib = C.CDLL(path)
        lib.refshim_matrix_create.restype = C.c_void_p
        lib.refshim_matrix_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, _u32p, _u32p, _u32p]
        lib.refshim_matrix_free.restype = None
        lib.refshim_matrix_free.argtypes = [C.c_void_p]
        lib.refshim_mxv_masked.restype = C.c_int
        lib.refshim_mxv_masked.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p, _u32p, C.c_uint32, C.c_int,
                                           _u32p, C.c_int, C.POINTER(C.c_double)]
        lib.refshim_vxm_masked.restype = C.c_int
        lib.refshim_vxm_masked.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, _u32p, _u32p, _u32p,
                                           C.c_uint32, _u32p, _u32p, C.c_uint32, _u32p, C.c_int, C.POINTER(C.c_double)]
        lib.refshim_bfs.restype = C.c_int
        lib.refshim_bfs.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        lib.refshim_sssp.restype = C.c_int
        lib.refshim_sssp.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_double)]
        lib.refshim_pr.restype = C.c_int
        lib.refshim_pr.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_double)]

    class Matrix:
        def __init__(self, ref, dtype, n_rows, n_cols, Ai, Aj, Ax):
            self.ref, self.dtype, self.n_rows, self.n_cols = ref, dtype, n_rows, n_cols
            Ai, Aj, Ax = map(_u32, (Ai, Aj, Ax))
            self.h = ref.lib.refshim_matrix_create(dtype, n_rows, n_cols, len(Ai), _p(Ai), _p(Aj), _p(Ax))
            assert self.h

        def __del__(self):
            if getattr(self, "h", None):
                self.ref.lib.refshim_matrix_free(self.h)
                self.h = None

    def matrix(self, dtype, n_rows, n_cols, Ai, Aj, Ax):
        return RefSpla.Matrix(self, dtype, n_rows, n_cols, Ai, Aj, Ax)

    def matrix_from_csr(self, dtype, n_cols, Ap, Aj, Ax):
        Ap = np.asarray(Ap, dtype=np.uint32)
        n_rows = len(Ap) - 1
        Ai = np.repeat(np.arange(n_rows, dtype=np.uint32), np.diff(Ap.astype(np.int64)))
        return self.matrix(dtype, n_rows, n_cols, Ai, Aj, Ax)

    def mxv_masked(self, M, op_mult, op_add, op_select, v, mask, init, early_exit=False, repeats=0):
        v, mask = _u32(v), _u32(mask)
        assert len(v) == M.n_cols and len(mask) == M.n_rows
        r = np.empty(M.n_rows, dtype=np.uint32)
        sec = C.c_double(0.0)
        rc = self.lib.refshim_mxv_masked(M.h, BIN[op_mult], BIN[op_add], SEL[op_select], _p(v), _p(mask),
                                         _bits(M.dtype, init), int(early_exit), _p(r), repeats, C.byref(sec))
        assert rc == 0, rc
        r = r.view(NP_DTYPE[M.dtype])
        return (r, sec.value) if repeats else r

    def vxm_masked(self, M, op_mult, op_add, op_select, vi, vx, mask, init=0, repeats=0):
        vi, vx, mask = _u32(vi), _u32(vx), _u32(mask)
        assert len(mask) == M.n_cols
        cap = M.n_cols
        ri = np.empty(max(cap, 1), dtype=np.uint32)
        rx = np.empty(max(cap, 1), dtype=np.uint32)
        nr = C.c_uint32(0)
        sec = C.c_double(0.0)
        rc = self.lib.refshim_vxm_masked(M.h, BIN[op_mult], BIN[op_add], SEL[op_select], len(vi), _p(vi), _p(vx),
                                         _p(mask), _bits(M.dtype, init), _p(ri), _p(rx), cap, C.byref(nr),
                                         repeats, C.byref(sec))
        assert rc == 0, rc
        out = (ri[:nr.value].copy(), rx[:nr.value].copy().view(NP_DTYPE[M.dtype]))
        return (out + (sec.value,)) if repeats else out

    def bfs(self, M, source, mode=2, front_factor=0.05):
        d = np.zeros(M.n_rows, dtype=np.int32)
        sec = C.c_double(0.0)
        rc = self.lib.refshim_bfs(M.h, source, mode, front_factor, d.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(sec))
        assert rc == 0, rc
        return d, sec.value

    def sssp(self, M, source, mode=2, front_factor=0.05):
        d = np.zeros(M.n_rows, dtype=np.float32)
        sec = C.c_double(0.0)
        rc = self.lib.refshim_sssp(M.h, source, mode, front_factor, d.ctypes.data_as(C.POINTER(C.c_float)), C.byref(sec))
        assert rc == 0, rc
        return d, sec.value

    def pr(self, M, alpha=0.85, eps=1e-6):
        p = np.zeros(M.n_rows, dtype=np.float32)
        sec = C.c_double(0.0)
        rc = self.lib.refshim_pr(M.h, alpha, eps, p.ctypes.data_as(C.POINTER(C.c_float)), C.byref(sec))
        assert rc == 0, rc
        return p, sec.value
