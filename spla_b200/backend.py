"""Host-side Python binding of the splacu C ABI (include/splacu.h) for the bench harness and the parity tests.

The product is the CUDA library; this module only marshals device pointers. torch is plumbing here (device
memory, streams, torch.distributed) -- no arithmetic of the hot path happens in torch, and there is NO fallback:
if libsplacu.so is missing or no CUDA device is present, construction fails loudly.

The call surface mirrors the reference's own entry points for the path:
    Backend.mxv_masked(...)  <->  spla::exec_mxv_masked / spla_Exec_mxv_masked  (reference include/spla/exec.hpp:157-167, include/spla.h:372)
    Backend.vxm_masked(...)  <->  spla::exec_vxm_masked / spla_Exec_vxm_masked  (reference include/spla/exec.hpp:187-197, include/spla.h:373)
with the same argument meaning (r, mask, M, v, op_multiply, op_add, op_select, init, desc.early_exit).
"""
import ctypes as C
import os

import torch

from . import build as _build

INT, UINT, FLOAT = 0, 1, 2

BIN_OPS = ["PLUS", "MINUS", "MULT", "DIV", "MINUS_POW2", "FIRST", "SECOND", "BONE",
           "MIN", "MAX", "LOR", "LAND", "BOR", "BAND", "BXOR"]
SEL_OPS = ["EQZERO", "NQZERO", "GTZERO", "GEZERO", "LTZERO", "LEZERO", "ALWAYS", "NEVER"]
BIN = {n: i for i, n in enumerate(BIN_OPS)}
SEL = {n: i for i, n in enumerate(SEL_OPS)}

# every symbol include/splacu.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "splacu_init", "splacu_finalize", "splacu_device_count", "splacu_device_name", "splacu_sm_count",
    "splacu_default_stream", "splacu_sync", "splacu_last_error", "splacu_launch_count",
    "splacu_set_option", "splacu_get_option", "splacu_csr_info", "splacu_csr_phases", "splacu_csr_row_classes",
    "splacu_malloc", "splacu_free", "splacu_malloc_host", "splacu_free_host",
    "splacu_memcpy_h2d", "splacu_memcpy_d2h", "splacu_memcpy_d2d", "splacu_fill", "splacu_publish_window",
    "splacu_csr_create", "splacu_csr_destroy", "splacu_mxv_masked",
    "splacu_workspace_create", "splacu_workspace_destroy", "splacu_workspace_reset", "splacu_workspace_info",
    "splacu_vxm_masked_begin", "splacu_vxm_masked_emit", "splacu_vxm_masked",
    "splacu_coo_to_dense", "splacu_dense_to_coo_count", "splacu_dense_to_coo_emit",
    "splacu_v_assign_masked_dense", "splacu_v_assign_masked_sparse", "splacu_v_count_mf_dense",
    "splacu_v_pack_bits", "splacu_v_unpack_bits",
    "splacu_v_eadd_fdb_dense", "splacu_v_eadd_fdb_sparse_begin", "splacu_v_eadd_fdb_sparse_emit",
    "splacu_v_eadd_dense", "splacu_v_reduce_dense",
    "splacu_mxv_masked_ops", "splacu_vxm_masked_begin_ops", "splacu_v_assign_masked_dense_ops", "splacu_v_assign_masked_sparse_ops",
    "splacu_coo_to_csr", "splacu_profile_enable", "splacu_profile_reset", "splacu_profile_dump",
    "splacu_vxm_masked_begin_async", "splacu_vxm_masked_begin_finish",
    "splacu_dist_create", "splacu_dist_destroy", "splacu_dist_info", "splacu_dcsr_create", "splacu_dcsr_destroy", "splacu_dcsr_bounds",
    "splacu_dist_mxv_masked", "splacu_dist_vxm_masked_begin", "splacu_dist_vxm_masked_emit",
    "splacu_v_eadd_dense_op", "splacu_v_eadd_fdb_dense_op", "splacu_v_eadd_fdb_sparse_begin_op", "splacu_jit_compile", "splacu_jit_compile_count",
    "splacu_csr_hub_cols", "splacu_mxv_masked_part", "splacu_v_gather", "splacu_v_scatter", "splacu_v_push_peers",
]


class Op(C.Structure):
    """splacu_op: a built-in id, or id = -1 with (name, source text "(T a, T b) { ... }") of a user-defined op"""
    _fields_ = [("id", C.c_int), ("name", C.c_char_p), ("source", C.c_char_p)]


def make_op(op, table):
    """'PLUS' -> built-in; ('my_plus', '(int a, int b) { return a + b + 1; }') -> user-defined (compiled with NVRTC at first use)"""
    if isinstance(op, str):
        return Op(table[op], None, None)
    name, source = op
    return Op(-1, name.encode(), source.encode())


def _user(*ops):
    return any(not isinstance(o, str) for o in ops)


class SplacuError(RuntimeError):
    pass


def load_library(build_if_missing=True):
    """dlopen spla_b200/lib/libsplacu.so; build it with nvcc first if it is not there. Never falls back."""
    path = os.environ.get("SPLACU_LIB") or _build.LIB  # SPLACU_LIB: an experimental build of the same sources (tools/ A/B runs)
    if not os.path.exists(path):
        if not build_if_missing:
            raise SplacuError(f"{path} is missing: run `python -m spla_b200.build`")
        _build.build_splacu()
    lib = C.CDLL(path)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
    pu32 = C.POINTER(C.c_uint32)
    pop = C.POINTER(Op)
    sig = {
        "splacu_init": [i32], "splacu_finalize": [], "splacu_device_count": [C.POINTER(C.c_int)],
        "splacu_device_name": [C.c_char_p, i32], "splacu_sm_count": [C.POINTER(C.c_int)],
        "splacu_sync": [vp], "splacu_launch_count": [C.POINTER(C.c_uint64)],
        "splacu_set_option": [C.c_char_p, C.c_int64], "splacu_get_option": [C.c_char_p, C.POINTER(C.c_int64)],
        "splacu_csr_info": [vp, pu32, pu32], "splacu_csr_phases": [vp, C.POINTER(C.c_int), pu32, i32],
        "splacu_csr_row_classes": [vp, C.POINTER(C.c_int), pu32, pu32, i32],
        "splacu_malloc": [C.POINTER(vp), sz], "splacu_free": [vp],
        "splacu_malloc_host": [C.POINTER(vp), sz], "splacu_free_host": [vp],
        "splacu_memcpy_h2d": [vp, vp, sz, vp], "splacu_memcpy_d2h": [vp, vp, sz, vp], "splacu_memcpy_d2d": [vp, vp, sz, vp],
        "splacu_fill": [vp, u32, sz, vp], "splacu_publish_window": [C.POINTER(vp), i32, i32, sz, sz, vp],
        "splacu_csr_create": [C.POINTER(vp), u32, u32, u32, vp, vp, vp, vp], "splacu_csr_destroy": [vp],
        "splacu_mxv_masked": [vp, i32, i32, i32, i32, vp, vp, vp, u32, i32, vp],
        "splacu_workspace_create": [C.POINTER(vp)], "splacu_workspace_destroy": [vp],
        "splacu_workspace_reset": [vp, vp], "splacu_workspace_info": [vp, C.POINTER(C.c_int)],
        "splacu_vxm_masked_begin": [vp, i32, i32, i32, i32, u32, vp, vp, vp, vp, pu32, vp],
        "splacu_vxm_masked_emit": [vp, vp, vp, vp],
        "splacu_vxm_masked": [vp, i32, i32, i32, i32, u32, vp, vp, vp, vp, vp, u32, pu32, vp, vp],
        "splacu_coo_to_dense": [u32, u32, u32, vp, vp, vp, vp],
        "splacu_dense_to_coo_count": [i32, u32, u32, vp, vp, pu32, vp],
        "splacu_dense_to_coo_emit": [i32, u32, u32, vp, vp, vp, vp, vp],
        "splacu_v_assign_masked_dense": [i32, i32, i32, u32, vp, vp, u32, vp],
        "splacu_v_assign_masked_sparse": [i32, i32, i32, vp, u32, vp, vp, u32, vp],
        "splacu_v_count_mf_dense": [i32, u32, vp, u32, vp, pu32, vp],
        "splacu_v_pack_bits": [i32, i32, u32, vp, vp, vp], "splacu_v_unpack_bits": [u32, vp, u32, u32, vp, vp],
        "splacu_v_eadd_fdb_dense": [i32, i32, u32, vp, vp, vp, u32, vp],
        "splacu_v_eadd_fdb_sparse_begin": [i32, i32, vp, u32, vp, vp, vp, pu32, vp],
        "splacu_v_eadd_fdb_sparse_emit": [vp, vp, vp, vp],
        "splacu_v_eadd_dense": [i32, i32, u32, vp, vp, vp, vp],
        "splacu_v_reduce_dense": [i32, i32, u32, vp, u32, vp, pu32, vp],
        "splacu_mxv_masked_ops": [vp, i32, pop, pop, pop, vp, vp, vp, u32, i32, vp],
        "splacu_vxm_masked_begin_ops": [vp, i32, pop, pop, pop, u32, vp, vp, vp, vp, pu32, vp],
        "splacu_v_assign_masked_dense_ops": [i32, pop, pop, u32, vp, vp, u32, vp],
        "splacu_v_assign_masked_sparse_ops": [i32, pop, pop, vp, u32, vp, vp, u32, vp],
        "splacu_v_eadd_dense_op": [i32, pop, u32, vp, vp, vp, vp],
        "splacu_v_eadd_fdb_dense_op": [i32, pop, u32, vp, vp, vp, u32, vp],
        "splacu_v_eadd_fdb_sparse_begin_op": [i32, pop, vp, u32, vp, vp, vp, pu32, vp],
        "splacu_coo_to_csr": [u32, u32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp],
        "splacu_profile_enable": [i32], "splacu_profile_reset": [], "splacu_profile_dump": [C.c_char_p, i32],
        "splacu_vxm_masked_begin_async": [vp, i32, i32, i32, i32, u32, vp, vp, vp, vp, vp], "splacu_vxm_masked_begin_finish": [vp, pu32, vp],
        "splacu_dist_create": [C.POINTER(vp), i32, C.POINTER(C.c_int)], "splacu_dist_destroy": [vp], "splacu_dist_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "splacu_dcsr_create": [C.POINTER(vp), vp, u32, u32, u32, vp, vp, vp, vp], "splacu_dcsr_destroy": [vp],
        "splacu_dcsr_bounds": [vp, C.POINTER(C.c_int), pu32, pu32],
        "splacu_dist_mxv_masked": [vp, i32, i32, i32, i32, vp, vp, vp, u32, i32, vp],
        "splacu_dist_vxm_masked_begin": [vp, i32, i32, i32, i32, u32, vp, vp, vp, pu32, vp], "splacu_dist_vxm_masked_emit": [vp, vp, vp, vp],
        "splacu_jit_compile": [i32, pop, pop, pop, C.POINTER(C.c_size_t)], "splacu_jit_compile_count": [C.POINTER(C.c_uint64)],
        "splacu_csr_hub_cols": [vp, pu32, C.POINTER(vp)], "splacu_mxv_masked_part": [vp, i32, i32, i32, i32, vp, vp, vp, vp, u32, i32, vp],
        "splacu_v_gather": [u32, vp, vp, vp, vp], "splacu_v_scatter": [u32, vp, vp, vp, u32, vp],
        "splacu_v_push_peers": [u32, vp, vp, vp, u32, vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.splacu_default_stream.argtypes = []
    lib.splacu_default_stream.restype = C.c_void_p
    lib.splacu_last_error.argtypes = []
    lib.splacu_last_error.restype = C.c_char_p
    return lib


def dtype_code(t):
    if t.dtype == torch.float32:
        return FLOAT
    if t.dtype == torch.int32:
        return INT
    if t.dtype == torch.uint32:
        return UINT
    raise TypeError(f"spla values are 4-byte INT/UINT/FLOAT, got {t.dtype}")


def scalar_bits(code, x):
    if code == FLOAT:
        return int(torch.tensor([float(x)], dtype=torch.float32).view(torch.int32).item()) & 0xFFFFFFFF
    return int(x) & 0xFFFFFFFF


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class CsrMatrix:
    """Device-resident CSR (Ap uint32[n_rows+1], Aj uint32[nnz], Ax T[nnz]) plus the backend's load-balancing
    metadata; the counterpart of the reference's CLCsr decoration (src/opencl/cl_formats.hpp:93-103)."""

    def __init__(self, backend, n_rows, n_cols, Ap, Aj, Ax):
        assert Ap.is_cuda and Aj.is_cuda and Ax.is_cuda
        assert Ap.dtype == torch.int32 and Aj.dtype == torch.int32 and Ap.numel() == n_rows + 1
        self.backend, self.n_rows, self.n_cols = backend, n_rows, n_cols
        self.Ap, self.Aj, self.Ax = Ap.contiguous(), Aj.contiguous(), Ax.contiguous()
        self.nnz = Aj.numel()
        self.dtype = dtype_code(Ax)
        h = C.c_void_p()
        backend._check(backend.lib.splacu_csr_create(C.byref(h), n_rows, n_cols, self.nnz, _ptr(self.Ap), _ptr(self.Aj),
                                                     _ptr(self.Ax), backend.stream_ptr))
        self.handle = h

    def __del__(self):
        if getattr(self, "handle", None) is not None:
            try:
                self.backend.lib.splacu_csr_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


class DistGroup:
    def __init__(self, backend, device_ids):
        self.backend = backend
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = C.c_void_p()
        backend._check(backend.lib.splacu_dist_create(C.byref(h), len(device_ids), ids))
        self.handle, self.n = h, len(device_ids)
        n, nc = C.c_int(0), C.c_int(0)
        backend.lib.splacu_dist_info(h, C.byref(n), C.byref(nc))
        self.uses_nccl = bool(nc.value)

    def csr(self, M):
        """shard a CsrMatrix of this backend over the group"""
        return DistCsr(self, M)

    def __del__(self):
        if getattr(self, "handle", None) is not None:
            try:
                self.backend.lib.splacu_dist_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


class DistCsr:
    def __init__(self, group, M):
        self.group, self.M, self.backend = group, M, group.backend
        h = C.c_void_p()
        be = self.backend
        be._check(be.lib.splacu_dcsr_create(C.byref(h), group.handle, M.n_rows, M.n_cols, M.nnz, _ptr(M.Ap), _ptr(M.Aj), _ptr(M.Ax), be.stream_ptr))
        self.handle = h
        self._nr = C.c_uint32(0)

    def bounds(self):
        n = C.c_int(0)
        rb, cb = (C.c_uint32 * 17)(), (C.c_uint32 * 17)()
        self.backend._check(self.backend.lib.splacu_dcsr_bounds(self.handle, C.byref(n), rb, cb))
        return [int(rb[p]) for p in range(n.value + 1)], [int(cb[p]) for p in range(n.value + 1)]

    def mxv_masked(self, v, mask, op_mult, op_add, op_select, init, early_exit=False, out=None):
        be, M = self.backend, self.M
        code = M.dtype
        if out is None:
            out = be.empty(M.n_rows, like=v)
        be._check(be.lib.splacu_dist_mxv_masked(self.handle, code, BIN[op_mult], BIN[op_add], SEL[op_select], _ptr(v), _ptr(mask), _ptr(out),
                                                scalar_bits(code, init), int(bool(early_exit)), be.stream_ptr))
        return out

    def vxm_masked(self, vi, vx, mask, op_mult, op_add, op_select):
        be, M = self.backend, self.M
        code = M.dtype
        be._check(be.lib.splacu_dist_vxm_masked_begin(self.handle, code, BIN[op_mult], BIN[op_add], SEL[op_select], vi.numel(), _ptr(vi), _ptr(vx),
                                                      _ptr(mask), C.byref(self._nr), be.stream_ptr))
        nr = self._nr.value
        with torch.cuda.stream(be.stream):
            ri = torch.empty(nr, dtype=torch.int32, device=be.device)
            rx = torch.empty(nr, dtype=M.Ax.dtype, device=be.device)
        be._check(be.lib.splacu_dist_vxm_masked_emit(self.handle, _ptr(ri), _ptr(rx), be.stream_ptr))
        return ri, rx

    def __del__(self):
        if getattr(self, "handle", None) is not None:
            try:
                self.backend.lib.splacu_dcsr_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


class Backend:
    """One backend per process / GPU (one process per GPU under torch.distributed)."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise SplacuError("spla_b200 needs a CUDA device: there is no CPU fallback on this path")
        self.lib = load_library()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._check(self.lib.splacu_init(device))
        sp = self.lib.splacu_default_stream()
        self.stream_ptr = C.c_void_p(sp)
        self.stream = torch.cuda.ExternalStream(sp, device=self.device)
        ws = C.c_void_p()
        self._check(self.lib.splacu_workspace_create(C.byref(ws)))
        self.ws = ws
        self._nr = C.c_uint32(0)

    # ---- plumbing ----
    def _check(self, rc):
        if rc != 0:
            raise SplacuError(f"splacu error {rc}: {self.lib.splacu_last_error().decode()}")

    def sync(self):
        self._check(self.lib.splacu_sync(self.stream_ptr))

    def launch_count(self):
        c = C.c_uint64(0)
        self.lib.splacu_launch_count(C.byref(c))
        return c.value

    def device_name(self):
        buf = C.create_string_buffer(256)
        self.lib.splacu_device_name(buf, 256)
        return buf.value.decode()

    def set_option(self, name, value):
        """tuning knob read at matrix creation, e.g. set_option("mxv_hub", 2)"""
        self._check(self.lib.splacu_set_option(name.encode(), int(value)))

    def csr_info(self, M):
        nt, nh = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.splacu_csr_info(M.handle, C.byref(nt), C.byref(nh)))
        n_ph, nnz_ph = C.c_int(0), (C.c_uint32 * 32)()
        self._check(self.lib.splacu_csr_phases(M.handle, C.byref(n_ph), nnz_ph, 32))
        n_rc, nnz_rc, rows_rc = C.c_int(0), (C.c_uint32 * 8)(), (C.c_uint32 * 8)()
        self._check(self.lib.splacu_csr_row_classes(M.handle, C.byref(n_rc), nnz_rc, rows_rc, 8))
        return {"n_tiles": nt.value, "n_hub": nh.value, "phase_nnz": [int(nnz_ph[p]) for p in range(n_ph.value)],
                "row_class_nnz": [int(nnz_rc[q]) for q in range(n_rc.value)], "row_class_rows": [int(rows_rc[q]) for q in range(n_rc.value)]}

    def csr(self, n_rows, n_cols, Ap, Aj, Ax):
        with torch.cuda.stream(self.stream):
            return CsrMatrix(self, n_rows, n_cols, Ap, Aj, Ax)

    def empty(self, n, like=None, dtype=torch.float32):
        with torch.cuda.stream(self.stream):
            return torch.empty(n, dtype=like.dtype if like is not None else dtype, device=self.device)

    # ---- the hot path ----
    def mxv_masked(self, M, v, mask, op_mult, op_add, op_select, init, early_exit=False, out=None):
        """r = M x v (pull). Mirrors exec_mxv_masked(r, mask, M, v, op_multiply, op_add, op_select, init, desc)."""
        code = M.dtype
        assert v.numel() == M.n_cols and dtype_code(v) == code
        if mask is not None:
            assert mask.numel() == M.n_rows and dtype_code(mask) == code
        if out is None:
            out = self.empty(M.n_rows, like=v)
        if _user(op_mult, op_add, op_select):
            om, oa, osel = make_op(op_mult, BIN), make_op(op_add, BIN), make_op(op_select, SEL)
            self._check(self.lib.splacu_mxv_masked_ops(M.handle, code, C.byref(om), C.byref(oa), C.byref(osel), _ptr(v), _ptr(mask),
                                                       _ptr(out), scalar_bits(code, init), int(bool(early_exit)), self.stream_ptr))
            return out
        self._check(self.lib.splacu_mxv_masked(M.handle, code, BIN[op_mult], BIN[op_add], SEL[op_select], _ptr(v), _ptr(mask),
                                               _ptr(out), scalar_bits(code, init), int(bool(early_exit)), self.stream_ptr))
        return out

    def mxv_masked_part(self, M, part, v, hub_vals, mask, op_mult, op_add, op_select, init, out, stream_ptr=None):
        """One part of the two-part pull product (splacu_mxv_masked_part): part 1 = mask pass + hub classes (reads hub_vals, or v when
        hub_vals is None), part 2 = everything that reads v + the fix-ups."""
        self._check(self.lib.splacu_mxv_masked_part(M.handle, M.dtype, BIN[op_mult], BIN[op_add], SEL[op_select], _ptr(v), _ptr(hub_vals), _ptr(mask),
                                                    _ptr(out), scalar_bits(M.dtype, init), int(part), stream_ptr if stream_ptr is not None else self.stream_ptr))
        return out

    def csr_hub_cols(self, M):
        """the hub columns of the handle (device tensor of column ids in slot order; empty when the matrix has no column classes)"""
        n, ptr = C.c_uint32(0), C.c_void_p(0)
        self._check(self.lib.splacu_csr_hub_cols(M.handle, C.byref(n), C.byref(ptr)))
        out = torch.empty(n.value, dtype=torch.int32, device=self.device)
        if n.value:
            self._check(self.lib.splacu_memcpy_d2d(_ptr(out), ptr, n.value * 4, self.stream_ptr))
            self.sync()
        return out

    def v_gather(self, idx, src, dst, stream_ptr=None):
        self._check(self.lib.splacu_v_gather(idx.numel(), _ptr(idx), _ptr(src), _ptr(dst), stream_ptr if stream_ptr is not None else self.stream_ptr))
        return dst

    def v_scatter(self, idx, src, dst, stream_ptr=None):
        self._check(self.lib.splacu_v_scatter(idx.numel(), _ptr(idx), _ptr(src), _ptr(dst), dst.numel(), stream_ptr if stream_ptr is not None else self.stream_ptr))
        return dst

    def v_push_peers(self, src_idx, dst_idx, seg_off, peer_ptrs, src, stream_ptr=None):
        """peer_ptrs[q][dst_idx[k]] = src[src_idx[k]] for the elements k of segment q (seg_off: int32 [n_peers + 1], peer_ptrs: int64
        device tensor of n_peers device pointers)"""
        self._check(self.lib.splacu_v_push_peers(src_idx.numel(), _ptr(src_idx), _ptr(dst_idx), _ptr(seg_off), peer_ptrs.numel(), _ptr(peer_ptrs), _ptr(src),
                                                 stream_ptr if stream_ptr is not None else self.stream_ptr))

    def vxm_masked(self, M, vi, vx, mask, op_mult, op_add, op_select, out=None):
        """r = v x M (push) over the sparse vector (vi, vx). Mirrors exec_vxm_masked(r, mask, v, M, ...).
        Returns (ri, rx) exactly-sized (or views of `out=(ri_buf, rx_buf)`)."""
        code = M.dtype
        nv = vi.numel()
        assert vx.numel() == nv and (nv == 0 or dtype_code(vx) == code) and vi.dtype == torch.int32
        if mask is not None:
            assert mask.numel() == M.n_cols and dtype_code(mask) == code
        if _user(op_mult, op_add, op_select):
            om, oa, osel = make_op(op_mult, BIN), make_op(op_add, BIN), make_op(op_select, SEL)
            self._check(self.lib.splacu_vxm_masked_begin_ops(M.handle, code, C.byref(om), C.byref(oa), C.byref(osel), nv, _ptr(vi), _ptr(vx),
                                                             _ptr(mask), self.ws, C.byref(self._nr), self.stream_ptr))
        else:
            self._check(self.lib.splacu_vxm_masked_begin(M.handle, code, BIN[op_mult], BIN[op_add], SEL[op_select], nv, _ptr(vi), _ptr(vx),
                                                         _ptr(mask), self.ws, C.byref(self._nr), self.stream_ptr))
        nr = self._nr.value
        if out is not None:
            ri, rx = out[0][:nr], out[1][:nr]
        else:
            with torch.cuda.stream(self.stream):
                ri = torch.empty(nr, dtype=torch.int32, device=self.device)
                rx = torch.empty(nr, dtype=M.Ax.dtype, device=self.device)
        self._check(self.lib.splacu_vxm_masked_emit(self.ws, _ptr(ri), _ptr(rx), self.stream_ptr))
        return ri, rx

    def profile(self, on=True):
        """cudaEvent timing of every C-ABI entry point under its label (NVTX ranges are always emitted)"""
        self._check(self.lib.splacu_profile_enable(int(bool(on))))

    def profile_dump(self, reset=True):
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.splacu_profile_dump(buf, len(buf)))
        if reset:
            self.lib.splacu_profile_reset()
        rows = {}
        for ln in buf.value.decode().splitlines()[1:]:
            label, calls, dev, host = [x.strip() for x in ln.split(",")]
            rows[label] = {"calls": int(calls), "device_ms": float(dev), "host_ms": float(host)}
        return rows

    def vxm_info(self):
        """{"struct_only": bool}: did the last vxm_masked take the structure-only path (see include/splacu.h)"""
        f = C.c_int(0)
        self._check(self.lib.splacu_workspace_info(self.ws, C.byref(f)))
        return {"struct_only": bool(f.value)}

    def reset_workspace(self):
        self._check(self.lib.splacu_workspace_reset(self.ws, self.stream_ptr))

    # ---- multi-GPU behind the C ABI (include/splacu.h "multi-GPU, single box") ----
    def dist_group(self, device_ids):
        """N shards, one per listed device (ids may repeat: shards share a device); device_ids[0] must be this backend's device"""
        return DistGroup(self, device_ids)

    def coo_to_csr(self, n_rows, Ai, Aj, Ax):
        """device-side ingest: triplets (int32 row ids, int32 column ids, values; any order) -> (Ap, Aj, Ax, was_sorted); sorted rows are
        converted in place (the returned Aj / Ax are the inputs)"""
        nnz = Ai.numel()
        with torch.cuda.stream(self.stream):
            Ap = torch.empty(n_rows + 1, dtype=torch.int32, device=self.device)
            oj, ox = torch.empty_like(Aj), torch.empty_like(Ax)
        flag = C.c_int(0)
        # first try in place; the unsorted path needs distinct outputs
        rc = self.lib.splacu_coo_to_csr(n_rows, nnz, _ptr(Ai), _ptr(Aj), _ptr(Ax), _ptr(Ap), _ptr(oj), _ptr(ox), self.ws, C.byref(flag), self.stream_ptr)
        self._check(rc)
        return Ap, oj, ox, bool(flag.value)

    # ---- format glue ----
    def coo_to_dense(self, n, fill, vi, vx, out=None):
        code = dtype_code(vx)
        if out is None:
            out = self.empty(n, like=vx)
        self._check(self.lib.splacu_coo_to_dense(n, scalar_bits(code, fill), vi.numel(), _ptr(vi), _ptr(vx), _ptr(out), self.stream_ptr))
        return out

    def dense_to_coo(self, dense, fill):
        code = dtype_code(dense)
        n = dense.numel()
        fb = scalar_bits(code, fill)
        self._check(self.lib.splacu_dense_to_coo_count(code, n, fb, _ptr(dense), self.ws, C.byref(self._nr), self.stream_ptr))
        nr = self._nr.value
        with torch.cuda.stream(self.stream):
            ri = torch.empty(nr, dtype=torch.int32, device=self.device)
            rx = torch.empty(nr, dtype=dense.dtype, device=self.device)
        self._check(self.lib.splacu_dense_to_coo_emit(code, n, fb, _ptr(dense), self.ws, _ptr(ri), _ptr(rx), self.stream_ptr))
        return ri, rx

    def fill(self, t, value):
        self._check(self.lib.splacu_fill(_ptr(t), scalar_bits(dtype_code(t), value), t.numel(), self.stream_ptr))
        return t

    # ---- neighbours ----
    def v_assign_masked(self, r, mask, value, op_assign, op_select):
        """dense mask: exec_v_assign_masked(r, mask, value, op_assign, op_select); mask may be (mi, mx) sparse."""
        code = dtype_code(r)
        vb = scalar_bits(code, value)
        if _user(op_assign, op_select):
            oa, osel = make_op(op_assign, BIN), make_op(op_select, SEL)
            if isinstance(mask, tuple):
                mi, mx = mask
                self._check(self.lib.splacu_v_assign_masked_sparse_ops(code, C.byref(oa), C.byref(osel), _ptr(r), mi.numel(), _ptr(mi), _ptr(mx), vb, self.stream_ptr))
            else:
                self._check(self.lib.splacu_v_assign_masked_dense_ops(code, C.byref(oa), C.byref(osel), r.numel(), _ptr(r), _ptr(mask), vb, self.stream_ptr))
            return r
        if isinstance(mask, tuple):
            mi, mx = mask
            self._check(self.lib.splacu_v_assign_masked_sparse(code, BIN[op_assign], SEL[op_select], _ptr(r), mi.numel(), _ptr(mi), _ptr(mx),
                                                               vb, self.stream_ptr))
        else:
            self._check(self.lib.splacu_v_assign_masked_dense(code, BIN[op_assign], SEL[op_select], r.numel(), _ptr(r), _ptr(mask), vb,
                                                              self.stream_ptr))
        return r

    def v_count_mf(self, v, fill):
        code = dtype_code(v)
        self._check(self.lib.splacu_v_count_mf_dense(code, v.numel(), _ptr(v), scalar_bits(code, fill), self.ws, C.byref(self._nr), self.stream_ptr))
        return self._nr.value

    def pack_bits(self, v, op_select, out=None):
        """bit i = op_select(v[i]) (int32 words, ceil(n / 32) of them): the structure-only form a frontier is exchanged in"""
        n = v.numel()
        if out is None:
            out = self.empty((n + 31) // 32, dtype=torch.int32)
        assert out.numel() >= (n + 31) // 32 and out.dtype == torch.int32
        self._check(self.lib.splacu_v_pack_bits(dtype_code(v), SEL[op_select], n, _ptr(v), _ptr(out), self.stream_ptr))
        return out

    def unpack_bits(self, bits, n, one, zero, out):
        """out[i] = bit i ? one : zero"""
        code = dtype_code(out)
        assert out.numel() >= n and bits.numel() >= (n + 31) // 32
        self._check(self.lib.splacu_v_unpack_bits(n, _ptr(bits), scalar_bits(code, one), scalar_bits(code, zero), _ptr(out), self.stream_ptr))
        return out

    def v_eadd_fdb_dense(self, r, v, op, fdb_fill, fdb=None):
        code = dtype_code(r)
        if fdb is None:
            fdb = self.empty(r.numel(), like=r)
        if _user(op):
            o = make_op(op, BIN)
            self._check(self.lib.splacu_v_eadd_fdb_dense_op(code, C.byref(o), r.numel(), _ptr(r), _ptr(v), _ptr(fdb), scalar_bits(code, fdb_fill), self.stream_ptr))
            return fdb
        self._check(self.lib.splacu_v_eadd_fdb_dense(code, BIN[op], r.numel(), _ptr(r), _ptr(v), _ptr(fdb), scalar_bits(code, fdb_fill), self.stream_ptr))
        return fdb

    def v_eadd_fdb_sparse(self, r, vi, vx, op):
        code = dtype_code(r)
        if _user(op):
            o = make_op(op, BIN)
            self._check(self.lib.splacu_v_eadd_fdb_sparse_begin_op(code, C.byref(o), _ptr(r), vi.numel(), _ptr(vi), _ptr(vx), self.ws, C.byref(self._nr),
                                                                   self.stream_ptr))
        else:
            self._check(self.lib.splacu_v_eadd_fdb_sparse_begin(code, BIN[op], _ptr(r), vi.numel(), _ptr(vi), _ptr(vx), self.ws, C.byref(self._nr),
                                                                self.stream_ptr))
        nf = self._nr.value
        with torch.cuda.stream(self.stream):
            fi = torch.empty(nf, dtype=torch.int32, device=self.device)
            fx = torch.empty(nf, dtype=r.dtype, device=self.device)
        self._check(self.lib.splacu_v_eadd_fdb_sparse_emit(self.ws, _ptr(fi), _ptr(fx), self.stream_ptr))
        return fi, fx

    def v_eadd(self, u, v, op, out=None):
        code = dtype_code(u)
        if out is None:
            out = self.empty(u.numel(), like=u)
        if _user(op):
            o = make_op(op, BIN)
            self._check(self.lib.splacu_v_eadd_dense_op(code, C.byref(o), u.numel(), _ptr(out), _ptr(u), _ptr(v), self.stream_ptr))
            return out
        self._check(self.lib.splacu_v_eadd_dense(code, BIN[op], u.numel(), _ptr(out), _ptr(u), _ptr(v), self.stream_ptr))
        return out

    def v_reduce(self, v, op, init):
        code = dtype_code(v)
        self._check(self.lib.splacu_v_reduce_dense(code, BIN[op], v.numel(), _ptr(v), scalar_bits(code, init), self.ws, C.byref(self._nr), self.stream_ptr))
        bits = self._nr.value
        if code == FLOAT:
            return float(torch.tensor([bits if bits < 2 ** 31 else bits - 2 ** 32], dtype=torch.int32).view(torch.float32).item())
        if code == INT:
            return bits if bits < 2 ** 31 else bits - 2 ** 32
        return bits
