"""numpy <-> device helpers for the GPU parity tests (torch is only the carrier of device memory).

The parity bar (north_star): bit-exact for INT / UINT and for every order-independent FLOAT op; FLOAT PLUS / MULT
reductions within 1e-5 RELATIVE to the reference, per element: |got - ref| <= 1e-5 * |ref|. There is no absolute floor.
The only escape is for elements where the REFERENCE's own sequential fp32 fold is not accurate to 1e-5 of the true value (hub rows
of 10^5 entries: the left-to-right fold loses up to d * 2^-24 relative; sums of signed terms that cancel): such an element passes
iff the device value is no further from the float64 value than the reference itself is (or within 1e-5 relative of the float64
value) AND lies within the rigorous rounding-error bound of ANY fp32 summation order around the exact (float64) sum,
        |got - exact| <= gamma_k * sum_i |t_i|,   gamma_k = k u / (1 - k u),  u = 2^-24,  k = number of additions
(Higham, Accuracy and Stability of Numerical Algorithms, eq. 4.4: valid for every ordering, hence for the reference's
left-to-right fold and for the device's segmented tree alike). The terms t_i = fl(mult(a, v)) are computed in fp32 exactly as
both sides compute them. Every use of the escape is counted and written to gpurun_out/parity_stats.jsonl beside the measured
maximum relative error.
"""
import json
import os

import numpy as np
import torch

TORCH = {np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.uint32, np.dtype(np.float32): torch.float32}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5
U32 = 2.0 ** -24


def to_dev(a, backend):
    """4-byte numpy array -> cuda tensor of the matching spla type (uint32 index arrays -> int32)."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.view(np.int32).copy()).to(backend.device)
    return t if a.dtype == np.int32 else t.view(TORCH[a.dtype])


def idx_dev(a, backend):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint32).view(np.int32).copy()).to(backend.device)


def to_np(t, np_dtype):
    return t.detach().view(torch.int32).cpu().numpy().view(np_dtype)


def make_csr(backend, n_rows, n_cols, Ap, Aj, Ax):
    return backend.csr(n_rows, n_cols, idx_dev(Ap, backend), idx_dev(Aj, backend), to_dev(Ax, backend))


def _log_stats(rec):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        try:
            with open(os.path.join(d, "parity_stats.jsonl"), "a") as f:
                f.write(json.dumps(rec) + "\n")
        except OSError:
            pass


# ---- float32 restatement of the built-in binary ops for the error bound (checked against the C oracle in tests/test_oracle.py) ----
def np_binop_f32(op, a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    with np.errstate(all="ignore"):
        if op == "PLUS":
            return a + b
        if op == "MINUS":
            return a - b
        if op == "MULT":
            return a * b
        if op == "DIV":
            return a / b
        if op == "MINUS_POW2":
            d = a - b
            return d * d
        if op == "FIRST":
            return a + np.zeros_like(b)
        if op == "SECOND":
            return b + np.zeros_like(a)
        if op == "BONE":
            return np.ones(np.broadcast(a, b).shape, dtype=np.float32)
        if op == "MIN":
            return np.where(b < a, b, a)
        if op == "MAX":
            return np.where(a < b, b, a)
        if op == "LOR":
            return ((a != 0) | (b != 0)).astype(np.float32)
        if op == "LAND":
            return ((a != 0) & (b != 0)).astype(np.float32)
    raise ValueError(op)


def np_select(op, x):
    x = np.asarray(x)
    return {"EQZERO": x == 0, "NQZERO": x != 0, "GTZERO": x > 0, "GEZERO": x >= 0, "LTZERO": x < 0, "LEZERO": x <= 0,
            "ALWAYS": np.ones(x.shape, bool), "NEVER": np.zeros(x.shape, bool)}[op]


def _gamma(k):
    k = np.asarray(k, dtype=np.float64)
    return k * U32 / (1.0 - k * U32)


def mxv_bound(op_mult, op_add, Ap, Aj, Ax, v, init):
    """(exact64, abs_bound) per row for FLOAT mxv with op_add PLUS, else None. exact = init + sum of the fp32 products in float64;
    abs_bound = gamma_k * (|init| + sum |t_i|) with k = row length additions (any summation order)."""
    if op_add != "PLUS":
        return None
    Ap64 = np.asarray(Ap).astype(np.int64)
    n_rows = len(Ap64) - 1
    t = np_binop_f32(op_mult, np.asarray(Ax, dtype=np.float32), np.asarray(v, dtype=np.float32)[np.asarray(Aj).astype(np.int64)]).astype(np.float64)
    rows = np.repeat(np.arange(n_rows), np.diff(Ap64))
    s = np.bincount(rows, weights=t, minlength=n_rows)
    sa = np.bincount(rows, weights=np.abs(t), minlength=n_rows)
    k = np.diff(Ap64)
    i64 = float(np.float32(init))
    return i64 + s, _gamma(k) * (abs(i64) + sa)


def vxm_bound(op_mult, op_add, Ap, Aj, Ax, n_cols, vi, vx, take_cols, ri=None):
    """(exact64, abs_bound) per COLUMN (dense, length n_cols; or at the columns `ri`) for FLOAT vxm with op_add PLUS, else None.
    take_cols: bool[n_cols] = select(mask[j])."""
    if op_add != "PLUS":
        return None
    Ap64 = np.asarray(Ap).astype(np.int64)
    vi = np.asarray(vi).astype(np.int64)
    deg = Ap64[vi + 1] - Ap64[vi]
    pos = np.concatenate([np.arange(Ap64[i], Ap64[i + 1]) for i in vi]) if len(vi) else np.zeros(0, np.int64)
    x = np.repeat(np.asarray(vx, dtype=np.float32), deg)
    cols = np.asarray(Aj).astype(np.int64)[pos]
    t = np_binop_f32(op_mult, x, np.asarray(Ax, dtype=np.float32)[pos]).astype(np.float64)
    keep = np.asarray(take_cols)[cols]
    cols, t = cols[keep], t[keep]
    s = np.bincount(cols, weights=t, minlength=n_cols)
    sa = np.bincount(cols, weights=np.abs(t), minlength=n_cols)
    k = np.bincount(cols, minlength=n_cols)
    ab = _gamma(np.maximum(k - 1, 0)) * sa
    if ri is not None:
        ri = np.asarray(ri).astype(np.int64)
        return s[ri], ab[ri]
    return s, ab


def assert_values(got, want, exact, rtol=RTOL, what="", bound=None):
    """bit-exact for integer / order-independent work; otherwise |got - want| <= rtol * |want| PER ELEMENT (north_star: 1e-5
    relative). `bound` = (exact64, abs_bound) arrays aligned with got: the derived escape for cancelling sums (module docstring)."""
    if exact:
        gb, wb = got.view(np.uint32), want.view(np.uint32)
        if got.dtype == np.float32:  # NaN payloads are not part of the contract (x86 and sm_100a produce different quiet NaNs)
            gn, wn = np.isnan(got), np.isnan(want)
            assert np.array_equal(gn, wn), f"{what}: NaN pattern differs"
            gb, wb = gb[~gn], wb[~wn]
            # the sign of a zero is not part of the contract either: -0 == +0 in value, and which of two equal
            # zeros a MIN / MAX fold keeps depends on the (sequential) arrival order
            zero = ((gb & 0x7fffffff) == 0) & ((wb & 0x7fffffff) == 0)
            gb, wb = gb[~zero], wb[~zero]
            got, want = got[~gn][~zero], want[~wn][~zero]
        bad = np.nonzero(gb != wb)[0]
        assert len(bad) == 0, f"{what}: not bit-exact at {bad[:8]}\n got {got[bad[:8]]}\nwant {want[bad[:8]]}"
        return
    g = got.astype(np.float64)
    w = want.astype(np.float64)
    same_special = (np.isnan(g) == np.isnan(w)).all() and (np.isinf(g) == np.isinf(w)).all()
    assert same_special, f"{what}: nan/inf pattern differs"
    fin = np.isfinite(w)
    diff = np.abs(g - w)
    with np.errstate(all="ignore"):
        rel = np.where(w != 0, diff / np.abs(w), np.where(diff == 0, 0.0, np.inf))
    strict = (diff <= rtol * np.abs(w)) | ~fin
    n_escape = 0
    if not strict.all():
        bad = np.nonzero(~strict)[0]
        if callable(bound):
            bound = bound()
        assert bound is not None, f"{what}: {len(bad)} elements beyond {rtol} relative, max rel err {rel[fin].max():.3e} at {bad[:6]}: got {got[bad[:6]]} want {want[bad[:6]]}"
        ex, ab = bound
        dev_err = np.abs(g[bad] - ex[bad])
        ref_err = np.abs(w[bad] - ex[bad])
        slack = 1e-12 * np.abs(ex[bad]) + 1e-300  # float64 rounding of the "exact" value itself
        # the element is off the reference by more than 1e-5 relative. Accepted only if the DEVICE value is (a) no further from the
        # float64 value than the reference's own fp32 fold is, or within 1e-5 relative of the float64 value, and (b) inside the
        # rigorous any-order fp32 summation bound (when one is given)
        ok = (dev_err <= ref_err + slack) | (dev_err <= rtol * np.abs(ex[bad]))
        if ab is not None:
            ok &= dev_err <= ab[bad] + slack
        assert ok.all(), (f"{what}: {int((~ok).sum())} elements beyond {rtol} relative of the reference AND not explained by the reference's own rounding error; "
                          f"first at {bad[~ok][:4]}: got {got[bad[~ok][:4]]} want {want[bad[~ok][:4]]} float64 {ex[bad[~ok][:4]]}")
        n_escape = len(bad)
        with np.errstate(all="ignore"):
            dev_rel = np.max(dev_err / np.abs(ex[bad]))
            ref_rel = np.max(ref_err / np.abs(ex[bad]))
        print(f"[parity] {what}: {n_escape} of {len(g)} element(s) differ from the reference by more than {rtol} relative (max {rel[bad].max():.3e}); against the "
              f"float64 value the device is off by at most {dev_rel:.3e} relative, the reference's own sequential fp32 fold by up to {ref_rel:.3e}: accepted "
              f"(device error <= reference error or <= {rtol} of the float64 value" + ("" if ab is None else ", and inside the any-order fp32 summation bound") + ")")
    finite_rel = rel[fin & strict]
    _log_stats({"what": what, "n": int(len(g)), "max_rel_err_strict": float(finite_rel.max(initial=0.0)), "rtol": rtol, "escaped_by_bound": int(n_escape),
                "max_rel_err_vs_reference": float(rel[fin].max(initial=0.0))})


def pagerank_float64(Ap, Aj, Ax, alpha, iters):
    """The reference's pr() loop (src/algorithm.cpp:278-335: p = A p_prev + (1 - alpha)/N, `iters` times from p = 1/N) in float64, on the
    fp32 matrix values: the 'true' value both fp32 implementations approximate."""
    import scipy.sparse as sp

    n = len(Ap) - 1
    A = sp.csr_matrix((np.asarray(Ax, dtype=np.float64), np.asarray(Aj).astype(np.int64), np.asarray(Ap).astype(np.int64)), shape=(n, n))
    p = np.full(n, float(np.float32(1.0 / n)))
    add = float(np.float32((np.float32(1.0) - np.float32(alpha)) / np.float32(n)))
    for _ in range(iters):
        p = A @ p + add
    return p
