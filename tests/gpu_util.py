"""numpy <-> device helpers for the GPU parity tests (torch is only the carrier of device memory)."""
import numpy as np
import torch

TORCH = {np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.uint32, np.dtype(np.float32): torch.float32}


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


def assert_values(got, want, exact, rtol=1e-5, what=""):
    """bit-exact for integer / order-independent work; otherwise the north_star float tolerance (1e-5 relative)."""
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
    else:
        g = got.astype(np.float64)
        w = want.astype(np.float64)
        same_special = (np.isnan(g) == np.isnan(w)).all() and (np.isinf(g) == np.isinf(w)).all()
        assert same_special, f"{what}: nan/inf pattern differs"
        fin = np.isfinite(w)
        scale = np.maximum(np.abs(w[fin]), 1e-30)
        err = np.abs(g[fin] - w[fin]) / scale
        # sums of signed terms can cancel: allow the tolerance relative to the magnitude of the terms as well
        ok = (err <= rtol) | (np.abs(g[fin] - w[fin]) <= rtol * max(1.0, float(np.abs(w[fin]).max(initial=0.0))))
        assert ok.all(), f"{what}: max rel err {err.max():.3e}"
