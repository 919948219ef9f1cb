"""Synthetic graph generators for the bench harness and the parity tests.

The reference ships no generator (only ``MtxLoader`` for .mtx files, reference src/io.cpp:50-233), so
inputs are produced here with fixed seeds and then post-processed exactly like ``MtxLoader::load``'s
defaults (symmetrise, drop self loops, sort by (i, j), de-duplicate; reference src/io.cpp:159-214) so that
the CPU oracle and the CUDA path see byte-identical row-sorted CSR matrices.

torch is used as plumbing only (device RNG, sort, unique); everything works on CPU tensors as well, which
is what the CPU tests use.
"""
import torch


def _finish(i, j, n, symmetrise=True, drop_loops=True):
    """(i, j) int64 edge lists -> row-sorted, de-duplicated CSR structure (Ap int64[n+1], Aj int32[nnz])."""
    if symmetrise:
        i, j = torch.cat([i, j]), torch.cat([j, i])
    if drop_loops:
        keep = i != j
        i, j = i[keep], j[keep]
    key = i * n + j
    del i, j
    key = torch.unique(key)  # sorted ascending == sorted by (i, j)
    rows = torch.div(key, n, rounding_mode="floor")
    cols = (key - rows * n).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=n)
    del rows
    Ap = torch.zeros(n + 1, dtype=torch.int64, device=cols.device)
    torch.cumsum(counts, 0, out=Ap[1:])
    return Ap, cols


def rmat(scale, edge_factor=16, seed=1, a=0.57, b=0.19, c=0.19, device="cpu",
         permute=True, symmetrise=True, chunk=1 << 26):
    """Graph500-style R-MAT / Kronecker graph: n = 2**scale, edge_factor * n generated directed edges.

    Returns (n, Ap, Aj) with Ap int64[n+1] and Aj int32[nnz] (torch tensors on `device`).
    """
    n = 1 << scale
    m = edge_factor * n
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ii = torch.empty(m, dtype=torch.int64, device=device)
    jj = torch.empty(m, dtype=torch.int64, device=device)
    for lo in range(0, m, chunk):
        hi = min(m, lo + chunk)
        i = torch.zeros(hi - lo, dtype=torch.int64, device=device)
        j = torch.zeros(hi - lo, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(hi - lo, generator=g, device=device)
            ibit = r >= (a + b)
            jbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            i = (i << 1) | ibit
            j = (j << 1) | jbit
        ii[lo:hi] = i
        jj[lo:hi] = j
    if permute:
        perm = torch.randperm(n, generator=g, device=device)
        ii, jj = perm[ii], perm[jj]
    Ap, Aj = _finish(ii, jj, n, symmetrise=symmetrise)
    return n, Ap, Aj


def uniform_random(scale, avg_degree=16, seed=4, device="cpu", symmetrise=True):
    """Erdos-Renyi-like graph with n = 2**scale and ~avg_degree*n directed edges before symmetrisation."""
    n = 1 << scale
    m = avg_degree * n
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    i = torch.randint(0, n, (m,), generator=g, device=device, dtype=torch.int64)
    j = torch.randint(0, n, (m,), generator=g, device=device, dtype=torch.int64)
    Ap, Aj = _finish(i, j, n, symmetrise=symmetrise)
    return n, Ap, Aj


def grid2d(side, device="cpu"):
    """side x side 4-neighbour grid (road-like): n = side*side, nnz = 4*side*(side-1), symmetric."""
    n = side * side
    idx = torch.arange(n, dtype=torch.int64, device=device)
    x = idx % side
    y = torch.div(idx, side, rounding_mode="floor")
    right = x < side - 1
    down = y < side - 1
    i = torch.cat([idx[right], idx[down]])
    j = torch.cat([idx[right] + 1, idx[down] + side])
    Ap, Aj = _finish(i, j, n, symmetrise=True)
    return n, Ap, Aj


def out_degrees(Ap):
    return Ap[1:] - Ap[:-1]


def pagerank_values(Ap, alpha=0.85):
    """A[i][j] = alpha / outdeg(i) for every stored entry, as the reference example builds it
    (reference examples/pr.cpp:81-88)."""
    deg = out_degrees(Ap)
    w = (alpha / deg.clamp(min=1).to(torch.float32))
    return torch.repeat_interleave(w, deg)


def uniform_weights(nnz, lo=1.0, hi=2.0, seed=3, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.rand(nnz, generator=g, device=device) * (hi - lo) + lo
