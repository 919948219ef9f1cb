"""bfs / sssp / pr driven through the C ABI exactly like the reference drives them through exec_* (reference
src/algorithm.cpp:45-120, 158-229, 278-335): the same sequence of v_assign_masked / vxm_masked / mxv_masked / v_count_mf /
v_eadd_fdb / v_eadd / v_reduce calls, the same push / pull decision (front_factor), the same op triples. Used by the bench
harness and the parity tests; the spla C++ plug-in (spla_b200/src/cuda) runs the reference's own algorithm.cpp instead.

A frontier is kept in whichever format the last op produced: ("coo", vi, vx) after a push, ("dense", x) after a pull;
conversions happen where the reference's storage manager would convert (AccCoo <-> AccDense)."""
import torch

FLT_MAX = 3.4028234663852886e38


class Frontier:
    def __init__(self, be, n, fill, coo=None, dense=None):
        self.be, self.n, self.fill, self.coo, self.dense = be, n, fill, coo, dense

    def as_coo(self):
        if self.coo is None:
            self.coo = self.be.dense_to_coo(self.dense, self.fill)
        return self.coo

    def as_dense(self):
        if self.dense is None:
            self.dense = self.be.coo_to_dense(self.n, self.fill, *self.coo)
        return self.dense

    def count(self):
        if self.coo is not None:  # structural count, reference src/cpu/cpu_v_count_mf.hpp:81-90
            return int(self.coo[0].numel())
        return self.be.v_count_mf(self.dense, self.fill)


def bfs(be, M, source, mode="push_pull", front_factor=0.05, trace=None):
    """Depth vector (int32, 0 = unreached, source = 1), reference src/algorithm.cpp:45-120 (BAND / BOR / EQZERO, early_exit)."""
    n = M.n_rows
    dev = be.device
    with torch.cuda.stream(be.stream):
        v = torch.zeros(n, dtype=torch.int32, device=dev)
        front = Frontier(be, n, 0, coo=(torch.tensor([source], dtype=torch.int32, device=dev), torch.ones(1, dtype=torch.int32, device=dev)))
        size, level = 1, 1
        while size:
            if front.coo is not None:
                be.v_assign_masked(v, front.coo, level, "SECOND", "NQZERO")
            else:
                be.v_assign_masked(v, front.dense, level, "SECOND", "NQZERO")
            push = mode == "push" or (mode == "push_pull" and size / n <= front_factor)
            if push:
                ri, rx = be.vxm_masked(M, *front.as_coo(), v, "BAND", "BOR", "EQZERO")
                front = Frontier(be, n, 0, coo=(ri, rx))
            else:
                r = be.mxv_masked(M, front.as_dense(), v, "BAND", "BOR", "EQZERO", 0, early_exit=True)
                front = Frontier(be, n, 0, dense=r)
            size = front.count()
            if trace is not None:
                trace.append(("push" if push else "pull", size))
            level += 1
        be.sync()
    return v


def sssp(be, M, source, mode="push_pull", front_factor=0.05):
    """Distances (float32, FLT_MAX = unreached), reference src/algorithm.cpp:158-229 (PLUS / MIN / ALWAYS, eadd_fdb MIN)."""
    n = M.n_rows
    dev = be.device
    with torch.cuda.stream(be.stream):
        v = torch.full((n,), FLT_MAX, dtype=torch.float32, device=dev)
        v[source] = 0.0
        fdb = Frontier(be, n, FLT_MAX, coo=(torch.tensor([source], dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.float32, device=dev)))
        size = 1
        while size:
            push = mode == "push" or (mode == "push_pull" and size / n <= front_factor)
            if push:
                ri, rx = be.vxm_masked(M, *fdb.as_coo(), None, "PLUS", "MIN", "ALWAYS")
                fi, fx = be.v_eadd_fdb_sparse(v, ri, rx, "MIN")
                fdb = Frontier(be, n, FLT_MAX, coo=(fi, fx))
            else:
                r = be.mxv_masked(M, fdb.as_dense(), None, "PLUS", "MIN", "ALWAYS", FLT_MAX)
                fdb = Frontier(be, n, FLT_MAX, dense=be.v_eadd_fdb_dense(v, r, "MIN", FLT_MAX))
            size = fdb.count()
        be.sync()
    return v


def pagerank(be, M, alpha=0.85, eps=1e-6, max_iter=1000):
    """Ranks (float32), reference src/algorithm.cpp:278-335; M[i][j] = alpha / outdeg(i) as examples/pr.cpp:81-88 builds it."""
    n = M.n_rows
    dev = be.device
    with torch.cuda.stream(be.stream):
        p_prev = torch.full((n,), 1.0 / n, dtype=torch.float32, device=dev)
        addition = torch.full((n,), (1.0 - alpha) / n, dtype=torch.float32, device=dev)
        p = torch.empty_like(p_prev)
        p_tmp = torch.empty_like(p_prev)
        errors = torch.empty_like(p_prev)
        error, it = eps + 0.1, 0
        while error > eps and it < max_iter:
            be.mxv_masked(M, p_prev, None, "MULT", "PLUS", "ALWAYS", 0.0, out=p_tmp)
            be.v_eadd(p_tmp, addition, "PLUS", out=p)
            be.v_eadd(p, p_prev, "MINUS_POW2", out=errors)
            error = be.v_reduce(errors, "PLUS", 0.0) ** 0.5
            p, p_prev = p_prev, p
            it += 1
        be.sync()
    return p_prev, it


# ---------------------------------------------------------------------------------------------------------
# Multi-GPU BFS (BASELINE config 4): one process per GPU, vertices owned in contiguous nnz-balanced windows.
def make_bfs_shard(be, n, Ap, Aj, Ax, rank, world, bitmap_exchange=False):
    """Rank `rank`'s share of a SYMMETRIC n x n matrix: window [b[rank], b[rank+1]) of the vertices (nnz-balanced on Ap, and --
    the matrix being symmetric -- on the columns too), M_rows = M[window, :] for the pull direction (row-sharded mxv)
    and M_cols = M[:, window] for the push direction (column-sharded vxm)."""
    from . import dist as sd

    b = sd.balanced_boundaries(Ap, world)
    w0, w1 = b[rank], b[rank + 1]
    rAp, rAj, rAx = sd.row_slice(Ap, Aj, Ax, w0, w1)
    cAp, cAj, cAx = sd.column_slice(Ap, Aj, Ax, w0, w1)
    # the dense frontier of the pull levels lives in the padded equal-window layout (dist.padded_layout): the uneven windows are
    # then all-gathered by ONE collective instead of one broadcast per owner; the row slice's column ids are mapped once
    if world > 1:
        W, shifts = sd.padded_layout(b)
        rAj = sd.to_padded_index(rAj, b, shifts)
        n_vec = world * W
    else:
        W, n_vec = n, n
    # all stored values 1 (an adjacency matrix): BAND / BOR frontiers then only hold 0 / 1 and CAN travel as an n-bit bitmap
    # (SURVEY 8e). Opt-in: on 2 B200s over NVLink the 64 MB dense all-gather is as fast as pack + 2 MB all-gather + unpack
    # (3.07 ms against 3.33 / 3.63 ms per RMAT-24 BFS, tools/bench_bfs_dist.py --ab); not measured at 8 GPUs.
    unit = bool(bitmap_exchange) and bool((Ax == 1).all().item())
    return {"n": n, "bounds": b, "rank": rank, "world": world, "w0": w0, "w1": w1, "W": W, "n_vec": n_vec, "unit_values": unit,
            "M_rows": be.csr(w1 - w0, n_vec, rAp, rAj.to(torch.int32), rAx),
            "M_cols": be.csr(n, w1 - w0, cAp, cAj, cAx)}


def _owner(bounds, vertex):
    """rank whose window [bounds[p], bounds[p+1]) holds the vertex"""
    for p in range(len(bounds) - 1):
        if bounds[p] <= vertex < bounds[p + 1]:
            return p
    raise ValueError("vertex outside the matrix")


def bfs_dist(be, shard, source, mode="push_pull", front_factor=0.05, group=None, trace=None):
    """BFS over a sharded matrix; returns this rank's window of the depth vector (int32, 0 = unreached, source = 1).

    Per level, exactly the reference's sequence (src/algorithm.cpp:45-120) on the owner's window -- v_assign_masked, then
    vxm_masked (push) or mxv_masked with early exit (pull), then the front size -- with one exchange step in between:
      push  all-gather of the sparse frontier pieces (dist.exchange_frontier); every rank expands the whole frontier against its
            column slice under its own window of the depth vector as the mask; results are disjoint windows
      pull  ONE all-gather of the dense frontier windows in the padded layout (dist.allgather_padded) -- optionally as an n-bit
            bitmap when the matrix holds only ones (make_bfs_shard(bitmap_exchange=True); 2 MB instead of 64 MB at scale 24,
            SURVEY 8e); every rank pulls its rows
    and an all-gather of the per-rank front sizes: their sum drives the same push / pull decision on every rank, and the sizes
    themselves spare the next push level the count exchange of its frontier all-gather."""
    import contextlib

    import torch.distributed as tdist

    from . import dist as sd

    n, b, rank, world = shard["n"], shard["bounds"], shard["rank"], shard["world"]
    w0, w1 = shard["w0"], shard["w1"]
    n_loc = w1 - w0
    W, p0 = shard["W"], (shard["rank"] * shard["W"] if shard["world"] > 1 else 0)
    dev = be.device
    ctx = torch.cuda.stream(be.stream) if getattr(be, "stream", None) is not None else contextlib.nullcontext()
    with ctx:
        depth = torch.zeros(n_loc, dtype=torch.int32, device=dev)
        full = torch.zeros(shard["n_vec"], dtype=torch.int32, device=dev)  # padding stays 0: no column id points at it
        as_bits = world > 1 and shard["unit_values"]  # W is a multiple of 32: every window starts on a bitmap word
        bits = torch.zeros(shard["n_vec"] // 32, dtype=torch.int32, device=dev) if as_bits else None
        own = w0 <= source < w1
        li = torch.tensor([source - w0] if own else [], dtype=torch.int32, device=dev)
        front = Frontier(be, n_loc, 0, coo=(li, torch.ones(li.numel(), dtype=torch.int32, device=dev)))
        size, level = 1, 1
        sizes = [int(p == _owner(b, source)) for p in range(world)]  # per-rank sizes of the current frontier pieces
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        cnts = torch.zeros(world, dtype=torch.int64, device=dev)
        while size:
            if front.coo is not None:
                if front.coo[0].numel():
                    be.v_assign_masked(depth, front.coo, level, "SECOND", "NQZERO")
            else:
                be.v_assign_masked(depth, front.dense, level, "SECOND", "NQZERO")
            push = mode == "push" or (mode == "push_pull" and size / n <= front_factor)
            if push:
                had_coo = front.coo is not None  # the cached per-rank sizes describe a frontier that was produced sparse
                vi, vx = sd.exchange_frontier(*front.as_coo(), w0, group=group, sizes=sizes if had_coo else None)
                if n_loc and vi.numel():
                    ri, rx = be.vxm_masked(shard["M_cols"], vi, vx, depth, "BAND", "BOR", "EQZERO")
                else:
                    ri, rx = vi[:0], vx[:0]
                front = Frontier(be, n_loc, 0, coo=(ri, rx))
            else:
                if as_bits:
                    if n_loc:
                        be.pack_bits(front.as_dense(), "NQZERO", out=bits[p0 // 32:p0 // 32 + (n_loc + 31) // 32])
                    sd.allgather_padded(bits, W // 32, group=group)
                    be.unpack_bits(bits, shard["n_vec"], 1, 0, out=full)
                else:
                    if n_loc:
                        full[p0:p0 + n_loc].copy_(front.as_dense())
                    if world > 1:
                        sd.allgather_padded(full, W, group=group)
                if n_loc:
                    r = be.mxv_masked(shard["M_rows"], full, depth, "BAND", "BOR", "EQZERO", 0, early_exit=True)
                else:
                    r = depth[:0]
                front = Frontier(be, n_loc, 0, dense=r)
            cnt[0] = front.count() if n_loc else 0
            if world > 1:
                tdist.all_gather_into_tensor(cnts, cnt, group=group)
                sizes = [int(c) for c in cnts.tolist()]
            else:
                sizes = [int(cnt.item())]
            size = sum(sizes)
            if trace is not None:
                trace.append(("push" if push else "pull", size))
            level += 1
        be.sync()
    return depth
