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
