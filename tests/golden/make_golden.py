"""Generates tests/golden/*.npz from the UNMODIFIED reference CPU backend.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py

It drives oracle/_ref/libspla_ref.so (compiled from /root/reference by oracle/Makefile) through
oracle/ref_shim.cpp and stores inputs + reference outputs of exec_mxv_masked / exec_vxm_masked and of
spla::bfs / sssp / pr on small seeded inputs. The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from cases import FLOAT, INT, UINT  # noqa: E402
from oracle.oracle import RefSpla  # noqa: E402


def main():
    ref = RefSpla()
    rng = np.random.default_rng(20261017)
    out = {}
    meta = []
    cid = 0

    def add_case(dtype, om, oa, osel, n_rows, n_cols, avg, skew, early_exit, kind="small"):
        nonlocal cid
        Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, avg, skew=skew, kind=kind)
        M = ref.matrix(dtype, n_rows, n_cols, cases.csr_to_coo_rows(Ap), Aj, Ax)
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask_r = cases.rand_values(rng, dtype, n_rows)
        init = cases.rand_values(rng, dtype, 1)[0]
        r = ref.mxv_masked(M, om, oa, osel, v, mask_r, init, early_exit)
        vi, vx = cases.rand_frontier(rng, dtype, n_rows, max(1, n_rows // 3), kind)
        mask_c = cases.rand_values(rng, dtype, n_cols)
        ri, rx = ref.vxm_masked(M, om, oa, osel, vi, vx, mask_c)
        p = f"c{cid}_"
        out.update({p + "Ap": Ap, p + "Aj": Aj, p + "Ax": Ax, p + "v": v, p + "mask_r": mask_r, p + "init": np.array([init]),
                    p + "r": r, p + "vi": vi, p + "vx": vx, p + "mask_c": mask_c, p + "ri": ri, p + "rx": rx})
        meta.append(f"{cid},{dtype},{om},{oa},{osel},{n_rows},{n_cols},{int(early_exit)}")
        cid += 1

    # the semirings the algorithms use, several shapes, early-exit on and off
    for (dtype, om, oa, osel) in cases.NAMED_SEMIRINGS:
        for (nr, nc, avg, skew) in [(37, 53, 3, False), (200, 150, 6, True)]:
            for ee in (False, True):
                kind = "positive" if (om, oa) == ("PLUS", "MIN") else "small"
                add_case(dtype, om, oa, osel, nr, nc, avg, skew, ee, kind)
    # a sweep over every op as mult and as add, every select (small shapes)
    for dtype in (INT, UINT, FLOAT):
        for op in cases.BIN_OPS:
            if not cases.op_valid(dtype, op) or (op == "DIV" and dtype != FLOAT):
                continue
            add_case(dtype, op, "PLUS", "NQZERO", 41, 29, 3, False, False)
            add_case(dtype, "MULT", op, "GEZERO", 41, 29, 3, False, bool(len(op) % 2))
        for osel in cases.SEL_OPS:
            add_case(dtype, "MULT", "PLUS", osel, 30, 30, 4, False, False)
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "mxv_vxm_reference.npz"), **out)
    print("mxv/vxm cases:", cid)

    # whole algorithms on a small symmetric graph (reference src/algorithm.cpp bfs / sssp / pr)
    import torch

    from spla_b200 import graphs

    n, Ap, Aj = graphs.rmat(10, edge_factor=8, seed=7)
    Ap = Ap.numpy().astype(np.uint32)
    Aj = Aj.numpy().astype(np.uint32)
    rows = cases.csr_to_coo_rows(Ap)
    alg = {"n": np.array([n]), "Ap": Ap, "Aj": Aj}
    deg = np.diff(Ap.astype(np.int64))
    src = int(np.argmax(deg > 0))
    alg["source"] = np.array([src])
    Mi = ref.matrix(INT, n, n, rows, Aj, np.ones(len(Aj), dtype=np.int32))
    for mode, name in ((0, "push"), (1, "pull"), (2, "pushpull")):
        d, _ = ref.bfs(Mi, src, mode, 0.05)
        alg["bfs_" + name] = d
    w = graphs.uniform_weights(len(Aj), seed=3).numpy().astype(np.float32)
    # symmetric weights so push (f x A) and pull (A x f) agree as the reference assumes
    key = np.minimum(rows.astype(np.int64), Aj.astype(np.int64)) * n + np.maximum(rows.astype(np.int64), Aj.astype(np.int64))
    _, inv = np.unique(key, return_inverse=True)
    w = w[:inv.max() + 1][inv]
    alg["w"] = w
    Mf = ref.matrix(FLOAT, n, n, rows, Aj, w)
    for mode, name in ((0, "push"), (1, "pull"), (2, "pushpull")):
        d, _ = ref.sssp(Mf, src, mode, 0.05)
        alg["sssp_" + name] = d
    pv = graphs.pagerank_values(torch.from_numpy(Ap.astype(np.int64))).numpy().astype(np.float32)
    alg["pr_values"] = pv
    Mp = ref.matrix(FLOAT, n, n, rows, Aj, pv)
    p, _ = ref.pr(Mp, 0.85, 1e-6)
    alg["pr"] = p
    np.savez_compressed(os.path.join(HERE, "algorithms_reference.npz"), **alg)
    print("algorithm fixtures: n =", n, "nnz =", len(Aj), "bfs levels =", int(alg["bfs_push"].max()))


if __name__ == "__main__":
    main()
