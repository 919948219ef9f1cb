"""Known-answer vectors the reference holds for masked mxv / vxm, transcribed by hand with their source.

Each entry: dict(kind, dtype, ops, M triples, n_rows, n_cols, v (dense list or sparse dict), mask (dense), init, expect).
For vxm `expect` is the dense read-back the reference test checks (untouched entries read as fill 0).
"""
import numpy as np

from cases import INT


def _perf_mxv():
    # reference tests/test_mxv.cpp:91-130: N rows, band of K ones per row, mask = (i % S ? 1 : 0), EQZERO
    N, K, S = 2000, 16, 10  # the reference uses N=1e6, K=256; same construction at a CPU-test size
    rows = np.repeat(np.arange(N), K)
    cols = (rows + np.tile(np.arange(K), N)) % N
    order = np.lexsort((cols, rows))
    return dict(kind="mxv", dtype=INT, ops=("MULT", "PLUS", "EQZERO"), n_rows=N, n_cols=N,
                Ai=rows[order], Aj=cols[order], Ax=np.ones(N * K, dtype=np.int32),
                v=np.ones(N, dtype=np.int32), mask=np.array([1 if i % S else 0 for i in range(N)], dtype=np.int32),
                init=0, expect=np.array([0 if i % S else K for i in range(N)], dtype=np.int32),
                source="tests/test_mxv.cpp:91-130")


def _perf_vxm(ops):
    # reference tests/test_vxm.cpp:91-138 (MULT/PLUS) and :140-187 (BAND/BOR): every K-th vertex is in the
    # frontier with value 1 (all other entries are stored explicitly with value 0) and has W out-edges
    N, K, W, S = 3000, 10, 64, 10
    result = np.zeros(N, dtype=np.int64)
    Ai, Aj = [], []
    for i in range(0, N, K):
        for w in range(W):
            j = (i + w) % N
            Ai.append(i)
            Aj.append(j)
            result[j] += 1
    Ai, Aj = np.array(Ai), np.array(Aj)
    order = np.lexsort((Aj, Ai))
    vi = np.arange(N, dtype=np.uint32)
    vx = np.array([1 if i % K == 0 else 0 for i in range(N)], dtype=np.int32)
    if ops[0] == "MULT":
        exp = np.array([0 if i % S else result[i] for i in range(N)], dtype=np.int32)
    else:
        exp = np.array([0 if i % S else (1 if result[i] else 0) for i in range(N)], dtype=np.int32)
    return dict(kind="vxm", dtype=INT, ops=ops, n_rows=N, n_cols=N, Ai=Ai[order], Aj=Aj[order],
                Ax=np.ones(len(Ai), dtype=np.int32), vi=vi, vx=vx,
                mask=np.array([1 if i % S else 0 for i in range(N)], dtype=np.int32), init=0, expect=exp,
                source="tests/test_vxm.cpp:91-187")


KNOWN = [
    # reference tests/test_mxv.cpp:33-89  mxv_masked.naive  -> r = [0, 14, 0, 1]
    dict(kind="mxv", dtype=INT, ops=("MULT", "PLUS", "EQZERO"), n_rows=4, n_cols=5,
         Ai=[0, 0, 1, 1, 2, 3], Aj=[1, 4, 0, 4, 2, 4], Ax=[2, -9, 2, -8, 3, -1],
         v=[3, 0, 3, 0, -1], mask=[1, 0, 1, 0], init=0, expect=[0, 14, 0, 1], source="tests/test_mxv.cpp:33-89"),
    # reference tests/test_vxm.cpp:33-89  vxm_masked.naive (transposed layout) -> [0, 14, 0, 1]
    dict(kind="vxm", dtype=INT, ops=("MULT", "PLUS", "EQZERO"), n_rows=5, n_cols=4,
         Ai=[0, 1, 2, 4, 4, 4], Aj=[1, 0, 2, 0, 1, 3], Ax=[2, 2, 3, -9, -8, -1],
         vi=[0, 1, 2, 3, 4], vx=[3, 0, 3, 0, -1], mask=[1, 0, 1, 0], init=0, expect=[0, 14, 0, 1],
         source="tests/test_vxm.cpp:33-89"),
    # reference python/pyspla/matrix.py:942-952  mxv docstring 1 -> [., 1, ., 1]
    dict(kind="mxv", dtype=INT, ops=("LAND", "LOR", "GTZERO"), n_rows=4, n_cols=4,
         Ai=[0, 1, 2, 2, 3], Aj=[1, 2, 0, 3, 2], Ax=[1, 2, 3, 4, 5],
         v=[0, 0, 1, 0], mask=[1, 1, 1, 1], init=0, expect=[0, 1, 0, 1], source="python/pyspla/matrix.py:942-952"),
    # reference python/pyspla/matrix.py:954-963  mxv docstring 2 -> [3, 8, 6, .]
    dict(kind="mxv", dtype=INT, ops=("MULT", "PLUS", "EQZERO"), n_rows=4, n_cols=4,
         Ai=[0, 1, 2], Aj=[1, 2, 0], Ax=[1, 2, 3],
         v=[2, 3, 4, 0], mask=[0, 0, 0, 0], init=0, expect=[3, 8, 6, 0], source="python/pyspla/matrix.py:954-963"),
    # reference python/pyspla/vector.py:479-489  vxm docstring 1 -> [1, ., ., 1]
    dict(kind="vxm", dtype=INT, ops=("LAND", "LOR", "GTZERO"), n_rows=4, n_cols=4,
         Ai=[0, 1, 2, 2, 3], Aj=[1, 2, 0, 3, 2], Ax=[1, 2, 3, 4, 5],
         vi=[2], vx=[1], mask=[1, 1, 1, 1], init=0, expect=[1, 0, 0, 1], source="python/pyspla/vector.py:479-489"),
    # reference python/pyspla/vector.py:491-500  vxm docstring 2 -> [12, 2, 6, .]
    dict(kind="vxm", dtype=INT, ops=("MULT", "PLUS", "EQZERO"), n_rows=4, n_cols=4,
         Ai=[0, 1, 2], Aj=[1, 2, 0], Ax=[1, 2, 3],
         vi=[0, 1, 2], vx=[2, 3, 4], mask=[0, 0, 0, 0], init=0, expect=[12, 2, 6, 0], source="python/pyspla/vector.py:491-500"),
    # SURVEY 8c probe on the built reference: MULT/LOR with a single contributing edge returns the RAW product 15
    dict(kind="vxm", dtype=INT, ops=("MULT", "LOR", "ALWAYS"), n_rows=4, n_cols=4,
         Ai=[0, 1, 2, 2, 3], Aj=[1, 2, 0, 3, 2], Ax=[1, 2, 3, 4, 5],
         vi=[3], vx=[3], mask=[0, 0, 0, 0], init=0, expect=[0, 0, 15, 0], source="SURVEY.md 8c (probe of libspla_x64.so)"),
    _perf_mxv(),
    _perf_vxm(("MULT", "PLUS", "EQZERO")),
    _perf_vxm(("BAND", "BOR", "EQZERO")),
]


def as_arrays(case):
    """-> (Ap, Aj, Ax) row-sorted CSR plus typed numpy vectors."""
    dt = {0: np.int32, 1: np.uint32, 2: np.float32}[case["dtype"]]
    Ai = np.asarray(case["Ai"], dtype=np.int64)
    Aj = np.asarray(case["Aj"], dtype=np.uint32)
    Ax = np.asarray(case["Ax"], dtype=dt)
    Ap = np.zeros(case["n_rows"] + 1, dtype=np.int64)
    np.add.at(Ap, Ai + 1, 1)
    Ap = np.cumsum(Ap).astype(np.uint32)
    return Ap, Aj, Ax, dt
