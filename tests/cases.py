"""Seeded input generators shared by the oracle tests, the golden-fixture script and the GPU parity tests."""
import numpy as np

INT, UINT, FLOAT = 0, 1, 2
NP = {INT: np.int32, UINT: np.uint32, FLOAT: np.float32}

BIN_OPS = ["PLUS", "MINUS", "MULT", "DIV", "MINUS_POW2", "FIRST", "SECOND", "BONE",
           "MIN", "MAX", "LOR", "LAND", "BOR", "BAND", "BXOR"]
SEL_OPS = ["EQZERO", "NQZERO", "GTZERO", "GEZERO", "LTZERO", "LEZERO", "ALWAYS", "NEVER"]
BITWISE = ("BOR", "BAND", "BXOR")
ASSOC = ("PLUS", "MULT", "MIN", "MAX", "LOR", "LAND", "BOR", "BAND", "BXOR")

# semirings the reference's algorithms and tests use on the path
# (src/algorithm.cpp:97-99,208-210,312; tests/test_mxv.cpp:74; tests/test_vxm.cpp:74,174; README.md:79)
NAMED_SEMIRINGS = [
    (INT, "BAND", "BOR", "EQZERO"),
    (INT, "LAND", "LOR", "EQZERO"),
    (INT, "MULT", "PLUS", "EQZERO"),
    (INT, "LAND", "LOR", "GTZERO"),
    (UINT, "BAND", "BOR", "EQZERO"),
    (UINT, "MULT", "PLUS", "NQZERO"),
    (FLOAT, "MULT", "PLUS", "ALWAYS"),
    (FLOAT, "PLUS", "MIN", "ALWAYS"),
    (FLOAT, "LAND", "LOR", "EQZERO"),
]


def op_valid(dtype, op):
    return not (dtype == FLOAT and op in BITWISE)


def exact_expected(dtype, op_mult, op_add):
    """True when the CUDA result must be bit-identical to the sequential CPU fold (SURVEY 8a note G):
    everything except FLOAT reductions whose value depends on the summation order."""
    if dtype != FLOAT:
        return True
    return op_add in ("MIN", "MAX", "LOR", "LAND", "FIRST", "SECOND", "BONE", "MINUS", "DIV", "MINUS_POW2")


def rand_values(rng, dtype, n, kind="small"):
    if dtype == FLOAT:
        if kind == "unit":
            return rng.uniform(0.0, 1.0, n).astype(np.float32)
        if kind == "positive":
            return rng.uniform(0.5, 2.0, n).astype(np.float32)
        return (rng.integers(-2, 3, n) * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    if dtype == INT:
        if kind == "positive":
            return rng.integers(1, 6, n).astype(np.int32)
        return rng.integers(-3, 4, n).astype(np.int32)
    if kind == "positive":
        return rng.integers(1, 6, n).astype(np.uint32)
    return rng.integers(0, 4, n).astype(np.uint32)


def rand_csr(rng, dtype, n_rows, n_cols, avg_deg, skew=False, kind="small", empty_frac=0.2):
    """Row-sorted CSR with ascending columns per row (what Matrix::build of sorted COO / set_* produces,
    reference src/cpu/cpu_format_lil.hpp:54-73). `skew` adds a few very long rows."""
    deg = rng.poisson(avg_deg, n_rows).astype(np.int64)
    deg[rng.random(n_rows) < empty_frac] = 0
    if skew and n_rows > 4:
        hubs = rng.choice(n_rows, size=max(1, n_rows // 64), replace=False)
        deg[hubs] = rng.integers(n_cols // 2, n_cols + 1, len(hubs))
    deg = np.minimum(deg, n_cols)
    Ap = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(deg, out=Ap[1:])
    Aj = np.empty(Ap[-1], dtype=np.uint32)
    for i in range(n_rows):
        d = deg[i]
        if d:
            if d > n_cols // 4:
                cols = rng.permutation(n_cols)[:d]
            else:
                cols = np.unique(rng.integers(0, n_cols, int(d * 1.3) + 4))
                while len(cols) < d:
                    cols = np.unique(np.concatenate([cols, rng.integers(0, n_cols, d)]))
                cols = rng.permutation(cols)[:d]
            Aj[Ap[i]:Ap[i + 1]] = np.sort(cols)
    Ax = rand_values(rng, dtype, int(Ap[-1]), kind)
    return Ap.astype(np.uint32), Aj, Ax


def csr_to_coo_rows(Ap):
    Ap = Ap.astype(np.int64)
    return np.repeat(np.arange(len(Ap) - 1, dtype=np.uint32), np.diff(Ap))


def rand_frontier(rng, dtype, n, nv, kind="small"):
    nv = min(nv, n)
    vi = np.sort(rng.choice(n, size=nv, replace=False)).astype(np.uint32)
    return vi, rand_values(rng, dtype, nv, kind)
