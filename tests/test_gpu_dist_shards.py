"""The multi-GPU path behind the C ABI (spla_b200/csrc/dist.cu: splacu_dist_* / splacu_dcsr_*): a matrix sharded over N shards --
row blocks for the pull, column windows for the push -- must give what the single-device entry points give: bit-exact for integer
and order-independent work, and the oracle's values.

Shards may share a device, so the whole logic (boundaries, slices, window copies, ordered concatenation of the push results) runs
on the single-GPU box with device lists like [0, 0, 0]; with more than one GPU visible (gpurun --gpus 2) the same tests also run
over distinct devices, where v travels by ncclBroadcast and the windows by peer copies."""
import zlib

import numpy as np
import pytest
import torch

import cases
from cases import FLOAT, INT, UINT
from gpu_util import assert_values, idx_dev, make_csr, mxv_bound, to_dev, to_np

pytestmark = pytest.mark.gpu


def device_lists():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 1
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], [0, 1, 1]]
    if n >= 4:
        lists += [[0, 1, 2, 3]]
    return lists


@pytest.mark.parametrize("devs", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("dtype,om,oa,osel", [(INT, "MULT", "PLUS", "EQZERO"), (INT, "BAND", "BOR", "EQZERO"), (FLOAT, "PLUS", "MIN", "ALWAYS"),
                                               (UINT, "MULT", "MAX", "NQZERO"), (FLOAT, "MULT", "PLUS", "NQZERO"), (INT, "FIRST", "SECOND", "GTZERO")])
def test_sharded_products_match_oracle(backend, oracle, devs, dtype, om, oa, osel):
    rng = np.random.default_rng(zlib.crc32(repr((dtype, om, oa, len(devs))).encode()))
    n_rows, n_cols = 5000, 4200
    kind = "positive" if (om, oa) == ("PLUS", "MIN") else ("unit" if dtype == FLOAT else "small")
    Ap, Aj, Ax = cases.rand_csr(rng, dtype, n_rows, n_cols, 9, skew=True, kind=kind)
    M = make_csr(backend, n_rows, n_cols, Ap, Aj, Ax)
    G = backend.dist_group(devs)
    D = G.csr(M)
    rb, _ = D.bounds()
    assert rb[0] == 0 and rb[-1] == n_rows and all(rb[p] <= rb[p + 1] for p in range(len(devs)))
    np_t = cases.NP[dtype]
    exact = cases.exact_expected(dtype, om, oa)
    for ee in (False, True):
        v = cases.rand_values(rng, dtype, n_cols, kind)
        mask = cases.rand_values(rng, dtype, n_rows)
        init = np.float32(3.0e38) if (om, oa) == ("PLUS", "MIN") else 2
        want = oracle.mxv_masked(dtype, om, oa, osel, Ap, Aj, Ax, v, mask, init, ee)
        got = D.mxv_masked(to_dev(v, backend), to_dev(mask, backend), om, oa, osel, init, early_exit=ee)
        backend.sync()
        assert_values(to_np(got, np_t), want, exact or ee, what=f"sharded mxv {devs} {om}/{oa} ee={ee}", bound=lambda: mxv_bound(om, oa, Ap, Aj, Ax, v, init))
    for nv in (0, 1, 70, 900, n_rows):
        vi, vx = cases.rand_frontier(rng, dtype, n_rows, nv, kind)
        maskc = cases.rand_values(rng, dtype, n_cols)
        wi, wx = oracle.vxm_masked(dtype, om, oa, osel, Ap, Aj, Ax, n_cols, vi, vx, maskc)
        gi, gx = D.vxm_masked(idx_dev(vi, backend), to_dev(vx, backend), to_dev(maskc, backend), om, oa, osel)
        backend.sync()
        assert np.array_equal(to_np(gi, np.uint32), wi), f"sharded vxm pattern {devs} nv={nv}"
        assert_values(to_np(gx, np_t), wx, exact, what=f"sharded vxm {devs} {om}/{oa} nv={nv}")
    _, cb = D.bounds()
    assert cb[0] == 0 and cb[-1] == n_cols
    # the single-device handle keeps working next to the sharded one (user ops, other callers)
    v = cases.rand_values(rng, dtype, n_cols, kind)
    mask = cases.rand_values(rng, dtype, n_rows)
    a = backend.mxv_masked(M, to_dev(v, backend), to_dev(mask, backend), om, oa, osel, 1)
    b = D.mxv_masked(to_dev(v, backend), to_dev(mask, backend), om, oa, osel, 1)
    backend.sync()
    if exact:
        assert torch.equal(a, b)


@pytest.mark.parametrize("devs", device_lists()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_bfs_step_large(backend, devs):
    """RMAT-18 with column classes on every shard and the structure-only push: a BFS-like sequence on the sharded handle equals the
    single-device handle bit for bit (pull with and without early exit, push at a large frontier)."""
    from spla_b200 import graphs

    n, Ap64, Aj = graphs.rmat(18, edge_factor=16, seed=3, device=backend.device)
    ones = torch.ones(Aj.numel(), dtype=torch.int32, device=backend.device)
    g = torch.Generator(device=backend.device)
    g.manual_seed(4)
    front = (torch.rand(n, generator=g, device=backend.device) < 0.2).to(torch.int32)
    visited = (torch.rand(n, generator=g, device=backend.device) < 0.4).to(torch.int32)
    vi = torch.nonzero(torch.rand(n, generator=g, device=backend.device) < 0.05).flatten().to(torch.int32)
    vx = torch.ones(vi.numel(), dtype=torch.int32, device=backend.device)
    torch.cuda.synchronize()
    try:
        backend.set_option("mxv_hub", 3)  # column classes on a matrix this small too (per shard)
        backend.set_option("mxv_phase_slots", 4096)
        M = backend.csr(n, n, Ap64.to(torch.int32), Aj, ones)
        D = backend.dist_group(devs).csr(M)
    finally:
        backend.set_option("mxv_hub", 1)
        backend.set_option("mxv_phase_slots", 45056)
    for ee in (False, True):
        a = backend.mxv_masked(M, front, visited, "BAND", "BOR", "EQZERO", 0, early_exit=ee)
        b = D.mxv_masked(front, visited, "BAND", "BOR", "EQZERO", 0, early_exit=ee)
        backend.sync()
        assert torch.equal(a, b), ee
    a = backend.mxv_masked(M, front, visited, "MULT", "PLUS", "NQZERO", 7)
    b = D.mxv_masked(front, visited, "MULT", "PLUS", "NQZERO", 7)
    ai, ax = backend.vxm_masked(M, vi, vx, visited, "BAND", "BOR", "EQZERO")
    bi, bx = D.vxm_masked(vi, vx, visited, "BAND", "BOR", "EQZERO")
    ci, cx = D.vxm_masked(vi, vx, visited, "MULT", "PLUS", "EQZERO")
    di, dx = backend.vxm_masked(M, vi, vx, visited, "MULT", "PLUS", "EQZERO")
    backend.sync()
    assert torch.equal(a, b)
    assert torch.equal(ai, bi) and torch.equal(ax, bx)
    assert torch.equal(ci, di) and torch.equal(cx, dx)
