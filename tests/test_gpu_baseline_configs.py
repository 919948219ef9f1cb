"""BASELINE.json's configs on their own inputs, CUDA path vs the UNMODIFIED reference CPU backend (oracle/_ref, which travels to
the GPU box as a prebuilt .so) PER ELEMENT at benchmark sizes:

  config 2  PageRank step  mxv_masked FLOAT MULT/PLUS (ALWAYS and the bench's NQZERO all-ones mask), A = 0.85/outdeg, v = 1/N, RMAT-20 and RMAT-22
            + the whole pr() loop on RMAT-18
  config 1  BFS push / pull / push-pull on RMAT-16 (INT BAND/BOR/EQZERO): depths bit-exact
  config 3  SSSP on a 4-neighbour grid with uniform [1, 2) weights (FLOAT PLUS/MIN): distances bit-exact
  config 4/5 shape  the BFS-semiring pull with a real mask and the push at a 1 % frontier on RMAT-20: bit-exact

Bar: tests/gpu_util.py (bit-exact; FLOAT sums |gpu - ref| <= 1e-5 |ref| per element, derived escape counted and printed).
The measured maximum relative errors are printed and appended to gpurun_out/parity_stats.jsonl.
"""
import numpy as np
import pytest
import torch

from gpu_util import assert_values, mxv_bound, pagerank_float64, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import oracle as orc

    if not orc.ref_available():
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    return orc.RefSpla()


def _host_csr(Ap64, Aj, Ax):
    return Ap64.cpu().numpy().astype(np.uint32), Aj.cpu().numpy().astype(np.uint32), Ax.cpu().numpy()


@pytest.mark.parametrize("scale", [20, 22])
def test_config2_pagerank_step_vs_reference(backend, ref, scale):
    from oracle import oracle as orc
    from spla_b200 import graphs

    n, Ap64, Aj = graphs.rmat(scale, edge_factor=16, seed=2, device=backend.device)
    Ax = graphs.pagerank_values(Ap64, 0.85)
    torch.cuda.synchronize()
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, Ax)
    v = torch.full((n,), 1.0 / n, dtype=torch.float32, device=backend.device)
    ones = torch.ones(n, dtype=torch.float32, device=backend.device)
    torch.cuda.synchronize()
    r_always = backend.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0)
    r_bench = backend.mxv_masked(M, v, ones, "MULT", "PLUS", "NQZERO", 0.0)
    backend.sync()
    hAp, hAj, hAx = _host_csr(Ap64, Aj, Ax)
    hv = np.full(n, 1.0 / n, dtype=np.float32)
    Mr = ref.matrix_from_csr(orc.FLOAT, n, hAp, hAj, hAx)
    want = ref.mxv_masked(Mr, "MULT", "PLUS", "ALWAYS", hv, np.ones(n, np.float32), 0.0)
    del Mr
    bound = lambda: mxv_bound("MULT", "PLUS", hAp, hAj, hAx, hv, 0.0)  # noqa: E731
    g = to_np(r_always, np.float32)
    assert_values(g, want, False, what=f"config2 RMAT-{scale} MULT/PLUS/ALWAYS vs oracle/_ref (n={n}, nnz={len(hAj)})", bound=bound)
    assert_values(to_np(r_bench, np.float32), want, False, what=f"config2 RMAT-{scale} bench workload MULT/PLUS/NQZERO all-ones mask vs oracle/_ref", bound=bound)
    rel = np.abs(g.astype(np.float64) - want) / np.maximum(np.abs(want.astype(np.float64)), 1e-300)
    print(f"[parity] config 2 RMAT-{scale}: max rel err vs the reference CPU backend {rel[want != 0].max():.3e} over {n} rows, {len(hAj)} entries")


def test_config2_pagerank_loop_vs_reference(backend, ref):
    from oracle import oracle as orc
    from spla_b200 import algorithms, graphs

    n, Ap64, Aj = graphs.rmat(18, edge_factor=16, seed=2, device=backend.device)
    Ax = graphs.pagerank_values(Ap64, 0.85)
    torch.cuda.synchronize()
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, Ax)
    p, iters = algorithms.pagerank(backend, M, 0.85, 1e-6)
    hAp, hAj, hAx = _host_csr(Ap64, Aj, Ax)
    want, _ = ref.pr(ref.matrix_from_csr(orc.FLOAT, n, hAp, hAj, hAx), 0.85, 1e-6)
    assert_values(to_np(p, np.float32), want, False, what=f"config2 pr() RMAT-18, {iters} iterations, vs oracle/_ref pr()",
                  bound=lambda: (pagerank_float64(hAp, hAj, hAx, 0.85, iters), None))


def test_config1_bfs_rmat16_vs_reference(backend, ref):
    from oracle import oracle as orc
    from spla_b200 import algorithms, graphs

    n, Ap64, Aj = graphs.rmat(16, edge_factor=16, seed=1, device=backend.device)
    torch.cuda.synchronize()
    ones = torch.ones(Aj.numel(), dtype=torch.int32, device=backend.device)
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, ones)
    deg = (Ap64[1:] - Ap64[:-1])
    src = int(torch.nonzero(deg > 0).flatten()[0].item())
    hAp, hAj, hAx = _host_csr(Ap64, Aj, ones)
    Mr = ref.matrix_from_csr(orc.INT, n, hAp, hAj, hAx)
    for mode_id, mode in enumerate(("push", "pull", "push_pull")):
        want, _ = ref.bfs(Mr, src, mode_id, 0.05)
        got = algorithms.bfs(backend, M, src, mode=mode, front_factor=0.05)
        np.testing.assert_array_equal(to_np(got, np.int32), want, err_msg=f"bfs {mode}")


def test_config3_sssp_grid_vs_reference(backend, ref):
    from oracle import oracle as orc
    from spla_b200 import algorithms, graphs

    side = 384
    n, Ap64, Aj = graphs.grid2d(side, device=backend.device)
    # symmetric weights uniform in [1, 2): w(i, j) = w(j, i), derived from the unordered pair
    rows = torch.repeat_interleave(torch.arange(n, device=backend.device), Ap64[1:] - Ap64[:-1])
    lo, hi = torch.minimum(rows, Aj.long()), torch.maximum(rows, Aj.long())
    h = ((lo * 2654435761 + hi * 40503) % 1000003).to(torch.float32) / 1000003.0
    w = (1.0 + h).to(torch.float32)
    torch.cuda.synchronize()
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, w)
    hAp, hAj, hAx = _host_csr(Ap64, Aj, w)
    Mr = ref.matrix_from_csr(orc.FLOAT, n, hAp, hAj, hAx)
    want, _ = ref.sssp(Mr, 0, 2, 0.05)
    got = algorithms.sssp(backend, M, 0, mode="push_pull", front_factor=0.05)
    np.testing.assert_array_equal(to_np(got, np.float32).view(np.uint32), want.view(np.uint32))  # MIN of single fp adds: bit-exact


def test_config45_bfs_semiring_rmat20_vs_reference(backend, ref):
    """mxv INT BAND/BOR/EQZERO with a 50 % mask (with and without early exit) and vxm at a 1 % frontier on RMAT-20, per element."""
    from oracle import oracle as orc
    from spla_b200 import graphs

    n, Ap64, Aj = graphs.rmat(20, edge_factor=16, seed=2, device=backend.device)
    g = torch.Generator(device=backend.device)
    g.manual_seed(9)
    ones = torch.ones(Aj.numel(), dtype=torch.int32, device=backend.device)
    visited = (torch.rand(n, generator=g, device=backend.device) < 0.5).to(torch.int32)
    front = (torch.rand(n, generator=g, device=backend.device) < 0.3).to(torch.int32)
    vi = torch.nonzero(torch.rand(n, generator=g, device=backend.device) < 0.01).flatten().to(torch.int32)
    vx = torch.ones(vi.numel(), dtype=torch.int32, device=backend.device)
    torch.cuda.synchronize()
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, ones)
    r0 = backend.mxv_masked(M, front, visited, "BAND", "BOR", "EQZERO", 0)
    r1 = backend.mxv_masked(M, front, visited, "BAND", "BOR", "EQZERO", 0, early_exit=True)
    ri, rx = backend.vxm_masked(M, vi, vx, visited, "BAND", "BOR", "EQZERO")
    backend.sync()
    hAp, hAj, hAx = _host_csr(Ap64, Aj, ones)
    Mr = ref.matrix_from_csr(orc.INT, n, hAp, hAj, hAx)
    hvis, hfront = visited.cpu().numpy(), front.cpu().numpy()
    np.testing.assert_array_equal(to_np(r0, np.int32), ref.mxv_masked(Mr, "BAND", "BOR", "EQZERO", hfront, hvis, 0))
    np.testing.assert_array_equal(to_np(r1, np.int32), ref.mxv_masked(Mr, "BAND", "BOR", "EQZERO", hfront, hvis, 0, early_exit=True))
    wi, wx = ref.vxm_masked(Mr, "BAND", "BOR", "EQZERO", vi.cpu().numpy().view(np.uint32), vx.cpu().numpy(), hvis)
    np.testing.assert_array_equal(to_np(ri, np.uint32), wi)
    np.testing.assert_array_equal(to_np(rx, np.int32), wx)
