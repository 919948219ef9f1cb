"""Whole traversals driven through the C ABI (spla_b200/algorithms.py mirrors reference src/algorithm.cpp call for call) against
outputs of the unmodified reference CPU backend (tests/golden/algorithms_reference.npz, generator tests/golden/make_golden.py),
plus size-independent properties at benchmark-like sizes (BASELINE configs 1-4)."""
import os

import numpy as np
import pytest
import torch

from gpu_util import assert_values, idx_dev, pagerank_float64, to_dev, to_np

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = {"push": "push", "pull": "pull", "pushpull": "push_pull"}


def golden_matrix(backend, z, values):
    n = int(z["n"][0])
    return n, backend.csr(n, n, idx_dev(z["Ap"], backend), idx_dev(z["Aj"], backend), to_dev(values, backend))


@pytest.mark.parametrize("name", list(MODES))
def test_bfs_matches_reference_depths(backend, name):
    from spla_b200 import algorithms

    z = np.load(os.path.join(GOLDEN, "algorithms_reference.npz"))
    n, M = golden_matrix(backend, z, np.ones(len(z["Aj"]), dtype=np.int32))
    d = algorithms.bfs(backend, M, int(z["source"][0]), mode=MODES[name], front_factor=0.05)
    np.testing.assert_array_equal(to_np(d, np.int32), z["bfs_" + name])  # bit-exact depths


@pytest.mark.parametrize("name", list(MODES))
def test_sssp_matches_reference_distances(backend, name):
    from spla_b200 import algorithms

    z = np.load(os.path.join(GOLDEN, "algorithms_reference.npz"))
    n, M = golden_matrix(backend, z, z["w"])
    d = algorithms.sssp(backend, M, int(z["source"][0]), mode=MODES[name], front_factor=0.05)
    # MIN over single fp adds: order independent => bit-exact (SURVEY 8a note G); 1e-5 relative is the stated tolerance
    np.testing.assert_array_equal(to_np(d, np.float32).view(np.uint32), z["sssp_" + name].view(np.uint32))


def test_pagerank_matches_reference_ranks(backend):
    from spla_b200 import algorithms

    z = np.load(os.path.join(GOLDEN, "algorithms_reference.npz"))
    n, M = golden_matrix(backend, z, z["pr_values"])
    p, iters = algorithms.pagerank(backend, M, 0.85, 1e-6)
    assert_values(to_np(p, np.float32), z["pr"], False, what=f"pr() golden, {iters} iterations",  # strict 1e-5 relative per element
                  bound=lambda: (pagerank_float64(z["Ap"], z["Aj"], z["pr_values"], 0.85, iters), None))


def test_bfs_properties_rmat20(backend):
    """RMAT scale 20 (~31 M edges): push, pull and push-pull give identical depths; depths form a valid BFS labelling."""
    from spla_b200 import algorithms, graphs

    n, Ap64, Aj = graphs.rmat(20, seed=1, device=backend.device)
    torch.cuda.synchronize()
    deg = Ap64[1:] - Ap64[:-1]
    src = int(torch.argmax(deg).item())
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, torch.ones(Aj.numel(), dtype=torch.int32, device=backend.device))
    d = {m: algorithms.bfs(backend, M, src, mode=m, front_factor=0.05) for m in ("push", "pull", "push_pull")}
    torch.cuda.synchronize()
    assert torch.equal(d["push"], d["pull"]) and torch.equal(d["push"], d["push_pull"])
    depth = d["push"].long()
    assert int(depth[src]) == 1
    rows = torch.repeat_interleave(torch.arange(n, device=backend.device), deg)
    du, dv = depth[rows], depth[Aj.long()]
    reached = (du > 0) & (dv > 0)
    assert bool(((du > 0) == (dv > 0)).all())  # symmetric graph: an edge never leaves the component
    assert int((du[reached] - dv[reached]).abs().max()) <= 1  # neighbouring depths differ by at most one
    # every reached vertex but the source has a neighbour one level closer
    has_parent = torch.zeros(n, dtype=torch.bool, device=backend.device)
    has_parent[rows[reached & (dv == du - 1)]] = True
    need = depth > 1
    assert bool(has_parent[need].all())


def test_sssp_grid_matches_manhattan(backend):
    """4-neighbour grid with unit weights (BASELINE config 3 shape, reduced side): distances are Manhattan distances, bit-exact."""
    from spla_b200 import algorithms, graphs

    side = 256
    n, Ap64, Aj = graphs.grid2d(side, device=backend.device)
    torch.cuda.synchronize()
    M = backend.csr(n, n, Ap64.to(torch.int32), Aj, torch.ones(Aj.numel(), dtype=torch.float32, device=backend.device))
    d = algorithms.sssp(backend, M, 0, mode="push_pull", front_factor=0.05)
    torch.cuda.synchronize()
    idx = torch.arange(n, device=backend.device)
    want = (idx % side + idx // side).float()
    assert torch.equal(d, want)
