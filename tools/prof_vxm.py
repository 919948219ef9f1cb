"""tools/prof_vxm.py -- push vxm calls on RMAT-24 (run under ncu): a 5 % random frontier and the frontier of the heaviest BFS level."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(24, 16, seed=2, device=dev)
nnz = Aj.numel()
ones = torch.ones(nnz, dtype=torch.int32, device=dev)
deg = (Ap[1:] - Ap[:-1])
g = torch.Generator(device=dev)
g.manual_seed(9)
torch.cuda.synchronize()
with torch.cuda.stream(be.stream):
    be.set_option("mxv_hub", 0)
    M = be.csr(n, n, Ap.to(torch.int32), Aj, ones)
    visited = torch.zeros(n, dtype=torch.int32, device=dev)
    ri = torch.empty(n, dtype=torch.int32, device=dev)
    rx = torch.empty(n, dtype=torch.int32, device=dev)
    for name, vi in (("random 5 %", torch.nonzero(torch.rand(n, generator=g, device=dev) < 0.05).flatten().to(torch.int32)),
                     ("400 K highest degrees", torch.sort(torch.topk(deg, 400000).indices).values.to(torch.int32))):
        vx = torch.ones(vi.numel(), dtype=torch.int32, device=dev)
        ef = int(deg[vi.long()].sum().item())
        torch.cuda.synchronize()
        for _ in range(2):
            nr = be.vxm_masked(M, vi, vx, visited, "BAND", "BOR", "EQZERO", out=(ri, rx))
        be.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(be.stream)
        for _ in range(5):
            be.vxm_masked(M, vi, vx, visited, "BAND", "BOR", "EQZERO", out=(ri, rx))
        e1.record(be.stream)
        be.sync()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name}: nv {vi.numel()} edges {ef}  {ms:.3f} ms  {ef / ms / 1e6:.1f} GTEPS", flush=True)
