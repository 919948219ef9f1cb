"""tools/exp_bfs.py -- whole-BFS timing on RMAT through the C ABI (experiment, not the bench)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import algorithms, graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
args = ap.parse_args()
be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(args.scale, 16, seed=1, device=dev)
nnz = Aj.numel()
deg = Ap[1:] - Ap[:-1]
ones = torch.ones(nnz, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
M = be.csr(n, n, Ap.to(torch.int32), Aj, ones)
src = int(torch.argmax(deg).item())
for selbits in (1,):
  for mode in ("push_pull", "pull", "push"):
    for rep in range(3):
        trace = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v = algorithms.bfs(be, M, src, mode=mode, front_factor=0.05, trace=trace)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    reached = int((v > 0).sum().item())
    edges = int(deg[v > 0].sum().item())
    print(f"bfs {mode:9s} scale {args.scale}: {dt * 1e3:8.3f} ms  levels {len(trace)}  reached {reached}  {edges / dt / 1e9:8.1f} GTEPS (edges in component / time)  trace {trace}", flush=True)
# per-op timing of one push_pull run (synchronising after every op: upper bounds)
import types


def timed(name, fn):
    def w(*a, **k):
        be.sync()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        be.sync()
        print(f"   {name:18s} {(time.perf_counter() - t0) * 1e3:8.3f} ms", flush=True)
        return r
    return w


for name in ("v_assign_masked", "vxm_masked", "mxv_masked", "dense_to_coo", "coo_to_dense", "v_count_mf"):
    setattr(be, name, timed(name, getattr(be, name)))
algorithms.bfs(be, M, src, mode="push_pull", front_factor=0.05)
