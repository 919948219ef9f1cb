"""tools/prof_phase.py -- one pull mxv call on RMAT with column-class phases (run under ncu)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--phases", type=int, default=4)
ap.add_argument("--masked", type=int, default=0)
ap.add_argument("--calls", type=int, default=1)
ap.add_argument("--density", type=float, default=1.0)
args = ap.parse_args()
be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(args.scale, 16, seed=2, device=dev)
Ax = graphs.pagerank_values(Ap, 0.85)
v = torch.rand(n, device=dev)
mask = (torch.rand(n, device=dev) < args.density).float()
torch.cuda.synchronize()
with torch.cuda.stream(be.stream):
    r = torch.empty(n, device=dev)
    be.set_option("mxv_hub", 3)
    be.set_option("mxv_phases", args.phases)
    M = be.csr(n, n, Ap.to(torch.int32), Aj, Ax)
    print(be.csr_info(M))
    for _ in range(args.calls):
        if args.masked in (1, 2):
            be.mxv_masked(M, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r)
        if args.masked in (0, 2):
            be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)
    be.sync()
