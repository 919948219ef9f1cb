"""tools/exp_mxv.py -- kernel experiments on the GPU box (not a bench): pull mxv on RMAT with the original and with
degree-relabelled column ids (hubs first), to size the benefit of a hub cache. Prints one line per variant."""
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
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()

be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(args.scale, 16, seed=2, device=dev)
Ax = graphs.pagerank_values(Ap, 0.85)
nnz = Aj.numel()
Ap32 = Ap.to(torch.int32)
deg = (Ap[1:] - Ap[:-1])
v = torch.rand(n, device=dev)
mask = torch.ones(n, device=dev)
torch.cuda.synchronize()


def timeit(fn):
    for _ in range(3):
        fn()
    be.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(be.stream)
    for _ in range(args.reps):
        fn()
    e1.record(be.stream)
    be.sync()
    return e0.elapsed_time(e1) / args.reps


def report(name, ms):
    alg = 4 * (n + 1) + 4 * n + 8 * nnz + 4 * min(n, nnz)
    print(f"{name:40s} {ms:8.3f} ms  {nnz / ms / 1e6:8.1f} GTEPS  {alg / ms / 1e6:8.1f} GB/s alg", flush=True)


with torch.cuda.stream(be.stream):
    r = torch.empty(n, device=dev)
    be.set_option("mxv_hub", 0)
    M = be.csr(n, n, Ap32, Aj, Ax)
    report("orig ALWAYS", timeit(lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)))
    report("orig NQZERO all-ones mask", timeit(lambda: be.mxv_masked(M, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r)))
    r0 = r.clone()
    be.set_option("mxv_hub", 1)
    Mh = be.csr(n, n, Ap32, Aj, Ax)
    for persist in (0, 1, 0, 1):
        be.set_option("mxv_l2_persist", persist)
        report(f"hub default, l2 persist {persist} ALWAYS", timeit(lambda: be.mxv_masked(Mh, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)))
        report(f"hub default, l2 persist {persist} NQZERO", timeit(lambda: be.mxv_masked(Mh, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r)))
    be.set_option("mxv_l2_persist", 0)
    del Mh
    be.set_option("mxv_hub", 1)
    be.set_option("mxv_hub_smem", 16384)
    # relabel columns by descending degree: new id = rank
    order = torch.argsort(deg, descending=True, stable=True)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(n, device=dev)
    Aj2 = rank[Aj.long()].to(torch.int32)
    v2 = v[order].contiguous()
    be.sync()
    torch.cuda.synchronize()
    M2 = be.csr(n, n, Ap32, Aj2, Ax)
    report("hub-first columns ALWAYS", timeit(lambda: be.mxv_masked(M2, v2, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)))
    be.sync()
    err = ((r - r0).abs() / r0.abs().clamp(min=1e-30)).max().item()
    print("max rel diff relabelled vs orig:", err)
    for dens in (0.5, 0.1, 0.01):
        m = (torch.rand(n, device=dev) < dens).float()
        sel = int(deg[m != 0].sum().item())
        torch.cuda.synchronize()
        ms = timeit(lambda: be.mxv_masked(M2, v2, m, "MULT", "PLUS", "NQZERO", 0.0, out=r))
        alg = 4 * (n + 1) + 8 * n + 8 * sel + 4 * min(n, sel)
        print(f"mask density {dens} (hub + relabel): {ms:.3f} ms  {sel / ms / 1e6:.1f} GTEPS(selected)  {alg / ms / 1e6:.1f} GB/s alg", flush=True)
