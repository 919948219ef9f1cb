"""tools/exp_tail_ranges.py -- pull mxv (FLOAT MULT/PLUS/ALWAYS) on graphs whose vector outgrows the L2, for several widths of
the tail-class column windows (option mxv_tail_range_log2; 31 = one tail class over all of v, the behaviour before the option)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", default="rmat:25,uniform:25,rmat:24,rmat:26")
ap.add_argument("--shifts", default="31,24,23,22")
ap.add_argument("--masked", action="store_true", help="NQZERO with an all-ones mask (the bench workload) instead of ALWAYS")
args = ap.parse_args()
be = Backend(0)
dev = be.device
for spec in args.graphs.split(","):
    kind, scale = spec.split(":")
    scale = int(scale)
    try:
        if kind == "rmat":
            n, Ap, Aj = graphs.rmat(scale, 16, seed=2, device=dev)
        else:
            n, Ap, Aj = graphs.uniform_random(scale, 16, seed=4, device=dev)
        nnz = int(Aj.numel())
        Ap32 = Ap.to(torch.int32)
        del Ap
        torch.cuda.empty_cache()
        g = torch.Generator(device=dev)
        g.manual_seed(4)
        Ax = torch.rand(nnz, generator=g, device=dev)
        v = torch.rand(n, generator=g, device=dev)
        r = torch.empty_like(v)
        first = None
        mask = torch.ones(n, dtype=torch.float32, device=dev) if args.masked else None
        sel = "NQZERO" if args.masked else "ALWAYS"
        for shift in [int(x) for x in args.shifts.split(",")]:
            be.set_option("mxv_tail_range_log2", shift)
            M = be.csr(n, n, Ap32, Aj, Ax)
            info = be.csr_info(M)
            for _ in range(2):
                be.mxv_masked(M, v, mask, "MULT", "PLUS", sel, 0.0, out=r)
            be.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20 if scale <= 24 else 5
            e0.record(be.stream)
            for _ in range(reps):
                be.mxv_masked(M, v, mask, "MULT", "PLUS", sel, 0.0, out=r)
            e1.record(be.stream)
            be.sync()
            ms = e0.elapsed_time(e1) / reps
            if first is None:
                first = r.clone()
                err = 0.0
            else:
                err = float(((r - first).abs() / first.abs().clamp(min=1e-30)).max().item())
            print(f"{kind}-{scale} nnz {nnz} window 2^{shift}: {ms:8.3f} ms  {nnz / ms / 1e6:7.1f} GTEPS  classes {len(info['phase_nnz'])} "
                  f"{[round(x / nnz, 3) for x in info['phase_nnz']]}  max rel diff vs first {err:.2e}", flush=True)
            del M
            torch.cuda.empty_cache()
        del Ap32, Aj, Ax, v, r, first
    except Exception as e:
        print(f"{kind}-{scale} failed: {str(e)[:300]}", flush=True)
    torch.cuda.empty_cache()
be.set_option("mxv_tail_range_log2", 24)
