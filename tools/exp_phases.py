"""tools/exp_phases.py -- GPU experiment (not a bench): pull mxv on RMAT with the column-class phases at several
(hub classes, slots per class) settings against the single-pass kernels. One line per variant."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--phases", type=str, default="1,2,4,6,8")
ap.add_argument("--slots", type=str, default="49152")
ap.add_argument("--per-phase", type=str, default="4")
args = ap.parse_args()
per_phase = [int(x) for x in args.per_phase.split(",")]

be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(args.scale, 16, seed=2, device=dev)
Ax = graphs.pagerank_values(Ap, 0.85)
nnz = Aj.numel()
Ap32 = Ap.to(torch.int32)
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


def report(name, ms, masked):
    alg = 4 * (n + 1) + 4 * n + 8 * nnz + 4 * min(n, nnz) + (4 * n if masked else 0)
    print(f"{name:46s} {ms:8.3f} ms  {nnz / ms / 1e6:8.1f} GTEPS  {alg / ms / 1e6:8.1f} GB/s alg", flush=True)


with torch.cuda.stream(be.stream):
    r = torch.empty(n, device=dev)
    be.set_option("mxv_hub", 0)
    M = be.csr(n, n, Ap32, Aj, Ax)
    report("plain ALWAYS", timeit(lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)), False)
    r0 = r.clone()
    del M
    be.set_option("mxv_hub", 2)
    M = be.csr(n, n, Ap32, Aj, Ax)
    report("single-pass hub cache ALWAYS", timeit(lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)), False)
    report("single-pass hub cache NQZERO", timeit(lambda: be.mxv_masked(M, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r)), True)
    del M
    be.set_option("mxv_hub", 3)
    for slots in [int(x) for x in args.slots.split(",")]:
        for ph in [int(x) for x in args.phases.split(",")]:
            be.set_option("mxv_phases", ph)
            be.set_option("mxv_phase_slots", slots)
            be.sync()
            t0 = time.perf_counter()
            M = be.csr(n, n, Ap32, Aj, Ax)
            be.sync()
            t_create = time.perf_counter() - t0
            info = be.csr_info(M)
            shares = " ".join(f"{x / nnz:.3f}" for x in info["phase_nnz"])
            print(f"phases={ph} slots={slots}: create {t_create * 1e3:.0f} ms, n_hub {info['n_hub']}, class shares {shares}", flush=True)
            report(f"  phases={ph} slots={slots} ALWAYS", timeit(lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r)), False)
            be.sync()
            err = ((r - r0).abs() / r0.abs().clamp(min=1e-30)).max().item()
            report(f"  phases={ph} slots={slots} NQZERO", timeit(lambda: be.mxv_masked(M, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r)), True)
            be.sync()
            err2 = ((r - r0).abs() / r0.abs().clamp(min=1e-30)).max().item()
            print(f"  max rel diff vs plain: {err:.2e} / {err2:.2e}", flush=True)
            if ph in per_phase:
                for dens in (0.9, 0.5, 0.1):
                    m = (torch.rand(n, device=dev) < dens).float()
                    torch.cuda.synchronize()
                    ms = timeit(lambda: be.mxv_masked(M, v, m, "MULT", "PLUS", "NQZERO", 0.0, out=r))
                    print(f"    mask density {dens}: {ms:.3f} ms", flush=True)
                for p in range(len(info["phase_nnz"])):
                    be.set_option("mxv_phase_only", p + 1)
                    ms_a = timeit(lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r))
                    ms_m = timeit(lambda: be.mxv_masked(M, v, mask, "MULT", "PLUS", "NQZERO", 0.0, out=r))
                    e = info["phase_nnz"][p]
                    print(f"    class {p}: {e} entries  ALWAYS {ms_a:.3f} ms ({e / ms_a / 1e6:.0f} GTEPS)  NQZERO {ms_m:.3f} ms", flush=True)
                be.set_option("mxv_phase_only", 0)
            del M
