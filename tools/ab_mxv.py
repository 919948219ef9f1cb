"""tools/ab_mxv.py -- A/B of the pull product on ONE graph in ONE process (GPU box): every --cfg is a comma list of
splacu options (build- and call-time); the handle is rebuilt per config, the PageRank step of bench.py is timed with CUDA events
and its result compared per element with the first config's. Not a bench: prints one line per config.

    python tools/ab_mxv.py --scale 24 --cfg "" --cfg "mxv_red=1" --cfg "mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=10"
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402

DEFAULTS = {"mxv_hub": 1, "mxv_phase_slots": 45056, "mxv_phases": 4, "mxv_red": 1, "mxv_row_classes": 1, "mxv_row_min_count": 64, "mxv_tail_range_log2": 24, "mxv_phase_only": 0, "mxv_l2_persist": 0, "mxv_fixup_merge": 2, "mxv_row_min_nnz": 25165824, "mxv_bank_order": 1, "mxv_pdl": 1}

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--cfg", action="append", default=[])
ap.add_argument("--select", default="NQZERO")
ap.add_argument("--out", default=None)
ap.add_argument("--shard", type=int, default=1, help="time the first of K nnz-balanced row blocks (what one rank of a K-GPU run holds)")
ap.add_argument("--profile", action="store_true", help="per-launch device times inside the stream (splacu_profile_enable)")
args = ap.parse_args()

be = Backend(0)
dev = be.device
n, Ap, Aj = graphs.rmat(args.scale, 16, seed=2, device=dev)
Ax = graphs.pagerank_values(Ap, 0.85)
nnz = Aj.numel()
n_rows = n
if args.shard > 1:
    from spla_b200 import dist as sd
    bounds = sd.balanced_boundaries(Ap, args.shard)
    Ap, Aj, Ax = sd.row_slice(Ap, Aj, Ax, bounds[0], bounds[1])
    n_rows = bounds[1] - bounds[0]
    nnz = Aj.numel()
Ap32 = Ap.to(torch.int32)
v = torch.full((n,), 1.0 / n, device=dev)
mask = torch.ones(n_rows, device=dev)
torch.cuda.synchronize()
alg = 4 * (n_rows + 1) + 4 * n_rows * (args.select != "ALWAYS") + 4 * n_rows + 8 * nnz + 4 * min(n, nnz)


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


ref = None
lines = []
with torch.cuda.stream(be.stream):
    r = torch.empty(n_rows, device=dev)
    for cfg in args.cfg or [""]:
        opts = dict(DEFAULTS)
        for kv in filter(None, cfg.split(",")):
            k, val = kv.split("=")
            opts[k.strip()] = int(val)
        for k, val in opts.items():
            try:
                be.set_option(k, val)
            except Exception:
                if val != DEFAULTS.get(k):
                    raise
        M = be.csr(n_rows, n, Ap32, Aj, Ax)
        info = be.csr_info(M)
        m = None if args.select == "ALWAYS" else mask
        l0 = be.launch_count()
        be.mxv_masked(M, v, m, "MULT", "PLUS", args.select, 0.0, out=r)
        be.sync()
        launches = be.launch_count() - l0
        ms = timeit(lambda: be.mxv_masked(M, v, m, "MULT", "PLUS", args.select, 0.0, out=r))
        be.sync()
        if args.profile:
            be.profile(True)
            for _ in range(args.reps):
                be.mxv_masked(M, v, m, "MULT", "PLUS", args.select, 0.0, out=r)
            be.sync()
            dump = be.profile_dump()
            be.profile(False)
            for lab, row in sorted(dump.items()):
                print(f"    {lab:34s} {1000.0 * row['device_ms'] / max(1, row['calls']):9.2f} us  x {row['calls'] // args.reps}", flush=True)
        if ref is None:
            ref = r.clone()
            err = 0.0
        else:
            err = ((r - ref).abs() / ref.abs().clamp(min=1e-30)).max().item()
        line = {"cfg": cfg, "ms": round(ms, 4), "gteps": round(nnz / ms / 1e6, 1), "frac_of_6454.6": round(alg / ms / 1e6 / 6454.6, 4), "launches": launches,
                "max_rel_diff_vs_first": err, "phase_nnz": info["phase_nnz"], "row_class_nnz": info["row_class_nnz"]}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del M
if args.out:
    with open(args.out, "w") as f:
        for line in lines:
            f.write(json.dumps(line) + "\n")
