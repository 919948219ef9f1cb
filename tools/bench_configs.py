"""tools/bench_configs.py -- the BASELINE.json configs other than the headline one, on ONE B200, through the C ABI.

  config 1  BFS push-pull (INT BAND/BOR/EQZERO, early exit) on RMAT scale-16, whole run
  config 2  PageRank step (FLOAT MULT/PLUS/ALWAYS) on RMAT scale-22 + the full pr() loop to eps 1e-6
  config 3  SSSP (FLOAT PLUS/MIN/ALWAYS) on the 4096 x 4096 grid, unit and uniform [1, 2) weights, whole run
  config 5  mxv sweep: uniform-random vs RMAT, scales 20..26, FLOAT MULT/PLUS/ALWAYS and INT BAND/BOR/EQZERO at mask densities

Writes one JSON object (gpurun_out/configs.json by default). Each mxv line carries the algorithmic bytes of SURVEY 8(d) and the
fraction of the measured HBM peak; every float product is checked by a size-independent property (sum(r) against a float64
sum of Ax[k] * v[Aj[k]], the linearity checksum), every BFS / SSSP result by a labelling property. Not the bench contract:
bench.py is. Usage:  python tools/bench_configs.py [--max-scale 26] [--out gpurun_out/configs.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import algorithms, graphs  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return float(j.get("hbm_gbs") or j["hbm"]["gbs"])
    except Exception:
        return 6454.6


def mxv_bytes(n_rows, n_cols, e_sel, reads_mask):
    return 4 * (n_rows + 1) + (4 * n_rows if reads_mask else 0) + 4 * n_rows + 8 * e_sel + 4 * min(n_cols, e_sel)


def timeit(be, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    be.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(be.stream)
    for _ in range(reps):
        fn()
    e1.record(be.stream)
    be.sync()
    return e0.elapsed_time(e1) / reps


def checksum64(Aj, Ax, v, chunk=1 << 27):
    tot = 0.0
    for lo in range(0, Aj.numel(), chunk):
        hi = min(Aj.numel(), lo + chunk)
        tot += float((Ax[lo:hi].double() * v[Aj[lo:hi].long()].double()).sum().item())
    return tot


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


def config1(be, out):
    dev = be.device
    n, Ap, Aj = graphs.rmat(16, 16, seed=1, device=dev)
    deg = Ap[1:] - Ap[:-1]
    M = be.csr(n, n, Ap.to(torch.int32), Aj, torch.ones(Aj.numel(), dtype=torch.int32, device=dev))
    g = torch.Generator(device="cpu")
    g.manual_seed(11)
    cand = torch.nonzero(deg > 0).flatten().cpu()
    srcs = cand[torch.randperm(cand.numel(), generator=g)[:16]].tolist()
    res = {}
    for mode in ("push_pull", "push", "pull"):
        teps, ms = [], []
        for s in srcs:
            algorithms.bfs(be, M, s, mode=mode, front_factor=0.05)
            v, dt = wall(lambda: algorithms.bfs(be, M, s, mode=mode, front_factor=0.05))
            edges = int(deg[v > 0].sum().item())
            teps.append(edges / dt)
            ms.append(dt * 1e3)
        gm = float(torch.tensor(teps).log().mean().exp().item())
        res[mode] = {"sources": len(srcs), "ms_mean": sum(ms) / len(ms), "gteps_geomean": gm / 1e9}
    out["config1_bfs_rmat16"] = {"n": n, "nnz": int(Aj.numel()), "front_factor": 0.05, "modes": res,
                                 "bound": "launch / host-sync latency (7-9 levels of a handful of launches and one 4-byte read each)"}


def config2(be, out, peak):
    dev = be.device
    n, Ap, Aj = graphs.rmat(22, 16, seed=2, device=dev)
    Ax = graphs.pagerank_values(Ap, 0.85)
    nnz = int(Aj.numel())
    M = be.csr(n, n, Ap.to(torch.int32), Aj, Ax)
    v = torch.full((n,), 1.0 / n, dtype=torch.float32, device=dev)
    r = torch.empty_like(v)
    ms = timeit(be, lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r), 50, warm=5)
    b = mxv_bytes(n, n, nnz, False)
    ref = checksum64(Aj, Ax, v)
    got = float(r.double().sum().item())
    (p, it), dt = wall(lambda: algorithms.pagerank(be, M, 0.85, 1e-6))
    (p, it), dt = wall(lambda: algorithms.pagerank(be, M, 0.85, 1e-6))
    out["config2_pagerank_rmat22"] = {
        "n": n, "nnz": nnz, "step_ms": ms, "gteps": nnz / ms / 1e6, "alg_bytes": b, "gbs": b / ms / 1e6, "frac_of_measured_hbm": b / ms / 1e6 / peak,
        "checksum_rel_err": abs(got - ref) / abs(ref), "pr_iterations": it, "pr_total_ms": dt * 1e3, "pr_ms_per_iteration": dt * 1e3 / max(it, 1),
        "pr_rank_sum": float(p.double().sum().item())}


def config3(be, out):
    dev = be.device
    side = 4096
    n, Ap, Aj = graphs.grid2d(side, device=dev)
    nnz = int(Aj.numel())
    res = {}
    for name in ("unit", "uniform_1_2"):
        Ax = torch.ones(nnz, dtype=torch.float32, device=dev)
        if name != "unit":  # symmetric weights: w(i, j) = w(j, i), as a road graph has
            rows = torch.repeat_interleave(torch.arange(n, device=dev), (Ap[1:] - Ap[:-1]))
            lo, hi = torch.minimum(rows, Aj.long()), torch.maximum(rows, Aj.long())
            h = ((lo * 2654435761 + hi * 40503) % 1000003).float() / 1000003.0
            Ax = (1.0 + h).float()
            del rows, lo, hi, h
        M = be.csr(n, n, Ap.to(torch.int32), Aj, Ax)
        calls = [0]
        orig = be.vxm_masked

        def counted(*a, **k):
            calls[0] += 1
            return orig(*a, **k)

        be.vxm_masked = counted
        d, dt = wall(lambda: algorithms.sssp(be, M, 0, mode="push_pull", front_factor=0.05))
        be.vxm_masked = orig
        ok = None
        if name == "unit":
            idx = torch.arange(n, device=dev)
            ok = bool(torch.equal(d, ((idx % side) + idx // side).float()))
        else:  # fixed point of the relaxation: no edge can still improve a distance
            rows = torch.repeat_interleave(torch.arange(n, device=dev), (Ap[1:] - Ap[:-1]))
            ok = bool((d[Aj.long()] <= d[rows] + Ax).all().item()) and float(d[0].item()) == 0.0
            del rows
        res[name] = {"iterations_push": calls[0], "total_ms": dt * 1e3, "us_per_iteration": dt * 1e6 / max(calls[0], 1),
                     "edges_relaxed_lower_bound": nnz, "mteps_nnz_over_time": nnz / dt / 1e6, "property_ok": ok,
                     "max_distance": float(d.max().item())}
        del M
    out["config3_sssp_grid4096"] = {"n": n, "nnz": nnz, "weights": res,
                                    "bound": "launch / host-sync latency: thousands of iterations over fronts of a few thousand vertices"}


def sweep_one(be, peak, kind, scale):
    dev = be.device
    if kind == "rmat":
        n, Ap, Aj = graphs.rmat(scale, 16, seed=2, device=dev)
    else:
        n, Ap, Aj = graphs.uniform_random(scale, 16, seed=4, device=dev)
    nnz = int(Aj.numel())
    if nnz >= 2 ** 31:
        raise RuntimeError(f"nnz = {nnz} >= 2^31: not attempted (uint32 Ap would hold it, the harness passes int32 tensors)")
    Ap32 = Ap.to(torch.int32)
    deg = Ap[1:] - Ap[:-1]
    del Ap
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    Ax = torch.rand(nnz, generator=g, device=dev)
    v = torch.rand(n, generator=g, device=dev)
    r = torch.empty_like(v)
    M = be.csr(n, n, Ap32, Aj, Ax)
    reps = 20 if scale <= 24 else 8
    ms = timeit(be, lambda: be.mxv_masked(M, v, None, "MULT", "PLUS", "ALWAYS", 0.0, out=r), reps)
    b = mxv_bytes(n, n, nnz, False)
    ref = checksum64(Aj, Ax, v)
    got = float(r.double().sum().item())
    row = {"graph": kind, "scale": scale, "n": n, "nnz": nnz, "max_degree": int(deg.max().item()), "float_mult_plus_always": {
        "ms": ms, "gteps": nnz / ms / 1e6, "gbs": b / ms / 1e6, "frac_of_measured_hbm": b / ms / 1e6 / peak,
        "checksum_rel_err": abs(got - ref) / abs(ref)}}
    del M, Ax, r, v
    torch.cuda.empty_cache()
    ones = torch.ones(nnz, dtype=torch.int32, device=dev)
    Mi = be.csr(n, n, Ap32, Aj, ones)
    front = (torch.rand(n, generator=g, device=dev) < 0.5).to(torch.int32)
    ri = torch.empty(n, dtype=torch.int32, device=dev)
    for density in (1.0, 0.5, 0.1, 0.01):
        visited = (torch.rand(n, generator=g, device=dev) >= density).to(torch.int32)  # EQZERO selects `density` of the rows
        e_sel = int(deg[visited == 0].sum().item())
        ms = timeit(be, lambda: be.mxv_masked(Mi, front, visited, "BAND", "BOR", "EQZERO", 0, out=ri), reps)
        b = mxv_bytes(n, n, e_sel, True)
        row[f"int_band_bor_eqzero_density{density}"] = {
            "ms": ms, "selected_edges": e_sel, "gteps_selected": e_sel / ms / 1e6, "gbs": b / ms / 1e6,
            "frac_of_measured_hbm": b / ms / 1e6 / peak,
            "unselected_rows_zero": bool((ri[visited != 0] == 0).all().item())}
    return row


def config5(be, out, peak, max_scale, min_scale=20):
    rows = out.setdefault("config5_mxv_sweep", [])
    scales = [s for s in (20, 22, 24, 25, 26) if min_scale <= s <= max_scale]
    for kind in ("rmat", "uniform"):
        for scale in scales:
            if kind == "uniform" and scale > 25:
                continue
            try:
                row = sweep_one(be, peak, kind, scale)
            except Exception as e:  # an out-of-memory at the largest scale is a result, not a crash
                row = {"graph": kind, "scale": scale, "error": str(e)[:300]}
            rows.append(row)
            print(json.dumps(row), flush=True)
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-scale", type=int, default=25)
    ap.add_argument("--min-scale", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--only", default="1,2,3,5")
    args = ap.parse_args()
    be = Backend(0)
    peak = hbm_peak()
    out = {"device": be.device_name(), "hbm_peak_gbs_measured": peak}
    only = set(args.only.split(","))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    for key, fn in (("1", lambda: config1(be, out)), ("2", lambda: config2(be, out, peak)), ("3", lambda: config3(be, out)),
                    ("5", lambda: config5(be, out, peak, args.max_scale, args.min_scale))):
        if key not in only:
            continue
        t0 = time.perf_counter()
        try:
            fn()
        except Exception as e:
            out[f"config{key}_error"] = str(e)[:500]
        print(f"config {key} done in {time.perf_counter() - t0:.1f} s", flush=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out)[:4000])


if __name__ == "__main__":
    main()
