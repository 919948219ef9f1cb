"""tools/bench_bfs_dist.py -- BASELINE config 4: BFS push-pull (INT BAND/BOR/EQZERO, early-exit pull) on RMAT scale-S at N GPUs,
one process per GPU, NCCL frontier exchange (spla_b200.algorithms.bfs_dist). Launch:

    python tools/bench_bfs_dist.py --scale 24                                    (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_bfs_dist.py --scale 24 --out gpurun_out/bfs_dist_N.json

Every rank generates the same graph on its GPU and keeps its row / column slices. Per source: barrier + synchronize, run, barrier +
synchronize; time = max over ranks; TEPS = entries of the rows reached / time (Graph500 style); the line reports the geometric
mean over the sources and the per-level trace of the first one. `--check` first runs scale 16 and compares the sharded depths
with the single-GPU bfs() of the same backend, bit for bit."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spla_b200 import algorithms, graphs  # noqa: E402
from spla_b200 import dist as sd  # noqa: E402
from spla_b200.backend import Backend  # noqa: E402


def pick_sources(deg, k, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    cand = torch.nonzero(deg > 0).flatten().cpu()
    return cand[torch.randperm(cand.numel(), generator=g)[:k]].tolist()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--sources", type=int, default=16)
    ap.add_argument("--front-factor", type=float, default=0.05)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--bitmap", action="store_true", help="pull levels exchange the frontier as an n-bit bitmap (opt-in)")
    ap.add_argument("--ab", action="store_true", help="time the pull levels with the bitmap exchange and with the dense int32 exchange")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    be = Backend(local)
    dev = be.device

    def build(scale):
        n, Ap, Aj = graphs.rmat(scale, 16, seed=1, device=dev)
        ones = torch.ones(Aj.numel(), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        return n, Ap, Aj, ones

    def barrier():
        be.sync()
        torch.cuda.synchronize()
        dist.barrier()

    result = {"n_gpus": world}
    if args.check:
        n, Ap, Aj, ones = build(16)
        shard = algorithms.make_bfs_shard(be, n, Ap, Aj, ones, rank, world, bitmap_exchange=args.bitmap)
        M = be.csr(n, n, Ap.to(torch.int32), Aj, ones)
        deg = Ap[1:] - Ap[:-1]
        ok = True
        for src in pick_sources(deg, 8, 5):
            want = algorithms.bfs(be, M, src, mode="push_pull", front_factor=args.front_factor)
            for mode in ("push_pull", "push", "pull"):
                mine = algorithms.bfs_dist(be, shard, src, mode=mode, front_factor=args.front_factor)
                ok = ok and bool(torch.equal(mine, want[shard["w0"]:shard["w1"]]))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        result["check_scale16_bit_exact_vs_single_gpu"] = bool(flag.item())
        del shard, M, Ap, Aj, ones
        torch.cuda.empty_cache()

    n, Ap, Aj, ones = build(args.scale)
    nnz = int(Aj.numel())
    deg = Ap[1:] - Ap[:-1]
    t0 = time.perf_counter()
    shard = algorithms.make_bfs_shard(be, n, Ap, Aj, ones, rank, world, bitmap_exchange=args.bitmap or args.ab)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    srcs = pick_sources(deg, args.sources, 11)
    w0, w1 = shard["w0"], shard["w1"]
    deg_loc = deg[w0:w1].clone()
    local_nnz = int(deg_loc.sum().item())
    del Ap, Aj, ones, deg
    torch.cuda.empty_cache()
    algorithms.bfs_dist(be, shard, srcs[0], front_factor=args.front_factor)  # warm-up (NCCL channels, allocator)
    if args.ab:
        unit = shard["unit_values"]
        ab = {}
        for name, flag in (("bitmap", unit), ("dense_int32", False), ("bitmap_again", unit)):
            shard["unit_values"] = flag
            algorithms.bfs_dist(be, shard, srcs[0], front_factor=args.front_factor)
            ts = []
            for s in srcs:
                barrier()
                t0 = time.perf_counter()
                algorithms.bfs_dist(be, shard, s, mode="push_pull", front_factor=args.front_factor)
                barrier()
                dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                ts.append(float(dt.item()) * 1e3)
            ts.sort()
            ab[name] = {"ms_mean": sum(ts) / len(ts), "ms_median": ts[len(ts) // 2], "ms_min": ts[0]}
        shard["unit_values"] = args.bitmap
        result["pull_exchange_ab"] = ab
    times, teps, traces = [], [], []
    for s in srcs:
        trace = []
        barrier()
        t0 = time.perf_counter()
        d = algorithms.bfs_dist(be, shard, s, mode="push_pull", front_factor=args.front_factor, trace=trace)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e = deg_loc[d > 0].sum().to(torch.int64).reshape(1)
        dist.all_reduce(e)
        times.append(float(dt.item()))
        teps.append(float(e.item()) / float(dt.item()))
        traces.append(trace)
    gm = float(torch.tensor(teps, dtype=torch.float64).log().mean().exp().item())
    nnz_all = torch.tensor([local_nnz], device=dev, dtype=torch.int64)
    gathered = [torch.zeros_like(nnz_all) for _ in range(world)]
    dist.all_gather(gathered, nnz_all)
    result.update({"graph": f"rmat-{args.scale}", "n": n, "nnz": nnz, "front_factor": args.front_factor, "sources": len(srcs),
                   "ms_mean": 1e3 * sum(times) / len(times), "ms_min": 1e3 * min(times), "gteps_geomean": gm / 1e9,
                   "levels_first_source": traces[0], "nnz_per_rank": [int(g.item()) for g in gathered], "shard_build_s": build_s,
                   "timing": "host wall clock between barrier + synchronize pairs, max over ranks (a BFS is host-driven: one 4-byte read per level)"})
    if rank == 0:
        line = json.dumps(result)
        print(line, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
