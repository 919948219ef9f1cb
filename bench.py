#!/usr/bin/env python
"""bench.py -- GTEPS of the masked semiring mxv hot path on synthetic RMAT graphs, 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scale S] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over the whole graph: one masked pull product
    r = mxv_masked(mask, A, v; FLOAT MULT / PLUS, select NQZERO on an all-ones mask, init 0)
i.e. the PageRank step of reference src/algorithm.cpp:312 with the mask really read (every row selected, E = nnz).
At N > 1 the rows are nnz-balanced across ranks (strong scaling on the fixed graph) and every step ends with the
all-gather of the result windows over NVLink, which is exactly the next step's input vector.

Rank 0 prints ONE JSON line (see the keys below). `value` is device-resident whole-job throughput; `e2e` is the same
metric through the C-ABI with HOST buffers (pinned h2d of v and mask, d2h of r inside the timed region); `roofline`
is the dominant kernel against the measured HBM peak; `cpu_baseline` is the reference's own CPU backend
(oracle/_ref, kind "reference") timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PEER_MODE = [False]
OVERLAP = [False]
METRIC = "GTEPS of masked mxv/vxm, RMAT-24, 1/2/4/8 B200; % of HBM roofline"
UNIT = "GTEPS"
OPS = ("MULT", "PLUS", "NQZERO")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=24, help="RMAT scale of the workload (BASELINE metric: 24)")
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--cpu-sample-nnz", type=int, default=64 << 20, help="entries of the bounded CPU-baseline sample (leading row block of the SAME graph)")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: CPU seconds the steps may take in total; the sample is the whole "
                                                                      "graph when it fits, else the leading row block that does")
    ap.add_argument("--no-vxm", action="store_true", help="skip the push (vxm) block of the line")
    ap.add_argument("--no-bfs", action="store_true", help="skip the BFS push-pull block (BASELINE config 4) of the line")
    ap.add_argument("--bfs-sources", type=int, default=8)
    ap.add_argument("--no-plugin", action="store_true", help="skip the leg that runs the workload through spla's public C++ API (libspla_cuda_x64.so)")
    ap.add_argument("--plugin-cpu-steps", type=int, default=1, help="steps of spla's CPU backend inside the plug-in leg (N = 1 only; 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time vxm / BFS-semiring variants (reported under 'extra')")
    return ap.parse_args()


def workload_config(args, n=None, nnz=None, world=1):
    cfg = {
        "workload": f"mxv_masked FLOAT {OPS[0]}/{OPS[1]}/{OPS[2]} all-ones mask (PageRank step, E = nnz) on RMAT scale-{args.scale} "
                    f"edge-factor {args.edge_factor}, symmetrised + dedup + no loops, A[i][j] = 0.85/outdeg(i), v = 1/N",
        "graph": f"rmat-{args.scale}",
        "parallelism": "single GPU" if world == 1 else f"rows nnz-balanced over {world} ranks, vector in the padded equal-window layout, " + ("windows published to the peers by one kernel of NVLink peer stores + device barrier per step" if PEER_MODE[0] else (("a step starts the exchange of its input (hub values first, then the windows as peer copies on the copy engines into symmetric memory) beside its own mask pass and hub class passes; one CUDA graph per direction of the ping-pong" if OVERLAP[0] == 2 else "hub values exchanged first (small all-to-all), the in-place ncclAllGather of the windows overlaps the hub class passes of the next step") if OVERLAP[0] else "one in-place ncclAllGather per step")),
        "cache": "inputs larger than L2 (CSR >> 126 MB), no flush between iterations",
    }
    if n is not None:
        cfg["n"] = n
        cfg["nnz"] = nnz
    return cfg


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


def mxv_alg_bytes(n_rows, n_cols, nnz_selected, reads_mask=True):
    """4(n_rows+1) [Ap] + 4 n_rows S [mask] + 4 n_rows [r] + 8 E [Aj+Ax] + 4 min(n_cols, E) [v gather]."""
    return 4 * (n_rows + 1) + (4 * n_rows if reads_mask else 0) + 4 * n_rows + 8 * nnz_selected + 4 * min(n_cols, nnz_selected)


# ---------------------------------------------------------------------------------------------------------
def make_graph(scale, edge_factor, device):
    import torch

    from spla_b200 import graphs

    n, Ap, Aj = graphs.rmat(scale, edge_factor=edge_factor, seed=2, device=device)
    Ax = graphs.pagerank_values(Ap, 0.85)
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()
    return n, Ap, Aj, Ax


CPU_GTEPS_GUESS = 0.15  # spla's CPU backend on one core, for sizing the bounded sample (measured 0.18-0.19 on the pool's hosts)


def cpu_reference_run(args, steps, warmup, graph=None, max_nnz=None):
    """The reference's own CPU implementation of the path (oracle/_ref, else the C port) on a bounded sample of the SAME workload:
    the leading row block [0, r) of the same RMAT scale-`--scale` matrix (all n columns kept, the full-length v), r chosen so that
    the block holds <= max_nnz entries -- the whole graph when it fits. Returns (gteps, seconds_per_step, info)."""
    import numpy as np
    import torch

    from oracle import oracle as orc

    if graph is None:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        graph = make_graph(args.scale, args.edge_factor, dev)
    n, Ap, Aj, Ax = graph
    nnz_all = int(Aj.numel())
    if max_nnz is None or max_nnz >= nnz_all:
        r = n
    else:
        r = int(torch.searchsorted(Ap.to(torch.int64).contiguous(), torch.tensor([max_nnz], device=Ap.device, dtype=torch.int64), right=True).item()) - 1
        r = max(1, min(n, r))
    nnz = int(Ap[r].item())
    Ap_h = Ap[:r + 1].cpu().numpy().astype(np.uint32)
    Aj_h = Aj[:nnz].cpu().numpy().astype(np.uint32)
    Ax_h = Ax[:nnz].cpu().numpy().astype(np.float32)
    v = np.full(n, 1.0 / n, dtype=np.float32)
    mask = np.ones(r, dtype=np.float32)
    whole = r == n
    sample = (f"{'the whole' if whole else 'leading row block [0, %d) of the' % r} RMAT scale-{args.scale} ef{args.edge_factor} matrix of the GPU arm "
              f"({r} x {n}, nnz={nnz} of {nnz_all}), same generator / seed / ops, {steps} timed calls after {warmup} warm-up")
    if orc.ref_available():
        ref = orc.RefSpla()
        rows = np.repeat(np.arange(r, dtype=np.uint32), np.diff(Ap_h.astype(np.int64)))
        t0 = time.perf_counter()
        M = ref.matrix(orc.FLOAT, r, n, rows, Aj_h, Ax_h)
        build_s = time.perf_counter() - t0
        del rows
        for _ in range(max(0, warmup - 1)):
            ref.mxv_masked(M, *OPS, v, mask, 0.0)
        _, sec = ref.mxv_masked(M, *OPS, v, mask, 0.0, repeats=steps)  # one untimed + `steps` timed calls inside the shim
        kind = "reference"
    else:
        o = orc.Oracle()
        build_s = 0.0
        for _ in range(warmup):
            o.mxv_masked(orc.FLOAT, *OPS, Ap_h, Aj_h, Ax_h, v, mask, 0.0)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.mxv_masked(orc.FLOAT, *OPS, Ap_h, Aj_h, Ax_h, v, mask, 0.0)
        sec = (time.perf_counter() - t0) / steps
        kind = "port"
    return nnz / sec / 1e9, sec, {"kind": kind, "cores": 1, "host_cores_total": os.cpu_count(), "sample": sample, "sample_rows": r, "sample_nnz": nnz,
                                  "whole_graph": whole, "matrix_build_s": round(build_s, 2)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    total = args.steps + max(1, args.warmup)
    max_nnz = int(args.ref_budget_s * CPU_GTEPS_GUESS * 1e9 / total)
    gteps, sec, info = cpu_reference_run(args, args.steps, max(1, args.warmup), max_nnz=max_nnz)
    cfg = workload_config(args, world=1)
    cfg["parallelism"] = "spla CPU backend, one host thread (the reference has no multi-threaded or multi-device path)"
    cfg["sample"] = info["sample"]
    cfg["same_config"] = bool(info["whole_graph"])
    line = {
        "impl": "reference", "metric": METRIC, "value": gteps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, args.warmup),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": dict(info, value=gteps, unit=UNIT),
        "e2e": {"value": gteps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "spla's CPU backend is single-threaded (reference src/cpu/cpu_mxv.hpp:53); each step is one exec_mxv_masked over the sample named in "
                "config.sample (GTEPS = its entries / its time; the per-entry cost of a row block equals the whole graph's)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    from spla_b200 import dist as sd
    from spla_b200.backend import Backend, scalar_bits, BIN, SEL, FLOAT

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    be = Backend(local_rank)
    dev = be.device

    # ---- workload: every rank generates the same graph (same seed, same device type) and keeps its row block ----
    n, Ap, Aj, Ax = make_graph(args.scale, args.edge_factor, dev)
    nnz = int(Aj.numel())
    bounds = sd.balanced_boundaries(Ap, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    Ap_l, Aj_l, Ax_l = sd.row_slice(Ap, Aj, Ax, r0, r1)
    nnz_l = int(Aj_l.numel())
    del Ax
    torch.cuda.empty_cache()  # Ap / Aj (the structure, 2 GB at scale 24) stay for the vxm / BFS blocks
    # N > 1: the vector lives in the padded layout of spla_b200.dist (equal windows, rank p at [p*W, p*W + rows_p)), so that a step
    # ends with one in-place all-gather; the column ids of the local slice are mapped once. N = 1: w0 = 0, n_vec = n.
    if world > 1:
        W, shifts = sd.padded_layout(bounds)
        Aj_l = sd.to_padded_index(Aj_l, bounds, shifts)
        n_vec, w0 = world * W, rank * W
    else:
        W, n_vec, w0 = n, n, 0
    # N > 1: a step ends with one in-place ncclAllGather of the equal windows. SPLA_B200_P2P=1 switches to peer-mapped copies of
    # the vector and one kernel of NVLink peer stores + a device barrier per step (splacu_publish_window); measured on 2 / 8 B200s it
    # is within noise of the NCCL call (560 vs 538, 1013 vs 1034 GTEPS: the step is bound by synchronisation, not bytes), so the
    # collective library stays the default.
    peer = None
    overlap = world > 1 and os.environ.get("SPLA_B200_OVERLAP", "1") == "1" and os.environ.get("SPLA_B200_P2P", "0") != "1"
    pvecs = None
    reserve = int(os.environ.get("SPLA_B200_RESERVE_SMS", "0"))
    if reserve:  # SMs the persistent class kernels leave to the kernels of a collective running beside them
        be.set_option("mxv_reserve_sms", reserve)
    if overlap and os.environ.get("SPLA_B200_OVERLAP_DMA", "1") == "1":
        try:  # peer-mapped vectors: the windows travel on the copy engines
            pvecs = [sd.PeerVector(be, n_vec), sd.PeerVector(be, n_vec)]
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print(f"bench: peer-mapped vectors unavailable ({ex}); the windows travel by ncclAllGather", file=sys.stderr)
            pvecs = None
        if world > 1:  # all ranks or none
            ok = torch.tensor([1 if pvecs else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if not bool(ok.item()):
                pvecs = None
    if world > 1 and os.environ.get("SPLA_B200_P2P", "0") == "1":
        try:
            pv = [sd.PeerVector(be, n_vec), sd.PeerVector(be, n_vec)]
            peer = {pv[0].tensor.data_ptr(): pv[0], pv[1].tensor.data_ptr(): pv[1]}
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print(f"bench: peer-mapped vectors unavailable ({ex}); using ncclAllGather", file=sys.stderr)
            peer = None
    PEER_MODE[0] = bool(peer)
    if peer:
        v, v_next = pv[0].tensor[:n_vec], pv[1].tensor[:n_vec]
        v.fill_(1.0 / n)
        v_next.fill_(1.0 / n)
    elif pvecs:
        v, v_next = pvecs[0].tensor[:n_vec], pvecs[1].tensor[:n_vec]
        v.fill_(1.0 / n)
        v_next.fill_(1.0 / n)
    else:
        v = torch.full((n_vec,), 1.0 / n, dtype=torch.float32, device=dev)
        v_next = torch.full((n_vec,), 1.0 / n, dtype=torch.float32, device=dev)
    mask_l = torch.ones(r1 - r0, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    M = be.csr(r1 - r0, n_vec, Ap_l, Aj_l, Ax_l)
    csr_info = be.csr_info(M)
    w1 = w0 + (r1 - r0)

    # N > 1 (default): the exchange of step k overlaps the hub class passes of step k + 1 (spla_b200.dist.PipelinedPull: the hub
    # values travel first in a small all-to-all, the all-gather of the windows runs beside part 1 of the next product).
    # SPLA_B200_OVERLAP=0: product, then all-gather, strictly one after the other.
    pp = None
    if overlap:
        pp = sd.PipelinedPull(be, M, W, w0, r1 - r0, OPS, 0.0, mask_l,
                              peers={pvecs[0].tensor.data_ptr(): pvecs[0], pvecs[1].tensor.data_ptr(): pvecs[1]} if pvecs else None)
        if not pp.enabled:
            pp = None
    OVERLAP[0] = (2 if pvecs else 1) if pp is not None else 0

    def finish():
        if pp:
            pp.finish()

    def step(src, dst):
        if pp:
            pp.step(src, dst)
            return
        be.mxv_masked(M, src, mask_l, *OPS, 0.0, out=dst[w0:w1])
        if peer:
            peer[dst.data_ptr()].publish(w0, r1 - r0)
        elif world > 1:
            sd.allgather_padded(dst, W)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    GRAPHS = False
    with torch.cuda.stream(be.stream):
        a, b = v, v_next
        if pp and pvecs:
            GRAPHS = pp.prepare(a, b)  # both directions of the ping-pong as CUDA graphs (one host call per step)
            a.fill_(1.0 / n)
            b.fill_(1.0 / n)
        for _ in range(max(3, args.warmup)):
            step(a, b)
            a, b = b, a
        finish()
        # ---- timed region: device-resident whole-job throughput ----
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = be.launch_count() + (pp.replay_launches if pp else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(be.stream)
        t_issue = time.perf_counter()
        for _ in range(args.steps):
            step(a, b)
            a, b = b, a
        finish()  # the last exchange is inside the timed region
        e1.record(be.stream)
        host_issue_ms = (time.perf_counter() - t_issue) * 1e3 / args.steps  # host time to ISSUE a step (no sync inside the loop)
        barrier()
        ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        launches = be.launch_count() + (pp.replay_launches if pp else 0) - l0  # replayed graphs: the launches counted at their capture
        clocks = sampler.stop() if rank == 0 else None

        # ---- dominant kernel alone (no collective): roofline ----
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(be.stream)
        for _ in range(args.steps):
            be.mxv_masked(M, a, mask_l, *OPS, 0.0, out=b[w0:w1])
        k1.record(be.stream)
        barrier()
        ms_kernel = k0.elapsed_time(k1) / args.steps
        alg_bytes = mxv_alg_bytes(r1 - r0, n, nnz_l, reads_mask=True)
        achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
        achieved_min = -max_over_ranks(-achieved)  # slowest rank's kernel bandwidth
        kernel_ms_ranks = [ms_kernel]
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device=dev)
            t[rank] = ms_kernel
            dist.all_reduce(t)
            kernel_ms_ranks = [round(float(x), 4) for x in t.tolist()]

        # ---- parity of the timed path at this N: one step from the canonical input v = 1/n; the sum of the result over all ranks
        #      against a float64 re-computation that does not depend on N or on the kernel: sum_i (A v)_i = (sum of all Ax) / n ----
        a.fill_(1.0 / n)
        b.zero_()  # the padding of the windows stays 0 on every rank
        step(a, b)
        finish()
        be.sync()
        sums = torch.stack([b[w0:w1].double().sum(), Ax_l.double().sum() / n,
                            (b[:w0].double().sum() + b[w1:].double().sum()) if world > 1 else torch.zeros((), dtype=torch.float64, device=dev)])
        if world > 1:
            part = sums[:2].clone()
            dist.all_reduce(part)
            # after the all-gather every rank holds every window: the foreign windows must add up to the other ranks' partial sums
            gathered_ok = abs(float(sums[0] + sums[2]) - float(part[0])) <= 1e-9 * abs(float(part[0]))
            sums = torch.stack([part[0], part[1], sums[2]])
        else:
            gathered_ok = True
        got_sum, want_sum = float(sums[0]), float(sums[1])
        parity_rel = abs(got_sum - want_sum) / abs(want_sum)
        parity_ok = parity_rel <= 1e-6 and gathered_ok
        if world > 1:
            flag = torch.tensor([1.0 if parity_ok else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity_ok = bool(flag.item() == 1.0)
        assert parity_ok, f"bench parity check failed at N={world}: sum(r) = {got_sum!r}, float64 expectation {want_sum!r} (rel {parity_rel:.3e}), gathered_ok={gathered_ok}"

        # ---- e2e: the C-ABI call with HOST buffers (pinned h2d of the rank's window of v and of its mask window, d2h of its result window;
        #      at N > 1 the uploaded windows are all-gathered over NVLink, so the job as a whole uploads v once, not N times) ----
        hv = torch.full((n_vec,), 1.0 / n, dtype=torch.float32).pin_memory()
        hm = torch.ones(r1 - r0, dtype=torch.float32).pin_memory()
        hr = torch.empty(r1 - r0, dtype=torch.float32).pin_memory()
        lib, sp = be.lib, be.stream_ptr
        e2e_steps = max(3, min(args.steps, 10))

        nw = r1 - r0  # this rank's window

        def upload_v(dst, stream_ptr):
            if world == 1:
                lib.splacu_memcpy_h2d(C.c_void_p(dst.data_ptr()), C.c_void_p(hv.data_ptr()), n_vec * 4, stream_ptr)
            else:
                lib.splacu_memcpy_h2d(C.c_void_p(dst[w0:w1].data_ptr()), C.c_void_p(hv[w0:w1].data_ptr()), nw * 4, stream_ptr)

        def e2e_step():
            upload_v(a, sp)
            if world > 1:
                sd.allgather_padded(a, W)
            lib.splacu_memcpy_h2d(C.c_void_p(mask_l.data_ptr()), C.c_void_p(hm.data_ptr()), (r1 - r0) * 4, sp)
            rc = lib.splacu_mxv_masked(M.handle, FLOAT, BIN[OPS[0]], BIN[OPS[1]], SEL[OPS[2]], C.c_void_p(a.data_ptr()), C.c_void_p(mask_l.data_ptr()),
                                       C.c_void_p(b[w0:w1].data_ptr()), scalar_bits(FLOAT, 0.0), 0, sp)
            assert rc == 0
            lib.splacu_memcpy_d2h(C.c_void_p(hr.data_ptr()), C.c_void_p(b[w0:w1].data_ptr()), (r1 - r0) * 4, sp)
            lib.splacu_sync(sp)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        t_serial = max_over_ranks((time.perf_counter() - t0) / e2e_steps)

        # The same calls, pipelined the way a host application streams batches through the C ABI: three streams (copy-in, the
        # backend's compute stream, copy-out) and two sets of device buffers, ordered by events. Every step still uploads its v and
        # mask from pinned host memory and downloads its r; PCIe is full duplex, so the upload of step k+1 overlaps the kernels and
        # the download of step k. All kernels stay on ONE stream: the matrix handle's scratch is not reentrant.
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        p_in, p_out = C.c_void_p(s_in.cuda_stream), C.c_void_p(s_out.cuda_stream)
        dv = [torch.empty(n_vec, dtype=torch.float32, device=dev) for _ in range(2)]
        dm = [torch.empty(r1 - r0, dtype=torch.float32, device=dev) for _ in range(2)]
        dr = [torch.empty(r1 - r0, dtype=torch.float32, device=dev) for _ in range(2)]
        hrs = [torch.empty(r1 - r0, dtype=torch.float32).pin_memory() for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_k = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        torch.cuda.synchronize()

        def e2e_pipelined(steps):
            for k in range(steps):
                q = k & 1
                if k >= 2:
                    s_in.wait_event(ev_k[q])    # the kernels of step k-2 have consumed dv[q] / dm[q]
                upload_v(dv[q], p_in)
                lib.splacu_memcpy_h2d(C.c_void_p(dm[q].data_ptr()), C.c_void_p(hm.data_ptr()), (r1 - r0) * 4, p_in)
                ev_in[q].record(s_in)
                be.stream.wait_event(ev_in[q])
                if world > 1:
                    sd.allgather_padded(dv[q], W)
                if k >= 2:
                    be.stream.wait_event(ev_out[q])  # dr[q] of step k-2 has been downloaded
                rc = lib.splacu_mxv_masked(M.handle, FLOAT, BIN[OPS[0]], BIN[OPS[1]], SEL[OPS[2]], C.c_void_p(dv[q].data_ptr()),
                                           C.c_void_p(dm[q].data_ptr()), C.c_void_p(dr[q].data_ptr()), scalar_bits(FLOAT, 0.0), 0, sp)
                assert rc == 0
                ev_k[q].record(be.stream)
                s_out.wait_event(ev_k[q])
                lib.splacu_memcpy_d2h(C.c_void_p(hrs[q].data_ptr()), C.c_void_p(dr[q].data_ptr()), (r1 - r0) * 4, p_out)
                ev_out[q].record(s_out)
            torch.cuda.synchronize()

        e2e_pipelined(2)
        barrier()
        pipe_steps = 2 * e2e_steps
        t0 = time.perf_counter()
        e2e_pipelined(pipe_steps)
        t_e2e = max_over_ranks((time.perf_counter() - t0) / pipe_steps)
        assert torch.equal(hrs[0], hr) and torch.equal(hrs[1], hr), "pipelined e2e result differs from the serial one"
        # whole-job bytes per step (all ranks): v once + every mask window up, every result window down
        h2d = (n + n) * 4 if world > 1 else (n_vec + (r1 - r0)) * 4
        d2h = n * 4
        checksum = float(hr.double().sum())

        extra = None
        if args.extra and world == 1:
            extra = extra_timings(be, M, n, nnz, args.steps)

    # ---- the other half of the metric: push (vxm) at two frontier sizes, N = 1; BFS push-pull (BASELINE config 4) at every N ----
    vxm = bfs = None
    if world == 1 and not (args.no_vxm and args.no_bfs):
        ones = torch.ones(nnz, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        Mi = be.csr(n, n, Ap.to(torch.int32), Aj, ones)
        if not args.no_vxm:
            vxm = vxm_block(be, Mi, n, Ap, max(3, args.steps // 2))
        if not args.no_bfs:
            bfs = bfs_block(be, args, n, Ap, Aj, ones, rank, world, single=Mi)
        del Mi, ones
    elif not args.no_bfs:
        ones = torch.ones(nnz, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        bfs = bfs_block(be, args, n, Ap, Aj, ones, rank, world)
        del ones

    # ---- through spla's public API with the plug-in (rank 0 drives it; at N > 1 the plug-in itself shards over the N GPUs while the
    #      other ranks wait at the barrier) ----
    plugin = None
    if not args.no_plugin:
        # the other ranks wait on the HOST (a file flag), not in a collective: an NCCL barrier kernel spinning on their GPUs would take
        # SMs away from the plug-in's persistent kernels
        flag = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"spla_b200_plugin_done_{os.environ.get('MASTER_PORT', '0')}")
        if rank == 0 and os.path.exists(flag):
            os.remove(flag)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        if rank == 0:
            plugin = plugin_block(args, n, Ap, Aj, world)
            if world > 1:
                open(flag, "w").close()
        elif world > 1:
            t_wait = time.time()
            while not os.path.exists(flag) and time.time() - t_wait < 1500:
                time.sleep(0.2)
        if world > 1:
            dist.barrier()
            if rank == 0 and os.path.exists(flag):
                os.remove(flag)

    # ---- CPU baseline: rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            g, sec, info = cpu_reference_run(args, args.cpu_steps, 2, graph=(n, Ap, Aj, Ax_l), max_nnz=args.cpu_sample_nnz)
            cpu = dict(info, value=g, unit=UNIT, ms_per_step=sec * 1e3)
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": f"failed: {ex}"}

    phase_nnz = csr_info.get("phase_nnz") or []
    if phase_nnz:
        shares = ", ".join(f"{x / max(1, nnz_l):.3f}" for x in phase_nnz)
        kernel_name = "mxv_seg_kernel"
        row_nnz = csr_info.get("row_class_nnz") or []
        row_desc = ""
        if row_nnz:
            row_desc = (f"; mxv_scat_kernel x {len(row_nnz)} row classes of the tail (entry shares " + ", ".join(f"{x / max(1, nnz_l):.3f}" for x in row_nnz) +
                        f"; {sum(csr_info.get('row_class_rows') or [])} rows accumulate in shared-memory tables while v streams) + merge kernel per class")
        kernel_desc = (f"mxv_seg_kernel x {len(phase_nnz)} column classes (entry shares {shares}; {csr_info['n_hub']} hub columns in shared-memory "
                       f"tables of 16-bit slots, tail class gathers v) + one two-launch fix-up of the rows that span tiles" + row_desc)
        roofline_note = (f"one step = one splacu_mxv_masked call = {len(phase_nnz)} launches of mxv_seg_kernel (one per column class), the two fix-up launches (chain sums, rows), "
                         f"{len(row_nnz)} launches of mxv_scat_kernel (row classes of the tail) with their merges, "
                         "the mask-count / fill pass (it also packs the hub values) and the gated CSR pass idling on a side stream; achieved = algorithmic bytes of the step / its device time, traffic = DRAM bytes "
                         "of all launches of the step")
    else:
        kernel_name = "mxv_wtile_kernel"
        kernel_desc = f"mxv_wtile_kernel, {csr_info['n_tiles']} warp tiles of 512 nnz, {csr_info['n_hub']} hub columns"
        roofline_note = "one launch = one step"
    if rank == 0:
        peak, peak_src = measured_peak()
        gteps = nnz / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": gteps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, n, nnz, world), kernel=kernel_desc),
            "clocks": clocks,
            "e2e": {"value": nnz / t_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3,
                    "steps": pipe_steps, "result_checksum": checksum, "ms_per_step_unpipelined": t_serial * 1e3,
                    "path": "per step: splacu_memcpy_h2d(v, mask) from pinned host memory -> splacu_mxv_masked -> splacu_memcpy_d2h(r); steps pipelined "
                            "over three streams and two device buffer sets (upload of step k+1 overlaps kernels and download of step k); "
                            "ms_per_step_unpipelined = the same calls back to back on one stream with a sync per step"},
            "gpu_launches": launches,
            "host_issue_ms_per_step": host_issue_ms,
            "cuda_graphs": bool(GRAPHS),
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved_min, "peak": peak, "unit": "GB/s", "frac": achieved_min / peak,
                         "traffic": ncu_traffic(kernel_name), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": ms_kernel, "kernel_ms_per_rank": kernel_ms_ranks, "note": ("rank-0 bytes, slowest rank's bandwidth; " if world > 1 else "") + roofline_note},
            "cpu_baseline": cpu,
            "parity_checked": True,
            "parity": {"sum_r_all_ranks": got_sum, "float64_expectation": want_sum, "rel_diff": parity_rel, "bar": 1e-6,
                       "what": "one step from v = 1/n at this N: sum over all ranks of r against (sum of all Ax)/n in float64; at N > 1 also: the "
                               "all-gathered foreign windows on every rank add up to the other ranks' partial sums; asserted before this line is printed. "
                               "Per-element parity at this size: tests/test_gpu_baseline_configs.py (vs the reference CPU backend)"},
        }
        if plugin:
            line["plugin"] = plugin
        if vxm:
            line["vxm"] = vxm
        if bfs:
            line["bfs"] = bfs
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def plugin_block(args, n, Ap, Aj, devices):
    """The same workload through spla's OWN public C++ API with the CUDA backend of this repository plugged in (tools/spla_bench.cpp ->
    spla_b200/lib/spla_bench, linked to libspla_cuda_x64.so): Matrix::build, the first call with the storage manager's format
    conversions and the handle build, the steady-state exec_mxv_masked step, spla::pr -- and, at N = 1, spla's CPU backend on the same
    Matrix object in the same process. devices > 1: the plug-in shards the products over that many GPUs (SPLA_CUDA_DEVICES)."""
    import numpy as np
    import torch

    from spla_b200 import graphs

    exe = os.path.join(ROOT, "spla_b200", "lib", "spla_bench")
    if not os.path.exists(exe):
        return {"unavailable": "spla_b200/lib/spla_bench is not built (needs the reference checkout at build time)"}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    prefix = os.path.join(tmp, f"spla_b200_bench_{os.getpid()}")
    nnz = int(Aj.numel())
    try:
        deg = (Ap[1:] - Ap[:-1])
        torch.repeat_interleave(torch.arange(n, device=Ap.device, dtype=torch.int32), deg).cpu().numpy().astype(np.uint32).tofile(prefix + "_Ai.bin")
        Aj.cpu().numpy().astype(np.uint32).tofile(prefix + "_Aj.bin")
        graphs.pagerank_values(Ap, 0.85).cpu().numpy().astype(np.float32).tofile(prefix + "_Ax.bin")
        torch.cuda.empty_cache()
        env = dict(os.environ)
        if devices > 1:
            env["SPLA_CUDA_DEVICES"] = str(devices)
        cpu_steps = args.plugin_cpu_steps if devices == 1 else 0
        t0 = time.perf_counter()
        p = subprocess.run([exe, prefix, str(n), str(nnz), str(max(3, args.steps)), str(cpu_steps), "1"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=900)
        wall = time.perf_counter() - t0
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not line:
            return {"unavailable": f"spla_bench failed (rc {p.returncode}): {(p.stdout + p.stderr)[-400:]}"}
        out = json.loads(line[-1])
        out["devices"] = devices
        out["wall_s"] = round(wall, 1)
        out["path"] = ("spla public C++ API (Matrix::build, exec_mxv_masked, exec_v_count_mf, spla::pr) -> Dispatcher -> *_cuda algorithms of "
                       "spla_b200/src/cuda -> C ABI; vectors stay device-resident between calls, as spla's storage manager keeps them")
        return out
    except Exception as ex:  # noqa: BLE001  (a reported leg, never a reason to lose the bench line)
        return {"unavailable": f"plug-in leg failed: {ex}"}
    finally:
        for suf in ("_Ai.bin", "_Aj.bin", "_Ax.bin"):
            try:
                os.remove(prefix + suf)
            except OSError:
                pass


def vxm_alg_bytes(nv, e_f, n_cols, nr, reads_mask=True, reads_ax=True):
    """8 nv [vi, vx] + 8 nv [Ap[i], Ap[i+1]] + 8 E_f [Aj + Ax; 4 E_f when Ax is provably not read] + 4 min(n_cols, E_f) S [mask] + 8 nr [ri, rx]."""
    return 16 * nv + (8 if reads_ax else 4) * e_f + (4 * min(n_cols, e_f) if reads_mask else 0) + 8 * nr


def vxm_block(be, Mi, n, Ap, reps):
    """exec_vxm_masked INT BAND/BOR/EQZERO (the BFS push step, reference tests/test_vxm.cpp:140-187 shape) over random frontiers of 5 % and
    1 % of the vertices of the bench graph, nothing visited yet (every reached column passes the mask and is a first touch)."""
    import torch

    dev = be.device
    peak, _ = measured_peak()
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    deg = (Ap[1:] - Ap[:-1]).to(torch.int64)
    out = {"ops": "INT BAND/BOR/EQZERO, frontier values 1, matrix values 1, mask all zero", "timing": "CUDA events on the backend's stream around "
           f"{reps} whole splacu_vxm_masked calls (offsets, expand, count with its 4-byte host read, ordered emit)"}
    ri = torch.empty(n, dtype=torch.int32, device=dev)
    rx = torch.empty(n, dtype=torch.int32, device=dev)
    visited = torch.zeros(n, dtype=torch.int32, device=dev)
    for frac in (0.05, 0.01):
        vi = torch.nonzero(torch.rand(n, generator=g, device=dev) < frac).flatten().to(torch.int32)
        vx = torch.ones(vi.numel(), dtype=torch.int32, device=dev)
        ef = int(deg[vi.long()].sum().item())
        torch.cuda.synchronize()
        with torch.cuda.stream(be.stream):
            oi, _ = be.vxm_masked(Mi, vi, vx, visited, "BAND", "BOR", "EQZERO", out=(ri, rx))
            nr = int(oi.numel())
            be.sync()
            l0 = be.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(be.stream)
            for _ in range(reps):
                be.vxm_masked(Mi, vi, vx, visited, "BAND", "BOR", "EQZERO", out=(ri, rx))
            e1.record(be.stream)
            be.sync()
        ms = e0.elapsed_time(e1) / reps
        info = be.vxm_info() if hasattr(be, "vxm_info") else {}
        reads_ax = not info.get("struct_only", False)
        bts = vxm_alg_bytes(int(vi.numel()), ef, n, nr, reads_mask=True, reads_ax=reads_ax)
        ach = bts / (ms * 1e-3) / 1e9
        out[f"front_{frac}"] = {"nv": int(vi.numel()), "edges": ef, "nr": nr, "ms": ms, "gteps": ef / ms / 1e6, "launches_per_call": (be.launch_count() - l0) / reps,
                                "algorithmic_bytes": bts, "ax_read": reads_ax, "achieved_gbs": ach, "frac_of_measured_hbm_peak": ach / peak,
                                "traffic": ncu_traffic(f"vxm_expand_kernel@{frac}")}
    return out


def bfs_block(be, args, n, Ap, Aj, ones, rank, world, single=None):
    """BASELINE config 4: BFS push-pull (front_factor 0.05; INT BAND/BOR/EQZERO push, early-exit pull) on the bench graph from random
    sources of non-zero degree; Graph500-style TEPS = entries of the reached rows / time, geometric mean over the sources. N = 1: the
    reference's own call sequence on one handle (spla_b200.algorithms.bfs); N > 1: vertices owned in nnz-balanced windows, row slice for
    the pull, column slice for the push, NCCL frontier exchange per level (algorithms.bfs_dist)."""
    import torch
    import torch.distributed as dist

    from spla_b200 import algorithms

    dev = be.device
    deg = Ap[1:] - Ap[:-1]
    gcpu = torch.Generator(device="cpu")
    gcpu.manual_seed(11)
    cand = torch.nonzero(deg > 0).flatten().cpu()
    srcs = cand[torch.randperm(cand.numel(), generator=gcpu)[:max(1, args.bfs_sources)]].tolist()
    if world > 1:
        shard = algorithms.make_bfs_shard(be, n, Ap, Aj, ones, rank, world)
        w0, w1 = shard["w0"], shard["w1"]
        run = lambda s, trace=None: algorithms.bfs_dist(be, shard, s, mode="push_pull", front_factor=0.05, trace=trace)  # noqa: E731
    else:
        w0, w1 = 0, n
        run = lambda s, trace=None: algorithms.bfs(be, single, s, mode="push_pull", front_factor=0.05, trace=trace)  # noqa: E731
    deg_loc = deg[w0:w1]

    def barrier():
        be.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    run(srcs[0])  # warm-up (allocator, NCCL channels)
    times, teps, trace0, reached0 = [], [], None, None
    for k, src in enumerate(srcs):
        trace = []
        barrier()
        t0 = time.perf_counter()
        d = run(src, trace)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        e = deg_loc[d > 0].sum().to(torch.int64).reshape(1)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(e)
        times.append(float(dt.item()))
        teps.append(float(e.item()) / float(dt.item()))
        if k == 0:
            trace0, reached0 = trace, int(e.item())
    gm = float(torch.tensor(teps, dtype=torch.float64).log().mean().exp().item())
    return {"config": "BASELINE config 4: BFS push-pull, front_factor 0.05, INT BAND/BOR/EQZERO, on the bench graph", "sources": len(srcs),
            "gteps_geomean": gm / 1e9, "ms_mean": 1e3 * sum(times) / len(times), "ms_min": 1e3 * min(times), "levels_first_source": trace0,
            "edges_reached_first_source": reached0,
            "exchange": "none (single GPU)" if world == 1 else "per level: all-gather of the sparse frontier pieces (push) or of the dense windows (pull) + front sizes, NCCL",
            "timing": "host wall clock between synchronize (+ barrier) pairs, max over ranks: a BFS is host-driven, one 4-byte read per level"}


def extra_timings(be, M, n, nnz, steps):
    """Secondary numbers (not the headline): BFS-semiring pull with a real mask, and push vxm at several frontier sizes."""
    import torch

    out = {}
    dev = be.device
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    ones = torch.ones(M.nnz, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    Mi = be.csr(M.n_rows, M.n_cols, M.Ap, M.Aj, ones)
    deg = (M.Ap[1:] - M.Ap[:-1]).to(torch.int64)

    def timeit(fn, reps):
        fn()
        be.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(be.stream)
        for _ in range(reps):
            fn()
        e1.record(be.stream)
        be.sync()
        return e0.elapsed_time(e1) / reps

    for density in (0.5, 0.1, 0.01):
        visited = (torch.rand(n, generator=g, device=dev) >= density).to(torch.int32)  # mask selects `density` of the rows
        front = (torch.rand(n, generator=g, device=dev) < 0.3).to(torch.int32)
        r = torch.empty(n, dtype=torch.int32, device=dev)
        sel_edges = int(deg[visited == 0].sum().item())
        torch.cuda.synchronize()
        for ee in (False, True):
            ms = timeit(lambda: be.mxv_masked(Mi, front, visited, "BAND", "BOR", "EQZERO", 0, early_exit=ee, out=r), steps)
            out[f"mxv_int_band_bor_select{density}_early{int(ee)}"] = {"ms": ms, "gteps_selected_edges": sel_edges / ms / 1e6}
    for frac in (1e-4, 1e-3, 1e-2, 5e-2):
        vi = torch.nonzero(torch.rand(n, generator=g, device=dev) < frac).flatten().to(torch.int32)
        vx = torch.ones(vi.numel(), dtype=torch.int32, device=dev)
        visited = torch.zeros(n, dtype=torch.int32, device=dev)
        ef = int(deg[vi.long()].sum().item())
        ri = torch.empty(n, dtype=torch.int32, device=dev)
        rx = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ms = timeit(lambda: be.vxm_masked(Mi, vi, vx, visited, "BAND", "BOR", "EQZERO", out=(ri, rx)), max(3, steps // 2))
        out[f"vxm_int_band_bor_front{frac}"] = {"ms": ms, "nv": int(vi.numel()), "edges": ef, "gteps": ef / ms / 1e6}
    return out


if __name__ == "__main__":
    main()
