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
    ap.add_argument("--cpu-scale", type=int, default=20, help="RMAT scale of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time vxm / BFS-semiring variants (reported under 'extra')")
    return ap.parse_args()


def workload_config(args, n=None, nnz=None, world=1):
    cfg = {
        "workload": f"mxv_masked FLOAT {OPS[0]}/{OPS[1]}/{OPS[2]} all-ones mask (PageRank step, E = nnz) on RMAT scale-{args.scale} "
                    f"edge-factor {args.edge_factor}, symmetrised + dedup + no loops, A[i][j] = 0.85/outdeg(i), v = 1/N",
        "graph": f"rmat-{args.scale}",
        "parallelism": "single GPU" if world == 1 else f"rows nnz-balanced over {world} ranks, vector in the padded equal-window layout, " + ("windows published to the peers by one kernel of NVLink peer stores + device barrier per step" if PEER_MODE[0] else "one in-place ncclAllGather per step"),
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


def cpu_reference_run(args, steps, warmup):
    """The reference's own CPU implementation of the path (oracle/_ref, else the C port) on a bounded sample:
    RMAT scale `--cpu-scale` from the same generator, same ops. Returns (gteps, seconds_per_step, info)."""
    import numpy as np
    import torch

    from oracle import oracle as orc

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n, Ap, Aj, Ax = make_graph(args.cpu_scale, args.edge_factor, dev)
    Ap_h = Ap.cpu().numpy().astype(np.uint32)
    Aj_h = Aj.cpu().numpy().astype(np.uint32)
    Ax_h = Ax.cpu().numpy().astype(np.float32)
    nnz = len(Aj_h)
    v = np.full(n, 1.0 / n, dtype=np.float32)
    mask = np.ones(n, dtype=np.float32)
    sample = f"RMAT scale-{args.cpu_scale} ef{args.edge_factor} (n={n}, nnz={nnz}), same generator/ops, {steps} timed calls after {warmup} warm-up"
    if orc.ref_available():
        ref = orc.RefSpla()
        rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(Ap_h.astype(np.int64)))
        M = ref.matrix(orc.FLOAT, n, n, rows, Aj_h, Ax_h)
        del rows
        for _ in range(max(0, warmup - 1)):
            ref.mxv_masked(M, *OPS, v, mask, 0.0)
        _, sec = ref.mxv_masked(M, *OPS, v, mask, 0.0, repeats=steps)  # one untimed + `steps` timed calls inside the shim
        kind = "reference"
    else:
        o = orc.Oracle()
        for _ in range(warmup):
            o.mxv_masked(orc.FLOAT, *OPS, Ap_h, Aj_h, Ax_h, v, mask, 0.0)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.mxv_masked(orc.FLOAT, *OPS, Ap_h, Aj_h, Ax_h, v, mask, 0.0)
        sec = (time.perf_counter() - t0) / steps
        kind = "port"
    return nnz / sec / 1e9, sec, {"kind": kind, "cores": 1, "host_cores_total": os.cpu_count(), "sample": sample}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    gteps, sec, info = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gteps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world=args.gpus),
        "cpu_baseline": dict(info, value=gteps, unit=UNIT),
        "e2e": {"value": gteps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "spla's CPU backend is single-threaded (reference src/cpu/cpu_mxv.hpp:53); each step is one exec_mxv_masked on the bounded sample",
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
    del Ap, Aj, Ax
    torch.cuda.empty_cache()
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
    else:
        v = torch.full((n_vec,), 1.0 / n, dtype=torch.float32, device=dev)
        v_next = torch.full((n_vec,), 1.0 / n, dtype=torch.float32, device=dev)
    mask_l = torch.ones(r1 - r0, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    M = be.csr(r1 - r0, n_vec, Ap_l, Aj_l, Ax_l)
    csr_info = be.csr_info(M)
    w1 = w0 + (r1 - r0)

    def step(src, dst):
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

    with torch.cuda.stream(be.stream):
        a, b = v, v_next
        for _ in range(max(3, args.warmup)):
            step(a, b)
            a, b = b, a
        # ---- timed region: device-resident whole-job throughput ----
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = be.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(be.stream)
        for _ in range(args.steps):
            step(a, b)
            a, b = b, a
        e1.record(be.stream)
        barrier()
        ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        launches = be.launch_count() - l0
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

        # ---- e2e: the C-ABI call with HOST buffers (pinned h2d of v and the mask window, d2h of the result window) ----
        hv = torch.full((n_vec,), 1.0 / n, dtype=torch.float32).pin_memory()
        hm = torch.ones(r1 - r0, dtype=torch.float32).pin_memory()
        hr = torch.empty(r1 - r0, dtype=torch.float32).pin_memory()
        lib, sp = be.lib, be.stream_ptr
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_step():
            lib.splacu_memcpy_h2d(C.c_void_p(a.data_ptr()), C.c_void_p(hv.data_ptr()), n_vec * 4, sp)
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
                lib.splacu_memcpy_h2d(C.c_void_p(dv[q].data_ptr()), C.c_void_p(hv.data_ptr()), n_vec * 4, p_in)
                lib.splacu_memcpy_h2d(C.c_void_p(dm[q].data_ptr()), C.c_void_p(hm.data_ptr()), (r1 - r0) * 4, p_in)
                ev_in[q].record(s_in)
                be.stream.wait_event(ev_in[q])
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
        h2d = (n_vec + (r1 - r0)) * 4
        d2h = (r1 - r0) * 4
        checksum = float(hr.double().sum())

        extra = None
        if args.extra and world == 1:
            extra = extra_timings(be, M, n, nnz, args.steps)

    # ---- CPU baseline: rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            g, sec, info = cpu_reference_run(args, args.cpu_steps, 2)
            cpu = dict(info, value=g, unit=UNIT, ms_per_step=sec * 1e3)
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": f"failed: {ex}"}

    phase_nnz = csr_info.get("phase_nnz") or []
    if phase_nnz:
        shares = ", ".join(f"{x / max(1, nnz_l):.3f}" for x in phase_nnz)
        kernel_name = "mxv_seg_kernel"
        kernel_desc = (f"mxv_seg_kernel x {len(phase_nnz)} column classes (entry shares {shares}; {csr_info['n_hub']} hub columns in shared-memory "
                       f"tables of 16-bit slots, tail class gathers v) + mxv_seg_fixup_kernel per class")
        roofline_note = (f"one step = one splacu_mxv_masked call = {len(phase_nnz)} launches of mxv_seg_kernel (one per column class) with their fix-ups, "
                         "the mask-count / fill pass and the hub pack; achieved = algorithmic bytes of the step / its device time, traffic = DRAM bytes "
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
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved_min, "peak": peak, "unit": "GB/s", "frac": achieved_min / peak,
                         "traffic": ncu_traffic(kernel_name), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": ms_kernel, "kernel_ms_per_rank": kernel_ms_ranks, "note": ("rank-0 bytes, slowest rank's bandwidth; " if world > 1 else "") + roofline_note},
            "cpu_baseline": cpu,
        }
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
