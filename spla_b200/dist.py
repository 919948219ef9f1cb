"""Single-box multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch on the GPU box, gloo in the CPU tests) as plumbing. Net-new relative to the reference, which is
single-device.

  mxv pull   rows are independent: contiguous row blocks chosen on the prefix sum of Ap so that every rank holds
             ~nnz/P entries (nnz-balanced, not n-balanced). Each rank owns r[rows_p] and mask[rows_p]; a step is the
             local masked mxv on the slice, written IN PLACE into the rank's window of a full-length vector, followed by
             an all-gather of the windows -- the gathered vector is exactly the next step's input.
  vxm push   column-sharded: rank p stores M[:, cols_p] as CSR over all rows (column ranges nnz-balanced), owns
             mask[cols_p] and produces r[cols_p]; the windows are disjoint so the concatenation is already sorted.
             The frontier exchange is an all-gather of the per-rank (count, indices, values).

The partition helpers are pure index arithmetic on tensors and run on CPU tensors too (used by the gloo tests).
"""
import os
import sys

import torch
import torch.distributed as dist


def balanced_boundaries(Ap, parts):
    """Row boundaries b[0..parts] with b[0]=0, b[parts]=n such that each [b[p], b[p+1]) holds ~nnz/parts entries.
    Ap: int64/int32 tensor [n+1] (any device)."""
    n = Ap.numel() - 1
    nnz = int(Ap[-1].item())
    targets = torch.arange(1, parts, device=Ap.device, dtype=torch.int64) * nnz // parts
    cuts = torch.searchsorted(Ap.to(torch.int64).contiguous(), targets, right=False).clamp(max=n)
    b = [0] + [int(c) for c in cuts.tolist()] + [n]
    for i in range(1, len(b)):  # keep it monotone when a huge row swallows several targets
        b[i] = max(b[i], b[i - 1])
    return b


def row_slice(Ap, Aj, Ax, r0, r1):
    """CSR slice of rows [r0, r1): (Ap_local int32 rebased to 0, Aj_local, Ax_local) -- all columns kept."""
    k0, k1 = int(Ap[r0].item()), int(Ap[r1].item())
    Ap_l = (Ap[r0:r1 + 1] - Ap[r0]).to(torch.int32).contiguous()
    # clone: fresh allocations are 16-byte aligned, which the 128-bit streaming loads of the pull kernel need
    return Ap_l, Aj[k0:k1].clone(), Ax[k0:k1].clone()


def column_slice(Ap, Aj, Ax, c0, c1):
    """CSR over ALL rows restricted to columns [c0, c1), columns rebased to 0 (for column-sharded vxm)."""
    n = Ap.numel() - 1
    keep = (Aj >= c0) & (Aj < c1)
    rows = torch.repeat_interleave(torch.arange(n, device=Ap.device), (Ap[1:] - Ap[:-1]).to(torch.int64))
    cnt = torch.bincount(rows[keep], minlength=n)
    Ap_l = torch.zeros(n + 1, dtype=torch.int64, device=Ap.device)
    torch.cumsum(cnt, 0, out=Ap_l[1:])
    return Ap_l.to(torch.int32), (Aj[keep] - c0).to(torch.int32).contiguous(), Ax[keep].contiguous()


def column_boundaries(Aj, n_cols, parts):
    """nnz-balanced column ranges for the column-sharded push."""
    cnt = torch.bincount(Aj.to(torch.int64), minlength=n_cols)
    Cp = torch.zeros(n_cols + 1, dtype=torch.int64, device=Aj.device)
    torch.cumsum(cnt, 0, out=Cp[1:])
    return balanced_boundaries(Cp, parts)


def allgather_windows(full, bounds, group=None):
    """In-place all-gather of the per-rank windows full[b[p]:b[p+1]] of one full-length vector.
    The windows are nnz-balanced, hence uneven: equal windows take one all_gather_into_tensor, uneven ones one broadcast per
    owner straight into place (no staging copy; issued together, on NVSwitch every broadcast runs at full link rate)."""
    world = dist.get_world_size(group)
    if world == 1:
        return full
    buf = full.view(torch.int32) if full.dtype == torch.uint32 else full
    sizes = [bounds[p + 1] - bounds[p] for p in range(world)]
    if len(set(sizes)) == 1 and sizes[0] > 0:
        rank = dist.get_rank(group)
        dist.all_gather_into_tensor(buf[bounds[0]:bounds[world]], buf[bounds[rank]:bounds[rank + 1]].clone(), group=group)
        return full
    works = []
    for p in range(world):
        if sizes[p] > 0:
            src = dist.get_global_rank(group, p) if group is not None else p
            works.append(dist.broadcast(buf[bounds[p]:bounds[p + 1]], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    return full


def padded_layout(bounds, align=32):
    """Distributed layout of a full-length vector with EQUAL windows: the rows [b[p], b[p+1]) of rank p live at
    [p*W, p*W + size_p) of a vector of length world*W, W = the largest window rounded up to `align` elements. The nnz-balanced
    windows are uneven, and an uneven all-gather costs either one broadcast per owner or staging copies; in this layout a step
    ends with ONE in-place ncclAllGather and no copy. Column ids of the matrix are mapped once (to_padded_index).
    Returns (W, shifts) with padded index = global index + shifts[owner]."""
    world = len(bounds) - 1
    w = max(bounds[p + 1] - bounds[p] for p in range(world))
    w = (max(w, 1) + align - 1) // align * align
    return w, [p * w - bounds[p] for p in range(world)]


def to_padded_index(idx, bounds, shifts):
    """global row / column ids -> ids in the padded layout (idx: integer tensor on any device)."""
    cuts = torch.tensor(bounds[1:-1], dtype=torch.int64, device=idx.device)
    owner = torch.bucketize(idx.to(torch.int64), cuts, right=True)
    sh = torch.tensor(shifts, dtype=torch.int64, device=idx.device)
    return (idx.to(torch.int64) + sh[owner]).to(idx.dtype)


def allgather_padded(full_padded, w, group=None):
    """In-place all-gather of the equal windows full_padded[p*w:(p+1)*w] (one collective, no staging on NCCL)."""
    world = dist.get_world_size(group)
    if world == 1:
        return full_padded
    rank = dist.get_rank(group)
    buf = full_padded.view(torch.int32) if full_padded.dtype == torch.uint32 else full_padded
    mine = buf[rank * w:(rank + 1) * w]
    if dist.get_backend(group) != "nccl":
        mine = mine.clone()  # gloo wants distinct buffers
    dist.all_gather_into_tensor(buf[:world * w], mine, group=group)
    return full_padded


class PeerVector:
    """A full-length 4-byte-element vector that every rank of a single box holds at the same offsets in peer-mapped memory
    (torch.distributed._symmetric_memory). publish(w0, count) writes this rank's window into every peer's copy with one kernel of
    NVLink peer stores (splacu_publish_window) and closes the step with the device-side barrier of the symmetric allocation --
    the all-gather of the row-sharded pull without a collective library call. NCCL / CUDA only."""

    def __init__(self, backend, n, dtype=torch.float32, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        self.backend, self.group = backend, group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n_pad = (n + 3) // 4 * 4
        self.tensor = symm.empty(n_pad, dtype=dtype, device=backend.device)
        self.handle = symm.rendezvous(self.tensor, self.group)
        ptrs = list(self.handle.buffer_ptrs)
        assert len(ptrs) == self.world and ptrs[self.rank] == self.tensor.data_ptr()
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self._C = C

    def publish(self, w0, count):
        """window [w0, w0 + count) of the local copy -> all peers; count is rounded up to 4 elements (padded layout: the tail of a
        window is padding). Runs on the backend's stream; returns after enqueueing."""
        C = self._C
        cnt = (count + 3) // 4 * 4
        rc = self.backend.lib.splacu_publish_window(self._ptrs, self.world, self.rank, C.c_size_t(w0), C.c_size_t(cnt), self.backend.stream_ptr)
        if rc != 0:
            raise RuntimeError(f"splacu_publish_window failed ({rc})")
        with torch.cuda.stream(self.backend.stream):
            self.handle.barrier(channel=0)


def exchange_frontier(vi_local, vx_local, offset, group=None, sizes=None):
    """All-gather of sparse frontier pieces. Each rank contributes (indices local to its window + offset, values);
    returns the concatenated global (vi, vx), sorted because windows are disjoint and ordered by rank.
    Pieces are uneven: counts are exchanged first (unless the caller already knows every rank's `sizes`, as a traversal does
    from the all-gather of the front sizes that ends its previous level), then the pieces travel padded to the largest one."""
    world = dist.get_world_size(group)
    vi_g = vi_local + offset if offset else vi_local
    if world == 1:
        return vi_g, vx_local
    dev = vi_local.device
    if sizes is None:
        cnt = torch.tensor([vi_local.numel()], dtype=torch.int64, device=dev)
        cnts = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cnts, cnt, group=group)
        sizes = [int(c) for c in cnts.tolist()]
    assert sizes[dist.get_rank(group)] == vi_local.numel()
    cap = max(max(sizes), 1)
    send = torch.zeros(2 * cap, dtype=torch.int32, device=dev)  # [0, cap): indices, [cap, 2 cap): value bit patterns
    send[:vi_g.numel()] = vi_g.to(torch.int32)
    send[cap:cap + vx_local.numel()] = vx_local.contiguous().view(torch.int32)
    recv = torch.empty(world * 2 * cap, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, 2, cap)
    vi = torch.cat([recv[p, 0, :sizes[p]] for p in range(world)])
    vx = torch.cat([recv[p, 1, :sizes[p]] for p in range(world)]).view(vx_local.dtype)
    return vi.to(vi_g.dtype), vx


# ---- the row-sharded pull with the exchange overlapped (hub values first) ------------------------------------------------------
def hub_exchange_plan(hub_cols, w, group=None):
    """Who sends which elements of its window so that every rank gets v[hub_cols] of ITS handle without waiting for the whole vector.
    hub_cols: this rank's hub column ids in the padded layout (slot order; owner of id c = c // w, offset in the owner's window = c % w).
    Returns a dict:
      order        int64 [n_hub]  received value j belongs to slot order[j] (values arrive grouped by owner, slot order inside a group)
      recv_counts  list [world]   values this rank receives from each owner
      req          int32 [...]    offsets into THIS rank's window that the other ranks ask for, grouped by requester
      send_counts  list [world]   how many of them go to each requester
      dst_slot     int32 [...]    for every element of req: the slot of the REQUESTER's hub table it fills (splacu_v_push_peers)
    Pure index arithmetic + two small all-to-all exchanges (runs on gloo with CPU tensors, too)."""
    world = dist.get_world_size(group)
    cols = hub_cols.to(torch.int64)
    owner = torch.div(cols, w, rounding_mode="floor")
    order = torch.argsort(owner, stable=True)
    recv_counts = torch.bincount(owner, minlength=world)
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts, group=group)
    offs = (cols[order] - owner[order] * w).to(torch.int32).contiguous()
    rc, sc = [int(x) for x in recv_counts.tolist()], [int(x) for x in send_counts.tolist()]
    req = torch.empty(sum(sc), dtype=torch.int32, device=hub_cols.device)
    dist.all_to_all_single(req, offs, output_split_sizes=sc, input_split_sizes=rc, group=group)
    # where the requester wants every value: slot order[j] of its hub table for its j-th received value (sent back grouped like req)
    dst_slot = torch.empty(sum(sc), dtype=torch.int32, device=hub_cols.device)
    dist.all_to_all_single(dst_slot, order.to(torch.int32).contiguous(), output_split_sizes=sc, input_split_sizes=rc, group=group)
    return {"order": order, "recv_counts": rc, "req": req, "send_counts": sc, "dst_slot": dst_slot}


class PipelinedPull:
    """Row-sharded pull (one process per GPU) whose exchange overlaps the product.

    A plain step is: local product -> all-gather of the windows -> next step. The hub classes of the product (63 % of the entries of
    an RMAT matrix) read only the ~180 K most referenced elements of v, though, so the owners' values of every rank's hub columns
    (< 1 MB per rank) travel first and the hub classes (splacu_mxv_masked_part) start on them while the windows are still in flight;
    the rest (row classes, tail classes, fix-ups) waits for the windows, the prologue (mask pass) for nothing. Same results as the plain
    step (the classes add onto r in the same order).

    Two flows:
      * peer-mapped vectors (`peers`: {tensor.data_ptr(): PeerVector}; the default of bench.py): EXCHANGE-FIRST steps. step(src, dst)
        begins with the exchange of src -- whose own window the previous step produced -- and computes dst's window beside it:
        hub values by one kernel of NVLink peer stores straight into the peers' hub tables (splacu_v_push_peers), windows as
        cudaMemcpyAsync peer copies on the copy engines, each closed by the device barrier of a symmetric allocation (no SM is taken from
        the persistent class kernels: an NCCL kernel beside them delays one class CTA per SM it holds). prepare(a, b) captures both
        directions of the ping-pong as CUDA graphs, so a step costs the host one cudaGraphLaunch instead of ~25 launches, copies and
        event operations. finish() exchanges the last result. Measured on RMAT-24 (profiles/r02_bench_multi_gpu.txt): 0.717 ms on 2,
        0.453 ms on 4, 0.3245 ms on 8 B200s against 1.302 ms on one.
      * without peer-mapped vectors: EXCHANGE-LAST steps over the collective library (a small all-to-all for the hub values + an in-place
        ncclAllGather of the windows on two communicators, started after the product and awaited by the next step's parts)."""

    def __init__(self, backend, M, w, w0, n_win, ops, init, mask_l, group=None, peers=None):
        import ctypes as C

        self.C = C
        self.peers = peers or {}
        self.be, self.M, self.w, self.w0, self.n_win = backend, M, w, w0, n_win
        self.ops, self.init, self.mask_l, self.group = ops, init, mask_l, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.hub_cols = backend.csr_hub_cols(M)
        n_hub = self.hub_cols.numel()
        dev = self.hub_cols.device
        # every rank must have column classes for the split to make sense (tiny shards have none: fall back to the plain step)
        stat = torch.tensor([n_hub if n_hub > 0 else 0, -n_hub], device=dev, dtype=torch.int64)
        dist.all_reduce(stat, op=dist.ReduceOp.MIN, group=group)
        self.enabled = int(stat[0].item()) > 0 and self.world > 1
        if not self.enabled:
            return
        n_hub_max = -int(stat[1].item())
        plan = hub_exchange_plan(self.hub_cols, w, group)
        self.order = plan["order"].to(torch.int32).contiguous()
        self.rc, self.sc, self.req = plan["recv_counts"], plan["send_counts"], plan["req"]
        self.send = torch.empty(max(1, sum(self.sc)), dtype=torch.float32, device=dev)[:sum(self.sc)]
        self.hub_vals = torch.empty(n_hub, dtype=torch.float32, device=dev)
        self.comm, self.dma = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.comm_ptr, self.dma_ptr = C.c_void_p(self.comm.cuda_stream), C.c_void_p(self.dma.cuda_stream)
        # further copy streams: the peer copies of one step are spread over them so that several copy engines work at once
        # (default 1: on 8 GPUs one stream of 7 copies measured 0.446 ms per step, four streams 0.469)
        n_dma = max(1, min(int(os.environ.get("SPLA_B200_DMA_STREAMS", "1")), self.world - 1))
        self.dma_more = [torch.cuda.Stream(device=dev) for _ in range(n_dma - 1)]
        self.dma_more_ev = [torch.cuda.Event() for _ in self.dma_more]
        self.ev_done, self.ev_hub, self.ev_full = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.primed, self.k = False, 0
        self.recv_sym = None
        self.graphs, self.last_dst = {}, None
        self.graph_launches, self.replay_launches = {}, 0
        # transport of the windows in the peer-mapped flow: "dma" = peer copies on the copy engines, "nccl" = in-place ncclAllGather
        # (default "dma": measured on 2 and 8 B200s; "nccl" inside the graph was only run on 2 GPUs -- 0.759 ms against 0.716 -- and is opt-in)
        self.win_nccl = os.environ.get("SPLA_B200_WIN", "dma") == "nccl"
        self.win_group = None
        self.use_graphs = os.environ.get("SPLA_B200_GRAPH", "1") == "1"
        if self.peers:
            # owner p writes its values for rank q at q's offset recv_off[q][p] of q's receive buffer (double-buffered by step parity:
            # a push for step k + 2 can only start after every rank has passed the hub barrier of step k + 1, i.e. after the scatter
            # of step k has read the buffer)
            off = torch.zeros(self.world, dtype=torch.int64, device=dev)
            off[1:] = torch.cumsum(torch.tensor(self.rc[:-1], dtype=torch.int64, device=dev), 0)
            table = torch.empty(self.world * self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(table, off, group=group)
            self.recv_off = [int(x) for x in table.view(self.world, self.world)[:, self.rank].tolist()]  # where MY values go at rank q
            self.send_off = [0] * self.world
            for q in range(1, self.world):
                self.send_off[q] = self.send_off[q - 1] + self.sc[q - 1]
            self.recv_sym = [PeerVector(backend, n_hub_max, group=group), PeerVector(backend, n_hub_max, group=group)]
            if self.win_nccl:  # its own communicator (and NCCL stream) beside whatever else the caller runs on the default group
                self.win_group = dist.new_group(backend=dist.get_backend(group))
            # direct form (default): one kernel stores the requested values straight into the requesters' hub tables (recv_sym used as the
            # table itself, slot order) -- no copies, no scatter; SPLA_B200_HUB_PUSH=0 keeps gather + copies + scatter
            self.hub_push = os.environ.get("SPLA_B200_HUB_PUSH", "1") == "1"
            self.dst_slot = plan["dst_slot"].contiguous()
            so = [0] * (self.world + 1)
            for q in range(self.world):
                so[q + 1] = so[q] + self.sc[q]
            self.seg_off = torch.tensor(so, dtype=torch.int32, device=dev)
            self.peer_tabs = [torch.tensor([int(x) for x in rs.handle.buffer_ptrs], dtype=torch.int64, device=dev) for rs in self.recv_sym]
        else:
            self.recv = torch.empty(n_hub, dtype=torch.float32, device=dev)
            self.small = dist.new_group(backend=dist.get_backend(group))  # its own communicator: runs beside the big all-gather

    def step(self, src, dst):
        """dst[w0 : w0 + n_win] = M_local x src, then the exchange of dst's windows is STARTED; the next step (or finish()) waits for it."""
        be, M, C = self.be, self.M, getattr(self, "C", None)
        out = dst[self.w0:self.w0 + self.n_win]
        if not self.enabled:
            be.mxv_masked(M, src, self.mask_l, *self.ops, self.init, out=out)
            allgather_padded(dst, self.w, self.group)
            return
        if self.recv_sym is not None and self.peers.get(src.data_ptr()) is not None and self.peers.get(dst.data_ptr()) is not None:
            self._step_peer(src, dst)
            return
        PRO, HUB, REST = 4, 1, 2
        be.mxv_masked_part(M, PRO, None, None, self.mask_l, *self.ops, self.init, out)
        if self.primed:
            be.stream.wait_event(self.ev_hub)
            be.mxv_masked_part(M, HUB, None, self.hub_vals, self.mask_l, *self.ops, self.init, out)
            be.stream.wait_event(self.ev_full)
        else:  # the first step finds src complete: the hub values come from it
            be.mxv_masked_part(M, HUB, src, None, self.mask_l, *self.ops, self.init, out)
        be.mxv_masked_part(M, REST, src, None, self.mask_l, *self.ops, self.init, out)
        self.ev_done.record(be.stream)
        pv = self.peers.get(dst.data_ptr())
        # (1) the hub values
        self.comm.wait_event(self.ev_done)
        be.v_gather(self.req, out, self.send, stream_ptr=self.comm_ptr)
        if self.recv_sym is not None:
            rs = self.recv_sym[self.k & 1]
            for d in range(self.world):
                q = (self.rank + d) % self.world
                if self.sc[q]:
                    be._check(be.lib.splacu_memcpy_d2d(C.c_void_p(rs._ptrs[q] + self.recv_off[q] * 4), C.c_void_p(self.send.data_ptr() + self.send_off[q] * 4),
                                                       self.sc[q] * 4, self.comm_ptr))
            with torch.cuda.stream(self.comm):
                rs.handle.barrier(channel=0)
            be.v_scatter(self.order, rs.tensor, self.hub_vals, stream_ptr=self.comm_ptr)
            self.ev_hub.record(self.comm)
        else:
            with torch.cuda.stream(self.comm):
                w_small = dist.all_to_all_single(self.recv, self.send, output_split_sizes=self.rc, input_split_sizes=self.sc, group=self.small, async_op=True)
                w_big = None
                if pv is None:
                    mine = dst[self.w0:self.w0 + self.w]
                    w_big = dist.all_gather_into_tensor(dst[:self.world * self.w], mine, group=self.group, async_op=True)
                w_small.wait()
                be.v_scatter(self.order, self.recv, self.hub_vals, stream_ptr=self.comm_ptr)
                self.ev_hub.record(self.comm)
                if w_big is not None:
                    w_big.wait()
                    self.ev_full.record(self.comm)
        # (2) the windows: peer copies on the copy engines, then the barrier of the symmetric allocation
        if pv is not None:
            self.dma.wait_event(self.ev_done)
            for st in self.dma_more:
                st.wait_event(self.ev_done)
            off, nbytes = self.w0 * 4, self.w * 4
            lanes = [self.dma] + self.dma_more
            for d in range(1, self.world):
                p = (self.rank + d) % self.world
                st = lanes[(d - 1) % len(lanes)]
                be._check(be.lib.splacu_memcpy_d2d(C.c_void_p(pv._ptrs[p] + off), C.c_void_p(dst.data_ptr() + off), nbytes, C.c_void_p(st.cuda_stream)))
            for st, ev in zip(self.dma_more, self.dma_more_ev):
                ev.record(st)
                self.dma.wait_event(ev)
            with torch.cuda.stream(self.dma):
                pv.handle.barrier(channel=1)
            self.ev_full.record(self.dma)
        elif self.recv_sym is not None:  # a vector that is not peer-mapped: the collective library moves the windows
            self.dma.wait_event(self.ev_done)
            with torch.cuda.stream(self.dma):
                allgather_padded(dst, self.w, self.group)
            self.ev_full.record(self.dma)
        self.primed = True
        self.k += 1

    # ---- peer-mapped vectors: a step = (exchange of src, started first) beside (product into dst); one CUDA graph per direction ----
    def _exchange_peer(self, vec, parity, hub):
        """This rank's window of `vec` -> every peer's copy (copy engines, spread over the copy streams), closed by the device barrier of
        the symmetric allocation on self.dma (-> ev_full); with `hub` also the owners' values of every rank's hub columns -> their
        receive buffers, barrier, scatter into self.hub_vals on self.comm (-> ev_hub). Forked from the backend stream at the call."""
        be, C = self.be, self.C
        pv = self.peers[vec.data_ptr()]
        self.ev_done.record(be.stream)
        lanes = [self.dma] + self.dma_more
        if hub and self.hub_push:
            self.comm.wait_event(self.ev_done)
            rs = self.recv_sym[parity]
            be.v_push_peers(self.req, self.dst_slot, self.seg_off, self.peer_tabs[parity], vec[self.w0:self.w0 + self.n_win], stream_ptr=self.comm_ptr)
            with torch.cuda.stream(self.comm):
                rs.handle.barrier(channel=0)
            self.ev_hub.record(self.comm)
            self.cur_hub = rs.tensor
        elif hub:
            self.cur_hub = self.hub_vals
            self.comm.wait_event(self.ev_done)
            be.v_gather(self.req, vec[self.w0:self.w0 + self.n_win], self.send, stream_ptr=self.comm_ptr)
            rs = self.recv_sym[parity]
            for d in range(self.world):
                q = (self.rank + d) % self.world
                if self.sc[q]:
                    be._check(be.lib.splacu_memcpy_d2d(C.c_void_p(rs._ptrs[q] + self.recv_off[q] * 4), C.c_void_p(self.send.data_ptr() + self.send_off[q] * 4),
                                                       self.sc[q] * 4, self.comm_ptr))
            with torch.cuda.stream(self.comm):
                rs.handle.barrier(channel=0)
            be.v_scatter(self.order, rs.tensor, self.hub_vals, stream_ptr=self.comm_ptr)
            self.ev_hub.record(self.comm)
        if self.win_nccl:
            # the windows by the collective library (in place, equal windows): on 8 GPUs its all-gather (NVLS multicast) is faster than
            # 7 peer copies per rank; captured into the step's graph like everything else
            self.dma.wait_event(self.ev_done)
            with torch.cuda.stream(self.dma):
                dist.all_gather_into_tensor(vec[:self.world * self.w], vec[self.w0:self.w0 + self.w], group=self.win_group)
            self.ev_full.record(self.dma)
            return
        for st in lanes:
            st.wait_event(self.ev_done)
        off, nbytes = self.w0 * 4, self.w * 4
        for d in range(1, self.world):
            p = (self.rank + d) % self.world
            st = lanes[(d - 1) % len(lanes)]
            be._check(be.lib.splacu_memcpy_d2d(C.c_void_p(pv._ptrs[p] + off), C.c_void_p(vec.data_ptr() + off), nbytes, C.c_void_p(st.cuda_stream)))
        for st, ev in zip(self.dma_more, self.dma_more_ev):
            ev.record(st)
            self.dma.wait_event(ev)
        if hub:
            # the two device barriers of a step spin on the peers: keep them in ONE order on every rank (hub barrier, then window barrier)
            # -- independent branches of a graph are not guaranteed to run concurrently, and two ranks that serialised them in opposite
            # orders would wait for each other forever
            self.dma.wait_event(self.ev_hub)
        with torch.cuda.stream(self.dma):
            pv.handle.barrier(channel=1)
        self.ev_full.record(self.dma)

    def _body_peer(self, src, dst, parity):
        be, M = self.be, self.M
        out = dst[self.w0:self.w0 + self.n_win]
        PRO, HUB, REST = 4, 1, 2
        self._exchange_peer(src, parity, hub=True)
        be.mxv_masked_part(M, PRO, None, None, self.mask_l, *self.ops, self.init, out)
        be.stream.wait_event(self.ev_hub)
        be.mxv_masked_part(M, HUB, None, self.cur_hub, self.mask_l, *self.ops, self.init, out)
        be.stream.wait_event(self.ev_full)
        be.mxv_masked_part(M, REST, src, None, self.mask_l, *self.ops, self.init, out)

    def _step_peer(self, src, dst):
        """Exchange-first step: src holds this rank's fresh window (the peers' windows may be stale); its exchange starts at once, the
        mask pass runs beside it, the hub classes wait for the hub values, the rest for the windows. The sequence is fixed per
        (src, dst), so after one eager run it is replayed as a CUDA graph: ~25 launches, copies and events per step cost one
        cudaGraphLaunch on the host (a rank of an 8-GPU run computes for ~0.25 ms per step -- the Python issue loop is longer)."""
        key = (src.data_ptr(), dst.data_ptr())
        g = self.graphs.get(key)
        if g is not None:
            g.replay()
            self.replay_launches += self.graph_launches[key]
        else:
            self._body_peer(src, dst, 0 if key[0] < key[1] else 1)
        self.last_dst = dst
        self.k += 1

    def prepare(self, a, b):
        """Collective: one eager step in each direction of the ping-pong (a -> b, b -> a; allocations and function attributes settle),
        then both directions are captured as CUDA graphs. Without it (or with SPLA_B200_GRAPH=0) the steps stay eager."""
        if not (self.enabled and self.use_graphs and self.recv_sym is not None and self.peers.get(a.data_ptr()) is not None
                and self.peers.get(b.data_ptr()) is not None):
            return False
        self._step_peer(a, b)
        self._step_peer(b, a)
        self.finish()
        ok = True
        for src, dst in ((a, b), (b, a)):
            key = (src.data_ptr(), dst.data_ptr())
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                l0 = self.be.launch_count()
                with torch.cuda.graph(g, stream=self.be.stream, capture_error_mode="relaxed"):
                    self._body_peer(src, dst, 0 if key[0] < key[1] else 1)
                self.graphs[key] = g
                self.graph_launches[key] = self.be.launch_count() - l0  # kernels of this library inside one replay
            except Exception as ex:  # noqa: BLE001
                ok = False
                self.graphs.pop(key, None)
                print(f"PipelinedPull rank {self.rank}: CUDA graph capture failed ({ex}); eager steps", file=sys.stderr)
                torch.cuda.synchronize()
        # all ranks or none: a rank replaying a graph and a rank issuing eagerly would still meet at the same barriers, but keep it uniform
        flag = torch.tensor([1 if ok else 0], device=self.hub_cols.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if not bool(flag.item()):
            self.graphs = {}
        return bool(self.graphs)

    def finish(self):
        """the backend stream waits for the exchange of the last step (the vector is complete after this point in the stream)"""
        if self.enabled and self.last_dst is not None:
            self._exchange_peer(self.last_dst, 0, hub=False)
            self.be.stream.wait_event(self.ev_full)
            self.last_dst = None
            return
        if self.enabled and self.primed:
            self.be.stream.wait_event(self.ev_hub)
            self.be.stream.wait_event(self.ev_full)
            self.primed = False
