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
