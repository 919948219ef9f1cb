"""Host-side logic of the single-box multi-GPU sharding (spla_b200/dist.py), on CPU with the gloo backend and world size 2:
nnz-balanced row blocks + all-gather of result windows for the pull product, nnz-balanced column blocks + frontier exchange for
the push product. The per-rank arithmetic is done by the C oracle here (no GPU in this test), so what is checked is exactly the
partition / exchange plumbing that bench.py --gpus N and a multi-GPU traversal rely on."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import INT, Oracle
        from spla_b200 import dist as sd
        from spla_b200 import graphs

        orc = Oracle()
        n, Ap, Aj = graphs.rmat(10, edge_factor=8, seed=5)  # every rank generates the same graph
        g = torch.Generator().manual_seed(3)
        Ax = torch.randint(1, 4, (Aj.numel(),), generator=g, dtype=torch.int32)
        v = torch.randint(0, 3, (n,), generator=g, dtype=torch.int32)
        mask = torch.randint(0, 2, (n,), generator=g, dtype=torch.int32)
        hAp, hAj, hAx = Ap.numpy().astype(np.uint32), Aj.numpy().astype(np.uint32), Ax.numpy()

        # ---- pull: row blocks, in-place windows, all-gather ----
        b = sd.balanced_boundaries(Ap, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        nnz_parts = [int(Ap[b[p + 1]] - Ap[b[p]]) for p in range(world)]
        assert max(nnz_parts) <= 0.6 * int(Ap[-1]) + int((Ap[1:] - Ap[:-1]).max())  # nnz-balanced, not n-balanced
        r0, r1 = b[rank], b[rank + 1]
        Ap_l, Aj_l, Ax_l = sd.row_slice(Ap, Aj, Ax, r0, r1)
        full = torch.full((n,), -7, dtype=torch.int32)
        part = orc.mxv_masked(INT, "MULT", "PLUS", "EQZERO", Ap_l.numpy().astype(np.uint32), Aj_l.numpy().astype(np.uint32), Ax_l.numpy(),
                              v.numpy(), mask[r0:r1].numpy(), 0)
        full[r0:r1] = torch.from_numpy(part)
        sd.allgather_windows(full, b)
        want = orc.mxv_masked(INT, "MULT", "PLUS", "EQZERO", hAp, hAj, hAx, v.numpy(), mask.numpy(), 0)
        assert np.array_equal(full.numpy(), want), "row-sharded mxv + all-gather differs from the single-device product"

        # ---- pull in the padded layout: equal windows, column ids mapped once, ONE in-place all-gather ----
        W, shifts = sd.padded_layout(b)
        assert W % 32 == 0 and all(b[p + 1] - b[p] <= W for p in range(world))
        Aj_p = sd.to_padded_index(Aj_l, b, shifts)
        v_p = torch.zeros(world * W, dtype=torch.int32)
        for p in range(world):
            v_p[p * W:p * W + (b[p + 1] - b[p])] = v[b[p]:b[p + 1]]
        part_p = orc.mxv_masked(INT, "MULT", "PLUS", "EQZERO", Ap_l.numpy().astype(np.uint32), Aj_p.numpy().astype(np.uint32), Ax_l.numpy(),
                                v_p.numpy(), mask[r0:r1].numpy(), 0)
        assert np.array_equal(part_p, part), "mapped column ids change the local product"
        full_p = torch.full((world * W,), -7, dtype=torch.int32)
        full_p[rank * W:rank * W + (r1 - r0)] = torch.from_numpy(part_p)
        sd.allgather_padded(full_p, W)
        got = torch.cat([full_p[p * W:p * W + (b[p + 1] - b[p])] for p in range(world)])
        assert np.array_equal(got.numpy(), want), "padded-layout all-gather differs from the single-device product"

        # ---- push: column blocks, frontier exchange ----
        cb = sd.column_boundaries(Aj, n, world)
        c0, c1 = cb[rank], cb[rank + 1]
        Ap_c, Aj_c, Ax_c = sd.column_slice(Ap, Aj, Ax, c0, c1)
        vi = torch.nonzero(v == 2).flatten().to(torch.int32)
        vx = torch.ones(vi.numel(), dtype=torch.int32)
        ri, rx = orc.vxm_masked(INT, "MULT", "PLUS", "EQZERO", Ap_c.numpy().astype(np.uint32), Aj_c.numpy().astype(np.uint32), Ax_c.numpy(),
                                c1 - c0, vi.numpy().astype(np.uint32), vx.numpy(), mask[c0:c1].numpy())
        gi, gx = sd.exchange_frontier(torch.from_numpy(ri.astype(np.int32)), torch.from_numpy(rx), c0)
        wi, wx = orc.vxm_masked(INT, "MULT", "PLUS", "EQZERO", hAp, hAj, hAx, n, vi.numpy().astype(np.uint32), vx.numpy(), mask.numpy())
        assert np.array_equal(gi.numpy().astype(np.uint32), wi) and np.array_equal(gx.numpy(), wx), "column-sharded vxm + frontier exchange differs"
        assert bool((gi[1:] > gi[:-1]).all())  # concatenation of disjoint ordered windows is sorted
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_row_and_column_sharding_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_balanced_boundaries_edge_cases():
    from spla_b200 import dist as sd

    Ap = torch.tensor([0, 0, 0, 10, 10, 11, 11], dtype=torch.int64)  # one hub row swallows most targets
    for parts in (1, 2, 3, 4, 8):
        b = sd.balanced_boundaries(Ap, parts)
        assert len(b) == parts + 1 and b[0] == 0 and b[-1] == 6 and all(b[i] <= b[i + 1] for i in range(parts))
    b = sd.balanced_boundaries(torch.zeros(5, dtype=torch.int64), 4)  # empty matrix
    assert b[0] == 0 and b[-1] == 4


# ---- multi-GPU BFS host logic (spla_b200.algorithms.bfs_dist) with the C oracle doing every rank's arithmetic ----
class _OracleBackend:
    """Stand-in for spla_b200.backend.Backend on CPU tensors: same method names and argument meaning, every op computed by the
    oracle. What the test exercises is bfs_dist's ownership / exchange / push-pull logic, not the kernels."""

    def __init__(self):
        from oracle.oracle import INT, Oracle

        self.orc, self.INT = Oracle(), INT
        self.device, self.stream = "cpu", None

    class _M:
        pass

    def csr(self, n_rows, n_cols, Ap, Aj, Ax):
        m = self._M()
        m.n_rows, m.n_cols = n_rows, n_cols
        m.Ap, m.Aj, m.Ax = (t.numpy().astype(np.uint32) for t in (Ap, Aj, Ax))
        return m

    def sync(self):
        pass

    def v_assign_masked(self, r, mask, value, op_assign, op_select):
        if isinstance(mask, tuple):
            out = self.orc.v_assign_masked_sparse(self.INT, op_assign, op_select, r.numpy(), mask[0].numpy(), mask[1].numpy(), value)
        else:
            out = self.orc.v_assign_masked_dense(self.INT, op_assign, op_select, r.numpy(), mask.numpy(), value)
        r.copy_(torch.from_numpy(out.astype(np.int32)))
        return r

    def vxm_masked(self, M, vi, vx, mask, op_mult, op_add, op_select):
        ri, rx = self.orc.vxm_masked(self.INT, op_mult, op_add, op_select, M.Ap, M.Aj, M.Ax, M.n_cols, vi.numpy().astype(np.uint32), vx.numpy(),
                                     mask.numpy())
        return torch.from_numpy(ri.astype(np.int32)), torch.from_numpy(rx.astype(np.int32))

    def mxv_masked(self, M, v, mask, op_mult, op_add, op_select, init, early_exit=False):
        return torch.from_numpy(self.orc.mxv_masked(self.INT, op_mult, op_add, op_select, M.Ap, M.Aj, M.Ax, v.numpy(), mask.numpy(), init,
                                                    early_exit).astype(np.int32))

    def v_count_mf(self, v, fill):
        return self.orc.v_count_mf_dense(self.INT, v.numpy(), fill)

    def pack_bits(self, v, op_select, out=None):
        assert op_select == "NQZERO"
        words = np.packbits((v.numpy() != 0).astype(np.uint8), bitorder="little")
        words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.int32)
        if out is None:
            return torch.from_numpy(words.copy())
        out[:len(words)] = torch.from_numpy(words.copy())
        return out

    def unpack_bits(self, bits, n, one, zero, out):
        b = np.unpackbits(bits.numpy().view(np.uint8), bitorder="little")[:n]
        out[:n] = torch.from_numpy(np.where(b != 0, one, zero).astype(np.int32))
        return out

    def dense_to_coo(self, dense, fill):
        idx = torch.nonzero(dense != fill).flatten()
        return idx.to(torch.int32), dense[idx]

    def coo_to_dense(self, n, fill, vi, vx):
        d = torch.full((n,), fill, dtype=vx.dtype)
        d[vi.long()] = vx
        return d


def _bfs_reference(n, Ap, Aj, source):
    depth = np.zeros(n, dtype=np.int32)
    depth[source] = 1
    front, level = [source], 1
    while front:
        level += 1
        nxt = []
        for u in front:
            for k in range(Ap[u], Ap[u + 1]):
                j = Aj[k]
                if depth[j] == 0:
                    depth[j] = level
                    nxt.append(j)
        front = nxt
    return depth


def _bfs_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spla_b200 import algorithms, graphs
        from spla_b200 import dist as sd

        be = _OracleBackend()
        n, Ap, Aj = graphs.rmat(9, edge_factor=6, seed=11)
        Ax = torch.ones(Aj.numel(), dtype=torch.int32)
        shards = [algorithms.make_bfs_shard(be, n, Ap, Aj, Ax, rank, world, bitmap_exchange=flag) for flag in (False, True)]
        assert [s["unit_values"] for s in shards] == [False, True]
        hAp, hAj = Ap.numpy(), Aj.numpy()
        deg = np.diff(hAp)
        sources = [int(np.argmax(deg)), int(np.nonzero(deg > 0)[0][0]), int(np.nonzero(deg == 0)[0][0]) if (deg == 0).any() else 0]
        for src in sources:
            want = _bfs_reference(n, hAp, hAj, src)
            for k, (mode, ff) in enumerate((("push", 0.05), ("pull", 0.05), ("push_pull", 0.05), ("push_pull", 0.3), ("pull", 0.05))):
                trace = []
                shard = shards[k & 1]  # dense int32 and bitmap exchange of the pull levels alternate; "pull" runs with both
                mine = algorithms.bfs_dist(be, shard, src, mode=mode, front_factor=ff, trace=trace)
                full = torch.zeros(n, dtype=torch.int32)
                full[shard["w0"]:shard["w1"]] = mine
                sd.allgather_windows(full, shard["bounds"])
                assert np.array_equal(full.numpy(), want), f"bfs_dist {mode} ff={ff} from {src} differs from the sequential BFS"
                if mode == "push_pull" and ff == 0.3 and src == sources[0]:
                    assert {t[0] for t in trace} == {"push", "pull"}, trace  # both directions were exercised
        open(os.path.join(out_dir, f"bfs_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_bfs_dist_matches_sequential_bfs(tmp_path, world):
    mp.spawn(_bfs_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"bfs_ok{r}").exists() for r in range(world))


def _hub_plan_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spla_b200 import dist as sd

        w = 96  # equal windows of the padded layout
        g = torch.Generator().manual_seed(11)
        v_full = torch.randint(-1000, 1000, (world * w,), generator=g, dtype=torch.int32)  # the same on every rank
        gr = torch.Generator().manual_seed(100 + rank)
        # every rank has its OWN hub list (its handle ranks the columns of its row block): a random subset, random slot order,
        # with owners that get nothing from this rank and, on rank 1, an empty window request to rank 0's last element
        n_hub = [40, 57, 13][rank % 3]
        hub_cols = torch.randperm(world * w, generator=gr)[:n_hub].to(torch.int32)
        if rank == 1:
            hub_cols = hub_cols[hub_cols >= w]  # nothing from owner 0
        plan = sd.hub_exchange_plan(hub_cols, w)
        assert sum(plan["recv_counts"]) == hub_cols.numel() and len(plan["recv_counts"]) == world
        assert plan["req"].numel() == sum(plan["send_counts"])
        window = v_full[rank * w:(rank + 1) * w]
        send = window[plan["req"].long()].contiguous()  # what splacu_v_gather does on the device
        recv = torch.empty(hub_cols.numel(), dtype=torch.int32)
        dist.all_to_all_single(recv, send, output_split_sizes=plan["recv_counts"], input_split_sizes=plan["send_counts"])
        hub_vals = torch.full((hub_cols.numel(),), -7, dtype=torch.int32)
        hub_vals[plan["order"]] = recv  # splacu_v_scatter
        assert torch.equal(hub_vals, v_full[hub_cols.long()]), "hub values gathered from their owners differ from v[hub_cols]"
        # the direct form (splacu_v_push_peers): the owner stores value k into slot dst_slot[k] of the requester's table; emulate the
        # peer stores by shipping the slots beside the values
        assert plan["dst_slot"].numel() == plan["req"].numel()
        slots = torch.empty(hub_cols.numel(), dtype=torch.int32)
        dist.all_to_all_single(slots, plan["dst_slot"].contiguous(), output_split_sizes=plan["recv_counts"], input_split_sizes=plan["send_counts"])
        table = torch.full((hub_cols.numel(),), -7, dtype=torch.int32)
        table[slots.long()] = recv
        assert torch.equal(table, v_full[hub_cols.long()]), "hub values pushed into the requesters' slots differ from v[hub_cols]"
        open(os.path.join(out_dir, f"hub_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_hub_value_exchange_plan(tmp_path, world):
    """The owner-side gather / all-to-all / slot scatter that feeds part 1 of the two-part pull product (dist.PipelinedPull) delivers
    exactly v[hub_cols] to every rank, for uneven per-rank hub lists and owners that contribute nothing."""
    mp.spawn(_hub_plan_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"hub_ok{r}")) for r in range(world))
