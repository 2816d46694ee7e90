"""The N>1 host logic on CPU with the gloo backend, world size 2 (no GPU needed).

The sharding layer (nbody_streams_b200/engine.py: ShardLayout, allgather_shards, gather_rows) is plain
torch.distributed code; here two CPU processes run a sharded KDK loop in which each rank owns the i-rows of
its shard, the per-step exchange is the same in-place all-gather of equal-size shards the GPU engine uses,
and the pair sums of the local rows come from the oracle (test infrastructure standing in for the sweep
kernel, which needs a GPU).  The result must equal the single-process loop bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from nbody_streams_b200.engine import ShardLayout, allgather_shards, gather_rows, group_info

N, STEPS, DT, H = 700, 4, 1e-3, 0.05
REC = 32   # bytes per particle slot in the emulated packed buffer: x, y, z, pad as float64


def _ic():
    rng = np.random.default_rng(123)
    pos = rng.standard_normal((N, 3))
    vel = rng.standard_normal((N, 3)) * 0.1
    mass = rng.uniform(0.5, 1.5, N)
    return pos, vel, mass


def _serial():
    pos, vel, mass = _ic()
    acc = oracle.compute_nbody_forces_cpu(pos, mass, H, 1.0, "plummer")
    for _ in range(STEPS):
        vel += acc * (DT / 2)
        pos += vel * DT
        acc = oracle.compute_nbody_forces_cpu(pos, mass, H, 1.0, "plummer")
        vel += acc * (DT / 2)
    return np.hstack([pos, vel])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert group_info() == (rank, world)
        L = ShardLayout(N, world)
        rows = L.rows(rank)
        pos_all, vel_all, mass = _ic()
        pos, vel = pos_all[rows].copy(), vel_all[rows].copy()
        # emulated packed j-buffer: one 32-byte record per slot, all shards equal size
        buf = torch.zeros(L.n_slots * REC, dtype=torch.uint8)
        view = buf.view(torch.float64).view(L.n_slots, 4)

        def publish():
            view[L.first(rank):L.first(rank) + len(pos), :3] = torch.from_numpy(pos)
            allgather_shards(buf, L, rank, L.shard * REC)
            return view[:N, :3].numpy().copy()

        def forces(p_all):
            return oracle.compute_nbody_forces_cpu(p_all, mass, H, 1.0, "plummer", rows=(rows.start, rows.stop))

        acc = forces(publish())
        for _ in range(STEPS):
            vel += acc * (DT / 2)
            pos += vel * DT
            acc = forces(publish())
            vel += acc * (DT / 2)
        full = gather_rows(torch.from_numpy(np.hstack([pos, vel])), L, rank)
        tot = torch.tensor([float(len(pos))])
        dist.all_reduce(tot)
        assert int(tot.item()) == N
        if rank == 0:
            out.put(full.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_sharded_kdk_world2_equals_serial():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert np.array_equal(res, _serial())


def test_gather_rows_world1_is_identity():
    L = ShardLayout(10, 1)
    x = torch.arange(30.0).reshape(10, 3)
    assert gather_rows(x, L, 0) is x
    assert allgather_shards(torch.zeros(L.n_slots * 4, dtype=torch.uint8), L, 0, L.shard * 4) is None


# ---------------------------------------------------------------------------
# pair-symmetric multi-rank schedule (engine.sym_plan + reduce_scatter_shards)
# ---------------------------------------------------------------------------
from nbody_streams_b200.engine import reduce_scatter_shards, sym_plan  # noqa: E402


def _plan_tile_pairs(L, rank, tpb):
    """(i-tile, j-tile, one_sided) work items of one rank, expanded from its plan the way nb_sym.cuh does:
    block ib sweeps the tiles of its own range from its first tile on (the first `tpb` of them are the
    one-sided 'diagonal' tiles), ranges A and B, and range C if ibC0 <= ib < ibC1."""
    p = sym_plan(L, rank, tpb)
    it0, it1 = p["own"]
    items = []
    for ib in range(p["n_ib"]):
        blk = list(range(it0 + tpb * ib, min(it0 + tpb * (ib + 1), it1)))
        js = [(t, t < it0 + tpb * (ib + 1)) for t in range(it0 + tpb * ib, it1)]
        js += [(t, False) for t in range(*p["A"])] + [(t, False) for t in range(*p["B"])]
        if p["ibC"][0] <= ib < p["ibC"][1]:
            js += [(t, False) for t in range(*p["C"])]
        for jt, diag in js:
            for t in blk:
                items.append((t, jt, diag))
    return items


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("n", [100, 256 * 7 + 3, 256 * 64, 256 * 131 + 17])
@pytest.mark.parametrize("tpb", [2, 4])
def test_sym_plan_covers_every_tile_pair_exactly_once(world, n, tpb):
    L = ShardLayout(n, world)
    R = L.tiles_total
    ordered = np.zeros((R, R), dtype=int)      # ordered[i, j]: how often particles of tile i receive from tile j
    for rank in range(world):
        for it, jt, diag in _plan_tile_pairs(L, rank, tpb):
            assert 0 <= it < R and 0 <= jt < R, "padding tiles must never be swept"
            ordered[it, jt] += 1               # i side
            if not diag:
                assert it != jt
                ordered[jt, it] += 1           # j side of a symmetric unit
    assert (ordered == 1).all()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sym_plan_balances_the_ranks(world):
    L = ShardLayout(256 * 64 * world, world)
    work = [sum(1 if d else 2 for _, _, d in _plan_tile_pairs(L, r, 4)) for r in range(world)]
    assert max(work) <= 1.05 * min(work)


def _sym_worker(rank, world, port, out):
    """Each rank accumulates the partial sums of BOTH sides of its tile pairs (pair sums from the oracle with
    the source masses restricted to one tile), then the shards are reduce-scattered: the data flow of
    DirectEngine._compute_sym without the GPU kernel."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, tile = 256 * 5 + 40, 256
        rng = np.random.default_rng(7)
        pos = rng.standard_normal((n, 3))
        mass = rng.uniform(0.5, 1.5, n)
        L = ShardLayout(n, world)

        def pull(it, jt):   # sum over j in tile jt of m_j K (r_j - r_i) for the rows of tile it
            m = np.zeros(n)
            m[jt * tile:(jt + 1) * tile] = mass[jt * tile:(jt + 1) * tile]
            rows = (it * tile, min((it + 1) * tile, n))
            return rows, oracle.compute_nbody_forces_cpu(pos, m, H, 1.0, "plummer", rows=rows)

        ws = torch.zeros((L.n_slots, 3), dtype=torch.float64)
        for it, jt, diag in _plan_tile_pairs(L, rank, 2):
            rows, a = pull(it, jt)
            ws[rows[0]:rows[1]] += torch.from_numpy(a)
            if not diag:
                rows, a = pull(jt, it)
                ws[rows[0]:rows[1]] += torch.from_numpy(a)
        mine = torch.zeros((L.shard, 3), dtype=torch.float64)
        reduce_scatter_shards(ws, mine, L, rank)
        full = gather_rows(mine[:L.count(rank)], L, rank)
        if rank == 0:
            out.put((full.numpy(), oracle.compute_nbody_forces_cpu(pos, mass, H, 1.0, "plummer")))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_symmetric_partial_sums_reduce_scatter_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sym_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


# ---------------------------------------------------------------------------
# staging-free row gather (engine.gather_rows_padded): what the fields API (gather='all' / 'root') and the
# root-only state gather of run_nbody_gpu's writer use
# ---------------------------------------------------------------------------
from nbody_streams_b200.engine import gather_rows_padded  # noqa: E402


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for n in (700, 513, 256, 3):        # ragged shards, a rank with one tile of padding only (n <= 256), tiny n
            L = ShardLayout(n, world)
            full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3) + 0.5
            rows = L.rows(rank)
            padded = torch.full((L.shard, 3), -1.0, dtype=torch.float64)   # padding rows must never show up
            padded[: L.count(rank)] = full[rows]
            got = gather_rows_padded(padded, L, rank)
            ok &= bool(torch.equal(got, full))
            local = padded[: L.count(rank)]
            got_root = gather_rows_padded(padded, L, rank, root_only=True, local=local)
            ok &= bool(torch.equal(got_root, full)) if rank == 0 else (got_root is local)
            vec = torch.full((L.shard,), -1.0, dtype=torch.float64)        # 1-D rows (phi)
            vec[: L.count(rank)] = full[rows, 0]
            ok &= bool(torch.equal(gather_rows_padded(vec, L, rank), full[:, 0]))
        flag = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(bool(flag.item() == 1.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_rows_padded_world2_all_and_root_only():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res is True


def test_gather_rows_padded_world1_drops_the_padding():
    L = ShardLayout(300, 1)
    padded = torch.zeros((L.shard, 3))
    assert L.shard == 512 and gather_rows_padded(padded, L, 0).shape == (300, 3)
