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
