"""Device-resident direct-summation engine: sharded state + the per-step launch sequence.

New layer (the reference is single-GPU only).  One process
per GPU; ``torch.distributed`` (NCCL over NVLink) is the plumbing.  i-particles
are sharded contiguously and tile-aligned across ranks; every rank keeps the
whole packed j-buffer (20 B/particle f32, 40 B f64).  Per step:

    kick+drift+repack own shard (1 fused kernel)
      -> all-gather of the packed shards on a side stream, in place
      -> overlap mode : sweep own j-tiles now, the remote tiles once the gather lands
         overlap=False: one sweep over all tiles in global j order after the gather
      -> finish (fixed-order fix-up, x G)
      -> half kick

With world size 1 the same code runs without any collective, which is the path
``run_nbody_gpu`` uses on a single GPU.

Host-side pieces that do not need a GPU (``ShardLayout``, ``allgather_shards``,
``gather_rows``) are plain functions so the N>1 logic is testable with the
``gloo`` backend on CPU tensors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import DT_F32, DT_F64, MODE_BOTH, MODE_FORCE, MODE_POT, PREC_F64

TILE_J = 256  # must equal nb_tile_j(); asserted when the library is loaded


# ---------------------------------------------------------------------------
# shard layout (pure Python)
# ---------------------------------------------------------------------------
@dataclass(frozen=True)
class ShardLayout:
    """Contiguous, tile-aligned split of N particles over `world` ranks.

    Every rank owns `tiles_per_rank` tiles = `shard` particle slots; only the
    tail of the global range is padding, so a particle's slot index equals its
    original index."""

    N: int
    world: int
    tile: int = TILE_J

    @property
    def tiles_total(self) -> int:
        return (self.N + self.tile - 1) // self.tile

    @property
    def tiles_per_rank(self) -> int:
        return max(1, (self.tiles_total + self.world - 1) // self.world)

    @property
    def shard(self) -> int:
        return self.tiles_per_rank * self.tile

    @property
    def n_slots(self) -> int:
        return self.shard * self.world

    def first(self, rank: int) -> int:
        return rank * self.shard

    def count(self, rank: int) -> int:
        return int(min(max(self.N - self.first(rank), 0), self.shard))

    def tile_range(self, rank: int) -> tuple[int, int]:
        return rank * self.tiles_per_rank, (rank + 1) * self.tiles_per_rank

    def rows(self, rank: int) -> slice:
        f = self.first(rank)
        return slice(f, f + self.count(rank))


def sym_plan(layout: ShardLayout, rank: int, tpb: int) -> dict:
    """Tile ranges one rank sweeps with the pair-symmetric kernel so that, over all ranks, every unordered
    pair of particles is evaluated exactly once (``nb_sym_sweep`` arguments, see include/nbody_sm100.h).

    Shards form a ring: rank r takes the pairs inside its own shard (``tri``) and the full rectangles
    (own shard) x (shard r+k), k = 1..(G-1)//2, as up to two contiguous tile ranges A and B (wrap-around).
    For an even ring the opposite shard o = r + G/2 is shared: the lower rank sweeps its FIRST half of
    i-blocks against all of shard o (range C restricted to i-blocks [0, h)), the upper rank sweeps all of
    its i-blocks against the tiles of the SECOND half of the lower rank's shard (range C, all i-blocks).
    All-padding tiles (>= layout.tiles_total) are never swept.  Pure Python: tested on CPU."""
    G, T, R = layout.world, layout.tiles_per_rank, layout.tiles_total

    def shard(s):
        b = min(s * T, R)
        return b, min((s + 1) * T, R)

    own = shard(rank)
    n_ib = (own[1] - own[0] + tpb - 1) // tpb
    plan = {"own": own, "n_ib": n_ib, "A": (0, 0), "B": (0, 0), "C": (0, 0), "ibC": (0, 0)}
    K = (G - 1) // 2
    if K > 0:
        first, last = rank + 1, rank + K                       # shards first..last (mod G), contiguous on the ring
        if last < G:
            plan["A"] = (shard(first)[0], shard(last)[1])
        elif first >= G:
            plan["A"] = (shard(first - G)[0], shard(last - G)[1])
        else:
            plan["A"] = (shard(first)[0], shard(G - 1)[1])
            plan["B"] = (shard(0)[0], shard(last - G)[1])
    if G % 2 == 0 and G > 1:
        half = G // 2
        if rank < half:
            plan["C"] = shard(rank + half)
            plan["ibC"] = (0, n_ib // 2)
        else:
            lo = shard(rank - half)
            n_ib_lo = (lo[1] - lo[0] + tpb - 1) // tpb
            plan["C"] = (min(lo[0] + (n_ib_lo // 2) * tpb, lo[1]), lo[1])
            plan["ibC"] = (0, n_ib)
    for k in ("A", "B", "C"):
        if plan[k][1] <= plan[k][0]:
            plan[k] = (0, 0)
    return plan


def _dist():
    import torch.distributed as dist

    return dist


def group_info(group=None) -> tuple[int, int]:
    """(rank, world) of the given process group; (0, 1) when torch.distributed is not initialised."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allgather_shards(buf, layout: ShardLayout, rank: int, shard_bytes: int, group=None):
    """In-place all-gather of the equal-size shards of a flat uint8 buffer
    (shard r occupies bytes [r*S, (r+1)*S), S = shard_bytes = whole tiles incl. their metadata)."""
    if layout.world == 1:
        return None
    S = int(shard_bytes)
    out = buf[: S * layout.world]
    inp = buf[rank * S:(rank + 1) * S]
    return _dist().all_gather_into_tensor(out, inp, group=group)


def reduce_scatter_shards(ws, out, layout: ShardLayout, rank: int, group=None):
    """out (shard rows) = sum over ranks of rows [rank*shard, (rank+1)*shard) of ws (n_slots rows).
    NCCL: one reduce-scatter; backends without it (gloo, CPU tests): all-reduce + slice."""
    dist = _dist()
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, ws, group=group)
    else:
        dist.all_reduce(ws, group=group)
        out.copy_(ws[rank * layout.shard:(rank + 1) * layout.shard])
    return out


def gather_rows(local, layout: ShardLayout, rank: int, group=None):
    """All ranks receive the (N, k) concatenation of every rank's real rows."""
    import torch

    if layout.world == 1:
        return local
    k = local.shape[1] if local.dim() > 1 else 1
    padded = torch.zeros((layout.shard, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local.reshape(local.shape[0], k)
    out = torch.empty((layout.n_slots, k), dtype=local.dtype, device=local.device)
    _dist().all_gather_into_tensor(out, padded, group=group)
    out = out[: layout.N]
    return out if local.dim() > 1 else out.reshape(-1)


def gather_rows_padded(padded, layout: ShardLayout, rank: int, group=None, root_only: bool = False, local=None):
    """Gather shard-sized (padded) row blocks WITHOUT a staging copy: `padded` has layout.shard rows of which the
    first layout.count(rank) are real.  All ranks (or only rank 0 with root_only; the others get `local` back)
    receive the first N rows of the concatenation."""
    import torch

    if layout.world == 1:
        return padded[: layout.N]
    dist = _dist()
    shape = (layout.n_slots,) + tuple(padded.shape[1:])
    if root_only:
        root = dist.get_global_rank(group, 0) if group is not None else 0
        if rank == 0:
            out = torch.empty(shape, dtype=padded.dtype, device=padded.device)
            dist.gather(padded, list(out.split(layout.shard)), dst=root, group=group)
            return out[: layout.N]
        dist.gather(padded, None, dst=root, group=group)
        return local
    out = torch.empty(shape, dtype=padded.dtype, device=padded.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[: layout.N]


# ---------------------------------------------------------------------------
# the engine
# ---------------------------------------------------------------------------
class DirectEngine:
    """Sharded device state and kernels for repeated force evaluations / KDK steps.

    Parameters
    ----------
    N : total particle count.
    mass, softening : scalar or (N,) host arrays (all particles).
    prec : library precision code; kernel_id : 0..4; G, eps2 : floats.
    device : torch CUDA device of this rank.
    group : torch.distributed process group (None = default group / single process).
    exchange : "nccl" (default) — in-place NCCL all-gather of the packed shards each step;
               "p2p" — the packed buffer lives in torch symmetric memory and the fused
               kick+drift kernel stores the refreshed records straight into every rank's buffer over
               NVLink (NVSwitch multicast when available, else one store per peer), bracketed by
               symmetric-memory barriers: no separate collective runs.
    algorithm : None/'auto' (pair-symmetric sweep from nb_sym_min_n() particles on), 'symmetric', 'onesided'.
              With the symmetric sweep on several ranks each rank evaluates 1/world of the unordered pairs
              (ring schedule, `sym_plan`) into partial sums for ALL particles it touched, and a
              reduce-scatter (NCCL, double) delivers every shard's sums to its owner.
    overlap : sweep the local j-shard while the all-gather is in flight (default);
              False sweeps all tiles in global (j-ascending) order after the gather.
              Either way results are bitwise reproducible for a given rank count and agree
              with the 1-GPU result to rounding (the stream-K pieces are cut at other tiles).
    """

    def __init__(self, N, mass, softening, prec, kernel_id, G, eps2, device, group=None, overlap=True,
                 exchange="nccl", algorithm=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.lib = _lib.load()
        assert self.lib.nb_tile_j() == TILE_J
        self.N, self.prec, self.kernel_id = int(N), int(prec), int(kernel_id)
        self.G, self.eps2 = float(G), float(eps2)
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = group_info(group)
        self.overlap = bool(overlap) and self.world > 1
        self.layout = ShardLayout(self.N, self.world)
        from .fields import use_symmetric

        # pair-symmetric sweep (every unordered pair once, partial sums of BOTH particles; large N) or the
        # one-sided sweep (bitwise reproducible, finer work units; small N)
        self.symmetric = use_symmetric(self.N, algorithm, self.world)
        self._sym_ws = None
        self._sym_out = None
        self.first = self.layout.first(self.rank)
        self.n_local = self.layout.count(self.rank)
        self.shard_bytes = int(self.lib.nb_packed_bytes(self.prec, self.layout.shard))
        self.dtype = torch.float64 if prec == PREC_F64 else torch.float32
        self.acc_dt = DT_F64 if prec == PREC_F64 else DT_F32

        self.exchange = exchange if self.world > 1 else "nccl"
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError(f"exchange must be 'nccl' or 'p2p', got {exchange!r}")
        self._symm = None
        with torch.cuda.device(self.device):
            if self.exchange == "p2p":
                import torch.distributed._symmetric_memory as symm_mem

                self.packed = symm_mem.empty(self.shard_bytes * self.world, dtype=torch.uint8, device=self.device)
                self.packed.zero_()
                self._symm = symm_mem.rendezvous(self.packed, group if group is not None else _dist().group.WORLD)
                ptrs = [int(p) for p in self._symm.buffer_ptrs]
                self._peer_ptrs = (C.c_void_p * self.world)(*ptrs)
                self._mc_ptr = int(getattr(self._symm, "multicast_ptr", 0) or 0)
                self.overlap = True     # the local sweep always runs between the push and its barrier
            else:
                self.packed = torch.zeros(self.shard_bytes * self.world, dtype=torch.uint8, device=self.device)
            self._grids = {}
            self._ws = None
            self._ws_bytes = 0
            self.pos = torch.zeros((self.n_local, 3), dtype=torch.float64, device=self.device)
            self.vel = torch.zeros((self.n_local, 3), dtype=torch.float64, device=self.device)
            self.acc = torch.zeros((self.n_local, 3), dtype=self.dtype, device=self.device)
            self.phi = torch.zeros((self.n_local,), dtype=self.dtype, device=self.device)
            self._e_part = torch.zeros(self.lib.nb_energy_workspace_bytes(max(self.n_local, 1)) // 8,
                                       dtype=torch.float64, device=self.device)
            self._e_out = torch.zeros(8, dtype=torch.float64, device=self.device)
            self.comm_stream = torch.cuda.Stream(self.device) if self.world > 1 else None

        # static per-particle data (host -> device once); masses in the KERNEL dtype semantics
        self._mass_host = None if np.isscalar(mass) else np.ascontiguousarray(mass)
        self._mass_scalar = float(mass) if np.isscalar(mass) else 0.0
        self._h_host = None if np.isscalar(softening) else np.ascontiguousarray(softening)
        self._h_scalar = float(softening) if np.isscalar(softening) else 0.0
        self.mass_local = self._local_mass_tensor()
        self.launches = 0   # kernels launched by this engine (bench.py gpu_launches)
        self.sweep_events = None   # set to [] to record (start, end) CUDA events around every sweep launch

    # -- helpers ------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _sweep(self, mode, jt0, jt1, jt2_0, jt2_1, ws, slot, grid):
        """Launch one nb_sweep on the current stream (optionally bracketed by CUDA events)."""
        torch = self.torch
        ev = None
        if self.sweep_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(self.device))
        _lib.check(self.lib.nb_sweep(self.prec, self.kernel_id, mode, self.packed.data_ptr(), self.N,
                                     jt0, jt1, jt2_0, jt2_1, self.first, self.n_local, self.eps2,
                                     ws.data_ptr(), self._ws_bytes, slot, grid, self._stream()))
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            self.sweep_events.append((ev[0], ev[1], (jt1 - jt0) + (jt2_1 - jt2_0)))
        self.launches += 1

    def _local_mass_tensor(self):
        torch = self.torch
        if self._mass_host is None:
            return torch.full((max(self.n_local, 1),), self._mass_scalar, dtype=torch.float64, device=self.device)[: self.n_local]
        m = np.asarray(self._mass_host[self.layout.rows(self.rank)], dtype=np.float64)
        return torch.from_numpy(np.ascontiguousarray(m)).to(self.device)

    def _sym_grid(self, mode):
        g = self._grids.get(("sym", mode))
        if g is None:
            out = C.c_int(0)
            with self.torch.cuda.device(self.device):
                _lib.check(self.lib.nb_sym_grid(self.prec, self.kernel_id, mode, C.byref(out)))
            g = self._grids[("sym", mode)] = int(out.value)
        return g

    def _sym_buffers(self):
        """Workspace of 4 doubles per particle slot (all shards) + this rank's reduce-scatter output."""
        torch = self.torch
        if self._sym_ws is None:
            if self.exchange == "p2p":
                # the workspace lives in symmetric memory: the owner of a shard reads its rows from every rank's
                # workspace over NVLink in the fused finish kernel (no reduce-scatter)
                import torch.distributed._symmetric_memory as symm_mem

                self._sym_ws = symm_mem.empty((self.layout.n_slots, 4), dtype=torch.float64, device=self.device)
                self._sym_ws.zero_()
                self._ws_symm = symm_mem.rendezvous(self._sym_ws, self.group if self.group is not None
                                                    else _dist().group.WORLD)
                self._ws_peer_ptrs = (C.c_void_p * self.world)(*[int(p) for p in self._ws_symm.buffer_ptrs])
            else:
                self._sym_ws = torch.zeros((self.layout.n_slots, 4), dtype=torch.float64, device=self.device)
                if self.world > 1:
                    self._sym_out = torch.zeros((self.layout.shard, 4), dtype=torch.float64, device=self.device)
        return self._sym_ws, self._sym_out

    def _sym_sweep(self, mode, plan, tri, rect, zero_first, grid, ws):
        """One nb_sym_sweep launch: the triangle of the own shard and/or the rectangle ranges of `plan`."""
        torch = self.torch
        A, B, Cr, ibC = (plan["A"], plan["B"], plan["C"], plan["ibC"]) if rect else ((0, 0),) * 4
        own = plan["own"]
        ev = None
        if self.sweep_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(self.device))
        _lib.check(self.lib.nb_sym_sweep(self.prec, self.kernel_id, mode, self.packed.data_ptr(), self.N,
                                         own[0], own[1], 1 if tri else 0, A[0], A[1], B[0], B[1], Cr[0], Cr[1],
                                         ibC[0], ibC[1], self.eps2, ws.data_ptr(), ws.numel() * 8,
                                         self.layout.n_slots, 1 if zero_first else 0, grid, self._stream()))
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            self.sweep_events.append((ev[0], ev[1], "tri" if tri and not rect else ("rect" if rect and not tri else "all")))
        self.launches += 1 + (1 if zero_first else 0)

    def _compute_sym(self, mode, positions_changed):
        torch, lib = self.torch, self.lib
        grid = self._sym_grid(mode)
        ws, out = self._sym_buffers()
        plan = sym_plan(self.layout, self.rank, lib.nb_sym_block_tiles(self.prec))
        has_own = plan["own"][1] > plan["own"][0]
        has_rect = has_own and any(plan[k][1] > plan[k][0] for k in ("A", "B", "C"))
        p2p = self.exchange == "p2p"
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if self.world > 1 and positions_changed and not p2p:
                self.comm_stream.wait_stream(cur)          # shard records are written on `cur`
                with torch.cuda.stream(self.comm_stream):
                    allgather_shards(self.packed, self.layout, self.rank, self.shard_bytes, self.group)

            if p2p and getattr(self, "_p2p_reads_pending", False) and not positions_changed:
                # two evaluations in a row without a kick_drift (whose barrier orders the ranks): peers may still be
                # reading this rank's workspace from the previous evaluation
                self._ws_symm.barrier(channel=3)
            self._p2p_reads_pending = False

            def wait_exchange():
                if self.world == 1:
                    return
                if p2p:
                    if positions_changed:
                        self._symm.barrier(channel=1)      # every rank's pushed records have landed
                else:
                    cur.wait_stream(self.comm_stream)

            if not has_own:
                ws.zero_()
                wait_exchange()
            elif self.overlap:
                self._sym_sweep(mode, plan, True, False, True, grid, ws)      # own shard: no dependency
                wait_exchange()
                if has_rect:
                    self._sym_sweep(mode, plan, False, True, False, grid, ws)
            else:
                wait_exchange()
                self._sym_sweep(mode, plan, True, has_rect, True, grid, ws)
            if p2p:
                # fused reduce + finish over peer memory: barrier (every rank's sweeps have completed), then each
                # owner adds the rows of its shard from all workspaces.  The workspaces are cleared again only after
                # the next step's first barrier (kick_drift), i.e. after every rank has finished reading.
                self._ws_symm.barrier(channel=2)
                if self.n_local > 0:
                    _lib.check(lib.nb_sym_finish_peers(self.prec, mode, self._ws_peer_ptrs, self.world, self.first,
                                                       self.n_local, self.G,
                                                       self.acc.data_ptr() if mode & MODE_FORCE else None,
                                                       self.phi.data_ptr() if mode & MODE_POT else None,
                                                       self._stream()))
                    self.launches += 1
                self._p2p_reads_pending = True
                return
            src, first = ws, self.first
            if self.world > 1:
                reduce_scatter_shards(ws, out, self.layout, self.rank, self.group)   # my shard's sums from every rank
                src, first = out, 0
            if self.n_local > 0:
                _lib.check(lib.nb_sym_finish(self.prec, mode, src.data_ptr(), first, self.n_local, self.G,
                                             self.acc.data_ptr() if mode & MODE_FORCE else None,
                                             self.phi.data_ptr() if mode & MODE_POT else None, self._stream()))
                self.launches += 1

    def _grid(self, mode):
        g = self._grids.get(mode)
        if g is None:
            out = C.c_int(0)
            with self.torch.cuda.device(self.device):
                _lib.check(self.lib.nb_sweep_grid(self.prec, self.kernel_id, mode, C.byref(out)))
            g = self._grids[mode] = int(out.value)
        return g

    def _workspace(self, grid):
        need = self.lib.nb_workspace_bytes(self.prec, max(self.n_local, 1), grid, 2)
        if self._ws is None or self._ws_bytes < need:
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
            self._ws_bytes = need
        return self._ws

    # -- state --------------------------------------------------------------
    def load_state(self, phase_space: np.ndarray) -> None:
        """Upload this rank's rows of the (N,6) float64 state and pack ALL particles
        (every rank has the full initial positions, so no collective is needed here)."""
        torch = self.torch
        xv = np.asarray(phase_space, dtype=np.float64)
        rows = self.layout.rows(self.rank)
        with torch.cuda.device(self.device):
            self.pos.copy_(torch.from_numpy(np.ascontiguousarray(xv[rows, :3])))
            self.vel.copy_(torch.from_numpy(np.ascontiguousarray(xv[rows, 3:6])))
            pos_all = torch.from_numpy(np.ascontiguousarray(xv[:, :3])).to(self.device)
            self.pack_all(pos_all)

    def pack_all(self, pos_all) -> None:
        """(Re)build every packed record from a device (N,3) position array."""
        torch = self.torch
        dt_mass = self.torch.float64
        mass_t = h_t = None
        if self._mass_host is not None:
            mass_t = torch.from_numpy(np.asarray(self._mass_host)).to(self.device)
            if mass_t.dtype not in (torch.float32, torch.float64):
                mass_t = mass_t.to(dt_mass)
        if self._h_host is not None:
            h_t = torch.from_numpy(np.asarray(self._h_host)).to(self.device)
            if h_t.dtype not in (torch.float32, torch.float64):
                h_t = h_t.to(dt_mass)
        code = lambda t: DT_F64 if t.dtype == torch.float64 else DT_F32  # noqa: E731
        pos_all = pos_all.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nb_pack(
                self.prec, self.kernel_id, self.packed.data_ptr(), 0,
                pos_all.data_ptr(), code(pos_all),
                mass_t.data_ptr() if mass_t is not None else None, code(mass_t) if mass_t is not None else 0,
                self._mass_scalar,
                h_t.data_ptr() if h_t is not None else None, code(h_t) if h_t is not None else 0,
                self._h_scalar,
                self.N, self.layout.n_slots, self.eps2, self._stream()))
            self.launches += 1
            # keep operand tensors alive until the kernel has consumed them
            torch.cuda.current_stream(self.device).synchronize()

    def state_device(self, root_only: bool = False):
        """(N,6) float64 device tensor of the whole system: on every rank, or (root_only) on rank 0 only — the
        other ranks then get None and send their rows without receiving anything.  No staging copy: the rows are
        gathered from a persistent shard-sized buffer."""
        torch = self.torch
        if self.world == 1:
            return torch.cat([self.pos, self.vel], dim=1)
        if getattr(self, "_xv_pad", None) is None:
            self._xv_pad = torch.zeros((self.layout.shard, 6), dtype=torch.float64, device=self.device)
        self._xv_pad[: self.n_local, :3] = self.pos
        self._xv_pad[: self.n_local, 3:] = self.vel
        full = gather_rows_padded(self._xv_pad, self.layout, self.rank, self.group, root_only=root_only, local=None)
        if root_only and self.rank != 0:
            return None
        return full

    def state_numpy(self) -> np.ndarray:
        """(N,6) float64 on the host, identical on all ranks."""
        return self.state_device().cpu().numpy()

    # -- force evaluation -----------------------------------------------------
    def compute(self, mode: int = MODE_FORCE, positions_changed: bool = True) -> None:
        """Sweep + finish into self.acc / self.phi for this rank's i-particles.
        Assumes the packed records of this rank's shard are current; all-gathers them
        first when world > 1 and `positions_changed`."""
        if self.symmetric:
            return self._compute_sym(mode, positions_changed)
        torch, lib = self.torch, self.lib
        grid = self._grid(mode)
        ws = self._workspace(grid)
        t0, t1 = self.layout.tile_range(self.rank)
        n_tiles = self.layout.tiles_per_rank * self.world
        n_jt = (C.c_int * 4)(0, 0, 0, 0)
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            p2p = self.exchange == "p2p"
            if self.world > 1 and positions_changed and not p2p:
                self.comm_stream.wait_stream(cur)          # shard records are written on `cur`
                with torch.cuda.stream(self.comm_stream):
                    allgather_shards(self.packed, self.layout, self.rank, self.shard_bytes, self.group)
            if self.n_local > 0:
                if self.overlap:
                    # slot 0: own shard (no dependency on the collective)
                    self._sweep(mode, t0, t1, 0, 0, ws, 0, grid)
                    n_jt[0] = t1 - t0
                    if p2p:
                        if positions_changed:
                            self._symm.barrier(channel=1)   # every rank's pushed records have landed
                    else:
                        cur.wait_stream(self.comm_stream)
                    # slot 1: every other rank's tiles, as two ranges around mine
                    self._sweep(mode, 0, t0, t1, n_tiles, ws, 1, grid)
                    n_jt[1] = n_tiles - (t1 - t0)
                    n_sweeps = 2
                else:
                    if self.world > 1:
                        cur.wait_stream(self.comm_stream)
                    self._sweep(mode, 0, n_tiles, 0, 0, ws, 0, grid)
                    n_jt[0] = n_tiles
                    n_sweeps = 1
                _lib.check(lib.nb_finish(self.prec, self.kernel_id, mode, self.n_local, grid, n_sweeps, n_jt,
                                         ws.data_ptr(), self.G,
                                         self.acc.data_ptr() if mode & MODE_FORCE else None,
                                         self.phi.data_ptr() if mode & MODE_POT else None, self._stream()))
                self.launches += 1
            elif self.world > 1:
                if p2p:
                    if positions_changed:
                        self._symm.barrier(channel=1)
                else:
                    cur.wait_stream(self.comm_stream)

    # -- KDK pieces -----------------------------------------------------------
    def kick_drift(self, dt: float, acc_ext=None) -> None:
        """vel += acc*dt/2 ; pos += vel*dt ; refresh own packed x,y,z (one fused kernel)."""
        if self.exchange == "p2p":
            with self.torch.cuda.device(self.device):
                self._symm.barrier(channel=0)       # every rank is done reading the previous records
                if self.n_local > 0:
                    _lib.check(self.lib.nb_kick_drift_push(
                        self.prec, self.pos.data_ptr(), self.vel.data_ptr(), self.acc.data_ptr(), self.acc_dt,
                        acc_ext.data_ptr() if acc_ext is not None else None,
                        self._peer_ptrs, self.world, self._mc_ptr if self._mc_ptr else None,
                        self.first, self.n_local, 0.5 * dt, dt, self._stream()))
                    self.launches += 1
            return
        if self.n_local == 0:
            return
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.nb_kick_drift_pack(
                self.prec, self.pos.data_ptr(), self.vel.data_ptr(), self.acc.data_ptr(), self.acc_dt,
                acc_ext.data_ptr() if acc_ext is not None else None,
                self.packed.data_ptr(), self.first, self.n_local, 0.5 * dt, dt, self._stream()))
        self.launches += 1

    def kick(self, dt: float, acc_ext=None) -> None:
        """vel += acc*dt/2."""
        if self.n_local == 0:
            return
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.nb_kick(self.vel.data_ptr(), self.acc.data_ptr(), self.acc_dt,
                                        acc_ext.data_ptr() if acc_ext is not None else None,
                                        self.n_local, 0.5 * dt, self._stream()))
        self.launches += 1

    def capture_step(self, mode: int, dt: float):
        """Capture one KDK step (kick+drift+repack, sweep, finish, kick) into a CUDA graph and return a
        callable that replays it: one graph launch per step instead of four kernel launches through
        Python/ctypes — what matters for N <~ 1e5 where a step is shorter than its launch overhead.
        Single rank, no hooks (their accelerations change every step)."""
        assert self.world == 1
        torch = self.torch
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            if self.symmetric:
                self._sym_grid(mode)
                self._sym_buffers()
            else:
                # allocate outside the capture, sized for every mode: a later compute() with another mode must not
                # reallocate (and thereby free) the buffer whose address is baked into the graph
                self._workspace(max(self._grid(m) for m in (MODE_FORCE, MODE_POT, MODE_BOTH)))
            graph = torch.cuda.CUDAGraph()
            saved_events, self.sweep_events = self.sweep_events, None
            with torch.cuda.graph(graph, stream=side):
                self.kick_drift(dt)
                self.compute(mode)
                self.kick(dt)
            self.sweep_events = saved_events
        per_step = 5 if self.symmetric else 4        # the symmetric sweep is preceded by a memset node

        def replay():
            graph.replay()
            self.launches += per_step

        self.launches -= per_step      # the capture pass recorded the launches without running them
        replay.graph = graph
        return replay

    def energy(self, with_phi: bool = True, with_acc: bool = False) -> np.ndarray:
        """[KE, PE, px, py, pz, fx, fy, fz] summed over all ranks (float64, host)."""
        torch = self.torch
        with torch.cuda.device(self.device):
            if self.n_local > 0:
                _lib.check(self.lib.nb_energy(
                    self.vel.data_ptr(), self.mass_local.data_ptr(), DT_F64,
                    self.phi.data_ptr() if with_phi else None, self.acc_dt,
                    self.acc.data_ptr() if with_acc else None, self.acc_dt,
                    self.n_local, self._e_part.data_ptr(), self._e_out.data_ptr(), self._stream()))
                self.launches += 2
            else:
                self._e_out.zero_()
            out = self._e_out.clone()
            if self.world > 1:
                _dist().all_reduce(out, group=self.group)
        return out.cpu().numpy()


__all__ = ["ShardLayout", "DirectEngine", "sym_plan", "reduce_scatter_shards", "gather_rows_padded", "allgather_shards", "gather_rows", "group_info", "TILE_J"]
