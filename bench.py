#!/usr/bin/env python3
"""bench.py — the reference's headline metric on B200: pairwise interactions/s (and ms/step) for the
direct-summation path at N = 1M (BASELINE.json), 1/2/4/8 GPUs.

    python bench.py --gpus N --steps K --warmup W            # this package (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # N > 1: one rank per GPU over NCCL

Workload (config 3 of BASELINE.json): make_plummer_sphere(1_000_000, M_total=1e9, a=1.0, seed=42069),
scalar softening 0.01, precision 'float32_kahan', kernel 'plummer', force AND potential.
One "step" = one KDK leapfrog step of the whole system = one fused kick+drift+repack, the all-gather of
the j-records (N > 1), ONE force+potential sweep over all pairs, the finish kernel (reduce-scatter of the
partial sums first when N > 1) and the closing half kick.
Interactions per step = N^2 (reference convention, fields.py:1317: every ordered pair counts).  The default
sweep is the pair-symmetric one: it evaluates each unordered pair once and applies the result to both
particles, i.e. N^2/2 kernel evaluations deliver the N^2 interactions (--algorithm onesided times the
round-1 sweep that evaluates all N^2).  Total work is fixed as GPUs are added ("scaling": "strong").

The JSON line carries, besides the driver contract keys:
  roofline      FP32 CUDA-core roofline of the dominant kernel (the pair sweep): 20 flop/interaction
                (BASELINE.json) x interactions per launch / CUDA-event duration of the launch, against the
                FFMA peak measured in the same process by nb_ubench (MEASURED_PEAKS.json has no FP32 figure);
  cpu_baseline  the oracle port of the reference's Numba path timed on this box's host cores on a bounded
                row sample of the same workload (N = 1 only);
  cpu_baseline_numba  the same sample through oracle/numba_port.py — the reference's CPU path restated in its own
                technology (Numba @njit(parallel=True); bit-identical to and as fast as the reference's functions
                where both can run) — in a subprocess; {"unavailable": ...} when Numba is missing;
  e2e           the same metric through the public host-array API (compute_nbody_forces_potential_gpu with
                NumPy in / NumPy out): H2D of positions and D2H of acc+phi inside the timed region;
  parity        outside the timed region: acc and phi of rank 0's first rows after the last step against the
                float64 oracle on the same (float32-rounded) positions — relative RMS, at every GPU count;
  ref_gpu_kernel  (1 GPU) the reference's own CUDA kernel, compiled for sm_100 with its flags into baseline/_ref
                by tools/make_ref_gpu_baseline.py, timed on the same GPU next to this package's sweep.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOPS_PER_INTERACTION = 20.0          # BASELINE.json convention
print_json = None
FP32_PEAK_NOMINAL_TF = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4: SMs x lanes x 2 x max SM clock
TRAFFIC_SYM_N1M = 391907328   # dram read + write bytes per launch of nb::sym_f32_kernel at N=1M (profiles/r02_ncu_sym_f32_bench_n1m.md)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--kernel", default="plummer")
    p.add_argument("--precision", default="float32_kahan")
    p.add_argument("--softening", type=float, default=0.01)
    p.add_argument("--dt", type=float, default=1e-4)
    p.add_argument("--no-overlap", action="store_true", help="no local-shard sweep while the exchange is in flight (N>1)")
    p.add_argument("--algorithm", default="auto", choices=["auto", "symmetric", "onesided"],
                   help="pair-symmetric sweep (default from nb_sym_min_n() particles on) or the one-sided sweep")
    p.add_argument("--parity-rows", type=int, default=2048, help="rows of rank 0 checked against the oracle (0 = skip)")
    p.add_argument("--skip-ref-gpu", action="store_true")
    p.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                   help="N>1: NCCL all-gather of the j-records, or the exchange fused into the drift kernel (NVLink stores)")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-numba", action="store_true", help="no cpu_baseline_numba key (Numba restatement in a subprocess)")
    p.add_argument("--skip-e2e", action="store_true")
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def config_dict(args, N) -> dict:
    """Identical for both arms (the driver compares the dicts)."""
    tag = "BASELINE config 3" if (N == 1_000_000 and args.precision == "float32_kahan") else (
        "BASELINE config 4 size" if (N == 262_144 and args.precision == "float64") else "non-default workload")
    return {"workload": f"Plummer N={N} direct {args.precision} force+potential, kernel={args.kernel}, "
                        f"softening={args.softening} ({tag})",
            "l2": "GPU arm: 256 MiB memset between steps (inside the timed region); CPU arm: not applicable"}


def workload(args):
    import nbody_streams_b200 as nb

    xv, m = nb.make_plummer_sphere(args.n, M_total=1e9, a=1.0, seed=42069)
    return xv, m


def cpu_sample(args, xv, m, seconds: float) -> dict:
    """Oracle port (fast build, all host threads) on rows [0, R) of the workload vs all N."""
    import oracle

    oracle.build()
    oracle.set_num_threads(len(os.sched_getaffinity(0)), fast=True)   # all host cores (torchrun sets OMP_NUM_THREADS=1)
    N = xv.shape[0]
    f64 = args.precision == "float64"
    pos = xv[:, :3] if f64 else xv[:, :3].astype(np.float32)
    kw = dict(kernel=args.kernel, precision="float64" if f64 else "float32", fast=True)

    def run(rows):
        t0 = time.perf_counter()
        oracle.compute_nbody_forces_cpu(pos, float(m[0]), args.softening, rows=(0, rows), **kw)
        oracle.compute_nbody_potential_cpu(pos, float(m[0]), args.softening, rows=(0, rows), **kw)
        return time.perf_counter() - t0

    run(min(N, 256))                                    # warms the OpenMP pool and the page cache
    probe_rows = min(N, 2048)
    t = run(probe_rows)                                 # ~0.4 s: rate estimate for sizing the sample
    rate = probe_rows * N / max(t, 1e-9)
    rows = probe_rows
    for _ in range(3):                                  # the short probes underestimate the rate: re-size until
        rows = int(min(N, max(probe_rows, seconds * rate / N)))   # the sample takes about `seconds`
        rows = max(1, (rows // 64) * 64) if rows >= 64 else rows
        t = run(rows)
        rate = rows * N / max(t, 1e-9)
        if t >= 0.7 * seconds or rows >= N:
            break
    return {"value": rows * N / t, "unit": "interactions/s", "cores": oracle.num_threads(True), "kind": "port",
            "sample": f"rows [0,{rows}) of the N={N} Plummer workload against all N, force pass + potential pass "
                      f"({args.kernel}, " + ("float64" if f64 else "float32 pair math / float64 sums") +
                      f" as the reference CPU path), {t:.1f} s",
            "seconds": t, "rows": rows}


def numba_sample(args, seconds: float) -> dict:
    """The reference's CPU path in its own technology: oracle/numba_port.py (a Numba @njit(parallel=True) restatement,
    bit-identical to and as fast as the reference's functions in the build container) on a bounded row
    sample, in a SUBPROCESS — a missing or misbehaving Numba on the box can only cost this one key."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "NUMBA_NUM_THREADS")}
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/nbody_b200_numba_cache")
    cmd = [sys.executable, "-m", "oracle.numba_port", "--n", str(args.n), "--kernel", args.kernel,
           "--precision", "float64" if args.precision == "float64" else "float32",
           "--softening", str(args.softening), "--seconds", str(seconds),
           "--threads", str(len(os.sched_getaffinity(0)))]
    try:
        r = subprocess.run(cmd, cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=60 + 6 * seconds)
        if r.returncode != 0:
            return {"unavailable": (r.stderr.strip().splitlines() or ["non-zero exit"])[-1][:200]}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001 - optional key
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    xv, m = workload(args)
    res = []
    per_step = max(1.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r = cpu_sample(args, xv, m, per_step)
        if i >= args.warmup:
            res.append(r)
    val = float(np.median([r["value"] for r in res]))
    N = xv.shape[0]
    line = {
        "impl": "reference", "metric": "pairwise interactions/s", "value": val, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.median([r["seconds"] for r in res])) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.precision == "float64" else "f32", "data": "synthetic",
        "config": config_dict(args, N),
        "details": {"note": "reference CPU algorithm (C/OpenMP port of the Numba path; the reference is Python and "
                            "cannot travel to the GPU box); a step of this arm is a bounded row sample of the workload "
                            "(cpu_baseline.sample), ms_per_step is the measured time of one sample, value the rate it "
                            "sustained, ms_per_full_evaluation that rate extrapolated to all N^2 interactions",
                    "ms_per_full_evaluation": N * N / val * 1e3,
                    "rows_per_step": int(np.median([r["rows"] for r in res]))},
        "cpu_baseline": {**{k: res[-1][k] for k in ("unit", "cores", "kind", "sample")}, "value": val},
        "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.skip_numba:
        line["cpu_baseline_numba"] = numba_sample(args, min(args.cpu_seconds, 5.0))
    print_json(line)


def main():
    args = parse()
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               str(Path(__file__).resolve()), *sys.argv[1:]]
        sys.exit(subprocess.call(cmd))
    # keep stdout for the ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_json
    print_json = lambda line: (json_out.write(json.dumps(line) + "\n"), json_out.flush())  # noqa: E731
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import nbody_streams_b200 as nb
    from nbody_streams_b200 import _lib
    from nbody_streams_b200._lib import MODE_BOTH
    from nbody_streams_b200.engine import DirectEngine
    from nbody_streams_b200.fields import EPS2_DEFAULT, KERNEL_MAP, _PRECISION_MAP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    xv, m = workload(args)
    N = xv.shape[0]
    prec, dtype_np = _PRECISION_MAP[args.precision]
    kid = KERNEL_MAP[args.kernel]

    # ---- FP32 / FP64 FMA peak of THIS device, measured now (roofline denominator) ----
    ops, ms = C.c_double(), C.c_double()
    which = 7 if args.precision == "float64" else 0
    _lib.check(lib.nb_ubench(which, 20000, C.byref(ops), C.byref(ms)))
    peak_tf = 2.0 * ops.value / 1e12

    eng = DirectEngine(N, float(m[0]), args.softening, prec, kid, nb.G_DEFAULT, EPS2_DEFAULT[args.precision], device,
                       overlap=not args.no_overlap, exchange=args.exchange, algorithm=args.algorithm)
    eng.load_state(xv)
    eng.compute(MODE_BOTH, positions_changed=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def step():
        flush.zero_()                      # L2 flush between steps (inside the timed region, ~0.05 ms)
        eng.kick_drift(args.dt)
        eng.compute(MODE_BOTH)
        eng.kick(args.dt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.sweep_events = []
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = e0.elapsed_time(e1)
    gpu_launches = eng.launches - launches0 + args.steps       # + the L2-flush memset per step (torch kernel)
    sweep_ms = [a.elapsed_time(b) for a, b, _ in eng.sweep_events]
    sweep_tags = [t for _, _, t in eng.sweep_events]
    eng.sweep_events = None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = args.steps * float(N) * float(N) / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (the pair sweep) on this rank: algorithmic flops per launch / duration.
    # Interactions credited to a launch = ordered pairs it delivers (reference convention): a one-sided launch
    # over T j-tiles delivers n_local x 256 T; a symmetric launch delivers (own shard)^2 for the triangle plus
    # 2 x (own shard) x (particles of the rectangle ranges).
    def real(t0, t1):
        return float(max(0, min(N, t1 * 256) - min(N, t0 * 256)))

    if eng.symmetric:
        from nbody_streams_b200.engine import sym_plan

        plan = sym_plan(eng.layout, eng.rank, lib.nb_sym_block_tiles(prec))
        own = real(*plan["own"])
        nb_ib = max(1, plan["n_ib"])
        rect = 2.0 * own * (real(*plan["A"]) + real(*plan["B"])) \
            + 2.0 * own * real(*plan["C"]) * (plan["ibC"][1] - plan["ibC"][0]) / nb_ib
        credit = {"tri": own * own, "rect": rect, "all": own * own + rect}
        inter_per_launch = [credit[tag] for tag in sweep_tags]
        kernel_name = "nb::sym_f32_kernel" if args.precision != "float64" else "nb::sym_f64_kernel"
    else:
        inter_per_launch = [eng.n_local * tiles * 256.0 for tiles in sweep_tags]
        kernel_name = "nb::sweep_f32_kernel" if args.precision != "float64" else "nb::sweep_f64_kernel"
    sweep_total_ms = float(np.sum(sweep_ms))
    achieved_tf = FLOPS_PER_INTERACTION * float(np.sum(inter_per_launch)) / (sweep_total_ms * 1e-3) / 1e12
    # per-slot launch times, max over ranks (slot = position of the launch inside a step)
    per_step = max(1, len(sweep_ms) // max(1, args.steps))
    slot_ms = [float(np.mean(sweep_ms[k::per_step])) for k in range(per_step)]
    slot_t = torch.tensor(slot_ms + [0.0] * (4 - len(slot_ms)), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(slot_t, op=dist.ReduceOp.MAX)
    slot_ms_max = [float(x) for x in slot_t.tolist()[:per_step]]
    nominal = FP32_PEAK_NOMINAL_TF if args.precision != "float64" else FP32_PEAK_NOMINAL_TF / 2
    traffic = None
    if N == 1_000_000 and world == 1 and args.precision != "float64":
        traffic = TRAFFIC_SYM_N1M if eng.symmetric else 33459712
    roofline = {
        "bound": "fp32" if args.precision != "float64" else "fp64",
        "kernel": kernel_name,
        "algorithm": "pair-symmetric: each unordered pair evaluated once, applied to both particles" if eng.symmetric
                     else "one-sided: every ordered pair evaluated",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": "FFMA (DFMA for float64) microbenchmark nb_ubench run in this process just before the timed "
                       "region; MEASURED_PEAKS.json carries no FP32/FP64 CUDA-core figure",
        "peak_nominal": nominal, "frac_of_nominal": achieved_tf / nominal,
        "flops_per_interaction": FLOPS_PER_INTERACTION,
        "note": ("achieved = 20 flop x ordered pairs delivered / launch time (the reference's N^2 convention); the "
                 "symmetric sweep executes 16 packed FP32 instructions per two interactions where the one-sided "
                 "sweep executes 24, so its FP32-pipe utilisation (ncu, profiles/) is lower than this fraction")
                if args.precision != "float64" else
                ("achieved = 20 flop x ordered pairs delivered / launch time (the reference's N^2 convention); the "
                 "symmetric sweep executes 18 (force) / 21 (force+potential) FP64 instructions per two interactions, "
                 "the one-sided sweep 2 x 15-17; a DFMA reading three different registers takes 3 issue cycles, not 2 "
                 "(profiles/r02_ubench_f64.md), which is what the FP64 pipe-active figure in profiles/ reflects"),
        "launch_ms_avg": sweep_total_ms / max(1, len(sweep_ms)), "launches_timed": len(sweep_ms),
        "launch_ms": [{"slot": k, "what": sweep_tags[k] if k < len(sweep_tags) else None, "ms_rank0": slot_ms[k],
                       "ms_max_over_ranks": slot_ms_max[k]} for k in range(per_step)],
        "sweep_share_of_step": sweep_total_ms / elapsed_ms,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full` capture of this
        # command at the default workload (profiles/); null for other shapes
        "traffic": traffic,
        "traffic_unit": "bytes per launch (ncu dram read+write); algorithmic bytes per launch: 20 MB of records read "
                        "+ 32 MB of double partial sums (L2-resident, atomically updated) for 2e13 algorithmic flop "
                        "-> " + ("FP32" if args.precision != "float64" else "FP64") + "-pipe bound",
    }

    # ---- parity (outside the timed region): rank 0's first rows after the last step against the float64 oracle
    parity = None
    if args.parity_rows > 0:
        xv_now = eng.state_numpy()                         # collective: every rank takes part
        if rank == 0:
            import oracle

            oracle.build()
            oracle.set_num_threads(len(os.sched_getaffinity(0)), fast=True)
            R = int(min(args.parity_rows, eng.n_local))
            p64 = xv_now[:, :3] if args.precision == "float64" else xv_now[:, :3].astype(np.float32).astype(np.float64)
            m64 = float(m[0]) if args.precision == "float64" else float(np.float32(m[0]))
            h64 = args.softening if args.precision == "float64" else float(np.float32(args.softening))
            kwo = dict(kernel=args.kernel, rows=(0, R), fast=True)
            a_ref = oracle.compute_nbody_forces_cpu(p64, m64, h64, nb.G_DEFAULT, **kwo)
            p_ref = oracle.compute_nbody_potential_cpu(p64, m64, h64, nb.G_DEFAULT, **kwo)
            a_gpu = eng.acc[:R].double().cpu().numpy()
            p_gpu = eng.phi[:R].double().cpu().numpy()
            rms = lambda x, r: float(np.sqrt(((x - r) ** 2).sum() / (r ** 2).sum()))  # noqa: E731
            parity = {"acc_rel_rms": rms(a_gpu, a_ref), "phi_rel_rms": rms(p_gpu, p_ref),
                      "acc_rel_max": float((np.linalg.norm(a_gpu - a_ref, axis=1) / np.linalg.norm(a_ref, axis=1)).max()),
                      "rows": R, "tolerance": 1e-12 if args.precision == "float64" else 1e-5,
                      "against": "float64 oracle (oracle/nbody_oracle.c) on the state after the last timed step"}
            parity["ok"] = bool(max(parity["acc_rel_rms"], parity["phi_rel_rms"]) <= parity["tolerance"])

    # ---- e2e: public host-array API, H2D + D2H inside the timed region ----
    e2e = None
    if not args.skip_e2e:
        pos_host = np.ascontiguousarray(xv[:, :3].astype(dtype_np))
        kw = dict(precision=args.precision, kernel=args.kernel)
        if world > 1:
            kw["group"] = dist.group.WORLD
        nb.compute_nbody_forces_potential_gpu(pos_host, float(m[0]), args.softening, **kw)   # warm-up
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            acc, phi = nb.compute_nbody_forces_potential_gpu(pos_host, float(m[0]), args.softening, **kw)
        barrier()
        dt_e2e = time.perf_counter() - t0
        t = torch.tensor([dt_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_e2e = float(t.item())
        e2e = {"value": n_e2e * float(N) * float(N) / dt_e2e, "unit": "interactions/s",
               # per rank: with several ranks each uploads only the rows of its shard (the packed records of the
               # other shards arrive over NVLink) and every rank downloads the full result (gather='all')
               "h2d_bytes_per_step": int(pos_host.nbytes // world), "d2h_bytes_per_step": int(acc.nbytes + phi.nbytes),
               "ms_per_call": dt_e2e / n_e2e * 1e3, "calls": n_e2e,
               "api": "compute_nbody_forces_potential_gpu(numpy pos, scalar mass, scalar softening) -> numpy"}

    # ---- context: the other sweeps of this workload on the same resident state (1 GPU, device-timed) ----
    sweeps = None
    if world == 1 and not args.skip_e2e:
        from nbody_streams_b200._lib import MODE_FORCE

        def time_engine(kname, mode, algorithm):
            e = DirectEngine(N, float(m[0]), args.softening, prec, KERNEL_MAP[kname], nb.G_DEFAULT,
                             EPS2_DEFAULT[args.precision], device, algorithm=algorithm)
            e.load_state(xv)
            e.compute(mode, positions_changed=False)          # warm-up
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                e.compute(mode, positions_changed=False)
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / 2

        sweeps = {}
        for kname, mode, mname in (("plummer", MODE_FORCE, "force"), ("spline", MODE_FORCE, "force"),
                                   ("spline", MODE_BOTH, "force+potential"), ("plummer", MODE_BOTH, "force+potential")):
            for alg in ("symmetric", "onesided"):
                if alg == "symmetric" and kname == args.kernel and mode == MODE_BOTH and eng.symmetric:
                    continue                                      # that is the timed step itself
                ms = time_engine(kname, mode, alg)
                rate = float(N) * N / (ms * 1e-3)
                sweeps[f"{kname}_{mname}_{alg}"] = {"ms": ms, "interactions_per_s": rate,
                                                    "frac_of_nominal_peak": rate * 20 / 1e12 / nominal,
                                                    "frac_of_measured_peak": rate * 20 / 1e12 / peak_tf}

    # ---- the reference's own CUDA kernel on this GPU (cubins built from its templates, baseline/_ref) ----
    ref_gpu = None
    if world == 1 and not args.skip_ref_gpu and args.precision != "float64":
        try:
            sys.path.insert(0, str(ROOT / "tests"))
            import ref_gpu_kernels as refk

            if refk.available():
                from nbody_streams_b200._lib import MODE_FORCE

                n_ref = min(N, 262_144)
                xr = torch.from_numpy(np.ascontiguousarray(xv[:n_ref, :3].astype(np.float32))).to(device)
                mr = torch.full((n_ref,), float(m[0]), dtype=torch.float32, device=device)
                hr = torch.full((n_ref,), args.softening, dtype=torch.float32, device=device)

                def med(fn, reps=5, warm=2):               # the reference's protocol, fields.py:1291-1314
                    ts = []
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    for r in range(warm + reps):
                        torch.cuda.synchronize(); a0.record(); fn(); a1.record(); torch.cuda.synchronize()
                        if r >= warm:
                            ts.append(a0.elapsed_time(a1))
                    return float(np.median(ts))

                from nbody_streams_b200.fields import _run_direct

                ref_gpu = {"n": n_ref, "precision": args.precision,
                           "what": "reference nbody_forces(_kahan)_kernel_float4 (cuda_kernels.py), its flags, sm_100, "
                                   "launch geometry of fields.py:823-852, kernel time only; ours = pack + sweep + finish"}
                for kname in ("plummer", "spline"):
                    kidr = KERNEL_MAP[kname]
                    t_ref = med(lambda: refk.launch("force", args.precision, xr, mr, hr, nb.G_DEFAULT, kidr))
                    t_our = med(lambda: _run_direct(xr, mr, 0.0, hr, 0.0, n_ref, prec, kidr, MODE_FORCE, nb.G_DEFAULT,
                                                    EPS2_DEFAULT[args.precision], dtype_np))
                    ref_gpu[f"force_{kname}"] = {"ref_ms": t_ref, "ours_ms": t_our, "speedup": t_ref / t_our}
        except Exception as e:  # noqa: BLE001 - a missing cubin or cuda-python must not cost the bench line
            ref_gpu = {"unavailable": repr(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_sample(args, xv, m, args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    cpu_numba = None
    if rank == 0 and world == 1 and not args.skip_cpu and not args.skip_numba:
        cpu_numba = numba_sample(args, min(args.cpu_seconds, 5.0))

    if rank == 0:
        line = {
            "metric": "pairwise interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision != "float64" else "f64", "data": "synthetic",
            "config": config_dict(args, N),
            "details": {
                "algorithm": ("pair-symmetric sweep (nb::sym_*): every unordered pair evaluated once and applied to both "
                              "particles, double atomics into a 32 B/particle workspace" if eng.symmetric else
                              "one-sided sweep (nb::sweep_*): every ordered pair evaluated, bitwise reproducible"),
                "step": "KDK leapfrog step: fused kick+drift+repack, all-gather of the j-records (N>1), one force+potential "
                        "sweep delivering all N^2 interactions, reduce-scatter of the partial sums (N>1, symmetric), "
                        "finish, half kick; float64 pos/vel resident on the device",
                "parallelism": f"i-particles sharded over {world} GPU(s), j-records "
                               + ("pushed to all ranks by the fused kick+drift kernel over NVLink peer memory"
                                  if (world > 1 and args.exchange == "p2p") else "all-gathered over NCCL")
                               + (", local-shard sweep overlapped with the gather" if world > 1 and not args.no_overlap else ""),
                "tuning_variant": int(lib.nb_get_variant(prec)),
            },
            "fraction_of_fp32_peak": value * FLOPS_PER_INTERACTION / 1e12 / (peak_tf * world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "parity": parity, "other_sweeps": sweeps, "ref_gpu_kernel": ref_gpu,
            "clocks": clocks,
        }
        if cpu_numba is not None:
            line["cpu_baseline_numba"] = cpu_numba
        print_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
