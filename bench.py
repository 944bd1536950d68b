#!/usr/bin/env python
"""bench.py — headline benchmark of the PBTE sweep: DOF*angle*band updates per second per source iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-strong]

Workload (BASELINE.json configs[1], SURVEY §8d "C2"): builtin unit-cube-hex 4x4x4 refined twice (4096 hexes), DG order 2
(ndof 27), 8 Gauss polar x 16 uniform azimuth directions (128), "gray" = num_spectral 1 (2 bands), walls -0.5 / top +0.5,
local matrices as the reference assembles them.  One step = one source iteration (pbte_cuda_iterate_async): upwind sweeps of
every (direction, band) over the mesh, one tail kernel that streams the flux for the moments and finishes the iteration (cell
integrals, residual, stop test) [for N > 1: moments, the library's NCCL all-reduce of the moment vector, finish].  The K timed
steps are enqueued back to back without host synchronisation, as pbte_cuda_solve runs them.
N > 1 (torchrun, one process per GPU): every rank sweeps its own block of directions over the whole mesh; the azimuth count
grows with N (16*N), so per-GPU work is fixed ("weak" scaling) and the one NCCL all-reduce per iteration, issued by the
library on its own communicator (pbte_cuda_comm_attach), is the only data-path collective.  Every N also carries a
`strong_scaling` record: BASELINE's non-gray 3-D Si case C3 at full size, its 128 directions split over the N GPUs.

Timed with CUDA events on the sweep stream, barrier + synchronize on both sides, max over ranks.  The flux (226 MB per
128 directions) and the matrix image (0.4 GB) together exceed the 126 MB L2 every step, so no explicit flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "DOF*angle*band updates/s per sweep"
UNIT = "updates/s"
BC3 = {1: -0.5, 2: -0.5, 3: -0.5, 4: -0.5, 5: -0.5, 6: 0.5}
WORKLOAD = dict(mesh="unit-cube-hex", order=2, refine=2, angle_dim=3, polar=8, azimuth=16, num_spectral=1)
WORKLOAD_NAME = "C2: builtin unit-cube-hex refine 2 (4096 hexes), DG order 2 (ndof 27), 8 Gauss polar x 16 uniform azimuth directions per GPU, gray SMRT = 2 bands"
# CPU legs: the SAME mesh, order, bands and polar rule; a step sweeps a bounded number of azimuth points (the cost of a direction
# does not depend on how many others there are), chosen so that the run ends within minutes and the reference's LU cache stays small
CPU_MAX_AZIMUTH = 4


def cpu_sample_azimuth(steps, warmup, budget_s=150.0, s_per_direction=0.04):
    """Azimuth points of the CPU sample: (steps + warmup) x 8 polar x azimuth x ~40 ms per direction within the budget."""
    return int(max(1, min(CPU_MAX_AZIMUTH, budget_s / (max(steps + warmup, 1) * 8 * s_per_direction))))
# --workload: the other BASELINE configs at their real sizes (extra measurements, profiles/README.md; the default stays C2, the
# configuration the metric is quoted on).  C2 is weak-scaled (128 directions per GPU); these keep the named quadrature and split
# its directions over the GPUs (strong scaling)
EXTRA_WORKLOADS = {
    "c3": dict(mesh="unit-cube-tet", order=2, refine=3, angle_dim=3, polar=8, azimuth=16, num_spectral=20,
               label="C3: builtin unit-cube-tet refine 3 (196608 tets), DG order 2 (ndof 10), 8 x 16 = 128 directions, non-gray Si = 40 bands"),
    "c5": dict(mesh="unit-square-quad", order=4, refine=5, angle_dim=2, polar=1, azimuth=64, num_spectral=20, bcs={1: -0.5, 2: -0.5, 3: 0.5, 4: -0.5},
               label="C5: builtin unit-square-quad refine 5 (65536 quads), DG order 4 (ndof 25), 64 azimuth directions (2-D), non-gray Si = 40 bands"),
    "c4full": dict(mesh="unit-cube-hex", order=3, refine=3, angle_dim=3, polar=16, azimuth=32, num_spectral=20,
                   label="C4 at full size: unit-cube-hex refine 3 = the 8-hex mesh file refined 4 times (32768 hexes), DG order 3 (ndof 64), 16 x 32 = 512 directions, 40 bands (4.29e10 updates per iteration, 344 GB of flux: 4 or 8 GPUs)"),
    "c4": dict(mesh="unit-cube-hex", order=3, refine=2, angle_dim=3, polar=16, azimuth=32, num_spectral=20,
               label="C4 (1-GPU strong-scaling point of SURVEY 8d): unit-cube-hex refine 2 (4096 hexes), DG order 3 (ndof 64), 16 x 32 = 512 directions, 40 bands"),
}


def flops_per_update(n, dim, faces):
    """Per-element LU path: algorithmic FP64 flops per update: LU + assembly + source/coupling products +
    triangular solves, per dof."""
    f_in = f_out = faces / 2.0
    per_solve = (2.0 / 3.0) * n ** 3 + 2.0 * (1 + dim + f_out) * n * n + (4 + 2 * f_in) * n * n
    return per_solve / n


def flops_per_update_shared(n, mean_inflow_faces):
    """Shared-operator path (restated algebra): u_e = P x_e + sum_{interior inflow f} Q_f u_nbr(f) + wall vectors:
    (1 + F_in) dense n x n products per element, F_in = measured mean number of interior inflow faces, plus forming x_e."""
    return 2.0 * n * (1.0 + mean_inflow_faces) + 2.0


def mean_interior_inflow_faces(P, k0, k1):
    """Mean over (owned direction, element) of the number of interior faces with n.Omega < 0 (exact count for this workload)."""
    nrm = P.face_normal[:, :, : P.dim]
    interior = P.face_nbr >= 0
    tot = 0
    for k in range(k0, k1):
        a = nrm @ P.dir[k, : P.dim]
        tot += int(np.count_nonzero((a < 0.0) & interior))
    return tot / float(max(k1 - k0, 1) * P.ne)


def load_measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


_SAMPLER_SRC = r"""
import sys, time, select
idx = int(sys.argv[1])
rows = []
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
    print("ready nvml", mx, flush=True)
    while not select.select([sys.stdin], [], [], 0.0002)[0]:
        rows.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(get_reasons(h))))
except Exception as ex:
    print("ready none 0", flush=True)
for r in rows:
    print(*r)
"""


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of the default run lasts ~10 ms, far below the
    100 ms period of the recipe's `nvidia-smi -lms` line (B200_PROFILING.md), so the same counters are read through NVML
    (nvidia_ml_py) as fast as the driver answers (a few queries per millisecond at best).  The polling runs in a CHILD process that keeps its samples in memory until it is told to stop:
    a polling thread inside this process competes with the launching thread for the interpreter lock and stretched the
    host-side gaps of rank 0 at N > 1 by 0.3 ms per step.  Samples are time-stamped; only those inside the timed region count."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.source = None
        self.max_mhz = None
        self.t0 = self.t1 = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def launch(self):
        """Start the child (before the warm-up: it needs ~0.2 s to import NVML); returns once it is polling."""
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self._physical_index())], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().split()
            self.source = first[1] if len(first) > 1 else "none"
            self.max_mhz = float(first[2]) if len(first) > 2 else None
        except Exception:
            self.proc = None

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML sampler could not start"], "samples": 0}
        try:
            out, _ = self.proc.communicate("stop\n", timeout=10)
        except Exception:
            self.proc.kill()
            out = ""
        sm, reasons = [], set()
        for line in out.splitlines():
            f = line.split()
            if len(f) != 3:
                continue
            t, mhz, mask = float(f[0]), float(f[1]), int(f[2])
            if self.t0 <= t <= self.t1:
                sm.append(mhz)
                for name, bit in self.BITS.items():
                    if mask & bit:
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm),
                "source": self.source, "window_ms": (self.t1 - self.t0) * 1e3}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, azimuth, threads_for_port=None):
    """The reference's own CPU implementation (its unmodified sources against the MFEM stand-in, oracle/_ref; serial, as the
    reference is) on the benchmark's mesh / order / bands with `azimuth` azimuth points; the C port when oracle/_ref is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    import pbte_b200 as pb
    wl = WORKLOAD
    mat = dict(pb.SI_MATERIAL)
    mat["num_spectral"] = wl["num_spectral"]
    times = []
    if orc.ref_available():
        kind, cores = "reference", 1
        rc = orc.RefCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], azimuth, True, False, BC3, mat)
        U = rc.ndof * rc.ne * rc.ndir * rc.nband
        rs = orc.RefSolver(rc, 1, 0)
        ctor = rs.ctor_seconds
        for i in range(warmup + steps):
            _, secs = rs.solve()
            if i >= warmup:
                times.append(secs)
        ne, ndir, nband = rc.ne, rc.ndir, rc.nband
        rs.close()
        rc.close()
    else:
        kind, cores = "port", threads_for_port or (os.cpu_count() or 1)
        hc = pb.HostCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], azimuth, True, False, BC3, num_spectral=wl["num_spectral"])
        P = hc.problem
        U = P.updates_per_sweep
        ctor = 0.0
        state = None
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = orc.solve(P, 1, 0.0, cache_lu=False, threads=cores, **(state or {}))
            secs = time.perf_counter() - t0
            state = dict(coeff=r["coeff"], Tc=r["Tc"], Tv=r["Tv"])
            if i >= warmup:
                times.append(secs)
        ne, ndir, nband = P.ne, P.ndir, P.nband
        hc.close()
    t = float(np.mean(times))
    sample = (f"{'reference sources (-O2, serial as the reference is; its MPI path is dead code: n/a)' if kind == 'reference' else 'C restatement, all host cores'}"
              f" on the benchmark's own mesh ({wl['mesh']} refine {wl['refine']}: {ne} hexes), order 2, 2 bands, 8 polar x {azimuth} azimuth = {ndir} of the 128"
              f" directions per step ({U} updates; a direction's sweep costs the same whatever the number of directions); {steps} timed PBTESolver::Solve"
              f" iterations after {warmup} warm-up; constructor with its FullLU cache {ctor:.1f} s, not timed")
    return {"value": U / t, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, t


def run_reference(args, rank, world):
    """--impl reference: the reference arm of the driver.  Rank 0 alone runs it."""
    if rank != 0:
        return
    base, t = cpu_reference_run(args.steps, args.warmup, cpu_sample_azimuth(args.steps, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "parallelism": "host cores of the GPU box", "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Rank 0, N = 1: the reference CPU path on a bounded sample of the same workload, timed beside the GPU numbers, and — as a
    labelled extra — the C port threaded over (direction, band) on all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    import pbte_b200 as pb
    out, _ = cpu_reference_run(3, 1, 2)
    if out["kind"] == "reference":
        wl = WORKLOAD
        hc = pb.HostCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], 4, True, False, BC3, num_spectral=wl["num_spectral"])
        cores = os.cpu_count() or 1
        orc.solve(hc.problem, 1, 0.0, cache_lu=False, threads=cores)
        t0 = time.perf_counter()
        orc.solve(hc.problem, 2, 0.0, cache_lu=False, threads=cores)
        tp = (time.perf_counter() - t0) / 2
        out["threaded_port"] = {"value": hc.problem.updates_per_sweep / tp, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "C restatement, LU on the fly, (direction, band) sweeps on all host cores, same mesh, 8 x 4 directions, 2 iterations"}
        hc.close()
    return out


class Plumbing:
    """torch.distributed for what is not the data path: barrier, max over ranks, shipping the communicator id."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch, self.rank, self.world, self.local_rank = torch, rank, world, local_rank
        self.dist = None
        self.saved_stdout = None
        if world > 1:
            # NCCL may print a banner on stdout when a communicator is created (NCCL_DEBUG=VERSION/INFO from the environment, which is
            # respected): stdout is kept for the one JSON line by pointing fd 1 at stderr until the warm-up is over
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)
            os.dup2(2, 1)
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def restore_stdout(self):
        if self.saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(self.saved_stdout, 1)
            os.close(self.saved_stdout)
            self.saved_stdout = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def attach(self, solver, pb):
        """The library's own NCCL communicator for the per-iteration all-reduce: rank 0's id goes to the others through torch."""
        if self.dist is None:
            return
        buf = self.torch.zeros(pb.COMM_ID_BYTES, dtype=self.torch.uint8, device=f"cuda:{self.local_rank}")
        if self.rank == 0:
            buf.copy_(self.torch.frombuffer(bytearray(pb.comm_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(buf, src=0)
        solver.comm_attach(bytes(buf.cpu().numpy().tobytes()), self.rank, self.world)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# With the reference's own face rule the source iteration DIVERGES on the order-2 hex mesh of C2 (the state grows
# by orders of magnitude per hundred iterations, in the reference's CPU code as well), so a long timed region re-zeroes the state
# in-stream every RESET_EVERY iterations to keep the numbers finite: memsets of the flux inside the timed region (~0.3 % of it).
RESET_EVERY = 64


def timed_iterations(solver, plumb, stream, steps, warmup, sampler=None):
    """W warm-up iterations, then EXACTLY `steps` iterations enqueued back to back (no host synchronisation between them),
    CUDA events on the sweep stream, barrier + synchronize on both sides; returns (ms of the timed region: max over ranks,
    residual of the last iteration, clocks, launches inside the timed region)."""
    torch = plumb.torch
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            solver.iterate_async()
        solver.iterate_sync()
        plumb.barrier()
        plumb.restore_stdout()
        if sampler is not None:
            sampler.start()
        st0 = solver.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(steps):
            if i and i % RESET_EVERY == 0:
                solver.reset_state()  # in-stream memsets, inside the timed region (see RESET_EVERY)
            solver.iterate_async()
        ev1.record(stream)
        _, res = solver.iterate_sync()
        plumb.barrier()
        clocks = sampler.stop() if sampler is not None else None
        st1 = solver.stats()
    return plumb.max(ev0.elapsed_time(ev1)), res, clocks, int(st1.kernel_launches - st0.kernel_launches)


def sampled_kernel_times(solver, stream, torch, n=10):
    """Duration of the sweep and of its dominant kernel, sampled over `n` more iterations of the same state with one
    synchronisation per iteration (the library's own CUDA events: before the sweep, behind the marching kernel, behind the tail)."""
    sweep_ms, march_ms = [], []
    solver.reset_state()
    with torch.cuda.stream(stream):
        for _ in range(n):
            solver.iterate_async()
            solver.iterate_sync()
            st = solver.stats()
            sweep_ms.append(st.last_sweep_ms)
            march_ms.append(st.last_march_ms)
    return float(np.mean(sweep_ms)), float(np.mean(march_ms))


def strong_scaling_leg(args, plumb, pb, stream):
    """BASELINE's non-gray 3-D Si case (C3: 196 608 tetrahedra order 2, 8 x 16 directions, 40 bands, 1.007e10 updates per
    iteration, 80.5 GB of flux) at its full size, the 128 directions split over the N GPUs: total work fixed = strong scaling."""
    torch, rank, world, local_rank = plumb.torch, plumb.rank, plumb.world, plumb.local_rank
    wl = EXTRA_WORKLOADS["c3"]
    t_setup = time.perf_counter()
    hc = pb.HostCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], wl["azimuth"], True, False, BC3, num_spectral=wl["num_spectral"])
    P = hc.problem
    k0, k1 = pb.partition_directions(P.ndir, world, rank)
    solver = pb.Solver(P, device=local_rank, dir_range=(k0, k1), cache_policy=0, kernel=0, stream=stream.cuda_stream)
    nclass, nlev = solver.build_schedule()
    plumb.attach(solver, pb)
    setup_s = time.perf_counter() - t_setup
    steps = 5
    total_ms, res, _, launches = timed_iterations(solver, plumb, stream, steps, 2)
    sweep_ms, march_ms = sampled_kernel_times(solver, stream, torch, 2)
    st = solver.stats()
    U = P.updates_per_sweep
    U_rank = P.ndof * P.ne * (k1 - k0) * P.nband
    peaks, peak_src = load_measured_peaks()
    kern_s = (march_ms if march_ms > 0 else sweep_ms) * 1e-3
    rec = {"workload": wl["label"], "scaling": "strong", "n_gpus": world, "value": U * steps / (total_ms * 1e-3), "unit": UNIT, "steps": steps,
           "ms_per_step": total_ms / steps, "updates_per_step": U, "directions_per_gpu": k1 - k0, "slots_per_gpu": (k1 - k0) * P.nband,
           "kernel_used": st.kernel_used, "operator_classes": st.n_operator_classes, "upwind_dags": nclass, "wavefront_levels": nlev, "tile_fill": st.march_fill,
           "flux_gb_per_gpu": st.flux_bytes / 1e9, "setup_seconds": setup_s, "residual": res, "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": 16.0 * U_rank / kern_s / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": 16.0 * U_rank / kern_s / 1e9 / peaks.get("hbm_gbs"), "kernel_ms": kern_s * 1e3, "peak_source": peak_src,
                        "note": "16 B/update (flux read once, written once) over the marching kernel's duration on rank 0"}}
    solver.close()
    hc.close()
    return rec


def run_gpu(args, rank, world, local_rank):
    import torch
    import pbte_b200 as pb
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    plumb = Plumbing(rank, world, local_rank)

    wl = EXTRA_WORKLOADS.get(args.workload, WORKLOAD)
    strong = args.workload in EXTRA_WORKLOADS
    azim = wl["azimuth"] * (1 if strong else world)
    t_setup = time.perf_counter()
    hc = pb.HostCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], azim, True, False, wl.get("bcs", BC3), num_spectral=wl["num_spectral"])
    host_setup_s = time.perf_counter() - t_setup
    P = hc.problem
    k0, k1 = pb.partition_directions(P.ndir, world, rank)
    if args.slice_of > 1:
        # one GPU's share of an N-GPU strong-scaling run, measured alone (no all-reduce): the per-GPU sweep time of that point
        k0, k1 = pb.partition_directions(P.ndir, args.slice_of, 0)
    stream = torch.cuda.Stream(device=local_rank)
    t_setup = time.perf_counter()
    solver = pb.Solver(P, device=local_rank, dir_range=(k0, k1), cache_policy=0, kernel=args.kernel, stream=stream.cuda_stream)
    nclass, nlev = solver.build_schedule()
    device_setup_s = time.perf_counter() - t_setup
    plumb.attach(solver, pb)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.launch()
    total_ms, res, clocks, launches = timed_iterations(solver, plumb, stream, args.steps, max(args.warmup, 3), sampler)
    U = P.updates_per_sweep  # all ranks together
    if args.slice_of > 1:
        U = P.ndof * P.ne * (k1 - k0) * P.nband  # the slice alone
    value = U * args.steps / (total_ms * 1e-3)
    sweep_ms, march_ms = sampled_kernel_times(solver, stream, torch)
    st1 = solver.stats()

    # ---- end-to-end leg: the same iteration with HOST buffers for what crosses the boundary every step ------------
    #   in : previous-iteration moments Tc, Tv from pinned host memory  (pbte_cuda_upload_moments)
    #   out: Tc, Qc, Tv, Qv and the residual back to host               (pbte_cuda_download_moments)
    # The flux is solver state and stays on the device, as in a Solve call of many iterations.
    block = P.ndof * P.ne
    pin_m = torch.zeros((1 + P.dim) * block, dtype=torch.float64).pin_memory()   # [Tc | Qc] as the library holds them
    pin_c = torch.zeros((1 + P.dim) * P.ne, dtype=torch.float64).pin_memory()    # [Tv | Qv]
    hm, hcell = pin_m.numpy(), pin_c.numpy()
    host = {"Tc": hm[:block], "Qc": hm[block:].reshape(P.dim, block), "Tv": hcell[: P.ne], "Qv": hcell[P.ne:].reshape(P.dim, P.ne)}
    solver.download_moments(host)
    e2e_steps = min(args.steps, 50)
    plumb.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(e2e_steps):
            solver.upload_moments(host["Tc"], host["Tv"])   # previous iterate from host memory
            solver.iterate_async()
            solver.iterate_sync()                            # (N > 1: the heat-flux moments are all-reduced here, once per read-out)
            solver.download_moments(host)                    # Tc, Qc, Tv, Qv back to host memory
    plumb.barrier()
    e2e_value = U * e2e_steps / plumb.max(time.perf_counter() - t0)
    h2d = (block + P.ne) * 8
    d2h = ((1 + P.dim) * block + (1 + P.dim) * P.ne + 1) * 8

    # ---- the kept C++ API call, PBTESolver::Solve(coeff, macro), with its full host-state upload/download (N = 1 only) --
    solve_call = None
    general = None
    if world == 1 and st1.kernel_used == 3 and not args.no_general and not strong:
        # the per-element LU path (any mesh) timed on the same workload, for the record
        solver.close()
        g = pb.Solver(P, device=local_rank, dir_range=(k0, k1), cache_policy=0, kernel=2, stream=stream.cuda_stream)
        g.build_schedule()
        with torch.cuda.stream(stream):
            for _ in range(3):
                g.sweep(); g.finish_iteration()
            ms = []
            for _ in range(5):
                g.sweep(); g.finish_iteration()
                ms.append(g.stats().last_sweep_ms)
        gs = float(np.mean(ms)) * 1e-3
        gf = flops_per_update(P.ndof, P.dim, P.nfaces)
        general = {"kernel": "warp-register (one pivoted LU per (element, direction, band); any mesh)", "value": U / gs, "unit": UNIT,
                   "sweep_ms": gs * 1e3, "flop_per_update": gf, "fp64_tflops": gf * U / gs / 1e12}
        g.close()
    if world == 1 and not strong:  # (the extra workloads hold up to 80 GB of flux: no host copy of it)
        solver.close()
        hs = pb.HostSolver(hc, 0)
        hs.solve(1)
        iters_per_call = 20
        t0 = time.perf_counter()
        _, iters, loop_ms = hs.solve(iters_per_call)
        dt = time.perf_counter() - t0
        solve_call = {"value": U * iters / dt, "unit": UNIT, "iterations_per_call": iters, "seconds": dt, "iteration_loop_ms": loop_ms,
                      "h2d_bytes_per_call": P.ndir * P.nband * block * 8 + h2d, "d2h_bytes_per_call": P.ndir * P.nband * block * 8 + d2h,
                      "note": "one PBTESolver::Solve(coeff, macro) call of the kept C++ API on host DenseMatrix blocks (views into the pinned slab that"
                              " CreateInitialCoefficients hands out): the whole flux uploaded at entry and downloaded at exit as two bulk transfers"}
        hs.close()

    # ---- seconds to converged T (the second half of BASELINE.json's metric), N = 1 -------------------------------------
    # With the reference's own face rule the source iteration DIVERGES on this order-2 hex mesh (Tc overflows
    # after a few hundred iterations, in the reference's CPU code as well: DESIGN.md section 4), so the time to a converged field is
    # measured on the same mesh and quadrature with consistent face traces, to the reference's default tolerance 1e-8.
    converged = None
    if world == 1 and not strong:
        t0 = time.perf_counter()
        hcc = pb.HostCase(wl["mesh"], wl["order"], wl["refine"], wl["angle_dim"], wl["polar"], azim, True, False, BC3, face_rule=1, num_spectral=wl["num_spectral"])
        t1 = time.perf_counter()
        sc = pb.Solver(hcc.problem, device=local_rank, cache_policy=0, kernel=args.kernel)
        sc.build_schedule()
        t2 = time.perf_counter()
        try:
            iters, resid, _ = sc.solve(5000, 1e-8)
            dt = time.perf_counter() - t2
            converged = {"seconds": dt, "iterations": int(iters), "residual": float(resid), "tol": 1e-8, "updates_per_s": U * iters / dt,
                         "face_rule": "consistent traces (the reference's own face rule, quirk Q1, diverges on this mesh)",
                         "setup_seconds": {"host_mesh_and_element_matrices": t1 - t0, "device_image_schedule_operators": t2 - t1},
                         "seconds_including_setup": time.perf_counter() - t0,
                         "what": "pbte_cuda_solve from zero initial state until |Tv - Tv_prev| / |Tv| < 1e-8; C2 mesh and quadrature"}
        except pb.PbteError as ex:
            converged = {"error": str(ex)}
        sc.close()
        hcc.close()
    if not strong:
        solver.close()
        hc.close()

    # ---- strong scaling of the non-gray 3-D Si case at its full size, at every N ---------------------------------------
    strong_rec = None
    if not strong and not args.no_strong:
        try:
            strong_rec = strong_scaling_leg(args, plumb, pb, stream)
        except Exception as ex:  # an extra must never take the headline down with it
            strong_rec = {"workload": EXTRA_WORKLOADS["c3"]["label"], "error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        plumb.close()
        return

    peaks, peak_src = load_measured_peaks()
    live = pb.measure_peaks(local_rank)
    U_rank = P.ndof * P.ne * (k1 - k0) * P.nband
    sweep_s = sweep_ms * 1e-3
    fp64_peak = max(live["fp64_fma_tflops"], live["fp64_mma_tflops"])
    kernel_names = {1: "generic", 2: "warp-register", 3: "shared-operator marching"}
    if st1.kernel_used == 3:
        fin = mean_interior_inflow_faces(P, k0, k1)
        fpu = flops_per_update_shared(P.ndof, fin)
        kname = ("sweep_march_regop_kernel" if st1.march_variant == 1 else
                 ("sweep_march_kernel" if st1.march_persistent else "sweep_march_classic_kernel"))
        khow = "persistent CTAs with work lists" if (st1.march_persistent and st1.march_variant != 1) else "one CTA per slot"
        kdesc = (f"{kname}<{P.ndof}> ({khow}; one launch per step: marching CTAs walk the {nlev} wavefront levels of every (direction, band) slot,"
                 f" {st1.n_operator_classes} operator class(es), tile fill {st1.march_fill:.2f}), followed by iteration_tail_kernel; sweep + tail {sweep_s * 1e3:.3f} ms per step"
                 f" by CUDA events on the sweep stream; algorithmic flops = 2n(1+F_in)+2 = {fpu:.1f} per update with F_in = {fin:.3f} interior inflow"
                 f" faces per element (shared-operator algebra, DESIGN.md); the per-element LU algebra of the reference would be"
                 f" {flops_per_update(P.ndof, P.dim, P.nfaces):.0f} flop/update")
        sched = {"classes": nclass, "wavefront_levels": nlev, "kernel": kernel_names[3], "operator_classes": st1.n_operator_classes,
                 "tiles": st1.march_tiles, "tile_fill": st1.march_fill, "operator_bytes": st1.operator_bytes, "cuda_graph": True}
    else:
        fpu = flops_per_update(P.ndof, P.dim, P.nfaces)
        kdesc = (f"sweep_{'warp' if st1.kernel_used == 2 else 'generic'}_kernel ({nlev} wavefront launches per step, {sweep_s * 1e3:.3f} ms per step by CUDA events"
                 f" on the sweep stream); algorithmic {fpu:.0f} flop/update (SURVEY 8d)")
        sched = {"classes": nclass, "wavefront_levels": nlev, "kernel": kernel_names.get(st1.kernel_used, "?"), "cuda_graph": True,
                 "pivot_cache": "FullLU policy: pivot lanes cached, LU recomputed in registers"}
    # duration of the dominant kernel: the marching kernel alone (event recorded right behind it inside the sweep), else the sweep
    march_s = march_ms * 1e-3 if (st1.kernel_used == 3 and march_ms > 0.0) else 0.0
    kern_s = march_s if march_s > 0.0 else sweep_s
    achieved_tf = fpu * U_rank / kern_s / 1e12
    hbm_gbs = 16.0 * U_rank / kern_s / 1e9
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    traffic = None
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    flux_mb = P.ndof * P.ne * (k1 - k0) * P.nband * 8 / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"{wl['label']} ({k1 - k0} directions on this GPU), reference-compatible local matrices; {U} updates per step" if strong else WORKLOAD_NAME)
                               + (f"; SLICE: one GPU's share of a {args.slice_of}-GPU run measured alone (value = this GPU's updates/s, no all-reduce)" if args.slice_of > 1 else ""),
                   "detail": f"{P.ne} elements, ndof {P.ndof}, {P.ndir} directions in all ({k1 - k0} per GPU), {P.nband} bands, reference-compatible local matrices; {U} updates per step",
                   "parallelism": (f"directions partitioned over {world} GPU(s), one FP64 all-reduce of the temperature moment Tc per step issued by the library (NCCL, inside the iteration's CUDA graph); the heat-flux moments Qc are all-reduced once, when the state is read out" if world > 1 else "single GPU"),
                   "l2_flush": f"none needed: the flux ({flux_mb:.0f} MB per GPU) is read and rewritten every step and exceeds the 126 MB L2",
                   "schedule": sched,
                   "timed_region": f"{args.steps} iterations enqueued back to back (pbte_cuda_iterate_async), one synchronisation at the end: {total_ms:.1f} ms;"
                                   f" state re-zeroed in-stream every {RESET_EVERY} iterations inside the timed region (the reference's face rule makes the iteration diverge on this mesh)",
                   "setup_seconds": {"host_mesh_and_element_matrices": host_setup_s, "device_image_schedule_operators": device_setup_s},
                   "residual_after_run": res},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "what": "per step: Tc,Tv uploaded from pinned host memory, one whole iteration through the C ABI, Tc,Qc,Tv,Qv,residual downloaded; flux stays device resident",
                "solve_call": solve_call},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "pipe": "FP64 tensor pipe (DMMA m8n8k4), not bf16: peak measured live, MEASURED_PEAKS.json has no FP64 entry", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak, "traffic": traffic,
                     "kernel_ms": kern_s * 1e3, "sweep_ms": sweep_s * 1e3, "frac_of_whole_sweep": fpu * U_rank / sweep_s / 1e12 / fp64_peak,
                     "note": f"dominant kernel = {kdesc} x {U_rank} updates; achieved = algorithmic flops / {kern_s * 1e3:.3f} ms ("
                             f"{'the marching kernel alone, CUDA event behind it, mean of 10 iterations sampled right after the timed region' if march_s > 0.0 else 'the whole sweep'}); peak = FP64 pipe measured live by pbte_cuda_measure_peaks"
                             f" (DFMA {live['fp64_fma_tflops']:.1f}, DMMA {live['fp64_mma_tflops']:.1f} TFLOP/s; the two share one pipe: profiles/r02_fp64_mix_probe.txt)",
                     "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm_gbs / peaks.get("hbm_gbs"),
                             "algorithmic_bytes_per_update": 16, "peak_source": peak_src, "live_copy_gbs": live["hbm_copy_gbs"]}},
    }
    if general is not None:
        line["general_path"] = general
    if converged is not None:
        line["time_to_converged_T"] = converged
    if strong_rec is not None:
        line["strong_scaling"] = strong_rec
    if world == 1 and not args.no_cpu_baseline and not strong:
        try:
            line["cpu_baseline"] = cpu_baseline_leg()
        except Exception as ex:  # the CPU leg must never take the GPU numbers down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    plumb.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300, help="timed source iterations (300 x 0.45 ms: a timed region above 100 ms)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2"] + sorted(EXTRA_WORKLOADS), help="c2 = the benchmark; c3/c4/c5 = extra full-size measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-general", action="store_true", help="skip the side measurement of the per-element LU kernel")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (full-size C3, directions split over the GPUs)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 warp-register, 3 shared-operator marching")
    ap.add_argument("--slice-of", type=int, default=1, help="with --workload: run one GPU's direction share of an N-GPU strong-scaling run alone")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29531", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--kernel", str(args.kernel), "--workload", args.workload] + (["--no-strong"] if args.no_strong else [])
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
