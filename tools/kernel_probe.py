"""Development probe (GPU): marching-kernel time and correctness on named workloads, for a library build chosen with PBTE_LIB_DIR
and run-time knobs taken from the environment (PBTE_MARCH_SPLIT, PBTE_MARCH_CTAS, PBTE_MARCH_REGOP, PBTE_TAIL_CHUNKS ...).
usage: kernel_probe.py <label> [workload ...]      workloads: c2 c2s c5s c3s c4s tri (default: c2)
Correctness: three iterations against the per-element LU kernel (kernel 2; ndof <= 32) or the generic one (kernel 1) on the same mesh."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import pbte_b200 as pb
from cases import BC2, BC3

WL = {
    # name: (mesh, order, refine, angle_dim, polar, azimuth, bcs, num_spectral, face_rule, check)
    "c2": ("unit-cube-hex", 2, 2, 3, 8, 16, BC3, 1, 0, True),
    "c2s": ("unit-cube-hex", 2, 1, 3, 8, 16, BC3, 1, 0, True),
    "c5s": ("unit-square-quad", 4, 4, 2, 1, 64, BC2, 4, 1, True),     # 16384 quads p4, 64 dirs x 8 bands = 512 slots
    "c3s": ("unit-cube-tet", 2, 2, 3, 8, 16, BC3, 2, 0, True),        # 24576 tets p2, 128 dirs x 4 bands = 512 slots
    "c4s": ("unit-cube-hex", 3, 1, 3, 8, 16, BC3, 1, 1, True),        # 512 hexes p3, 256 slots

}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


label = sys.argv[1]
for name in (sys.argv[2:] or ["c2"]):
    mesh, order, refine, adim, polar, azim, bcs, nspec, rule, check = WL[name]
    hc = pb.HostCase(mesh, order, refine, adim, polar, azim, True, False, bcs, face_rule=rule, num_spectral=nspec)
    P = hc.problem
    s = pb.Solver(P, kernel=3)
    s.build_schedule()
    s.solve(3, 0.0)
    got = s.download_moments()
    u = s.download_coeff(P.ndir // 3, P.nband - 1)
    err = None
    if check:
        g = pb.Solver(P, kernel=2 if P.ndof <= 32 else 1)
        g.build_schedule()
        g.solve(3, 0.0)
        want = g.download_moments()
        err = max(rel(got["Tc"], want["Tc"]), rel(got["Qc"], want["Qc"]), rel(u, g.download_coeff(P.ndir // 3, P.nband - 1)))
        g.close()
    s.reset_state()
    march, sweep = [], []
    for i in range(12):
        s.iterate_async()
        s.iterate_sync()
        st = s.stats()
        march.append(st.last_march_ms)
        sweep.append(st.last_sweep_ms)
    # back-to-back iterations, as the benchmark times them
    import time
    s.reset_state()
    s.iterate_sync()
    n = 50
    t0 = time.perf_counter()
    for i in range(n):
        s.iterate_async()
    s.iterate_sync()
    step = (time.perf_counter() - t0) / n * 1e3
    tm, ts = float(np.median(march[3:])), float(np.median(sweep[3:]))
    U = P.updates_per_sweep
    print(f"[{label}] {name}: ndof {P.ndof} ne {P.ne} slots {P.ndir * P.nband} variant {st.march_variant} fill {st.march_fill:.2f}: march {tm:8.4f} ms  sweep+tail {ts:8.4f} ms  "
          f"step {step:8.4f} ms = {U / step / 1e6:8.1f} M updates/ms   err vs LU kernel {err if err is None else f'{err:.1e}'}", flush=True)
    s.close()
    hc.close()
