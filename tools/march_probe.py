"""Development probe: sweep time of the marching kernel on the C2 mesh for different slot counts (CTAs per SM)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import pbte_b200 as pb
from cases import BC3

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for azim in [int(a) for a in (sys.argv[2:] or [4, 8, 16, 32])]:
    hc = pb.HostCase("unit-cube-hex", 2, refine, 3, 8, azim, True, False, BC3, num_spectral=1)
    P = hc.problem
    s = pb.Solver(P, kernel=3)
    s.build_schedule()
    ms = []
    for i in range(8):
        s.sweep(); s.finish_iteration()
        ms.append(s.stats().last_sweep_ms)
    st = s.stats()
    U = P.updates_per_sweep
    t = float(np.median(ms[3:]))
    print(f"refine {refine} azimuth {azim:3d}: slots {P.ndir * P.nband:5d}  sweep {t:8.4f} ms  {U / t / 1e6:9.1f} M updates/ms  fill {st.march_fill:.2f}", flush=True)
    s.close(); hc.close()
