"""Development probe: iterations and seconds of pbte_cuda_solve to a residual tolerance on C2 (seconds to converged T)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import pbte_b200 as pb
from cases import BC3

tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-8
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
face_rule = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # 0 = the reference's face rule (quirk Q1), 1 = consistent traces
hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, 16, True, False, BC3, face_rule=face_rule, num_spectral=1)
s = pb.Solver(hc.problem, kernel=0)
s.build_schedule()
t0 = time.perf_counter()
try:
    it, res, hist = s.solve(cap, tol)
except pb.PbteError as ex:
    print(f"C2 face_rule {face_rule}: solve stopped: {ex}")
    m = s.download_moments()
    print("   |Tc|max", float(np.nanmax(np.abs(m["Tc"]))), "finite", bool(np.isfinite(m["Tc"]).all()))
    sys.exit(0)
dt = time.perf_counter() - t0
print(f"C2 face_rule {face_rule}: tol {tol:g}: {it} iterations, residual {res:.3e}, {dt:.3f} s ({dt / max(it, 1) * 1e3:.3f} ms per iteration); residual history every 10 %:",
      [float(f"{hist[i]:.2e}") for i in range(0, it, max(1, it // 10))])
s.close(); hc.close()
