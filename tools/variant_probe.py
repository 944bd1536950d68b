"""Development probe: sweep time of the two marching variants (operators in shared memory / in registers) on a builtin mesh.
usage: variant_probe.py <mesh> <order> <refine> <angle_dim> <polar> <azimuth> [num_spectral]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import pbte_b200 as pb
from cases import BC3, BC2

mesh, order, refine, adim, polar, azim = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
nspec = int(sys.argv[7]) if len(sys.argv) > 7 else 1
bc = BC3 if "cube" in mesh else BC2
hc = pb.HostCase(mesh, order, refine, adim, polar, azim, True, False, bc, num_spectral=nspec)
P = hc.problem
ref = None
for regop in ("0", "1"):
    os.environ["PBTE_MARCH_REGOP"] = regop
    s = pb.Solver(P, kernel=3)
    s.build_schedule()
    ms = []
    for i in range(7):
        s.sweep(); s.finish_iteration()
        ms.append(s.stats().last_sweep_ms)
    st = s.stats()
    got = s.download_moments()
    t = float(np.median(ms[2:]))
    err = 0.0 if ref is None else float(np.abs(got["Tc"] - ref).max() / np.abs(ref).max())
    ref = got["Tc"] if ref is None else ref
    print(f"{mesh} p{order} r{refine} ndof {P.ndof} ne {P.ne} slots {P.ndir * P.nband} variant {st.march_variant}: sweep {t:8.4f} ms  {P.updates_per_sweep / t / 1e6:8.1f} M updates/ms  "
          f"levels {st.max_levels} fill {st.march_fill:.2f}  Tc vs variant 0: {err:.1e}", flush=True)
    s.close()
hc.close()
