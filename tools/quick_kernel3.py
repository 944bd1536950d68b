"""Manual GPU check of the shared-operator path on a few cases (not collected by pytest)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import oracle as orc
import pbte_b200 as pb
from cases import CASES, host_case

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

for name in sys.argv[1:] or ["hex_p1", "c2_hex_p2_small", "tet_p2", "c1_tri_p1", "quad_p2"]:
    hc = host_case(name)
    P = hc.problem
    iters = min(CASES[name]["iters"], 4)
    want = orc.solve(P, iters, 0.0, cache_lu=True, threads=8)
    t = time.time()
    s = pb.Solver(P, kernel=3)
    s.build_schedule()
    it, res, hist = s.solve(iters, 0.0)
    got = s.download_moments()
    coeff = s.download_all_coeff()
    st = s.stats()
    print(name, "kernel", st.kernel_used, "oclasses", st.n_operator_classes, "fill %.2f" % st.march_fill,
          {k: rel(got[k], want[k]) for k in ("Tc", "Qc")}, "coeff", rel(coeff, want["coeff"]), "%.2fs" % (time.time() - t), flush=True)
    s.close(); hc.close()
