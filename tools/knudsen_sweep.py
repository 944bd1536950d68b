"""BASELINE config C5 "ballistic-to-diffusive Knudsen sweep" (SURVEY 8d): the thin-film mesh (builtin unit-square-quad, DG order 4, 64
uniform azimuth directions, non-gray Si = 40 bands) solved for reference_length in {1e-8 .. 1e-4} m, i.e. from ballistic (mean free
paths far above the film size) to diffusive transport.  The source iteration is run to the reference's stop test (tol 1e-8) or to
the iteration cap; the pseudo-time iteration needs O(1/Kn^2) iterations in the diffusive limit.
usage: knudsen_sweep.py [refine=4] [max_iter=3000] [face_rule=1]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import pbte_b200 as pb
from cases import BC2

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 4
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
rule = int(sys.argv[3]) if len(sys.argv) > 3 else 1
print(f"C5 Knudsen sweep: unit-square-quad refine {refine}, order 4, 64 azimuth directions, 40 Si bands, face rule {rule}, tol 1e-8, cap {max_iter} iterations")
for ref_len in (1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
    t0 = time.perf_counter()
    hc = pb.HostCase("unit-square-quad", 4, refine, 2, 1, 64, True, False, BC2, face_rule=rule, num_spectral=20, ref_len=ref_len)
    P = hc.problem
    s = pb.Solver(P)
    s.build_schedule()
    t1 = time.perf_counter()
    try:
        it, res, hist = s.solve(max_iter, 1e-8)
        dt = time.perf_counter() - t1
        m = s.download_moments()
        st = s.stats()
        # Knudsen number of the band with the longest mean free path: Lambda = v / nu (nu = inv_kn is a rate here), film size = ref_len
        kn = float((P.vg / P.inv_kn).max() / ref_len)
        print(f"reference_length {ref_len:.0e} m (max Kn {kn:9.3e}): {P.ne} quads, {P.ndir * P.nband} slots, {it:5d} iterations, residual {res:.3e}, "
              f"solve {dt:7.3f} s ({P.updates_per_sweep * it / dt:.3e} updates/s), setup {t1 - t0:5.1f} s, Tv in [{m["Tv"].min():+.3e}, {m["Tv"].max():+.3e}], "
              f"kernel {st.kernel_used} variant {st.march_variant} persistent {st.march_persistent} fill {st.march_fill:.2f}", flush=True)
    except pb.PbteError as ex:
        print(f"reference_length {ref_len:.0e} m: {ex}", flush=True)
    s.close()
    hc.close()
