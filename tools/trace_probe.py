"""Development probe: per-tile clock stamps of CTA 0 of the marching kernel on the C2 mesh (needs the -DPBTE_TRACE build of the
CUDA library: PBTE_LIB_DIR=tools/_trace).  usage: trace_probe.py <azimuth> <out file> [tet]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pbte_b200 as pb
from cases import BC3

azim, out = int(sys.argv[1]), sys.argv[2]
if len(sys.argv) > 3 and sys.argv[3] == "tet":  # a slice of C3 one level coarser: 24576 tets order 2, 2 x azim directions x 40 bands
    hc = pb.HostCase("unit-cube-tet", 2, 2, 3, 2, azim, True, False, BC3, num_spectral=20)
else:
    hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, azim, True, False, BC3, num_spectral=1)
s = pb.Solver(hc.problem, kernel=3)
s.build_schedule()
for i in range(4):
    s.sweep(); s.finish_iteration()
os.environ["PBTE_TRACE_FILE"] = out
s.sweep(); s.finish_iteration()
del os.environ["PBTE_TRACE_FILE"]
st = s.stats()
print("variant", st.march_variant, "sweep ms (traced)", st.last_sweep_ms)
s.close(); hc.close()
