import sys, time, os
for p in ("dg-solver-for-pbte-with-mfem_b200","tests","oracle","."): sys.path.insert(0,p)
import numpy as np, pbte_b200 as pb
from cases import BC3
hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, 16, True, False, BC3, num_spectral=1); P=hc.problem
ref=None
for policy in (0,1):
    s=pb.Solver(P, kernel=2, cache_policy=policy); s.build_schedule(); s.solve(3,0.0)
    got=s.download_moments()
    t=[]
    for i in range(6):
        s.iterate_async(); s.iterate_sync(); t.append(s.stats().last_sweep_ms)
    st=s.stats(); print(sys.argv[1],"policy",policy,"kernel",st.kernel_used,"sweep ms %.3f"%min(t),"updates/s %.3e"%(P.ndof*P.ne*P.ndir*P.nband/(min(t)*1e-3)), "Tc checksum %.17g"%float(np.abs(got["Tc"]).sum()))
    s.close()
