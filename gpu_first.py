import sys, time, numpy as np
sys.path.insert(0, "dg-solver-for-pbte-with-mfem_b200"); sys.path.insert(0, "oracle")
import pbte_b200 as pb, oracle as orc
print("devices", pb.device_count())
print("peaks", pb.measure_peaks())
bc3={1:-0.5,2:-0.5,3:-0.5,4:-0.5,5:-0.5,6:0.5}; bc2={1:-0.5,2:-0.5,3:0.5,4:-0.5}
cases=[("unit-square-tri",1,0,2,1,16,1,bc2),("unit-square-quad",2,0,2,1,8,2,bc2),("unit-cube-tet",2,0,3,2,4,1,bc3),("unit-cube-hex",2,0,3,4,8,1,bc3),("unit-square-quad",4,1,2,1,8,1,bc2),("unit-cube-hex",3,0,3,2,4,1,bc3)]
for mesh,order,ref,ad,npol,naz,nspec,bcs in cases:
    hc=pb.HostCase(mesh, order, ref, ad, npol, naz, True, False, bcs, num_spectral=nspec)
    P=hc.problem
    o=orc.solve(P, 3, 0.0, True, 8)
    oo=orc.sweep_orders(P)
    for kern,pol in [(1,1),(2,0),(2,1)]:
        try:
            s=pb.Solver(P, kernel=kern, cache_policy=pol)
        except pb.PbteError as e:
            print(mesh, order, "kernel",kern,"unavailable:", e); continue
        try:
            nc,ml=s.build_schedule()
            ok=all((s.order(k)==oo[k]).all() for k in range(P.ndir))
            it,res,hist=s.solve(3,0.0)
            m=s.download_moments(); cf=s.download_all_coeff()
            rel=lambda a,b: np.abs(a-b).max()/np.abs(b).max()
            st=s.stats()
            print(f"{mesh} p{order} n={P.ndof} ne={P.ne} kern={kern} pol={pol} classes={nc} levels={ml} orders_ok={ok} Tc {rel(m['Tc'],o['Tc']):.2e} Qc {rel(m['Qc'],o['Qc']):.2e} coeff {rel(cf,o['coeff']):.2e} res {hist} vs {o['residuals']} sweep_ms {st.last_sweep_ms:.3f}")
        except pb.PbteError as e:
            print(mesh, order, "kernel",kern,"FAILED:", e)
        s.close()
