"""Development probe: sweep throughput of the marching kernel (and, optionally, the per-element LU kernel) on scaled-down
versions of the BASELINE configs C3-C5 on one GPU.  Prints one line per configuration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import pbte_b200 as pb
from cases import BC2, BC3

CONFIGS = {
    # name: (mesh, order, refine, angle_dim, polar, azimuth, bcs, num_spectral)
    "c3_tet_p2_r2_40bands_16dirs": ("unit-cube-tet", 2, 2, 3, 2, 8, BC3, 20),     # C3 slice of one GPU (16 directions x 40 bands), one level coarser
    "c4_hex_p3_r1_40bands_32dirs": ("unit-cube-hex", 3, 1, 3, 4, 8, BC3, 20),     # C4 two levels coarser
    "c5_quad_p4_r3_40bands_64dirs": ("unit-square-quad", 4, 3, 2, 1, 64, BC2, 20),  # C5 two levels coarser
    "c2_hex_p2_r2": ("unit-cube-hex", 2, 2, 3, 8, 16, BC3, 1),
    # full-size single-GPU points
    "c3_tet_p2_r3_40bands_16dirs_FULL_GPU_SLICE": ("unit-cube-tet", 2, 3, 3, 2, 8, BC3, 20),   # exactly one GPU's share of C3: 196608 tets, 640 slots, 10 GB
    "c4_hex_p3_r2_40bands_512dirs_1GPU_POINT": ("unit-cube-hex", 3, 2, 3, 16, 32, BC3, 20),    # C4's 1-GPU strong-scaling point: 4096 hexes, 43 GB
    "c5_quad_p4_r5_40bands_64dirs_FULL": ("unit-square-quad", 4, 5, 2, 1, 64, BC2, 20),        # all of C5: 65536 quads, 2560 slots, 33.5 GB
}
names = sys.argv[1:] or list(CONFIGS)
for name in names:
    mesh, order, refine, adim, polar, azim, bcs, nspec = CONFIGS[name]
    t0 = time.time()
    hc = pb.HostCase(mesh, order, refine, adim, polar, azim, True, False, bcs, num_spectral=nspec)
    P = hc.problem
    t_host = time.time() - t0
    for kernel in (3, 2):
        if kernel == 2 and (P.ndof > 32 or os.environ.get("PROBE_SKIP_LU")):
            continue
        t0 = time.time()
        try:
            s = pb.Solver(P, kernel=kernel)
            s.build_schedule()
        except pb.PbteError as ex:
            print(f"{name}: kernel {kernel} unavailable: {ex}", flush=True)
            continue
        t_setup = time.time() - t0
        ms = []
        for i in range(5):
            s.sweep(); s.finish_iteration()
            ms.append(s.stats().last_sweep_ms)
        st = s.stats()
        t = float(np.median(ms[2:]))
        tm = st.last_march_ms
        U = P.updates_per_sweep
        print(f"{name}: ne {P.ne} ndof {P.ndof} slots {P.ndir * P.nband} kernel {st.kernel_used} dag-classes {st.n_classes} levels {st.max_levels} "
              f"op-classes {st.n_operator_classes} fill {st.march_fill:.2f} flux {st.flux_bytes / 1e9:.2f} GB ops {st.operator_bytes / 1e6:.0f} MB | "
              f"variant {st.march_variant} host {t_host:.1f}s setup {t_setup:.1f}s sweep {t:.3f} ms (marching kernel {tm:.3f}) -> {U / t / 1e6:.1f} M updates/ms", flush=True)
        s.close()
    hc.close()
