"""Development probe (GPU): per-kernel error against the oracle on the ill-conditioned Q1 cases (quad order 4, hex order 3),
next to the condition numbers of their local systems.  Feeds the per-kernel tolerances of tests/test_gpu_parity.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("dg-solver-for-pbte-with-mfem_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np

import oracle as orc
import pbte_b200 as pb
from cases import CASES, host_case


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


for name in ("quad_p4", "hex_p3", "c2_hex_p2_small", "quad_p2"):
    hc = host_case(name)
    P = hc.problem
    iters = CASES[name]["iters"]
    conds = []
    rng = np.random.default_rng(1)
    for _ in range(64):
        k, b, e = int(rng.integers(P.ndir)), int(rng.integers(P.nband)), int(rng.integers(P.ne))
        conds.append(np.linalg.cond(orc.assemble_A(P, k, b, e)))
    print(f"{name}: ndof {P.ndof}, cond(A) over 64 sampled systems: median {np.median(conds):.3e} max {np.max(conds):.3e}", flush=True)
    for per_iter in (False, True):
        for kernel, variant in ((1, None), (2, None), (3, 0), (3, 1)):
            if kernel == 2 and P.ndof > 32:
                continue
            if variant is not None:
                os.environ["PBTE_MARCH_REGOP"] = str(variant)
            else:
                os.environ.pop("PBTE_MARCH_REGOP", None)
            s = pb.Solver(P, kernel=kernel)
            s.build_schedule()
            if not per_iter:
                want = orc.solve(P, iters, 0.0, cache_lu=True, threads=8)
                s.solve(iters, 0.0)
                got = s.download_moments()
                errs = {k: rel(got[k], want[k]) for k in ("Tc", "Qc")}
                errs["coeff"] = rel(s.download_all_coeff(), want["coeff"])
            else:
                coeff = np.zeros((P.ndir, P.nband, P.ndof * P.ne))
                Tc, Tv = np.zeros(P.ndof * P.ne), np.zeros(P.ne)
                errs = {"Tc": 0.0, "Qc": 0.0, "coeff": 0.0}
                for it in range(iters):
                    want = orc.solve(P, 1, 0.0, cache_lu=False, threads=8, coeff=coeff, Tc=Tc, Tv=Tv)
                    s.sweep()
                    s.finish_iteration()
                    got = s.download_moments()
                    got["coeff"] = s.download_all_coeff()
                    for k in errs:
                        errs[k] = max(errs[k], rel(got[k], want[k]))
                    coeff, Tc, Tv = got["coeff"], got["Tc"], got["Tv"]
            s.close()
            print(f"  {'per-iteration (same state)' if per_iter else 'whole run'} kernel {kernel} variant {variant}: " +
                  " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
    hc.close()
