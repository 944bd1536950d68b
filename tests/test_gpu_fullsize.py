"""GPU tests at BASELINE.json's benchmark size (config C2: 4096 hexes, order 2, 128 directions, 2 bands) through
size-independent properties — the CPU oracle needs minutes per iteration there."""
import numpy as np
import pytest

import pbte_b200 as pb
from cases import BC3

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, 16, True, False, BC3, num_spectral=1)
    yield hc
    hc.close()


def test_c2_orders_are_topological_and_sorted_by_pass(c2):
    P = c2.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (4096, 27, 128, 2)
    s = pb.Solver(P)
    nclass, nlev = s.build_schedule()
    assert nclass == 8 and nlev == 46  # 8 octants; 3*(16-1)+1 wavefronts
    for k in (0, 17, 64, 127):
        order, lev = s.order(k), s.levels(k)
        assert sorted(order) == list(range(P.ne))
        pos = np.empty(P.ne, dtype=np.int64)
        pos[order] = np.arange(P.ne)
        a = np.einsum("efd,d->ef", P.face_normal, P.dir[k])
        for e in range(0, P.ne, 7):
            for f in range(P.nfaces):
                nb = P.face_nbr[e, f]
                if nb >= 0 and a[e, f] < 0:
                    assert pos[nb] < pos[e] and lev[nb] < lev[e]
    s.close()


@pytest.mark.parametrize("kernel", [2, 3])
def test_c2_linearity_and_partition_invariance(c2, kernel):
    """From a cold start the iterate is linear in the wall temperatures: doubling them doubles every number exactly
    (power-of-two scaling commutes with FP64 rounding); two direction halves sum to the full moments."""
    P = c2.problem
    s = pb.Solver(P, kernel=kernel)
    s.build_schedule()
    _, _, h1 = s.solve(3, 0.0)
    m1 = s.download_moments()
    u1 = s.download_coeff(37, 1)
    s.close()
    P2 = pb.ProblemArrays(P.dim, P.ne, P.ndof, P.nfaces, P.ndir, P.nband, P.omega, P.heat_cap_v,
                          **{f: getattr(P, f) for f in pb.ProblemArrays.FIELDS if f != "face_tbc"}, face_tbc=2.0 * P.face_tbc)
    s = pb.Solver(P2, kernel=kernel)
    s.build_schedule()
    _, _, h2 = s.solve(3, 0.0)
    m2 = s.download_moments()
    assert np.array_equal(m2["Tc"], 2.0 * m1["Tc"]) and np.array_equal(m2["Qc"], 2.0 * m1["Qc"])
    assert np.array_equal(s.download_coeff(37, 1), 2.0 * u1)
    assert np.array_equal(h1, h2)
    s.close()
    # one sweep split over two direction blocks
    full = pb.Solver(P, kernel=kernel)
    full.build_schedule()
    full.sweep()
    want = full.download_moments()
    full.close()
    tc = np.zeros_like(want["Tc"])
    for r in range(2):
        sv = pb.Solver(P, kernel=kernel, dir_range=pb.partition_directions(P.ndir, 2, r))
        sv.build_schedule()
        sv.sweep()
        tc += sv.download_moments()["Tc"]
        sv.close()
    assert np.abs(tc - want["Tc"]).max() <= 1e-13 * np.abs(want["Tc"]).max()


@pytest.mark.parametrize("kernel", [2, 3])
def test_c2_local_systems_are_solved_to_backward_stability(c2, kernel):
    """Spot check of A u = r on sampled (direction, band, element) systems after one sweep, assembled independently here."""
    P = c2.problem
    n, F, dim = P.ndof, P.nfaces, P.dim
    s = pb.Solver(P, kernel=kernel)
    s.build_schedule()
    s.sweep()
    assert s.stats().kernel_used == kernel
    nu_max = P.inv_kn.max()
    rng = np.random.default_rng(3)
    for _ in range(12):
        k, b, e = int(rng.integers(P.ndir)), int(rng.integers(P.nband)), int(rng.integers(P.ne))
        u = s.download_coeff(k, b).reshape(P.ne, n)
        M = P.mass[e].reshape(n, n).T
        A = nu_max * M
        for d in range(dim):
            A = A - P.vg[b] * P.dir[k, d] * P.stiff[e, d].reshape(n, n).T
        rhs = np.zeros(n)  # cold start: T_old = 0, u_old = 0
        for f in range(F):
            a = float(P.face_normal[e, f] @ P.dir[k])
            if a > 0:
                A = A + P.vg[b] * a * P.face_mass[e, f].reshape(n, n).T
            elif a < 0:
                cin = P.vg[b] * a
                nb = P.face_nbr[e, f]
                if nb >= 0:
                    rhs -= cin * (P.face_coupling[e, f].reshape(n, n).T @ u[nb])
                else:
                    rhs += -cin * P.heat_cap[b] / P.omega * P.face_tbc[e, f] * P.face_integral[e, f]
        r = A @ u[e] - rhs
        assert np.abs(r).max() <= 1e-12 * (np.abs(A).max() * np.abs(u[e]).max() * n + np.abs(rhs).max())
    s.close()
