"""GPU parity at BASELINE.json's benchmark sizes (-m gpu): the CUDA path against the CPU oracle, iteration by iteration.

The threaded oracle needs ~1 s per source iteration of full C2 on the GPU box's host cores, so the configuration the
benchmark times (4096 hexes order 2, 128 directions, 2 bands: 46 wavefront levels, 532 tiles per slot, ring wrap-around,
tile watermarks, work split over persistent CTAs) is compared directly, with every sweep kernel of the library:
  kernel 3 (shared-operator marching, both variants: operators in shared memory / in registers) and kernel 2 (per-element LU).
C3 and C5 are compared one refinement level down (24 576 tetrahedra order 2; 4096 quads order 4) with their own quadratures.
Bar: Tc, Qc, Tv, Qv and the whole angular flux within 1e-10 relative (max norm) after EVERY iteration; with the reference's
own face rule (quirk Q1) the bound follows the measured conditioning of the local systems (tests/parity_tol.py).
"""
import os

import numpy as np
import pytest

import oracle as orc
import pbte_b200 as pb
from cases import BC2, BC3
from parity_tol import parity_tolerance

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8
TOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def stepwise_parity(P, iters, tol=TOL, **solver_kw):
    """Iteration by iteration from the same state: oracle and GPU both advance `it` -> `it + 1` and are compared on every field."""
    s = pb.Solver(P, **solver_kw)
    s.build_schedule()
    coeff = np.zeros((P.ndir, P.nband, P.ndof * P.ne))
    Tc, Tv = np.zeros(P.ndof * P.ne), np.zeros(P.ne)
    worst = {}
    for it in range(iters):
        want = orc.solve(P, 1, 0.0, cache_lu=False, threads=THREADS, coeff=coeff, Tc=Tc, Tv=Tv)
        s.sweep()
        r = s.finish_iteration()
        got = s.download_moments()
        got["coeff"] = s.download_all_coeff()
        for k in ("Tc", "Qc", "Tv", "Qv", "coeff"):
            worst[k] = max(worst.get(k, 0.0), rel(got[k], want[k]))
        worst["residual"] = max(worst.get("residual", 0.0), abs(r - want["residual"]) / want["residual"])
        # the GPU continues from ITS OWN state; the oracle is restarted from the GPU's state, so errors cannot hide by drifting together
        coeff, Tc, Tv = got["coeff"], got["Tc"], got["Tv"]
    st = s.stats()
    s.close()
    print(f"stepwise parity, kernel {st.kernel_used} variant {st.march_variant}, {iters} iterations, tol {tol:.2e}: worst relative errors " +
          " ".join(f"{k}={v:.2e}" for k, v in worst.items()))
    for k in ("Tc", "Qc", "Tv", "Qv", "coeff"):
        assert worst[k] <= tol, (k, worst[k], tol)
    assert worst["residual"] <= max(1e-9, 10 * tol)
    return st, worst


@pytest.fixture(scope="module")
def c2():
    hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, 16, True, False, BC3, num_spectral=1)
    yield hc
    hc.close()


@pytest.fixture(scope="module")
def c2_consistent():
    hc = pb.HostCase("unit-cube-hex", 2, 2, 3, 8, 16, True, False, BC3, face_rule=1, num_spectral=1)
    yield hc
    hc.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_c2_full_size_marching_kernel_vs_oracle(c2, variant, monkeypatch):
    """Full C2 with the reference's own element matrices (quirk Q1), three iterations, both marching variants."""
    monkeypatch.setenv("PBTE_MARCH_REGOP", str(variant))
    P = c2.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (4096, 27, 128, 2)
    st, worst = stepwise_parity(P, 3, tol=parity_tolerance(P, "C2 full size, reference face rule"), kernel=3)
    assert st.kernel_used == 3 and st.march_variant == variant


def test_c2_full_size_per_element_kernel_vs_oracle(c2):
    st, worst = stepwise_parity(c2.problem, 2, tol=parity_tolerance(c2.problem, "C2 full size, reference face rule"), kernel=2)
    assert st.kernel_used == 2


def test_c2_full_size_consistent_traces_vs_oracle(c2_consistent):
    """The form the time-to-converged-T leg of the benchmark runs (consistent face traces), library's own kernel choice."""
    st, worst = stepwise_parity(c2_consistent.problem, 3, kernel=0)
    assert st.kernel_used == 3


def test_c2_full_size_solve_loop_vs_oracle(c2):
    """pbte_cuda_solve (iterations in flight, device-side stop test) against three oracle iterations in one go."""
    P = c2.problem
    want = orc.solve(P, 3, 0.0, cache_lu=False, threads=THREADS)
    s = pb.Solver(P)
    s.build_schedule()
    it, res, hist = s.solve(3, 0.0)
    got = s.download_moments()
    tol = parity_tolerance(P, "C2 full size, reference face rule")
    assert it == 3 and np.allclose(hist, want["residuals"], rtol=max(1e-9, 10 * tol), atol=0)
    for k in ("Tc", "Qc", "Tv", "Qv"):
        assert rel(got[k], want[k]) <= 3 * tol, (k, rel(got[k], want[k]))  # three iterations without a restart from the same state
    assert rel(s.download_all_coeff(), want["coeff"]) <= 3 * tol
    s.close()


def test_c3_one_level_down_vs_oracle():
    """C3 one refinement level down: 24 576 tetrahedra order 2 (72 operator classes as at full size), 4 x 4 directions, 4 bands."""
    hc = pb.HostCase("unit-cube-tet", 2, 2, 3, 4, 4, True, False, BC3, num_spectral=2)
    P = hc.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (24576, 10, 16, 4)
    tol = parity_tolerance(P, "C3 one level down")
    st, worst = stepwise_parity(P, 2, tol=tol, kernel=3)
    assert st.kernel_used == 3 and st.n_operator_classes == 72
    st, worst = stepwise_parity(P, 1, tol=tol, kernel=2)
    hc.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_c5_one_level_down_vs_oracle(variant, monkeypatch):
    """C5 two refinement levels down: 4096 quads order 4 (ndof 25), 64 uniform azimuth points, 4 Si bands, consistent traces
    (with the reference's face rule the order-4 local systems are nearly singular: tests/test_gpu_parity.py states that bound)."""
    monkeypatch.setenv("PBTE_MARCH_REGOP", str(variant))
    hc = pb.HostCase("unit-square-quad", 4, 3, 2, 1, 64, True, False, BC2, face_rule=1, num_spectral=2)
    P = hc.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (4096, 25, 64, 4)
    st, worst = stepwise_parity(P, 3, tol=parity_tolerance(P, "C5 down-scaled, consistent traces"), kernel=3)
    assert st.kernel_used == 3 and st.march_variant == variant
    hc.close()


def test_c3_full_size_mesh_vs_oracle():
    """The C3 MESH at its benchmark size (196 608 tetrahedra order 2: the 72 operator classes, list-scheduled tiles at their real fill)
    with a quadrature the oracle finishes in seconds (2 x 4 directions, 2 bands): one iteration with the marching kernel."""
    hc = pb.HostCase("unit-cube-tet", 2, 3, 3, 2, 4, True, False, BC3, num_spectral=1)
    P = hc.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (196608, 10, 8, 2)
    st, worst = stepwise_parity(P, 1, tol=parity_tolerance(P, "C3 full-size mesh"), kernel=3)
    assert st.kernel_used == 3 and st.n_operator_classes == 72 and st.march_fill > 0.9
    hc.close()


def test_c5_full_size_mesh_vs_oracle():
    """The C5 MESH at its benchmark size (65 536 quads order 4) with 8 azimuth points and 2 bands, consistent traces: two iterations
    with the library's own kernel choice (the persistent marching kernel when there are more slots than resident CTAs, else the classic)."""
    hc = pb.HostCase("unit-square-quad", 4, 5, 2, 1, 8, True, False, BC2, face_rule=1, num_spectral=1)
    P = hc.problem
    assert (P.ne, P.ndof, P.ndir, P.nband) == (65536, 25, 8, 2)
    st, worst = stepwise_parity(P, 2, tol=parity_tolerance(P, "C5 full-size mesh, consistent traces"), kernel=0)
    assert st.kernel_used == 3
    hc.close()


@pytest.mark.parametrize("peer,fused", [("1", "1"), ("1", "0"), ("0", "1")])
def test_two_devices_equal_one_device(c2_consistent, peer, fused, monkeypatch):
    """Two handles on two GPUs of one process, directions split in halves, the exchange done by the library — over NVLink peer memory
    inside the one tail launch of an iteration (the default), in a finish kernel of its own (PBTE_PEER_FUSED=0), or as an NCCL
    all-reduce (PBTE_PEER_ALLREDUCE=0): same iterates as one handle on one GPU (the sum order of the moments differs: 1e-13)."""
    if pb.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    monkeypatch.setenv("PBTE_PEER_ALLREDUCE", peer)
    monkeypatch.setenv("PBTE_PEER_FUSED", fused)
    P = c2_consistent.problem  # well-conditioned local systems: the two summation orders agree to rounding
    one = pb.Solver(P, device=0)
    one.build_schedule()
    it1, res1, hist1 = one.solve(4, 0.0)
    want = one.download_moments()
    halves = []
    for r in range(2):
        s = pb.Solver(P, device=r, dir_range=pb.partition_directions(P.ndir, 2, r))
        s.build_schedule()
        halves.append(s)
    pb.comm_attach_all(halves)
    seen = []
    it2, res2, hist2 = pb.solve_multi(halves, 4, 0.0, on_iteration=lambda i, res, tv, tc: seen.append((i, res)))
    assert it2 == 4 and [i for i, _ in seen] == [1, 2, 3, 4]
    assert np.allclose(hist2, hist1, rtol=1e-11, atol=0)
    for s in halves:
        got = s.download_moments()
        for k in ("Tc", "Qc", "Tv", "Qv"):
            assert rel(got[k], want[k]) <= 1e-12, (k, rel(got[k], want[k]))
    k0, k1 = halves[1].dir_range
    assert rel(halves[1].download_coeff(k0, 1), one.download_coeff(k0, 1)) <= 1e-12  # a slot's flux sees the partition only through Tc
    # the stop test: every device forms the same sums, so all of them stop at the iteration one device stops at
    if hist1[2] < min(hist1[0], hist1[1]):
        tol_stop = 0.5 * (hist1[2] + min(hist1[0], hist1[1]))
        for s in halves:
            s.reset_state()
        it3, res3, hist3 = pb.solve_multi(halves, 10, tol_stop)
        assert it3 == 3 and abs(res3 - hist1[2]) <= 1e-11 * hist1[2]
        for s in halves:
            assert rel(s.download_moments()["Tc"], halves[0].download_moments()["Tc"]) == 0.0  # identical on every device
    for s in halves:
        s.close()
    one.close()
