"""CPU tests (no GPU) of the shared-operator planner: csrc/shared_plan.hpp (operator classes, canonical faces, tile schedule)
is exercised through tests/native/plan_check.cpp, a CPU emulation that walks the schedule exactly as sweep_march_kernel does,
and the result is compared with the oracle.  This pins the planner and the operator algebra; the kernel itself is pinned by
tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as orc
from cases import CASES, ROOT, host_case

NATIVE = os.path.join(ROOT, "tests", "native")
PKG = os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200")


@pytest.fixture(scope="module")
def plan_lib():
    so = os.path.join(NATIVE, "libplan_check.so")
    srcs = [os.path.join(NATIVE, "plan_check.cpp"), os.path.join(PKG, "csrc", "shared_plan.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(PKG, "csrc"), srcs[0], "-o", so])
    return C.CDLL(so)


def run_plan(lib, P, iters, tol=1e-12, max_classes=512, want_coeff=False):
    block = P.ndof * P.ne
    Tc, Qc = np.zeros(block), np.zeros((P.dim, block))
    coeff = np.zeros((P.ndir, P.nband, block)) if want_coeff else None
    info = np.zeros(8, dtype=np.int32)
    fill = C.c_double(0)
    st = P.struct()
    dp = C.POINTER(C.c_double)
    rc = lib.plan_check_run(C.byref(st), C.c_double(tol), max_classes, iters, Tc.ctypes.data_as(dp), Qc.ctypes.data_as(dp),
                            coeff.ctypes.data_as(dp) if want_coeff else None, info.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(fill))
    return rc, dict(Tc=Tc, Qc=Qc, coeff=coeff, noclass=int(info[0]), naclass=int(info[1]), nclass=int(info[2]), ntiles=int(info[3]),
                    max_levels=int(info[4]), close8=int(info[5]), close16=int(info[6]), mask_mismatch=int(info[7]), fill=fill.value)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# expected operator classes on the uniform builtin meshes: one for Cartesian quads/hexes, few for simplices
EXPECT_CLASSES = {"c1_tri_p1": 2, "quad_p2": 1, "tet_p1": 6, "tet_p2": 6, "hex_p1": 1, "c2_hex_p2_small": 1, "quad_p1_dim3": 1, "tri_p3_si": 2,
                  "c0_default": 2, "c0_refined_dim3": 2}


@pytest.mark.parametrize("name", ["c0_default", "c1_tri_p1", "quad_p2", "tet_p1", "hex_p1", "quad_p1_dim3", "tri_p3_si", "tet_p2", "c2_hex_p2_small"])
def test_marching_plan_reproduces_oracle(plan_lib, name):
    hc = host_case(name)
    P = hc.problem
    iters = min(CASES[name]["iters"], 6)
    rc, r = run_plan(plan_lib, P, iters, want_coeff=True)
    assert rc == 0
    assert r["noclass"] == EXPECT_CLASSES[name]
    o = orc.solve(P, iters, 0.0, cache_lu=True, threads=2)
    assert rel(r["Tc"], o["Tc"]) < 1e-11
    assert rel(r["Qc"], o["Qc"]) < 1e-10
    assert rel(r["coeff"], o["coeff"]) < 1e-10
    hc.close()


def test_refined_meshes_keep_few_classes(plan_lib):
    for name, refine, bound in (("hex_p1", 1, 1), ("tet_p1", 1, 64), ("c1_tri_p1", 1, 4), ("quad_p2", 2, 1)):
        hc = host_case(name, refine=refine)
        rc, r = run_plan(plan_lib, hc.problem, 0)
        assert rc == 0 and 1 <= r["noclass"] <= bound, (name, r)
        assert r["fill"] > 0.2
        hc.close()


def test_distorted_mesh_is_declined(plan_lib):
    """Random per-element scaling of the matrices -> every element is its own class -> the planner gives up cleanly."""
    hc = host_case("hex_p1", refine=1)
    P = hc.problem
    rng = np.random.default_rng(7)
    P.mass *= (1.0 + 0.1 * rng.random((P.ne, 1)))
    rc, r = run_plan(plan_lib, P, 0, max_classes=64)
    assert rc == 1 and r["noclass"] == -1
    hc.close()


def test_class_specific_couplings_split_classes(plan_lib):
    """Elements sharing A but not K_f must land in different operator classes (and still reproduce the oracle)."""
    hc = host_case("quad_p2")
    P = hc.problem
    P.face_coupling[: P.ne // 2] *= 1.25   # half of the elements get a different upwind coupling
    rc, r = run_plan(plan_lib, P, 3, want_coeff=True)
    assert rc == 0 and r["naclass"] == 1 and r["noclass"] > 1
    o = orc.solve(P, 3, 0.0, cache_lu=True, threads=2)
    assert rel(r["coeff"], o["coeff"]) < 1e-10
    hc.close()


@pytest.mark.parametrize("delay", ["0", "3", "8"])
@pytest.mark.parametrize("name,refine", [("tet_p2", 1), ("c2_hex_p2_small", 0), ("c1_tri_p1", 1), ("quad_p2", 1)])
def test_list_scheduled_tiles_reproduce_oracle(plan_lib, name, refine, delay, monkeypatch):
    """Tiles filled across wavefront levels (list scheduling, shared_plan.hpp: schedule_class_merged) instead of level by level:
    walked in tile order, every upwind neighbour must already be computed, so the sweep still reproduces the oracle; the tiles are
    at least as full as those of the level schedule when no eligibility delay is asked for."""
    hc = host_case(name, refine=refine)
    P = hc.problem
    monkeypatch.setenv("PBTE_TILE_MERGE", "0")
    rc, by_level = run_plan(plan_lib, P, 0)
    monkeypatch.setenv("PBTE_TILE_MERGE", "1")
    monkeypatch.setenv("PBTE_TILE_DELAY", delay)
    rc, r = run_plan(plan_lib, P, 2, want_coeff=True)
    assert rc == 0
    if delay == "0":
        assert r["ntiles"] <= by_level["ntiles"]
    o = orc.solve(P, 2, 0.0, cache_lu=False, threads=4)
    assert rel(r["Tc"], o["Tc"]) < 1e-11 and rel(r["coeff"], o["coeff"]) < 1e-10
    hc.close()


def test_axis_aligned_direction_and_sign_flip_inside_a_class(plan_lib):
    """A direction exactly tangential to faces of a Cartesian mesh is consistent (n.Omega == 0 for every element of the class); an
    element whose normal differs from its class representative's by less than the class tolerance but flips the sign of n.Omega is
    counted, so that the library declines the shared-operator path instead of dropping that element's inflow term."""
    hc = host_case("quad_p2")
    P = hc.problem
    P.dir[0] = (0.0, 1.0, 0.0)   # tangential to the x-faces
    P.dir[1] = (1.0, 0.0, 0.0)
    rc, r = run_plan(plan_lib, P, 2, want_coeff=True)
    assert rc == 0 and r["mask_mismatch"] == 0
    o = orc.solve(P, 2, 0.0, cache_lu=False, threads=2)
    assert rel(r["coeff"], o["coeff"]) < 1e-10
    # one element's x-normals tilt by 1e-14: same operator class (tolerance 1e-12), but for direction 0 its n.Omega is -1e-14 / +1e-14
    e = P.ne // 2
    for f in range(P.nfaces):
        if abs(P.face_normal[e, f, 0]) == 1.0:
            P.face_normal[e, f, 1] = -1e-14 * P.face_normal[e, f, 0]
    rc, r = run_plan(plan_lib, P, 0)
    assert r["noclass"] == 1 and r["mask_mismatch"] > 0
    hc.close()
