"""CPU tests: the C-ABI library loads and exports every declared symbol; host-side logic (no compute calls on a device)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pbte_b200 as pb
from cases import host_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pbte_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(pbte_cuda_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(pb.CUDA_ABI_SYMBOLS)
    lib = pb.cuda_lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pbte_cuda_abi_version() == 1


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations: 8 int32 + 9 pointers + ... (natural alignment)
    assert C.sizeof(pb.Problem) == 8 * 4 + 9 * 8 + 2 * 8 + 8 + 4 * 8 + 8
    assert C.sizeof(pb.Options) == 6 * 4 + 8 + 8 + 4 + 5 * 4
    assert C.sizeof(pb.Stats) == 2 * 8 + 2 * 8 + 2 * 8 + 6 * 4 + 8 + 8 + 8 + 2 * 4 + 8 + 8 + 2 * 4
    assert pb.Options.class_tol.offset == 32 and pb.Stats.march_fill.offset == 72


def test_no_device_fails_loudly():
    """Without a GPU every compute entry point reports PBTE_ERR_CUDA: there is no CPU fallback."""
    if pb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    hc = host_case("c0_default")
    with pytest.raises(pb.PbteError) as ei:
        pb.Solver(hc.problem)
    assert ei.value.status == -2
    with pytest.raises(pb.PbteError):
        hc.solve(1)
    hc.close()


def test_partition_directions_covers_everything():
    for ndir in (0, 1, 7, 16, 128, 513):
        for ws in (1, 2, 3, 8):
            parts = [pb.partition_directions(ndir, ws, r) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == ndir
            for a, b in zip(parts[:-1], parts[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_face_slots_are_ascending_global_face_ids():
    """Interior couplings are mutual and normals opposite: the topology part of the problem image is consistent."""
    for name in ("c1_tri_p1", "tet_p1", "hex_p1"):
        hc = host_case(name)
        P = hc.problem
        for e in range(P.ne):
            for f in range(P.nfaces):
                nb = P.face_nbr[e, f]
                if nb >= 0:
                    back = [g for g in range(P.nfaces) if P.face_nbr[nb, g] == e]
                    assert len(back) >= 1
                    assert any(np.array_equal(P.face_normal[nb, g], -P.face_normal[e, f]) for g in back)
        # unit normals
        used = P.face_nbr != -2
        assert np.allclose(np.linalg.norm(P.face_normal[used], axis=1), 1.0, atol=1e-14)
        hc.close()


def test_element_matrices_sanity():
    """Mass matrix sums to the element volume, basis integrals too; stiffness rows sum to zero (d/dx of a constant)."""
    for name in ("c1_tri_p1", "quad_p2", "tet_p2", "c2_hex_p2_small"):
        hc = host_case(name)
        P = hc.problem
        n = P.ndof
        vol = P.basis_integrals.sum(axis=1)
        assert np.allclose(P.mass.sum(axis=1), vol, rtol=1e-12)
        total = 1e-6 ** P.dim
        assert abs(vol.sum() / total - 1) < 1e-12
        G = P.stiff.reshape(P.ne, P.dim, n, n)  # [j][i] column-major: G[e,d,j,i] = int d_d(phi_i) phi_j
        assert np.abs(G.sum(axis=3)).max() <= 1e-12 * np.abs(G).max()
        hc.close()


def test_consistent_face_rule_gives_convergent_iteration():
    """With mathematically consistent trace integrals the source iteration contracts (the reference's default face rule,
    quirk Q1, does not on quads/hexes); this pins the optional rule of the host integrator."""
    import oracle as orc
    hc = host_case("quad_p2", face_rule=1)
    r = orc.solve(hc.problem, 12, 0.0)
    res = r["residuals"]
    assert all(res[i + 1] < res[i] for i in range(len(res) - 1))
    assert r["Tc"].min() > -0.51 and r["Tc"].max() < 0.51
    hc.close()
