"""Generates the golden fixtures of tests/golden/ from the REFERENCE's own code.

Run in the build container (needs /root/reference and oracle/_ref/libpbte_ref.so, built by `make -C oracle`):
    python tests/golden/make_golden.py
Every fixture <case>.npz holds what the reference computes for the named case of tests/cases.py:
    orders  [ndir][ne] int32      AngularSweepOrder::Build
    Tc Qc Tv Qv coeff residual    PBTESolver::Solve after `iters` iterations from a cold start (FullLU policy)
    img_*                         the flat problem image taken from the reference's objects (small cases only)
Also decodes the reference's shipped ParaView file and sweep log into shipped_default.npz.
"""
import base64
import os
import re
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
import pbte_b200 as pb  # noqa: E402
from cases import CASES  # noqa: E402

REF = "/root/reference"
IMAGE_LIMIT = 600_000  # doubles; larger images are not stored (the host layer regenerates them bit-exactly)


def make_case(name, c):
    mat = dict(pb.SI_MATERIAL)
    mat["num_spectral"] = c["num_spectral"]
    rc = orc.RefCase(c["mesh"], c["order"], c["refine"], c["angle_dim"], c["polar"], c["azimuth"], c["polar_gauss"], c["azimuth_gauss"], c["bcs"], mat)
    out = dict(orders=rc.orders())
    sol = rc.solve(c["iters"], 0.0, 0)
    for k in ("Tc", "Qc", "Tv", "Qv", "residual"):
        out[k] = sol[k]
    if sol["coeff"].size <= IMAGE_LIMIT:
        out["coeff"] = sol["coeff"]
    img = rc.arrays()
    if sum(np.size(v) for v in img.values()) <= IMAGE_LIMIT:
        for k, v in img.items():
            out["img_" + k] = v
    else:  # keep the small tables and a checksum of the big ones
        for k in ("dir", "dir_w", "inv_kn", "vg", "heat_cap", "freq_w", "omega", "heat_cap_v"):
            out["img_" + k] = img[k]
        for k in ("mass", "stiff", "face_mass", "face_coupling", "face_normal", "face_integral", "basis_integrals"):
            out["sum_" + k] = np.array([img[k].sum(), np.abs(img[k]).sum()])
        out["img_face_nbr"] = img["face_nbr"]
    out["sizes"] = np.array([rc.dim, rc.ne, rc.ndof, rc.nfaces, rc.ndir, rc.nband], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: ne={rc.ne} ndof={rc.ndof} ndir={rc.ndir} nband={rc.nband} residual={sol['residual']:.6e}")
    rc.close()


def decode_vtu_array(text, name):
    m = re.search(r'Name="%s"[^>]*>\s*([A-Za-z0-9+/=]+)\s*</DataArray>' % name, text)
    raw = m.group(1)
    head = base64.b64decode(raw[:8])
    nbytes = struct.unpack("<I", head[:4])[0]
    body = base64.b64decode(raw[8:])
    return np.frombuffer(body[:nbytes], dtype="<f8").copy()


def shipped():
    """Golden vectors the reference ships: decoded VTU T/Q, sweep logs, Tc_all (6 digits)."""
    out = {}
    vtu = open(os.path.join(REF, "output/vis/pbte_fields/Cycle000000/proc000000.vtu")).read()
    out["vtu_T"] = decode_vtu_array(vtu, "T")
    out["vtu_Q"] = decode_vtu_array(vtu, "Q")
    for tag in ("dim2", "dim3"):
        rows = []
        for line in open(os.path.join(REF, f"output/log/sweep_{tag}_np24_gauss_na24_gauss.txt")):
            if line.startswith("dir "):
                rows.append([int(x) for x in line.split("order:")[1].split()])
        out[f"sweep_{tag}"] = np.array(rows, dtype=np.int32)
    tc = []
    for line in open(os.path.join(REF, "output/log/Tc_all.txt")):
        if not line.startswith("#") and not line.startswith("elem"):
            tc.append([float(x) for x in line.split()])
    out["Tc_all_6digits"] = np.array(tc)
    sl = np.loadtxt(os.path.join(REF, "output/2D/results/T_slice.txt"), skiprows=2)
    out["T_slice"] = sl
    np.savez_compressed(os.path.join(HERE, "shipped_default.npz"), **out)
    print("shipped_default: T", out["vtu_T"], "sweep_dim3", out["sweep_dim3"].shape)


if __name__ == "__main__":
    shipped()
    for name, c in CASES.items():
        make_case(name, c)
