"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle.

  oracle/_build/libpbte_oracle.so   C restatement of the reference's algorithm (pbte_oracle.c)
  oracle/_ref/libpbte_ref.so        the reference's own sources compiled against the MFEM stand-in (ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may import this
module.  The product (dg-solver-for-pbte-with-mfem_b200/) never does.
"""
import contextlib
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "_build", "libpbte_oracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        lib.oracle_residual.restype = C.c_double
        _oracle = lib
    return _oracle


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libpbte_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libpbte_ref.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: it is built from /root/reference by `make -C oracle` in the build container")
        lib = C.CDLL(path)
        lib.ref_last_error.restype = C.c_char_p
        _ref = lib
    return _ref


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


# ---------------------------------------------------------------------------------------------------------
# C restatement
# ---------------------------------------------------------------------------------------------------------
def sweep_orders(problem):
    """AngularSweepOrder::Build restated: orders[ndir][ne] (raises on a stalled ordering)."""
    lib = oracle_lib()
    st = problem.struct()
    out = np.zeros((problem.ndir, problem.ne), dtype=np.int32)
    for k in range(problem.ndir):
        rc = lib.oracle_sweep_order(C.byref(st), k, _ip(out[k]))
        if rc != 0:
            raise RuntimeError(f"oracle_sweep_order: status {rc}")
    return out


def assemble_A(problem, k, band, e):
    lib = oracle_lib()
    st = problem.struct()
    n = problem.ndof
    A = np.zeros(n * n)
    lib.oracle_assemble_A(C.byref(st), C.c_double(float(problem.inv_kn.max())), k, band, e, _dp(A))
    return A.reshape(n, n).T.copy()  # row-major view of the column-major matrix


def solve(problem, max_iter, tol=0.0, cache_lu=True, threads=1, coeff=None, Tc=None, Tv=None):
    """PBTESolver::Solve restated.  Returns dict(coeff, Tc, Qc, Tv, Qv, residuals, iters)."""
    lib = oracle_lib()
    st = problem.struct()
    p = problem
    block = p.ndof * p.ne
    coeff = np.zeros((p.ndir, p.nband, block)) if coeff is None else np.ascontiguousarray(coeff, dtype=np.float64).copy()
    Tc = np.zeros(block) if Tc is None else np.ascontiguousarray(Tc, dtype=np.float64).copy()
    Tv = np.zeros(p.ne) if Tv is None else np.ascontiguousarray(Tv, dtype=np.float64).copy()
    Qc = np.zeros((p.dim, block))
    Qv = np.zeros((p.dim, p.ne))
    hist = np.zeros(max(max_iter, 1))
    last = C.c_double(0)
    rc = lib.oracle_solve(C.byref(st), max_iter, C.c_double(tol), int(cache_lu), threads, _dp(coeff), _dp(Tc), _dp(Qc), _dp(Tv), _dp(Qv),
                          _dp(hist), C.byref(last))
    if rc < 0:
        raise RuntimeError(f"oracle_solve: status {rc}")
    return dict(coeff=coeff, Tc=Tc, Qc=Qc, Tv=Tv, Qv=Qv, residuals=hist[:rc], iters=rc, residual=last.value)


def moments(problem, coeff):
    """MacroscopicQuantities Reset/Accumulate/Finalize restated on given coefficient blocks [ndir][nband][ndof*ne]."""
    lib = oracle_lib()
    st = problem.struct()
    p = problem
    block = p.ndof * p.ne
    coeff = np.ascontiguousarray(coeff, dtype=np.float64)
    Tc, Qc = np.zeros(block), np.zeros((p.dim, block))
    for k in range(p.ndir):
        for b in range(p.nband):
            lib.oracle_accumulate(C.byref(st), k, b, _dp(coeff[k, b]), _dp(Tc), _dp(Qc))
    Tv, Qv = np.zeros(p.ne), np.zeros((p.dim, p.ne))
    lib.oracle_finalize(C.byref(st), _dp(Tc), _dp(Qc), _dp(Tv), _dp(Qv))
    return dict(Tc=Tc, Qc=Qc, Tv=Tv, Qv=Qv)


def residual(Tv, prev_Tv):
    Tv = np.ascontiguousarray(Tv, dtype=np.float64)
    prev_Tv = np.ascontiguousarray(prev_Tv, dtype=np.float64)
    return oracle_lib().oracle_residual(_dp(Tv), _dp(prev_Tv), len(Tv))


# ---------------------------------------------------------------------------------------------------------
# the reference itself
# ---------------------------------------------------------------------------------------------------------
CONFIG_TEMPLATE = """mesh:
  path: {mesh}
boundary_conditions:
{bcs}
angles:
  dimension: {angle_dim}
  polar_points: {polar_points}
  azimuth_points: {azimuth_points}
  polar_scheme: {polar_scheme}
  azimuth_scheme: {azimuth_scheme}
numerical:
  tolerance: 1e-7
  max_iter: 101
"""

SI_TEMPLATE = """model: non-gray-smrt
num_spectral: {num_spectral}
reference_temperature: {ref_temp!r}
reference_length: {ref_len!r}
num_branches: 2
C_LA: [{C_LA[0]!r}, {C_LA[1]!r}]
C_TA: [{C_TA[0]!r}, {C_TA[1]!r}]
lattice_dist: {lattice_dist!r}
Ai: {Ai!r}
BL: {BL!r}
BT: {BT!r}
BU: {BU!r}
"""


@contextlib.contextmanager
def _chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class RefCase:
    """The reference pipeline (SpatialMesh -> ... -> PBTESolver) on config files written to a scratch directory."""

    def __init__(self, mesh, order, refine, angle_dim, polar_points, azimuth_points, polar_gauss, azimuth_gauss, bcs, material):
        from types import SimpleNamespace
        lib = ref_lib()
        self._tmp = tempfile.TemporaryDirectory(prefix="pbte_ref_")
        cfg = os.path.join(self._tmp.name, "config.yaml")
        si = os.path.join(self._tmp.name, "si.yaml")
        with open(cfg, "w") as f:
            f.write(CONFIG_TEMPLATE.format(mesh=mesh, angle_dim=angle_dim, polar_points=polar_points, azimuth_points=azimuth_points,
                                           polar_scheme="gauss" if polar_gauss else "uniform",
                                           azimuth_scheme="gauss" if azimuth_gauss else "uniform",
                                           bcs="\n".join(f"  - attr: {a}\n    temperature: {t!r}" for a, t in sorted(bcs.items()))))
        with open(si, "w") as f:
            f.write(SI_TEMPLATE.format(**material))
        self._h = C.c_void_p()
        with _chdir(self._tmp.name):  # the reference writes output/log/* relative to the working directory
            rc = lib.ref_case_create(cfg.encode(), si.encode(), order, refine, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"ref_case_create: {lib.ref_last_error().decode()}")
        v = [C.c_int(0) for _ in range(7)]
        lib.ref_case_sizes(self._h, *[C.byref(x) for x in v])
        self.dim, self.ne, self.ndof, self.nfaces, self.ndir, self.nbranch, self.nspec = [x.value for x in v]
        self.nband = self.nbranch * self.nspec
        self.sizes = SimpleNamespace(dim=self.dim, ne=self.ne, ndof=self.ndof, nfaces=self.nfaces, ndir=self.ndir, nband=self.nband)

    def close(self):
        if self._h:
            ref_lib().ref_case_destroy(self._h)
            self._h = C.c_void_p()

    def arrays(self):
        """The flat problem image taken from the reference's own objects (dict of numpy arrays + scalars)."""
        n, ne, F, dim, nd, nb = self.ndof, self.ne, self.nfaces, self.dim, self.ndir, self.nband
        a = dict(basis_integrals=np.zeros((ne, n)), mass=np.zeros((ne, n * n)), stiff=np.zeros((ne, dim, n * n)),
                 face_mass=np.zeros((ne, F, n * n)), face_coupling=np.zeros((ne, F, n * n)), face_integral=np.zeros((ne, F, n)),
                 face_normal=np.zeros((ne, F, 3)), face_nbr=np.zeros((ne, F), dtype=np.int32), face_tbc=np.zeros((ne, F)),
                 dir=np.zeros((nd, 3)), dir_w=np.zeros(nd), inv_kn=np.zeros(nb), vg=np.zeros(nb), heat_cap=np.zeros(nb), freq_w=np.zeros(nb))
        omega, hcv = C.c_double(0), C.c_double(0)
        polar, azim = np.zeros(nd), np.zeros(nd)
        rc = ref_lib().ref_case_fill(self._h, _dp(a["basis_integrals"]), _dp(a["mass"]), _dp(a["stiff"]), _dp(a["face_mass"]),
                                     _dp(a["face_coupling"]), _dp(a["face_integral"]), _dp(a["face_normal"]), _ip(a["face_nbr"]),
                                     _dp(a["face_tbc"]), _dp(a["dir"]), _dp(a["dir_w"]), C.byref(omega), _dp(a["inv_kn"]), _dp(a["vg"]),
                                     _dp(a["heat_cap"]), _dp(a["freq_w"]), C.byref(hcv), _dp(polar), _dp(azim))
        if rc != 0:
            raise RuntimeError(f"ref_case_fill: {ref_lib().ref_last_error().decode()}")
        a.update(omega=omega.value, heat_cap_v=hcv.value, polar=polar, azimuth=azim)
        return a

    def orders(self):
        out = np.zeros((self.ndir, self.ne), dtype=np.int32)
        rc = ref_lib().ref_case_orders(self._h, _ip(out))
        if rc != 0:
            raise RuntimeError(f"ref_case_orders: {ref_lib().ref_last_error().decode()}")
        return out

    def solve(self, max_iter, tol=0.0, policy=0, coeff=None):
        block = self.ndof * self.ne
        coeff = np.zeros((self.ndir, self.nband, block)) if coeff is None else np.ascontiguousarray(coeff, dtype=np.float64).copy()
        out = dict(Tc=np.zeros(block), Qc=np.zeros((self.dim, block)), Tv=np.zeros(self.ne), Qv=np.zeros((self.dim, self.ne)))
        res = C.c_double(0)
        secs = np.zeros(2)
        with _chdir(self._tmp.name):
            rc = ref_lib().ref_case_solve(self._h, max_iter, C.c_double(tol), policy, _dp(coeff), _dp(out["Tc"]), _dp(out["Qc"]),
                                          _dp(out["Tv"]), _dp(out["Qv"]), C.byref(res), _dp(secs))
        if rc != 0:
            err = RuntimeError(f"ref_case_solve: {ref_lib().ref_last_error().decode()}")
            err.status = rc
            raise err
        out.update(coeff=coeff, residual=res.value, seconds_ctor=secs[0], seconds_solve=secs[1])
        return out


class RefSolver:
    """The reference's PBTESolver kept alive across Solve calls (constructor incl. its LU cache timed separately)."""

    def __init__(self, case, iters_per_call=1, policy=0):
        self.case = case
        self._h = C.c_void_p()
        secs = C.c_double(0)
        with _chdir(case._tmp.name):
            rc = ref_lib().ref_solver_create(case._h, iters_per_call, policy, C.byref(self._h), C.byref(secs))
        if rc != 0:
            raise RuntimeError(f"ref_solver_create: {ref_lib().ref_last_error().decode()}")
        self.ctor_seconds = secs.value

    def solve(self):
        res, secs = C.c_double(0), C.c_double(0)
        with _chdir(self.case._tmp.name):
            rc = ref_lib().ref_solver_solve(self._h, C.byref(res), C.byref(secs))
        if rc != 0:
            raise RuntimeError(f"ref_solver_solve: {ref_lib().ref_last_error().decode()}")
        return res.value, secs.value

    def close(self):
        if self._h:
            ref_lib().ref_solver_destroy(self._h)
            self._h = C.c_void_p()
