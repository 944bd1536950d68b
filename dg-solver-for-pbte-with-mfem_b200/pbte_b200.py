"""ctypes access to the two in-tree libraries of the package.

  lib/libpbte_cuda.so  - the C ABI of include/pbte_cuda.h (CUDA kernels, sm_100a)
  lib/libpbte_host.so  - the host layer (kept pbte:: C++ classes on the MFEM stand-in) behind host/host_capi.h

This module is plumbing for tests/ and bench.py; it holds no numerics of its own.  Loading fails loudly when
a library is missing (build with `python -c "import __graft_entry__ as g; g.build()"`): there is no fallback.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("PBTE_LIB_DIR") or os.path.join(PKG_DIR, "lib")  # PBTE_LIB_DIR: development builds (tools/)

PBTE_OK = 0
ERR = {-1: "PBTE_ERR_INVALID", -2: "PBTE_ERR_CUDA", -3: "PBTE_ERR_SWEEP_STALL", -4: "PBTE_ERR_MISSING_BC",
       -5: "PBTE_ERR_ZERO_NORM", -6: "PBTE_ERR_SINGULAR", -7: "PBTE_ERR_NOMEM", -8: "PBTE_ERR_STATE"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PbteError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{ERR.get(status, status)}: {msg}")
        self.status = status


class Problem(C.Structure):
    """pbte_problem_t (include/pbte_cuda.h)."""
    _fields_ = [("abi_version", C.c_int32), ("dim", C.c_int32), ("ne", C.c_int32), ("ndof", C.c_int32),
                ("nfaces", C.c_int32), ("ndir", C.c_int32), ("nband", C.c_int32), ("reserved0", C.c_int32),
                ("basis_integrals", c_double_p), ("mass", c_double_p), ("stiff", c_double_p),
                ("face_mass", c_double_p), ("face_coupling", c_double_p), ("face_integral", c_double_p),
                ("face_normal", c_double_p), ("face_nbr", c_int32_p), ("face_tbc", c_double_p),
                ("dir", c_double_p), ("dir_w", c_double_p), ("omega", C.c_double),
                ("inv_kn", c_double_p), ("vg", c_double_p), ("heat_cap", c_double_p), ("freq_w", c_double_p),
                ("heat_cap_v", C.c_double)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("dir_begin", C.c_int32), ("dir_end", C.c_int32), ("cache_policy", C.c_int32),
                ("kernel", C.c_int32), ("use_graph", C.c_int32), ("stream", C.c_void_p), ("class_tol", C.c_double),
                ("max_classes", C.c_int32), ("reserved", C.c_int32 * 5)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("sweeps", C.c_int64), ("last_sweep_ms", C.c_double),
                ("last_finish_ms", C.c_double), ("flux_bytes", C.c_int64), ("launches_per_sweep", C.c_int64),
                ("kernel_used", C.c_int32), ("n_classes", C.c_int32), ("max_levels", C.c_int32),
                ("n_operator_classes", C.c_int32), ("march_tiles", C.c_int32), ("march_variant", C.c_int32),
                ("march_fill", C.c_double), ("operator_bytes", C.c_int64), ("last_march_ms", C.c_double),
                ("march_persistent", C.c_int32), ("march_ctas", C.c_int32), ("image_bytes", C.c_int64), ("create_ms", C.c_double),
                ("image_reps_only", C.c_int32), ("march_mask_mismatch", C.c_int32)]


class HostSpec(C.Structure):
    """pbte_host_spec_t (host/host_capi.h)."""
    _fields_ = [("mesh", C.c_char_p), ("order", C.c_int32), ("refine", C.c_int32), ("angle_dim", C.c_int32),
                ("polar_points", C.c_int32), ("azimuth_points", C.c_int32), ("polar_gauss", C.c_int32),
                ("azimuth_gauss", C.c_int32), ("face_rule", C.c_int32), ("n_bc", C.c_int32),
                ("bc_attr", c_int32_p), ("bc_temp", c_double_p), ("num_spectral", C.c_int32),
                ("ref_temp", C.c_double), ("ref_len", C.c_double), ("lattice_dist", C.c_double), ("Ai", C.c_double),
                ("BL", C.c_double), ("BT", C.c_double), ("BU", C.c_double), ("C_LA", C.c_double * 2), ("C_TA", C.c_double * 2)]


# every symbol include/pbte_cuda.h declares (checked by the CPU test-suite against the built library)
CUDA_ABI_SYMBOLS = [
    "pbte_cuda_abi_version", "pbte_cuda_last_error", "pbte_cuda_device_count", "pbte_cuda_default_options",
    "pbte_cuda_create", "pbte_cuda_destroy", "pbte_cuda_build_schedule", "pbte_cuda_get_order", "pbte_cuda_get_levels",
    "pbte_cuda_schedule_info", "pbte_cuda_reset_state", "pbte_cuda_upload_coeff", "pbte_cuda_download_coeff",
    "pbte_cuda_upload_moments", "pbte_cuda_download_moments", "pbte_cuda_sweep", "pbte_cuda_moment_buffer",
    "pbte_cuda_finish_iteration", "pbte_cuda_solve", "pbte_cuda_last_norms", "pbte_cuda_macro_reset",
    "pbte_cuda_macro_accumulate", "pbte_cuda_macro_finalize", "pbte_cuda_macro_residual", "pbte_cuda_get_stats",
    "pbte_cuda_get_stream", "pbte_cuda_measure_peaks", "pbte_cuda_iterate_async", "pbte_cuda_iterate_sync", "pbte_cuda_solve_ex",
    "pbte_cuda_comm_get_unique_id", "pbte_cuda_comm_attach", "pbte_cuda_comm_attach_all", "pbte_cuda_comm_detach",
    "pbte_cuda_upload_coeff_all", "pbte_cuda_download_coeff_all", "pbte_cuda_host_alloc", "pbte_cuda_host_free",
]
COMM_ID_BYTES = 128
ITER_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double)

_cuda = None
_host = None


def _load(name):
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build the package first (__graft_entry__.build()); there is no fallback path")
    return C.CDLL(path)


def cuda_lib():
    global _cuda
    if _cuda is None:
        lib = _load("libpbte_cuda.so")
        lib.pbte_cuda_last_error.restype = C.c_char_p
        lib.pbte_cuda_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(C.c_void_p)]
        lib.pbte_cuda_destroy.argtypes = [C.c_void_p]
        lib.pbte_cuda_build_schedule.argtypes = [C.c_void_p]
        lib.pbte_cuda_get_order.argtypes = [C.c_void_p, C.c_int32, c_int32_p]
        lib.pbte_cuda_get_levels.argtypes = [C.c_void_p, C.c_int32, c_int32_p]
        lib.pbte_cuda_schedule_info.argtypes = [C.c_void_p, c_int32_p, c_int32_p]
        lib.pbte_cuda_reset_state.argtypes = [C.c_void_p]
        lib.pbte_cuda_upload_coeff.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_double_p]
        lib.pbte_cuda_download_coeff.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_double_p]
        lib.pbte_cuda_upload_moments.argtypes = [C.c_void_p, c_double_p, c_double_p]
        lib.pbte_cuda_download_moments.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
        lib.pbte_cuda_sweep.argtypes = [C.c_void_p]
        lib.pbte_cuda_moment_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.pbte_cuda_finish_iteration.argtypes = [C.c_void_p, c_double_p]
        lib.pbte_cuda_solve.argtypes = [C.c_void_p, C.c_int32, C.c_double, c_int32_p, c_double_p, c_double_p]
        lib.pbte_cuda_last_norms.argtypes = [C.c_void_p, c_double_p, c_double_p]
        lib.pbte_cuda_macro_reset.argtypes = [C.c_void_p]
        lib.pbte_cuda_macro_accumulate.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_double_p]
        lib.pbte_cuda_macro_finalize.argtypes = [C.c_void_p, c_double_p]
        lib.pbte_cuda_macro_residual.argtypes = [C.c_void_p, c_double_p, c_double_p]
        lib.pbte_cuda_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.pbte_cuda_get_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.pbte_cuda_measure_peaks.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
        lib.pbte_cuda_upload_coeff_all.argtypes = [C.c_void_p, c_double_p]
        lib.pbte_cuda_download_coeff_all.argtypes = [C.c_void_p, c_double_p]
        lib.pbte_cuda_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        lib.pbte_cuda_host_free.argtypes = [C.c_void_p]
        lib.pbte_cuda_iterate_async.argtypes = [C.c_void_p]
        lib.pbte_cuda_iterate_sync.argtypes = [C.c_void_p, C.POINTER(C.c_int64), c_double_p]
        lib.pbte_cuda_solve_ex.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_double, c_int32_p, c_double_p, c_double_p,
                                           ITER_CALLBACK, C.c_void_p]
        lib.pbte_cuda_comm_get_unique_id.argtypes = [C.c_void_p]
        lib.pbte_cuda_comm_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        lib.pbte_cuda_comm_attach_all.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
        lib.pbte_cuda_comm_detach.argtypes = [C.c_void_p]
        lib.pbte_cuda_default_options.argtypes = [C.POINTER(Options)]
        lib.pbte_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
        _cuda = lib
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        lib = _load("libpbte_host.so")
        lib.pbte_host_last_error.restype = C.c_char_p
        lib.pbte_host_case_create.argtypes = [C.POINTER(HostSpec), C.POINTER(C.c_void_p)]
        lib.pbte_host_case_destroy.argtypes = [C.c_void_p]
        lib.pbte_host_case_problem.argtypes = [C.c_void_p]
        lib.pbte_host_case_problem.restype = C.POINTER(Problem)
        lib.pbte_host_case_band_table.argtypes = [C.c_void_p, C.c_int, c_double_p]
        lib.pbte_host_case_solve.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32, c_double_p, c_int32_p, c_double_p,
                                             c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p]
        lib.pbte_host_case_macro.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        lib.pbte_host_case_vertex_values.argtypes = [C.c_void_p, c_double_p, c_double_p, c_int32_p]
        lib.pbte_host_solver_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        lib.pbte_host_solver_solve.argtypes = [C.c_void_p, C.c_int32, C.c_double, c_double_p, c_int32_p, c_double_p]
        lib.pbte_host_solver_moments.argtypes = [C.c_void_p, c_double_p, c_double_p]
        lib.pbte_host_solver_destroy.argtypes = [C.c_void_p]
        _host = lib
    return _host


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def _ck(status):
    if status != PBTE_OK:
        raise PbteError(status, cuda_lib().pbte_cuda_last_error().decode())


# material constants of config/si.yaml (silicon)
SI_MATERIAL = dict(ref_temp=300.0, ref_len=1.0e-6, lattice_dist=5.43e-10, Ai=1.498e-45, BL=1.18e-24, BT=8.708e-13,
                   BU=2.890e-18, C_LA=(9.01e3, -2.0e-7), C_TA=(5.23e3, -2.26e-7), num_spectral=20)


class ProblemArrays:
    """A problem image held in numpy arrays (keeps them alive) plus its pbte_problem_t view."""
    FIELDS = ["basis_integrals", "mass", "stiff", "face_mass", "face_coupling", "face_integral", "face_normal",
              "face_nbr", "face_tbc", "dir", "dir_w", "inv_kn", "vg", "heat_cap", "freq_w"]

    def __init__(self, dim, ne, ndof, nfaces, ndir, nband, omega, heat_cap_v, **arrays):
        self.dim, self.ne, self.ndof, self.nfaces, self.ndir, self.nband = dim, ne, ndof, nfaces, ndir, nband
        self.omega, self.heat_cap_v = float(omega), float(heat_cap_v)
        for f in self.FIELDS:
            a = arrays.get(f)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.int32 if f == "face_nbr" else np.float64)
            setattr(self, f, a)

    @classmethod
    def from_struct(cls, p):
        """Deep copy of a C problem image (e.g. the one a host case owns)."""
        n, ne, F, dim, nd, nb = p.ndof, p.ne, p.nfaces, p.dim, p.ndir, p.nband
        shapes = dict(basis_integrals=(ne, n), mass=(ne, n * n), stiff=(ne, dim, n * n), face_mass=(ne, F, n * n),
                      face_coupling=(ne, F, n * n), face_integral=(ne, F, n), face_normal=(ne, F, 3), face_nbr=(ne, F),
                      face_tbc=(ne, F), dir=(nd, 3), dir_w=(nd,), inv_kn=(nb,), vg=(nb,), heat_cap=(nb,), freq_w=(nb,))
        arrays = {}
        for f, shp in shapes.items():
            ptr = getattr(p, f)
            arrays[f] = np.ctypeslib.as_array(ptr, shape=shp).copy() if ptr and int(np.prod(shp)) > 0 else None
        return cls(dim, ne, n, F, nd, nb, p.omega, p.heat_cap_v, **arrays)

    def struct(self):
        p = Problem()
        p.abi_version = 1
        p.dim, p.ne, p.ndof, p.nfaces, p.ndir, p.nband = self.dim, self.ne, self.ndof, self.nfaces, self.ndir, self.nband
        p.omega, p.heat_cap_v = self.omega, self.heat_cap_v
        for f in self.FIELDS:
            a = getattr(self, f)
            if a is not None:
                setattr(p, f, a.ctypes.data_as(c_int32_p if f == "face_nbr" else c_double_p))
        return p

    @property
    def updates_per_sweep(self):
        return self.ndof * self.ne * self.ndir * self.nband


class HostCase:
    """Mesh + FE setup through the host layer; owns the C++ objects and exposes the flattened problem image."""

    def __init__(self, mesh, order, refine=0, angle_dim=3, polar_points=8, azimuth_points=16, polar_gauss=True,
                 azimuth_gauss=False, bcs=None, face_rule=0, **material):
        mat = dict(SI_MATERIAL)
        mat.update(material)
        lib = host_lib()
        s = HostSpec()
        s.mesh = mesh.encode()
        s.order, s.refine, s.angle_dim = order, refine, angle_dim
        s.polar_points, s.azimuth_points = polar_points, azimuth_points
        s.polar_gauss, s.azimuth_gauss, s.face_rule = int(polar_gauss), int(azimuth_gauss), face_rule
        bcs = bcs or {}
        self._attr = np.array(sorted(bcs), dtype=np.int32)
        self._temp = np.array([bcs[a] for a in sorted(bcs)], dtype=np.float64)
        s.n_bc = len(self._attr)
        s.bc_attr, s.bc_temp = _ip(self._attr), _dp(self._temp)
        s.num_spectral = mat["num_spectral"]
        for k in ("ref_temp", "ref_len", "lattice_dist", "Ai", "BL", "BT", "BU"):
            setattr(s, k, mat[k])
        s.C_LA = (C.c_double * 2)(*mat["C_LA"])
        s.C_TA = (C.c_double * 2)(*mat["C_TA"])
        self._h = C.c_void_p()
        st = lib.pbte_host_case_create(C.byref(s), C.byref(self._h))
        if st != PBTE_OK:
            raise PbteError(st, lib.pbte_host_last_error().decode())
        self.problem = ProblemArrays.from_struct(lib.pbte_host_case_problem(self._h).contents)

    def close(self):
        if self._h:
            host_lib().pbte_host_case_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def band_table(self, which):
        out = np.zeros(self.problem.nband)
        host_lib().pbte_host_case_band_table(self._h, which, _dp(out))
        return out

    def solve(self, max_iter, tol=0.0, cache_policy=0, want_coeff=False):
        """The kept C++ API end to end (AngularSweepOrder::Build, PBTESolver, Solve) on the GPU."""
        p = self.problem
        block = p.ndof * p.ne
        out = dict(Tc=np.zeros(block), Qc=np.zeros((p.dim, block)), Tv=np.zeros(p.ne), Qv=np.zeros((p.dim, p.ne)),
                   orders=np.zeros((p.ndir, p.ne), dtype=np.int32))
        coeff = np.zeros((p.ndir, p.nband, block)) if want_coeff else None
        res, iters = C.c_double(0), C.c_int32(0)
        lib = host_lib()
        st = lib.pbte_host_case_solve(self._h, max_iter, tol, cache_policy, C.byref(res), C.byref(iters), _dp(out["Tc"]), _dp(out["Qc"]),
                                      _dp(out["Tv"]), _dp(out["Qv"]), _ip(out["orders"]), _dp(coeff))
        if st != PBTE_OK:
            raise PbteError(st, lib.pbte_host_last_error().decode())
        out.update(residual=res.value, iters=iters.value, coeff=coeff)
        return out

    def macro(self, coeff, prev_tv=None):
        p = self.problem
        block = p.ndof * p.ne
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        out = dict(Tc=np.zeros(block), Qc=np.zeros((p.dim, block)), Tv=np.zeros(p.ne), Qv=np.zeros((p.dim, p.ne)))
        res = C.c_double(0)
        pt = None if prev_tv is None else np.ascontiguousarray(prev_tv, dtype=np.float64)
        lib = host_lib()
        st = lib.pbte_host_case_macro(self._h, _dp(coeff), _dp(pt), _dp(out["Tc"]), _dp(out["Qc"]), _dp(out["Tv"]), _dp(out["Qv"]),
                                      C.byref(res) if pt is not None else None)
        if st != PBTE_OK:
            raise PbteError(st, lib.pbte_host_last_error().decode())
        out["residual"] = res.value
        return out

    def vertex_values(self, field):
        p = self.problem
        field = np.ascontiguousarray(field, dtype=np.float64)
        out = np.zeros((p.ne, 8))
        nv = C.c_int32(0)
        st = host_lib().pbte_host_case_vertex_values(self._h, _dp(field), _dp(out), C.byref(nv))
        if st != PBTE_OK:
            raise PbteError(st, host_lib().pbte_host_last_error().decode())
        return out.reshape(-1)[: p.ne * nv.value].reshape(p.ne, nv.value)


class HostSolver:
    """The kept C++ objects (AngularSweepOrder, PBTESolver, MacroscopicQuantities, host coefficient blocks) kept alive across
    Solve calls; each solve() is PBTESolver::Solve(coeff, macro) with host buffers in and out."""

    def __init__(self, case, cache_policy=0):
        self.case = case
        self._h = C.c_void_p()
        st = host_lib().pbte_host_solver_create(case._h, cache_policy, C.byref(self._h))
        if st != PBTE_OK:
            raise PbteError(st, host_lib().pbte_host_last_error().decode())

    def solve(self, max_iter, tol=0.0):
        res, it, ms = C.c_double(0), C.c_int32(0), C.c_double(0)
        st = host_lib().pbte_host_solver_solve(self._h, max_iter, tol, C.byref(res), C.byref(it), C.byref(ms))
        if st != PBTE_OK:
            raise PbteError(st, host_lib().pbte_host_last_error().decode())
        return res.value, it.value, ms.value

    def moments(self):
        p = self.case.problem
        Tc, Tv = np.zeros(p.ndof * p.ne), np.zeros(p.ne)
        host_lib().pbte_host_solver_moments(self._h, _dp(Tc), _dp(Tv))
        return Tc, Tv

    def close(self):
        if self._h:
            host_lib().pbte_host_solver_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    """One device handle of the C ABI (pbte_cuda_create .. pbte_cuda_destroy)."""

    def __init__(self, problem, device=-1, dir_range=None, cache_policy=0, kernel=0, use_graph=1, stream=None, class_tol=None):
        self.lib = cuda_lib()
        self.problem = problem
        self._struct = problem.struct()
        opt = Options()
        self.lib.pbte_cuda_default_options(C.byref(opt))
        opt.device, opt.cache_policy, opt.kernel, opt.use_graph = device, cache_policy, kernel, use_graph
        if dir_range is not None:
            opt.dir_begin, opt.dir_end = dir_range
        if stream is not None:
            opt.stream = stream
        if class_tol is not None:
            opt.class_tol = class_tol
        self.dir_range = dir_range if dir_range is not None else (0, problem.ndir)
        self._h = C.c_void_p()
        _ck(self.lib.pbte_cuda_create(C.byref(self._struct), C.byref(opt), C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.pbte_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_schedule(self):
        _ck(self.lib.pbte_cuda_build_schedule(self._h))
        nc, ml = C.c_int32(0), C.c_int32(0)
        _ck(self.lib.pbte_cuda_schedule_info(self._h, C.byref(nc), C.byref(ml)))
        return nc.value, ml.value

    def order(self, k):
        out = np.zeros(self.problem.ne, dtype=np.int32)
        _ck(self.lib.pbte_cuda_get_order(self._h, k, _ip(out)))
        return out

    def levels(self, k):
        out = np.zeros(self.problem.ne, dtype=np.int32)
        _ck(self.lib.pbte_cuda_get_levels(self._h, k, _ip(out)))
        return out

    def reset_state(self):
        _ck(self.lib.pbte_cuda_reset_state(self._h))

    def upload_coeff(self, k, band, block):
        block = np.ascontiguousarray(block, dtype=np.float64)
        _ck(self.lib.pbte_cuda_upload_coeff(self._h, k, band, _dp(block)))

    def download_coeff(self, k, band):
        out = np.zeros(self.problem.ndof * self.problem.ne)
        _ck(self.lib.pbte_cuda_download_coeff(self._h, k, band, _dp(out)))
        return out

    def download_all_coeff(self, per_block=False):
        """All owned blocks [k1-k0][nband][ndof*ne]: one bulk call (pbte_cuda_download_coeff_all) or block by block."""
        p = self.problem
        k0, k1 = self.dir_range
        out = np.zeros((k1 - k0, p.nband, p.ndof * p.ne))
        if not per_block:
            _ck(self.lib.pbte_cuda_download_coeff_all(self._h, _dp(out)))
            return out
        for k in range(k0, k1):
            for b in range(p.nband):
                _ck(self.lib.pbte_cuda_download_coeff(self._h, k, b, _dp(out[k - k0, b])))
        return out

    def upload_all_coeff(self, coeff):
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        k0, k1 = self.dir_range
        assert coeff.size == (k1 - k0) * self.problem.nband * self.problem.ndof * self.problem.ne
        _ck(self.lib.pbte_cuda_upload_coeff_all(self._h, _dp(coeff)))

    def upload_moments(self, Tc=None, Tv=None):
        Tc = None if Tc is None else np.ascontiguousarray(Tc, dtype=np.float64)
        Tv = None if Tv is None else np.ascontiguousarray(Tv, dtype=np.float64)
        _ck(self.lib.pbte_cuda_upload_moments(self._h, _dp(Tc), _dp(Tv)))

    def download_moments(self, out=None):
        """Tc, Qc, Tv, Qv of the last iteration; `out` may hold preallocated (e.g. pinned) float64 arrays of the right sizes."""
        p = self.problem
        block = p.ndof * p.ne
        if out is None:
            out = dict(Tc=np.zeros(block), Qc=np.zeros((p.dim, block)), Tv=np.zeros(p.ne), Qv=np.zeros((p.dim, p.ne)))
        _ck(self.lib.pbte_cuda_download_moments(self._h, _dp(out["Tc"]), _dp(out["Qc"]), _dp(out["Tv"]), _dp(out["Qv"])))
        return out

    def sweep(self):
        _ck(self.lib.pbte_cuda_sweep(self._h))

    def finish_iteration(self):
        r = C.c_double(0)
        _ck(self.lib.pbte_cuda_finish_iteration(self._h, C.byref(r)))
        return r.value

    def moment_buffer(self):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _ck(self.lib.pbte_cuda_moment_buffer(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def solve(self, max_iter, tol=0.0):
        hist = np.zeros(max(max_iter, 1))
        it, r = C.c_int32(0), C.c_double(0)
        _ck(self.lib.pbte_cuda_solve(self._h, max_iter, tol, C.byref(it), C.byref(r), _dp(hist)))
        return it.value, r.value, hist[: it.value]

    def iterate_async(self):
        """One whole source iteration enqueued (sweep, all-reduce when a communicator is attached, finish); no wait."""
        _ck(self.lib.pbte_cuda_iterate_async(self._h))

    def iterate_sync(self):
        """Waits for everything enqueued; (iterations finished since the counters were reset, residual of the last one)."""
        n, r = C.c_int64(0), C.c_double(0)
        _ck(self.lib.pbte_cuda_iterate_sync(self._h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def comm_attach(self, uid, rank, world_size):
        """Collective over the ranks: NCCL communicator for the per-iteration all-reduce of [Tc | Qc]; uid = comm_unique_id() of rank 0."""
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(bytes(uid))
        _ck(self.lib.pbte_cuda_comm_attach(self._h, buf, rank, world_size))

    def comm_detach(self):
        _ck(self.lib.pbte_cuda_comm_detach(self._h))

    def stats(self):
        s = Stats()
        _ck(self.lib.pbte_cuda_get_stats(self._h, C.byref(s)))
        return s

    def stream(self):
        p = C.c_void_p()
        _ck(self.lib.pbte_cuda_get_stream(self._h, C.byref(p)))
        return p.value

    def macro_reset(self):
        _ck(self.lib.pbte_cuda_macro_reset(self._h))

    def macro_accumulate(self, k, band, block):
        block = np.ascontiguousarray(block, dtype=np.float64)
        _ck(self.lib.pbte_cuda_macro_accumulate(self._h, k, band, _dp(block)))

    def macro_finalize(self, basis=None):
        basis = None if basis is None else np.ascontiguousarray(basis, dtype=np.float64)
        _ck(self.lib.pbte_cuda_macro_finalize(self._h, _dp(basis)))

    def macro_residual(self, prev_tv):
        prev_tv = np.ascontiguousarray(prev_tv, dtype=np.float64)
        r = C.c_double(0)
        _ck(self.lib.pbte_cuda_macro_residual(self._h, _dp(prev_tv), C.byref(r)))
        return r.value


def comm_unique_id():
    """128 bytes that rank 0 creates and ships to the other ranks (pbte_cuda_comm_get_unique_id)."""
    buf = (C.c_char * COMM_ID_BYTES)()
    _ck(cuda_lib().pbte_cuda_comm_get_unique_id(buf))
    return bytes(buf)


def comm_attach_all(solvers):
    """One process, several devices: a communicator set over the handles of `solvers` (one device each)."""
    arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    _ck(cuda_lib().pbte_cuda_comm_attach_all(arr, len(solvers)))


def solve_multi(solvers, max_iter, tol=0.0, on_iteration=None):
    """pbte_cuda_solve_ex over the handles of `solvers` (a single solver is allowed): (iterations, residual, history)."""
    arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    hist = np.zeros(max(max_iter, 1))
    it, r = C.c_int32(0), C.c_double(0)
    cb = ITER_CALLBACK(lambda user, i, res, tv, tc: on_iteration(i, res, tv, tc)) if on_iteration else ITER_CALLBACK()
    _ck(cuda_lib().pbte_cuda_solve_ex(arr, len(solvers), max_iter, tol, C.byref(it), C.byref(r), _dp(hist), cb, None))
    return it.value, r.value, hist[: it.value]


def measure_peaks(device=-1):
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _ck(cuda_lib().pbte_cuda_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c)))
    return dict(fp64_fma_tflops=a.value, fp64_mma_tflops=b.value, hbm_copy_gbs=c.value)


def device_count():
    n = C.c_int(0)
    st = cuda_lib().pbte_cuda_device_count(C.byref(n))
    return n.value if st == PBTE_OK else 0


def partition_directions(ndir, world_size, rank):
    """Contiguous block of directions owned by `rank` (the (angle, band) space is split by direction blocks so
    that the bands of a direction stay on one GPU).  Blocks differ by at most one direction; trailing ranks may be empty."""
    base, extra = divmod(ndir, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
