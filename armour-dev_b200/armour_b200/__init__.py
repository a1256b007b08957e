"""ctypes binding of libarmour_b200.so — the C ABI declared in include/armour_b200.h.

Method names follow the reference's Ipopt TNLP callbacks (armtd_NLP, KPR/NLPclass.cu) and its
PZsparseArray members so that tests read like the reference's own harness (KPR/PZ_tests.cu).
There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA
library and a B200; a missing library raises ImportError-like RuntimeError loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libarmour_b200.so")
NJ = 7
NF = 7
COMB = 36

TABLES = {"cos_q": 0, "sin_q": 1, "R": 2, "R_t": 3, "qd_des": 4, "qda_des": 5, "qdda_des": 6, "links": 7, "u_nom": 8, "u_nom_int": 9}
PZ_OPS = {"mul": 0, "add": 1, "sub": 2, "cross": 3, "simplify": 4, "add_one_dim0": 7, "add_one_dim1": 8, "add_one_dim2": 9, "cross_const_first": 10, "cross_const_second": 11}

EXPORTS = [
    "armour_default_config", "armour_create", "armour_destroy", "armour_last_error", "armour_build", "armour_build_batch",
    "armour_select_problem", "armour_build_armtd", "armour_get_nlp_info", "armour_get_bounds_info", "armour_get_starting_point", "armour_eval_f",
    "armour_eval_grad_f", "armour_eval_g", "armour_eval_jac_g", "armour_eval_g_jac", "armour_jac_structure", "armour_release_host_buffers", "armour_pinned_buffer_count", "armour_check_feasible",
    "armour_get_torque_radius", "armour_get_link_generators", "armour_get_link_sliced_center", "armour_get_hyperplanes",
    "armour_get_taylor_remainders", "armour_get_pz", "armour_pz_binary", "armour_last_build_ms", "armour_last_eval_ms",
    "armour_kernel_launches", "armour_set_kernel_timing", "armour_last_eval_host_us", "armour_upload_problems", "armour_build_resident", "armour_eval_resident", "armour_eval_resident_burst", "armour_eval_batch", "armour_last_eval_batch_ms", "armour_eval_batch_resident", "armour_batch_jacobian_device", "armour_get_batch_jacobian", "armour_upload_x", "armour_standin_solve", "armour_debug_phase_cycles", "armour_debug_canaries_verified", "armour_measure_fp64_peak",
]


class ArmourConfig(C.Structure):
    _fields_ = [
        ("num_time_steps", C.c_int),
        ("k_range", C.c_double * 7),
        ("mass_uncertainty", C.c_double),
        ("inertia_uncertainty", C.c_double),
        ("simplify_threshold", C.c_double),
        ("max_obstacles", C.c_int),
        ("max_monomials", C.c_int),
        ("max_entries", C.c_int),
        ("threads_per_cta", C.c_int),
        ("device", C.c_int),
        ("batch", C.c_int),
        ("pin_user_buffers", C.c_int),
        ("export_trajectory_tables", C.c_int),
    ]


class ArmourError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("armour_b200 error %d: %s" % (code, msg))
        self.code = code


def build_library(force=False):
    """Compile libarmour_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(PKG_DIR, "csrc", f) for f in os.listdir(os.path.join(PKG_DIR, "csrc")) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "armour_b200.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", PKG_DIR, "-s", "-j4"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libarmour_b200.so is missing (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.armour_last_error.restype = C.c_char_p
        L.armour_create.argtypes = [C.POINTER(ArmourConfig), C.POINTER(C.c_void_p)]
        L.armour_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _vec(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def default_config():
    cfg = ArmourConfig()
    lib().armour_default_config(C.byref(cfg))
    return cfg


def measure_fp64_peak(device=-1):
    v = C.c_double()
    rc = lib().armour_measure_fp64_peak(C.c_int(device), C.byref(v))
    if rc != 0:
        raise ArmourError(rc, lib().armour_last_error().decode())
    return v.value


class Planner:
    """One handle = the reference's BezierCurve + KinematicsDynamics + Obstacles + armtd_NLP for one
    planning problem (or a batch of independent ones)."""

    def __init__(self, T=128, k_range=None, mass_uncertainty=0.03, inertia_uncertainty=0.03, threshold=5e-4, max_obstacles=40,
                 max_monomials=0, max_entries=0, threads_per_cta=0, device=-1, batch=1, pin_user_buffers=False, export_trajectory_tables=0):
        self.L = lib()
        cfg = default_config()
        cfg.num_time_steps = T
        if k_range is not None:
            for i in range(7):
                cfg.k_range[i] = k_range[i]
        cfg.mass_uncertainty = mass_uncertainty
        cfg.inertia_uncertainty = inertia_uncertainty
        cfg.simplify_threshold = threshold
        cfg.max_obstacles = max_obstacles
        cfg.max_monomials = max_monomials
        cfg.max_entries = max_entries
        cfg.threads_per_cta = threads_per_cta
        cfg.device = device
        cfg.batch = batch
        cfg.pin_user_buffers = 1 if pin_user_buffers else 0
        cfg.export_trajectory_tables = int(export_trajectory_tables)
        self.T = T
        self.batch = batch
        self.pin_user_buffers = bool(pin_user_buffers)
        self._own = None
        self.k_range = np.array([cfg.k_range[i] for i in range(7)])
        self.n_obs = 0
        self.h = C.c_void_p()
        rc = self.L.armour_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            msg = self.L.armour_last_error().decode()
            if self.h:
                self.L.armour_destroy(self.h)
                self.h = None
            raise ArmourError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.armour_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise ArmourError(rc, self.L.armour_last_error().decode())
        return rc

    # ---- reach-set build (stages A-D) ----
    def build(self, q0, qd0, qdd0, obstacles):
        obs = _vec(obstacles)
        assert obs.size % 12 == 0
        self.n_obs = obs.size // 12
        self._ck(self.L.armour_build(self.h, _dp(_vec(q0, 7)), _dp(_vec(qd0, 7)), _dp(_vec(qdd0, 7)), _dp(obs) if obs.size else None, C.c_int(self.n_obs)))
        return self.last_build_ms()[0]

    def build_armtd(self, q0, qd0, jrs, k_range, obstacles):
        """ARMTD comparison planner: jrs[6, 7, T] offline JRS tables, k_range[7]."""
        obs = _vec(obstacles)
        self.n_obs = obs.size // 12
        self.k_range = np.array(k_range, dtype=float)
        self._ck(self.L.armour_build_armtd(self.h, _dp(_vec(q0, 7)), _dp(_vec(qd0, 7)), _dp(_vec(jrs, 6 * 7 * self.T)), _dp(_vec(k_range, 7)),
                                           _dp(obs) if obs.size else None, C.c_int(self.n_obs)))
        return self.last_build_ms()[0]

    def build_batch(self, q0, qd0, qdd0, obstacles, n_obs):
        q0, qd0, qdd0 = _vec(q0), _vec(qd0), _vec(qdd0)
        count = q0.size // 7
        obs = _vec(obstacles)
        assert obs.size == count * n_obs * 12
        self.n_obs = n_obs
        self._ck(self.L.armour_build_batch(self.h, C.c_int(count), _dp(q0), _dp(qd0), _dp(qdd0), _dp(obs) if obs.size else None, C.c_int(n_obs)))
        return self.last_build_ms()[0]

    def upload_problems(self, q0, qd0, qdd0, obstacles, n_obs):
        q0, qd0, qdd0 = _vec(q0), _vec(qd0), _vec(qdd0)
        count = q0.size // 7
        obs = _vec(obstacles)
        self.n_obs = n_obs
        self._ck(self.L.armour_upload_problems(self.h, C.c_int(count), _dp(q0), _dp(qd0), _dp(qdd0), _dp(obs) if obs.size else None, C.c_int(n_obs)))

    def build_resident(self):
        self._ck(self.L.armour_build_resident(self.h))
        return self.last_build_ms()[0]

    def eval_resident(self, x=None):
        self._ck(self.L.armour_eval_resident(self.h, _dp(_vec(x, 7)) if x is not None else None))

    def upload_x(self, x):
        self._ck(self.L.armour_upload_x(self.h, _dp(_vec(x, 7))))

    def select_problem(self, p):
        self._ck(self.L.armour_select_problem(self.h, C.c_int(p)))

    def last_build_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.armour_last_build_ms(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def last_eval_ms(self):
        a = C.c_float()
        self._ck(self.L.armour_last_eval_ms(self.h, C.byref(a)))
        return a.value

    def kernel_launches(self):
        a = C.c_uint64()
        self._ck(self.L.armour_kernel_launches(self.h, C.byref(a)))
        return a.value

    # ---- TNLP callbacks ----
    def get_nlp_info(self):
        n, m, nnz, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.armour_get_nlp_info(self.h, C.byref(n), C.byref(m), C.byref(nnz), C.byref(nh)))
        return n.value, m.value, nnz.value, nh.value

    @property
    def m(self):
        return self.get_nlp_info()[1]

    def get_bounds_info(self):
        m = self.m
        xl, xu, gl, gu = np.zeros(7), np.zeros(7), np.zeros(m), np.zeros(m)
        self._ck(self.L.armour_get_bounds_info(self.h, _dp(xl), _dp(xu), _dp(gl), _dp(gu)))
        return xl, xu, gl, gu

    def get_starting_point(self):
        x = np.ones(7)
        self._ck(self.L.armour_get_starting_point(self.h, _dp(x)))
        return x

    def eval_f(self, q_des, t_plan, x):
        f = C.c_double()
        self._ck(self.L.armour_eval_f(self.h, _dp(_vec(q_des, 7)), C.c_double(t_plan), _dp(_vec(x, 7)), C.byref(f)))
        return f.value

    def eval_grad_f(self, q_des, t_plan, x):
        g = np.zeros(7)
        self._ck(self.L.armour_eval_grad_f(self.h, _dp(_vec(q_des, 7)), C.c_double(t_plan), _dp(_vec(x, 7)), _dp(g)))
        return g

    def eval_g(self, x):
        g = np.zeros(self.m)
        self._ck(self.L.armour_eval_g(self.h, _dp(_vec(x, 7)), _dp(g)))
        return g

    def eval_jac_g(self, x):
        v = np.zeros(self.m * 7)
        self._ck(self.L.armour_eval_jac_g(self.h, _dp(_vec(x, 7)), _dp(v)))
        return v.reshape(self.m, 7)

    def eval_g_jac(self, x, g=None, values=None):
        """Fused eval_g + eval_jac_g.  Under pin_user_buffers the library page-locks the arrays it is handed and keeps them
        registered: when the caller passes none, results land in buffers this object owns for its lifetime (never in
        temporaries that the garbage collector could free while still registered) and copies are returned."""
        m = self.m
        own = g is None or values is None
        if own and self.pin_user_buffers:
            if self._own is None or self._own[0].size != m:
                if self._own is not None:
                    self._ck(self.L.armour_release_host_buffers(self.h))
                self._own = (np.zeros(m), np.zeros(m * 7))
            self._ck(self.L.armour_eval_g_jac(self.h, _dp(_vec(x, 7)), _dp(self._own[0]), _dp(self._own[1])))
            if g is not None:
                g[:] = self._own[0]
            if values is not None:
                values[:] = self._own[1]
            return (self._own[0].copy() if g is None else g), (self._own[1].copy() if values is None else values).reshape(m, 7)
        g = np.zeros(m) if g is None else g
        values = np.zeros(m * 7) if values is None else values
        self._ck(self.L.armour_eval_g_jac(self.h, _dp(_vec(x, 7)), _dp(g), _dp(values)))
        return g, values.reshape(m, 7)

    def eval_batch(self, xs, g=None, values=None, first=0, want_jac=False):
        """One launch for problems first .. first+len(xs)-1 of the last batch build, problem y at xs[y] (armour_eval_batch).
        g: (count, m) array written in place (allocated when None); values: (count, m * 7) array, or allocated when
        want_jac.  Under pin_user_buffers the arrays handed in must stay alive until release_host_buffers()/close()."""
        xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, 7)
        n, m = xs.shape[0], self.m
        if g is None:
            g = np.zeros((n, m))
        if values is None and want_jac:
            values = np.zeros((n, m * 7))
        assert g.dtype == np.float64 and g.flags.c_contiguous and g.size == n * m
        assert values is None or (values.dtype == np.float64 and values.flags.c_contiguous and values.size == n * m * 7)
        self._ck(self.L.armour_eval_batch(self.h, C.c_int(first), C.c_int(n), _dp(xs), _dp(g), _dp(values) if values is not None else None))
        return (g, values) if values is not None else g

    def eval_batch_resident(self, xs, g=None, first=0):
        """eval_batch with the Jacobians left on the device (armour_eval_batch_resident); g: (count, m) array or None."""
        xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, 7)
        assert g is None or (g.dtype == np.float64 and g.flags.c_contiguous and g.size == xs.shape[0] * self.m)
        self._ck(self.L.armour_eval_batch_resident(self.h, C.c_int(first), C.c_int(xs.shape[0]), _dp(xs), _dp(g) if g is not None else None))
        return g

    def get_batch_jacobian(self, y):
        v = np.zeros(self.m * 7)
        self._ck(self.L.armour_get_batch_jacobian(self.h, C.c_int(y), _dp(v)))
        return v

    def last_eval_batch_ms(self):
        v = C.c_float()
        self._ck(self.L.armour_last_eval_batch_ms(self.h, C.byref(v)))
        return v.value

    def pinned_buffer_count(self):
        v = C.c_int()
        self._ck(self.L.armour_pinned_buffer_count(self.h, C.byref(v)))
        return v.value

    def eval_resident_burst(self, x, launches=20):
        """average device time (ms) of `launches` back-to-back device-resident evaluations"""
        v = C.c_float()
        self._ck(self.L.armour_eval_resident_burst(self.h, _dp(_vec(x, 7)), C.c_int(launches), C.byref(v)))
        return v.value

    def last_eval_host_us(self):
        v = C.c_double()
        self._ck(self.L.armour_last_eval_host_us(self.h, C.byref(v)))
        return v.value

    def set_kernel_timing(self, enabled):
        self._ck(self.L.armour_set_kernel_timing(self.h, C.c_int(1 if enabled else 0)))

    def jac_structure(self):
        m = self.m
        ir, jc = np.zeros(m * 7, dtype=np.int32), np.zeros(m * 7, dtype=np.int32)
        self._ck(self.L.armour_jac_structure(self.h, _ip(ir), _ip(jc)))
        return ir, jc

    def check_feasible(self, g):
        f = C.c_int()
        self._ck(self.L.armour_check_feasible(self.h, _dp(_vec(g, self.m)), C.byref(f)))
        return bool(f.value)

    def standin_solve(self, q_des, t_plan=0.5):
        """Stand-in for the Ipopt solve (Ipopt is not installed): returns k, feasible, iterations, evaluations."""
        k = np.zeros(7)
        f, it, ev = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.armour_standin_solve(self.h, _dp(_vec(q_des, 7)), C.c_double(t_plan), _dp(k), C.byref(f), C.byref(it), C.byref(ev)))
        return k, bool(f.value), it.value, ev.value

    # ---- tables ----
    def get_pz(self, which, idx, s):
        w = TABLES[which] if isinstance(which, str) else which
        dims = np.zeros(2, dtype=np.int32)
        keys = np.zeros(256, dtype=np.uint64)
        coeffs = np.zeros(256 * 9)
        center, indep = np.zeros(9), np.zeros(9)
        idx, s = int(idx), int(s)
        n = self._ck(self.L.armour_get_pz(self.h, w, idx, s, _ip(dims), _up(keys), _dp(coeffs), _dp(center), _dp(indep)))
        dim = int(dims[0] * dims[1])
        return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n].copy(), coeffs=coeffs[: n * dim].reshape(n, dim).copy(),
                    center=center[:dim].copy(), independent=indep[:dim].copy())

    def torque_radius(self):
        out = np.zeros(self.T * NF)
        self._ck(self.L.armour_get_torque_radius(self.h, _dp(out)))
        return out.reshape(self.T, NF)

    def link_generators(self):
        out = np.zeros(self.T * NJ * 18)
        self._ck(self.L.armour_get_link_generators(self.h, _dp(out)))
        return out.reshape(self.T, NJ, 6, 3).transpose(0, 1, 3, 2)

    def taylor_remainders(self):
        c, s = np.zeros(NF * self.T * 2), np.zeros(NF * self.T * 2)
        self._ck(self.L.armour_get_taylor_remainders(self.h, _dp(c), _dp(s)))
        return c.reshape(NF, self.T, 2), s.reshape(NF, self.T, 2)

    def hyperplanes(self):
        n = self.T * NJ * self.n_obs * COMB
        A, d, dl = np.zeros(max(n, 1) * 3), np.zeros(max(n, 1)), np.zeros(max(n, 1))
        self._ck(self.L.armour_get_hyperplanes(self.h, _dp(A), _dp(d), _dp(dl)))
        shp = (self.T, NJ, self.n_obs, COMB)
        return A[: n * 3].reshape(shp + (3,)), d[:n].reshape(shp), dl[:n].reshape(shp)

    def link_sliced_center(self):
        out = np.zeros(self.T * NJ * 3)
        self._ck(self.L.armour_get_link_sliced_center(self.h, _dp(out)))
        return out.reshape(self.T, NJ, 3)

    # ---- stand-alone PZsparse arithmetic on the device ----
    def pz_binary(self, op, a, b, cap=1 << 14):
        def flat(z):
            keys = np.ascontiguousarray(z["keys"], dtype=np.uint64)
            co = np.ascontiguousarray(z["coeffs"], dtype=np.float64)
            ce, ind = _vec(z["center"]), _vec(z["independent"])
            return (z["rows"], z["cols"], len(keys), _up(keys), _dp(co), _dp(ce), _dp(ind)), [keys, co, ce, ind]

        fa, ka = flat(a)
        fb, kb = flat(b)
        dims = np.zeros(2, dtype=np.int32)
        keys = np.zeros(cap, dtype=np.uint64)
        coeffs = np.zeros(cap * 9)
        center, indep = np.zeros(9), np.zeros(9)
        n = self._ck(self.L.armour_pz_binary(self.h, PZ_OPS[op], *fa, *fb, cap, _ip(dims), _up(keys), _dp(coeffs), _dp(center), _dp(indep)))
        dim = int(dims[0] * dims[1])
        return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n].copy(), coeffs=coeffs[: n * dim].reshape(n, dim).copy(),
                    center=center[:dim].copy(), independent=indep[:dim].copy())
