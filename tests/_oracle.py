"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE — the CPU restatement of the
reference algorithm; see oracle/README.md).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

NJ = 7
NF = 7
COMB = 36


class OracleConfig(C.Structure):
    _fields_ = [
        ("num_time_steps", C.c_int),
        ("k_range", C.c_double * 7),
        ("mass_uncertainty", C.c_double),
        ("inertia_uncertainty", C.c_double),
        ("simplify_threshold", C.c_double),
        ("num_threads", C.c_int),
    ]


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle_armour.cpp", "oracle_controller.cpp", "oracle_pz.hpp", "oracle_interval.hpp", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "liboracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_build_ms.restype = C.c_double
        L.oracle_build_ms.argtypes = [C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_controller_create.restype = C.c_void_p
        L.oracle_controller_create.argtypes = [C.c_char_p, C.c_double]
        L.oracle_controller_destroy.argtypes = [C.c_void_p]
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


TABLES = {"cos_q": 0, "sin_q": 1, "R": 2, "R_t": 3, "qd_des": 4, "qda_des": 5, "qdda_des": 6, "links": 7, "u_nom": 8, "u_nom_int": 9}


class Oracle:
    """One planning problem on the CPU oracle.  Method names follow the reference's TNLP callbacks."""

    def __init__(self, T=128, k_range=None, mass_uncertainty=0.03, inertia_uncertainty=0.03, threshold=5e-4, num_threads=0):
        self.L = lib()
        cfg = OracleConfig()
        cfg.num_time_steps = T
        kr = [np.pi / 48] * 7 if k_range is None else list(k_range)
        for i in range(7):
            cfg.k_range[i] = kr[i]
        cfg.mass_uncertainty = mass_uncertainty
        cfg.inertia_uncertainty = inertia_uncertainty
        cfg.simplify_threshold = threshold
        cfg.num_threads = num_threads
        self.T = T
        self.k_range = np.array(kr)
        self.h = C.c_void_p(self.L.oracle_create(C.byref(cfg)))
        self.n_obs = 0

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build(self, q0, qd0, qdd0, obstacles):
        obs = _vec(obstacles)
        assert obs.size % 12 == 0
        self.n_obs = obs.size // 12
        rc = self.L.oracle_build(self.h, _dp(_vec(q0, 7)), _dp(_vec(qd0, 7)), _dp(_vec(qdd0, 7)), _dp(obs), C.c_int(self.n_obs))
        if rc != 0:
            raise RuntimeError("oracle_build failed: %d" % rc)
        return self.L.oracle_build_ms(self.h)

    def build_armtd(self, q0, qd0, jrs, k_range, obstacles):
        """ARMTD comparison planner (KPA/armtd_main.cu): jrs[6, 7, T] offline JRS tables, k_range[7]."""
        obs = _vec(obstacles)
        self.n_obs = obs.size // 12
        self.k_range = np.array(k_range, dtype=float)
        rc = self.L.oracle_build_armtd(self.h, _dp(_vec(q0, 7)), _dp(_vec(qd0, 7)), _dp(_vec(jrs, 6 * 7 * self.T)), _dp(_vec(k_range, 7)), _dp(obs), C.c_int(self.n_obs))
        if rc != 0:
            raise RuntimeError("oracle_build_armtd failed: %d" % rc)
        return self.L.oracle_build_ms(self.h)

    def op_stats(self):
        out = np.zeros(8, dtype=np.uint64)
        self.L.oracle_op_stats(self.h, _up(out))
        names = ["n_mul", "pair_products", "flops", "n_simplify", "simplify_in", "simplify_out", "max_simplify_in", "max_simplify_out"]
        return dict(zip(names, (int(v) for v in out)))

    def get_nlp_info(self):
        n, m, nnz, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.oracle_get_nlp_info(self.h, C.byref(n), C.byref(m), C.byref(nnz), C.byref(nh))
        return n.value, m.value, nnz.value, nh.value

    @property
    def m(self):
        return self.get_nlp_info()[1]

    def get_bounds_info(self):
        m = self.m
        xl, xu, gl, gu = np.zeros(7), np.zeros(7), np.zeros(m), np.zeros(m)
        self.L.oracle_get_bounds_info(self.h, _dp(xl), _dp(xu), _dp(gl), _dp(gu))
        return xl, xu, gl, gu

    def get_starting_point(self):
        x = np.ones(7)
        self.L.oracle_get_starting_point(self.h, _dp(x))
        return x

    def eval_f(self, q_des, t_plan, x):
        f = C.c_double()
        self.L.oracle_eval_f(self.h, _dp(_vec(q_des, 7)), C.c_double(t_plan), _dp(_vec(x, 7)), C.byref(f))
        return f.value

    def eval_grad_f(self, q_des, t_plan, x):
        g = np.zeros(7)
        self.L.oracle_eval_grad_f(self.h, _dp(_vec(q_des, 7)), C.c_double(t_plan), _dp(_vec(x, 7)), _dp(g))
        return g

    def eval_g(self, x):
        g = np.zeros(self.m)
        self.L.oracle_eval_g(self.h, _dp(_vec(x, 7)), _dp(g))
        return g

    def eval_jac_g(self, x):
        v = np.zeros(self.m * 7)
        self.L.oracle_eval_jac_g(self.h, _dp(_vec(x, 7)), _dp(v))
        return v.reshape(self.m, 7)

    def jac_structure(self):
        m = self.m
        ir, jc = np.zeros(m * 7, dtype=np.int32), np.zeros(m * 7, dtype=np.int32)
        self.L.oracle_jac_structure(self.h, _ip(ir), _ip(jc))
        return ir, jc

    def check_feasible(self, g):
        return bool(self.L.oracle_check_feasible(self.h, _dp(_vec(g, self.m))))

    def get_pz(self, which, idx, s):
        w = TABLES[which] if isinstance(which, str) else which
        dims = np.zeros(2, dtype=np.int32)
        idx, s = int(idx), int(s)
        n = self.L.oracle_get_pz(self.h, w, idx, s, _ip(dims), None, None, None, None)
        dim = int(dims[0] * dims[1])
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        coeffs = np.zeros((max(n, 1), dim))
        center, indep = np.zeros(dim), np.zeros(dim)
        self.L.oracle_get_pz(self.h, w, idx, s, _ip(dims), _up(keys), _dp(coeffs), _dp(center), _dp(indep))
        return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n], coeffs=coeffs[:n], center=center, independent=indep)

    def torque_radius(self):
        out = np.zeros(self.T * NF)
        self.L.oracle_get_torque_radius(self.h, _dp(out))
        return out.reshape(self.T, NF)  # [t, joint]

    def link_generators(self):
        out = np.zeros(self.T * NJ * 18)
        self.L.oracle_get_link_generators(self.h, _dp(out))
        return out.reshape(self.T, NJ, 6, 3).transpose(0, 1, 3, 2)  # [t, link, row, col]

    def taylor_remainders(self):
        c, s = np.zeros(NF * self.T * 2), np.zeros(NF * self.T * 2)
        self.L.oracle_get_taylor_remainders(self.h, _dp(c), _dp(s))
        return c.reshape(NF, self.T, 2), s.reshape(NF, self.T, 2)

    def hyperplanes(self):
        n = self.T * NJ * self.n_obs * COMB
        A, d, dl = np.zeros(n * 3), np.zeros(n), np.zeros(n)
        self.L.oracle_get_hyperplanes(self.h, _dp(A), _dp(d), _dp(dl))
        shp = (self.T, NJ, self.n_obs, COMB)
        return A.reshape(shp + (3,)), d.reshape(shp), dl.reshape(shp)

    def link_sliced_center(self):
        out = np.zeros(self.T * NJ * 3)
        self.L.oracle_get_link_sliced_center(self.h, _dp(out))
        return out.reshape(self.T, NJ, 3)

    def trace_interval(self, s, cap=20000):
        out = np.zeros(cap * 8, dtype=np.int32)
        n = self.L.oracle_trace_interval(self.h, s, cap, _ip(out))
        return out.reshape(cap, 8)[: min(n, cap)]


PZ_OPS = {"mul": 0, "add": 1, "sub": 2, "cross": 3, "simplify": 4, "reduce": 5, "transpose": 6, "add_one_dim0": 7, "add_one_dim1": 8, "add_one_dim2": 9,
          "cross_const_first": 10, "cross_const_second": 11}


def pz_binary(op, a, b=None, threshold=5e-4, cap=1 << 16):
    """a, b: dicts rows, cols, keys, coeffs[n, dim] (column-major flattening), center, independent."""
    L = lib()

    def flat(z):
        if z is None:
            return (0, 0, 0, None, None, None, None), []
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
    n = L.oracle_pz_binary(PZ_OPS[op], C.c_double(threshold), *fa, *fb, cap, _ip(dims), _up(keys), _dp(coeffs), _dp(center), _dp(indep))
    if n < 0:
        raise RuntimeError("oracle_pz_binary: %d" % n)
    dim = int(dims[0] * dims[1])
    return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n].copy(), coeffs=coeffs[: n * dim].reshape(n, dim).copy(),
                center=center[:dim].copy(), independent=indep[:dim].copy())



class OracleController:
    """CPU restatement of the robust-controller MEX (oracle/oracle_controller.cpp); same method names as
    armour_b200.controller.RobustController."""

    def __init__(self, robot_model_file, model_uncertainty=0.03, num_threads=1):
        h = lib().oracle_controller_create(str(robot_model_file).encode(), float(model_uncertainty))
        if not h:
            raise RuntimeError("oracle_controller_create failed for %s" % robot_model_file)
        self._h = C.c_void_p(h)
        self.n = lib().oracle_controller_num_joints(self._h)
        self.num_threads = num_threads

    def __del__(self):
        try:
            if self._h:
                lib().oracle_controller_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def model(self):
        a = np.zeros((self.n, 40))
        lib().oracle_controller_get_model(self._h, _dp(a))
        return a

    def _update(self, method, Kr, par, q, q_d, qd, qd_d, qd_dd):
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1, self.n) for a in (q, q_d, qd, qd_d, qd_dd)]
        single = np.asarray(q).ndim == 1
        count = arrs[0].shape[0]
        Kr = np.ascontiguousarray(np.broadcast_to(np.asarray(Kr, dtype=np.float64), (self.n,)))
        par = np.ascontiguousarray(np.asarray(par, dtype=np.float64))
        u, un, v = (np.zeros((count, self.n)) for _ in range(3))
        ui = np.zeros((count, self.n, 2))
        Vs = np.zeros(count)
        outside = lib().oracle_controller_update(self._h, C.c_int(method), C.c_int(count), _dp(Kr), _dp(par), *[_dp(a) for a in arrs],
                                                 _dp(u), _dp(un), _dp(v), _dp(ui), _dp(Vs), C.c_int(self.num_threads))
        if single:
            return u[0], un[0], v[0], ui[0], Vs[0], outside
        return u, un, v, ui, Vs, outside

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, q_d, qd, qd_d, qd_dd):
        return self._update(0, Kr, [alpha, V_max, r_norm_threshold], q, q_d, qd, qd_d, qd_dd)

    def update_althoff(self, Kr, Kp, Ki, max_error, q, q_d, qd, qd_d, qd_dd):
        return self._update(1, Kr, [Kp[0], Kp[1], 0.0], q, q_d, qd, qd_d, qd_dd)

    def rnea(self, q, qd, qda, qdd, gravity=True, interval=False):
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (q, qd, qda, qdd)]
        if interval:
            out = np.zeros((self.n, 2))
            lib().oracle_controller_rnea_interval(self._h, *[_dp(x) for x in a], C.c_int(1 if gravity else 0), _dp(out))
        else:
            out = np.zeros(self.n)
            lib().oracle_controller_rnea(self._h, *[_dp(x) for x in a], C.c_int(1 if gravity else 0), _dp(out))
        return out


REF_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref.so")
REFERENCE_TREE = "/root/reference/kinova_src/kinova_simulator_interfaces/kinova_planner_realtime"


def reference_available():
    """oracle/_ref/libref.so = the reference's own PZsparse/Trajectory/Dynamics sources compiled against the stand-in
    Eigen/Boost headers (oracle/Makefile target ref).  Built here when /root/reference is present; on the GPU box only
    the prebuilt file is used."""
    if os.path.isdir(REFERENCE_TREE):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "ref"])
    return os.path.exists(REF_LIB_PATH)


def ref_variant_path(base, variant=None):
    """oracle/_ref/<base>[_<variant>].so.  Variants (oracle/Makefile): "T512u5" = NUM_TIME_STEPS 512 and 5 % mass / inertia
    uncertainty (BASELINE.json configs[3]), "k24" = k_range pi/24 (KPR/debug_script.m:35), "nofma" (CUDA build only) = the
    reference's device code without FMA contraction."""
    return os.path.join(ORACLE_DIR, "_ref", base + ("_" + variant if variant else "") + ".so")


class Reference:
    """The reference's own reach-set build (KPR/armour_main.cu:94-205 call sequence) through oracle/ref_driver.cpp.
    T, k_range and the thresholds are the reference's compile-time values (T = 128, pi/48, 5e-4; other values through the
    temporarily patched variant builds, see ref_variant_path)."""

    def __init__(self, num_threads=None, variant=None):
        self.L = C.CDLL(ref_variant_path("libref", variant))
        self.L.ref_build.restype = C.c_void_p
        self.L.ref_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self.L.ref_destroy.argtypes = [C.c_void_p]
        self.L.ref_get_pz.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        self.L.ref_get_torque_radius.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_get_link_generators.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_slice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.ref_k_range.restype = C.c_double
        self.T = self.L.ref_num_time_steps()
        self.k_range = np.array([self.L.ref_k_range(i) for i in range(NF)])
        self.num_threads = num_threads or len(os.sched_getaffinity(0))
        self.h = None

    def build(self, q0, qd0, qdd0):
        self.close()
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (q0, qd0, qdd0)]
        h = self.L.ref_build(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, self.num_threads)
        if not h:
            raise RuntimeError("the reference build threw")
        self.h = C.c_void_p(h)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_pz(self, which, idx, s):
        w = TABLES[which] if isinstance(which, str) else which
        dims = np.zeros(2, dtype=np.int32)
        idx, s = int(idx), int(s)
        n = self.L.ref_get_pz(self.h, w, idx, s, dims.ctypes.data, None, None, None, None)
        dim = int(dims[0] * dims[1])
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        coeffs = np.zeros((max(n, 1), dim))
        center, indep = np.zeros(dim), np.zeros(dim)
        self.L.ref_get_pz(self.h, w, idx, s, dims.ctypes.data, keys.ctypes.data, coeffs.ctypes.data, center.ctypes.data, indep.ctypes.data)
        return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n], coeffs=coeffs[:n], center=center, independent=indep)

    def torque_radius(self):
        out = np.zeros(self.T * NF)
        self.L.ref_get_torque_radius(self.h, out.ctypes.data)
        return out.reshape(self.T, NF)

    def link_generators(self):
        out = np.zeros(self.T * NJ * 18)
        self.L.ref_get_link_generators(self.h, out.ctypes.data)
        return out.reshape(self.T, NJ, 6, 3).transpose(0, 1, 3, 2)

    def slice(self, which, idx, s, k):
        w = TABLES[which] if isinstance(which, str) else which
        k = np.ascontiguousarray(np.asarray(k, dtype=np.float64))
        lo, hi = np.zeros(9), np.zeros(9)
        n = self.L.ref_slice(self.h, w, int(idx), int(s), k.ctypes.data, lo.ctypes.data, hi.ctypes.data)
        return lo[:n], hi[:n]


REF_CUDA_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_cuda.so")


class ReferenceCuda:
    """The reference's complete path, its own CUDA kernels and TNLP callbacks included (oracle/ref_cuda_driver.cu over
    KPR/*.cu built by nvcc).  Needs a GPU.  Method names follow armtd_NLP."""

    def __init__(self, num_threads=None, variant=None):
        self.L = C.CDLL(ref_variant_path("libref_cuda", variant))
        self.L.refcuda_build.restype = C.c_void_p
        self.L.refcuda_build.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_void_p, C.c_int, C.c_int]
        self.L.refcuda_destroy.argtypes = [C.c_void_p]
        self.L.refcuda_k_range.restype = C.c_double
        self.L.refcuda_mass_uncertainty.restype = C.c_double
        self.L.refcuda_last_build_ms.restype = C.c_double
        self.L.refcuda_last_build_ms.argtypes = [C.c_void_p]
        self.T = self.L.refcuda_num_time_steps()
        self.k_range = np.array([self.L.refcuda_k_range(i) for i in range(NF)])
        self.mass_uncertainty = self.L.refcuda_mass_uncertainty()
        self.num_threads = num_threads or len(os.sched_getaffinity(0))
        self.h = None

    def build(self, q0, qd0, qdd0, q_des, obstacles, t_plan=0.5):
        self.close()
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (q0, qd0, qdd0, q_des)]
        obs = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 12))
        h = self.L.refcuda_build(*[x.ctypes.data for x in a], float(t_plan), obs.ctypes.data if obs.size else None, obs.shape[0], self.num_threads)
        if not h:
            raise RuntimeError("the reference build failed")
        self.h = C.c_void_p(h)
        n, m, a_, b_ = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert self.L.refcuda_get_nlp_info(self.h, C.byref(n), C.byref(m), C.byref(a_), C.byref(b_)) == 0
        self.n, self.m, self.nnz_jac_g = n.value, m.value, a_.value

    def close(self):
        if self.h:
            self.L.refcuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_bounds_info(self):
        xl, xu, gl, gu = np.zeros(self.n), np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        assert self.L.refcuda_get_bounds_info(self.h, self.n, self.m, _dp(xl), _dp(xu), _dp(gl), _dp(gu)) == 0
        return xl, xu, gl, gu

    def get_starting_point(self):
        x = np.zeros(self.n)
        assert self.L.refcuda_get_starting_point(self.h, self.n, self.m, _dp(x)) == 0
        return x

    def eval_f(self, x):
        f, g = C.c_double(), np.zeros(self.n)
        assert self.L.refcuda_eval_f(self.h, _dp(_vec(x, 7)), C.byref(f), _dp(g)) == 0
        return f.value, g

    def eval_g(self, x):
        g = np.zeros(self.m)
        assert self.L.refcuda_eval_g(self.h, _dp(_vec(x, 7)), self.m, _dp(g)) == 0
        return g

    def eval_jac_g(self, x):
        v = np.zeros(self.m * 7)
        assert self.L.refcuda_eval_jac_g(self.h, _dp(_vec(x, 7)), self.m, _dp(v)) == 0
        return v.reshape(self.m, 7)

    def check_feasible(self, x, g):
        g = np.ascontiguousarray(np.asarray(g, dtype=np.float64))
        return bool(self.L.refcuda_check_feasible(self.h, _dp(_vec(x, 7)), self.m, _dp(g)))

    def link_sliced_center(self):
        out = np.zeros(self.T * NJ * 3)
        self.L.refcuda_get_link_sliced_center(self.h, _dp(out))
        return out.reshape(self.T, NJ, 3)

    def last_build_ms(self):
        """The reference's own timed span (KPR/armour_main.cu:89-226): starts after the Obstacles constructor."""
        return self.L.refcuda_last_build_ms(self.h)


REF_ARMTD_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_armtd_cuda.so")


class ReferenceArmtd:
    """The reference's ARMTD comparison planner (KPA/*.cu built by nvcc, oracle/ref_armtd_driver.cu).  Needs a GPU."""

    def __init__(self, num_threads=None):
        self.L = C.CDLL(REF_ARMTD_LIB_PATH)
        self.L.refarmtd_build.restype = C.c_void_p
        self.L.refarmtd_build.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int]
        self.L.refarmtd_destroy.argtypes = [C.c_void_p]
        self.L.refarmtd_get_pz.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        self.L.refarmtd_get_link_generators.argtypes = [C.c_void_p, C.c_void_p]
        self.T = self.L.refarmtd_num_time_steps()
        self.num_threads = num_threads or len(os.sched_getaffinity(0))
        self.h = None

    def build(self, q0, qd0, jrs, k_range, q_des, obstacles):
        self.close()
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (q0, qd0, jrs, k_range, q_des)]
        assert a[2].size == 6 * 7 * self.T
        obs = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 12))
        h = self.L.refarmtd_build(*[x.ctypes.data for x in a], obs.ctypes.data if obs.size else None, obs.shape[0], self.num_threads)
        if not h:
            raise RuntimeError("the reference build failed")
        self.h = C.c_void_p(h)
        n, m, a_, b_ = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert self.L.refarmtd_get_nlp_info(self.h, C.byref(n), C.byref(m), C.byref(a_), C.byref(b_)) == 0
        self.n, self.m, self.nnz_jac_g = n.value, m.value, a_.value

    def close(self):
        if self.h:
            self.L.refarmtd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_pz(self, which, idx, s):
        w = TABLES[which] if isinstance(which, str) else which
        dims = np.zeros(2, dtype=np.int32)
        n = self.L.refarmtd_get_pz(self.h, w, int(idx), int(s), dims.ctypes.data, None, None, None, None)
        dim = int(dims[0] * dims[1])
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        coeffs = np.zeros((max(n, 1), dim))
        center, indep = np.zeros(dim), np.zeros(dim)
        self.L.refarmtd_get_pz(self.h, w, int(idx), int(s), dims.ctypes.data, keys.ctypes.data, coeffs.ctypes.data, center.ctypes.data, indep.ctypes.data)
        return dict(rows=int(dims[0]), cols=int(dims[1]), keys=keys[:n], coeffs=coeffs[:n], center=center, independent=indep)

    def link_generators(self):
        out = np.zeros(self.T * NJ * 18)
        self.L.refarmtd_get_link_generators(self.h, out.ctypes.data)
        return out.reshape(self.T, NJ, 6, 3).transpose(0, 1, 3, 2)

    def get_bounds_info(self):
        xl, xu, gl, gu = np.zeros(self.n), np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        assert self.L.refarmtd_get_bounds_info(self.h, self.n, self.m, _dp(xl), _dp(xu), _dp(gl), _dp(gu)) == 0
        return xl, xu, gl, gu

    def eval_f(self, x):
        f, g = C.c_double(), np.zeros(self.n)
        assert self.L.refarmtd_eval_f(self.h, _dp(_vec(x, 7)), C.byref(f), _dp(g)) == 0
        return f.value, g

    def eval_g(self, x):
        g = np.zeros(self.m)
        assert self.L.refarmtd_eval_g(self.h, _dp(_vec(x, 7)), self.m, _dp(g)) == 0
        return g

    def eval_jac_g(self, x):
        v = np.zeros(self.m * 7)
        assert self.L.refarmtd_eval_jac_g(self.h, _dp(_vec(x, 7)), self.m, _dp(v)) == 0
        return v.reshape(self.m, 7)

    def check_feasible(self, x, g):
        g = np.ascontiguousarray(np.asarray(g, dtype=np.float64))
        return bool(self.L.refarmtd_check_feasible(self.h, _dp(_vec(x, 7)), self.m, _dp(g)))


REF_CONTROLLER_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_controller.so")


class ReferenceController:
    """The reference's own robust controller (KRC/*.cpp compiled against the stand-in headers, oracle/ref_controller_driver.cpp):
    what the two MEX gateways compute per tick."""

    def __init__(self, robot_model_file, model_uncertainty=0.03):
        self.L = C.CDLL(REF_CONTROLLER_LIB_PATH)
        self.L.refctrl_create.restype = C.c_void_p
        self.L.refctrl_create.argtypes = [C.c_char_p, C.c_double]
        self.L.refctrl_destroy.argtypes = [C.c_void_p]
        h = self.L.refctrl_create(str(robot_model_file).encode(), float(model_uncertainty))
        if not h:
            raise RuntimeError("the reference could not load %s" % robot_model_file)
        self.h = C.c_void_p(h)
        self.n = self.L.refctrl_num_joints(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.refctrl_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _update(self, method, Kr, par, q, q_d, qd, qd_d, qd_dd):
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1, self.n) for a in (q, q_d, qd, qd_d, qd_dd)]
        count = arrs[0].shape[0]
        Kr = np.ascontiguousarray(np.broadcast_to(np.asarray(Kr, dtype=np.float64), (self.n,)))
        par = np.ascontiguousarray(np.asarray(par, dtype=np.float64))
        u, un, v = (np.zeros((count, self.n)) for _ in range(3))
        self.L.refctrl_update(self.h, C.c_int(method), C.c_int(count), _dp(Kr), _dp(par), *[_dp(a) for a in arrs], _dp(u), _dp(un), _dp(v))
        return u, un, v

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, q_d, qd, qd_d, qd_dd):
        return self._update(0, Kr, [alpha, V_max, r_norm_threshold], q, q_d, qd, qd_d, qd_dd)

    def update_althoff(self, Kr, Kp, Ki, max_error, q, q_d, qd, qd_d, qd_dd):
        return self._update(1, Kr, [Kp[0], Kp[1], Ki[0], Ki[1], max_error], q, q_d, qd, qd_d, qd_dd)

    def rnea(self, q, qd, qda, qdd, gravity=True):
        a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in (q, qd, qda, qdd)]
        tau, ti = np.zeros(self.n), np.zeros((self.n, 2))
        self.L.refctrl_rnea(self.h, *[_dp(x) for x in a], C.c_int(1 if gravity else 0), _dp(tau), _dp(ti))
        return tau, ti
