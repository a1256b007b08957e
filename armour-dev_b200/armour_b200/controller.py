"""ctypes binding of the robust-controller C ABI (include/armour_controller_b200.h).

Mirrors the reference's two MEX entry points (KRC/kinova_controller.cpp, kinova_controller_ALTHOFF.cpp):
`RobustController.update(Kr, alpha, V_max, r_norm_threshold, q, q_d, qd, qd_d, qd_dd)` returns (u, u_nominal, v) like
`[u, tau, v] = kinova_controller(...)`, for one sample ([n]) or a batch ([count][n]).  No CPU fallback."""
import ctypes as C

import numpy as np

from . import lib, _dp

EXPORTS = [
    "armour_controller_create", "armour_controller_destroy", "armour_controller_num_joints", "armour_controller_update",
    "armour_controller_update_althoff", "armour_controller_rnea", "armour_controller_upload", "armour_controller_update_resident",
    "armour_controller_download", "armour_controller_last_ms", "armour_controller_release_host_buffers",
]


class ControllerError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("armour_controller error %d: %s" % (code, text))
        self.code = code


def _check(rc):
    if rc != 0:
        raise ControllerError(rc, lib().armour_last_error().decode())


def _arr(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    single = a.ndim == 1
    a = a.reshape(-1, n)
    return a, single


class RobustController:
    def __init__(self, robot_model_file, model_uncertainty=0.03, device=-1):
        self._h = C.c_void_p()
        L = lib()
        L.armour_controller_create.argtypes = [C.c_char_p, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
        L.armour_controller_destroy.argtypes = [C.c_void_p]
        _check(L.armour_controller_create(str(robot_model_file).encode(), float(model_uncertainty), int(device), C.byref(self._h)))
        n = C.c_int()
        _check(L.armour_controller_num_joints(self._h, C.byref(n)))
        self.n = n.value

    def release_host_buffers(self):
        _check(lib().armour_controller_release_host_buffers(self._h))

    def close(self):
        if self._h:
            lib().armour_controller_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _states(self, q, q_d, qd, qd_d, qd_dd):
        arrs = [_arr(a, self.n) for a in (q, q_d, qd, qd_d, qd_dd)]
        single = arrs[0][1]
        arrs = [a for a, _ in arrs]
        count = arrs[0].shape[0]
        assert all(a.shape == (count, self.n) for a in arrs)
        return arrs, count, single

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, q_d, qd, qd_d, qd_dd, debug=False, out=None):
        """kinova_controller: ARMOUR robust input.  debug=True also returns (u_interval [..., n, 2], V_sup, outside).
        out=(u, u_nominal, v): reusable [count, n] float64 arrays (large batches page-lock the arrays they are given; arrays
        allocated here are released again before returning)."""
        arrs, count, single = self._states(q, q_d, qd, qd_d, qd_dd)
        Kr = np.ascontiguousarray(np.broadcast_to(np.asarray(Kr, dtype=np.float64), (self.n,)))
        if out is not None:
            u, un, v = out
            assert all(a.shape == (count, self.n) and a.dtype == np.float64 and a.flags.c_contiguous for a in out)
        else:
            u, un, v = (np.zeros((count, self.n)) for _ in range(3))
        ui = np.zeros((count, self.n, 2)) if debug else None
        Vs = np.zeros(count) if debug else None
        outside = C.c_int()
        rc = lib().armour_controller_update(self._h, C.c_int(count), _dp(Kr), C.c_double(alpha), C.c_double(V_max), C.c_double(r_norm_threshold),
                                            *[_dp(a) for a in arrs], _dp(u), _dp(un), _dp(v), _dp(ui) if debug else None, _dp(Vs) if debug else None,
                                            C.byref(outside))
        if out is None or debug:
            self.release_host_buffers()   # arrays allocated here must not stay page-locked after they are freed
        _check(rc)
        res = (u, un, v) if not single else (u[0], un[0], v[0])
        if debug:
            return res + ((ui, Vs, outside.value) if not single else (ui[0], Vs[0], outside.value))
        return res

    def update_althoff(self, Kr, Kp, Ki, max_error, q, q_d, qd, qd_d, qd_dd, debug=False):
        """kinova_controller_ALTHOFF."""
        arrs, count, single = self._states(q, q_d, qd, qd_d, qd_dd)
        Kr = np.ascontiguousarray(np.broadcast_to(np.asarray(Kr, dtype=np.float64), (self.n,)))
        Kp = np.ascontiguousarray(np.asarray(Kp, dtype=np.float64)); Ki = np.ascontiguousarray(np.asarray(Ki, dtype=np.float64))
        u, un, v = (np.zeros((count, self.n)) for _ in range(3))
        ui = np.zeros((count, self.n, 2)) if debug else None
        outside = C.c_int()
        rc = lib().armour_controller_update_althoff(self._h, C.c_int(count), _dp(Kr), _dp(Kp), _dp(Ki), C.c_double(max_error), *[_dp(a) for a in arrs],
                                                    _dp(u), _dp(un), _dp(v), _dp(ui) if debug else None, C.byref(outside))
        self.release_host_buffers()
        _check(rc)
        out = (u, un, v) if not single else (u[0], un[0], v[0])
        if debug:
            return out + ((ui, outside.value) if not single else (ui[0], outside.value))
        return out

    def rnea(self, q, qd, qda, qdd, gravity=True, interval=False):
        """passRNEA (interval=False) or passRNEA_Int (interval=True: [..., n, 2] lower/upper)."""
        arrs = [_arr(a, self.n) for a in (q, qd, qda, qdd)]
        single = arrs[0][1]
        arrs = [a for a, _ in arrs]
        count = arrs[0].shape[0]
        out = np.zeros((count, self.n, 2)) if interval else np.zeros((count, self.n))
        rc = lib().armour_controller_rnea(self._h, C.c_int(count), *[_dp(a) for a in arrs], C.c_int(1 if gravity else 0),
                                          None if interval else _dp(out), _dp(out) if interval else None)
        _check(rc)
        return out[0] if single else out

    # device-resident path (bench.py `value`)
    def upload(self, q, q_d, qd, qd_d, qd_dd):
        arrs, count, _ = self._states(q, q_d, qd, qd_d, qd_dd)
        st = np.ascontiguousarray(np.stack(arrs))
        _check(lib().armour_controller_upload(self._h, C.c_int(count), _dp(st)))
        self._resident = count

    def update_resident(self, Kr, alpha, V_max, r_norm_threshold):
        Kr = np.ascontiguousarray(np.broadcast_to(np.asarray(Kr, dtype=np.float64), (self.n,)))
        _check(lib().armour_controller_update_resident(self._h, _dp(Kr), C.c_double(alpha), C.c_double(V_max), C.c_double(r_norm_threshold)))

    def download(self):
        out = np.zeros((3, self._resident, self.n))
        outside = C.c_int()
        _check(lib().armour_controller_download(self._h, _dp(out), C.byref(outside)))
        return out[0], out[1], out[2], outside.value

    def last_ms(self):
        v = C.c_double()
        _check(lib().armour_controller_last_ms(self._h, C.byref(v)))
        return v.value
