"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle.so (fp64 restatement) and, when present,
oracle/_ref/libodelcp.so (the reference's own ODE LCP solver compiled from /root/reference).
May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp")))
    if force or not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_model_create.restype = ctypes.c_void_p
    return _LIB


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


class OracleWorld:
    """One fp64 world built from a RawModel (nimblephysics_b200.modelspec.RawModel)."""

    def __init__(self, raw):
        L = lib()
        self.raw = raw
        self.n = raw.ndof
        self.na = len(raw.action_map)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = [i(raw.parent), i(raw.jtype), i(raw.dof_off), i(raw.mobile), f(raw.axis), f(raw.Tpj), f(raw.Tcj),
                      f(raw.mass), f(raw.com), f(raw.moment), f(raw.damping), f(raw.spring), f(raw.rest),
                      f(raw.pos_lo), f(raw.pos_hi), f(raw.vel_lo), f(raw.vel_hi), f(raw.force_lo), f(raw.force_hi),
                      f(raw.gravity), i(raw.action_map)]
        k = self._keep
        I = ctypes.c_int
        self.h = ctypes.c_void_p(L.orc_model_create(
            I(raw.nb), I(raw.ndof), _p(k[0], I), _p(k[1], I), _p(k[2], I), _p(k[3], I), _p(k[4]), _p(k[5]), _p(k[6]),
            _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]), _p(k[14]), _p(k[15]), _p(k[16]),
            _p(k[17]), _p(k[18]), _p(k[19]), ctypes.c_double(raw.dt), I(self.na), _p(k[20], I)))

    def __del__(self):
        try:
            lib().orc_model_destroy(self.h)
        except Exception:
            pass

    def step(self, state, action, want_qdd=False):
        s = np.ascontiguousarray(state, np.float64)
        a = np.ascontiguousarray(action, np.float64)
        out = np.empty(2 * self.n)
        qdd = np.empty(self.n)
        lib().orc_step(self.h, _p(s), _p(a), _p(out), _p(qdd))
        return (out, qdd) if want_qdd else out

    def step_f32(self, state, action):
        s = np.ascontiguousarray(state, np.float64)
        a = np.ascontiguousarray(action, np.float64)
        out = np.empty(2 * self.n)
        lib().orc_step_f32(self.h, _p(s), _p(a), _p(out))
        return out

    def jacobian(self, state, action):
        """d[q+;v+]/d[q;v;tau]  [2n,3n]"""
        s = np.ascontiguousarray(state, np.float64)
        a = np.ascontiguousarray(action, np.float64)
        J = np.empty((2 * self.n, 3 * self.n))
        lib().orc_jacobian(self.h, _p(s), _p(a), _p(J))
        return J

    def backprop(self, state, action, grad_next):
        s = np.ascontiguousarray(state, np.float64)
        a = np.ascontiguousarray(action, np.float64)
        g = np.ascontiguousarray(grad_next, np.float64)
        gs = np.empty(2 * self.n)
        ga = np.empty(self.na)
        lib().orc_backprop(self.h, _p(s), _p(a), _p(g), _p(gs), _p(ga))
        return gs, ga
