"""TEST HARNESS ONLY: runs csrc/nb2_dyn.cuh compiled for the host (tests/host_emul/emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

from nimblephysics_b200._cabi import Nb2ModelDesc, make_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.join(_HERE, "..", "..")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libemul.so")
        srcs = [os.path.join(_HERE, "emul.cpp")] + [os.path.join(_ROOT, "nimblephysics_b200", "csrc", f)
                                                      for f in ("nb2_dyn.cuh", "nb2_math.cuh", "nb2_model.h", "nb2_host_model.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                                   os.path.join(_HERE, "emul.cpp")])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class EmulWorld:
    def __init__(self, cm):
        self.cm = cm
        self.desc, self._keep = make_desc(cm)
        self.n, self.na = cm.ndof, len(cm.action_map)
        self.sw = lib().emul_saved_words(ctypes.byref(self.desc))

    def forward(self, state, action, fp64=False):
        state = np.ascontiguousarray(state, np.float32)
        action = np.ascontiguousarray(action, np.float32)
        B = state.shape[0]
        nxt = np.empty_like(state)
        saved = np.zeros((self.sw, B), np.float32)
        rc = lib().emul_forward(ctypes.byref(self.desc), B, _p(state), _p(action), _p(nxt), _p(saved), int(fp64))
        assert rc == 0
        return nxt, saved

    def backward(self, state, action, saved, gnext, fp64=False):
        state = np.ascontiguousarray(state, np.float32)
        action = np.ascontiguousarray(action, np.float32)
        gnext = np.ascontiguousarray(gnext, np.float32)
        B = state.shape[0]
        gs = np.empty_like(state)
        ga = np.empty_like(action)
        rc = lib().emul_backward(ctypes.byref(self.desc), B, _p(state), _p(action), _p(saved), _p(gnext), _p(gs),
                                 _p(ga), int(fp64))
        assert rc == 0
        return gs, ga
