"""The C-ABI shared library loads and exports every symbol include/nb2.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from tests.util import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nb2.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nb2_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from nimblephysics_b200 import _cabi

    L = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _declared_symbols()
    assert "nb2_step_forward" in syms and "nb2_step_backward" in syms and len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/nb2.h but not exported by libnb2.so"
    assert b"nb2" in ctypes.c_char_p(ctypes.cast(L.nb2_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_model_create_without_gpu_fails_loudly_not_silently():
    """On a GPU-less box nb2_model_create must report an error (no CPU fallback); on a GPU box it must succeed."""
    import torch

    import nimblephysics_b200 as nb
    from nimblephysics_b200 import _cabi
    from tests.util import load_raw

    cm = nb.compile_model(load_raw("cartpole"))
    if torch.cuda.is_available():
        dm = nb.DeviceModel(cm)
        assert dm.saved_words == 2 * 21 + 2
    else:
        with pytest.raises(_cabi.Nb2Error):
            nb.DeviceModel(cm)


def test_unsupported_model_is_rejected():
    import numpy as np

    import nimblephysics_b200 as nb
    from nimblephysics_b200 import _cabi

    sk = nb.Skeleton()
    parent = None
    for _ in range(70):  # more moving bodies than NB2_MAX_BODIES
        _, parent = sk.createRevoluteJointAndBodyNodePair(parent)
    w = nb.World()
    w.addSkeleton(sk)
    cm = nb.compile_model(nb.flatten_world(w))
    with pytest.raises(_cabi.Nb2Error):
        nb.DeviceModel(cm)
