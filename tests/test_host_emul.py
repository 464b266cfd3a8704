"""Kernel math on the CPU: the per-world device functions (csrc/nb2_dyn.cuh) compiled as host code
(tests/host_emul) vs the fp64 oracle, on the same seeded inputs.  This is a development harness — the product
path never runs on the CPU — but it lets the GPU-less container check the exact code the kernels execute.
Tolerance: fp32 within 1e-4 rel of the fp64 reference path (BASELINE.json north_star)."""
import numpy as np
import pytest

import nimblephysics_b200 as nb
from tests.host_emul.binding import EmulWorld
from tests.util import load_raw, rel_err, sample_inputs

TOL = 1e-4


@pytest.mark.parametrize("name", ["cartpole", "half_cheetah", "atlas"])
@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("lanes", [1, 2, 4, 8])
def test_forward_backward_parity(oracle_mod, name, fp64, lanes):
    """lanes > 1: several threads cooperate on one world (trunk/limb stages, emulated with the lanes of odd worlds
    running in reverse order so that a cross-lane dependency inside one stage cannot hide)."""
    raw = load_raw(name)
    cm = nb.compile_model(raw, lanes=lanes)
    ow, ew = oracle_mod.OracleWorld(raw), EmulWorld(cm)
    B = 6
    s, a, g = sample_inputs(raw, B, seed=21)
    nxt, saved = ew.forward(s, a, fp64)
    gs, ga = ew.backward(s, a, saved, g, fp64)
    for w in range(B):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        ref = ow.step(s64, a64)
        rgs, rga = ow.backprop(s64, a64, g64)
        assert rel_err(nxt[w], ref) < TOL
        assert rel_err(gs[w], rgs) < TOL
        assert rel_err(ga[w], rga) < TOL


def test_weld_folding_and_canonical_frames_preserve_dynamics(oracle_mod):
    """compile_model folds welds / re-frames bodies; the oracle keeps the reference parametrisation."""
    from tests.test_oracle import _tree_world

    raw = nb.flatten_world(_tree_world())
    cm = nb.compile_model(raw, lanes=2)
    assert cm.nb == raw.nb - 1  # one weld folded
    ow, ew = oracle_mod.OracleWorld(raw), EmulWorld(cm)
    s, a, g = sample_inputs(raw, 3, seed=2)
    nxt, saved = ew.forward(s, a, True)
    gs, ga = ew.backward(s, a, saved, g, True)
    for w in range(3):
        ref = ow.step(s[w].astype(np.float64), a[w].astype(np.float64))
        rgs, rga = ow.backprop(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rel_err(nxt[w], ref) < 1e-6 and rel_err(gs[w], rgs) < 1e-6 and rel_err(ga[w], rga) < 1e-6


def test_gradient_clipping_at_bounds(oracle_mod):
    raw = load_raw("cartpole")
    cm = nb.compile_model(raw)
    ow, ew = oracle_mod.OracleWorld(raw), EmulWorld(cm)
    s, a, g = sample_inputs(raw, 4, seed=3)
    s[:, 0] = np.float32(raw.pos_hi[0])  # cart exactly on its upper position limit (15)
    s[2:, 0] = np.float32(raw.pos_lo[0])
    nxt, saved = ew.forward(s, a)
    gs, ga = ew.backward(s, a, saved, g)
    for w in range(4):
        rgs, rga = ow.backprop(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert (gs[w][0] == 0.0) == (rgs[0] == 0.0)  # bit-exact clipping decision
        assert rel_err(gs[w], rgs) < TOL


def test_action_map_subset(oracle_mod):
    raw = load_raw("atlas")
    raw.action_map = np.arange(6, raw.ndof, dtype=np.int32)  # root is unactuated
    cm = nb.compile_model(raw)
    ow, ew = oracle_mod.OracleWorld(raw), EmulWorld(cm)
    s, a, g = sample_inputs(raw, 2, seed=5, tau_scale=30.0)
    nxt, saved = ew.forward(s, a)
    gs, ga = ew.backward(s, a, saved, g)
    assert ga.shape[1] == raw.ndof - 6
    for w in range(2):
        ref = ow.step(s[w].astype(np.float64), a[w].astype(np.float64))
        rgs, rga = ow.backprop(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rel_err(nxt[w], ref) < TOL and rel_err(gs[w], rgs) < TOL and rel_err(ga[w], rga) < TOL


def test_mass_gradient_matches_finite_differences_of_the_oracle(oracle_mod):
    """a17: lossWrtMass = massVel^T g_v' (BackpropSnapshot.cpp:177-178) for every WrtMassBodyNodeEntryType the builder
    carries.  The kernel emits dL/d(m, h, Ibar) per canonical body; modelspec.inertia_param_jacobian maps that onto the
    registered mass vector (welded bodies are part of their owner).  Checked against central differences of the fp64
    oracle stepped with perturbed masses."""
    import copy

    from nimblephysics_b200 import modelspec as ms
    from tests.test_oracle import _tree_world

    for raw, entries in ((load_raw("atlas"), [(3, ms.INERTIA_MASS), (9, ms.INERTIA_COM), (14, ms.INERTIA_FULL), (20, ms.INERTIA_DIAGONAL),
                                               (25, ms.INERTIA_OFF_DIAGONAL), (0, ms.INERTIA_MASS)]),
                         (nb.flatten_world(_tree_world()), None)):
        cm = nb.compile_model(raw, lanes=2)
        if entries is None:  # every body of the small tree (one of them is welded into its parent), two kinds each
            entries = [(i, ms.INERTIA_MASS) for i in range(raw.nb) if cm.body_owner[i] >= 0] + \
                      [(i, ms.INERTIA_COM) for i in range(raw.nb) if cm.body_owner[i] >= 0]
            assert any(cm.orig_body[cm.body_owner[i]] != i for i, _ in entries)  # a welded body is among them
        ew = EmulWorld(cm)
        s, a, g = sample_inputs(raw, 2, seed=77)
        g[:, :raw.ndof] = 0.3 * g[:, :raw.ndof]
        nxt, saved = ew.forward(s, a, True)
        gs, ga, gi = ew.backward(s, a, saved, g, True, want_inertia_grad=True)
        P = ms.inertia_param_jacobian(raw, cm, entries)
        for w in range(2):
            gm = P @ gi[:, w].astype(np.float64)
            s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)

            def loss_at(row, h):
                r = copy.deepcopy(raw)
                k = 0
                for (bi, kind) in entries:
                    x = ms._mass_entry_value(kind, r.mass[bi], r.com[bi], r.moment[bi]).astype(np.float64)
                    for j in range(len(x)):
                        if k == row:
                            x[j] += h
                        k += 1
                    r.mass[bi], r.com[bi], r.moment[bi] = ms._apply_mass_entry(kind, x, r.mass[bi], r.com[bi], r.moment[bi])
                return float(g64 @ oracle_mod.OracleWorld(r).step(s64, a64))

            fd = np.array([(loss_at(j, 1e-5) - loss_at(j, -1e-5)) / 2e-5 for j in range(P.shape[0])])
            assert rel_err(gm, fd) < 2e-5, (gm, fd)


def test_sdf_loaded_atlas_parity(oracle_mod):
    """Atlas from its SDF description (loader.load_sdf_skeleton, creation order of SdfParser.cpp:843-880): kernels vs oracle."""
    raw = load_raw("atlas_sdf")
    assert raw.nb == 28 and raw.ndof == 33
    assert raw.body_names[:5] == ["pelvis", "ltorso", "mtorso", "utorso", "l_clav"]  # alphabetical, missing parents first
    cm = nb.compile_model(raw, lanes=4)
    ow, ew = oracle_mod.OracleWorld(raw), EmulWorld(cm)
    s, a, g = sample_inputs(raw, 4, seed=23)
    nxt, saved = ew.forward(s, a)
    gs, ga = ew.backward(s, a, saved, g)
    for w in range(4):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        rgs, rga = ow.backprop(s64, a64, g64)
        assert rel_err(nxt[w], ow.step(s64, a64)) < TOL and rel_err(gs[w], rgs) < TOL and rel_err(ga[w], rga) < TOL
