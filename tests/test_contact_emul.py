"""Contact / boxed-LCP stage: the device code (csrc/nb2_contact.cuh, compiled for the host by tests/host_emul) vs the
fp64 oracle over several chained steps with the LCP cache flowing from step to step.
Bar: contact set, LCP dimension, per-row ConstraintMapping labels and the solver-branch status word are BIT-EXACT;
impulses and next state within 1e-4 relative (inputs are fp32 rows on the device side)."""
import numpy as np
import pytest

import nimblephysics_b200 as nb
from oracle import binding as ob
from tests.host_emul.binding import EmulWorld
from tests.util import contact_inputs, load_raw, rel_err


@pytest.mark.parametrize("name", ["half_cheetah", "atlas_ground"])
def test_contact_forward_chain_matches_oracle(oracle_mod, name):
    raw = load_raw(name)
    cm = nb.compile_model(raw)
    ow, ew = ob.OracleContactWorld(raw), EmulWorld(cm)
    B, T = 6, 6
    s, a = contact_inputs(raw, name, B, seed=3)
    so, se = s.astype(np.float64).copy(), s.copy()
    xo = [None] * B
    xe = me = None
    seen_status = set()
    for t in range(T):
        r = ew.forward_contact(se, a, xe, me)
        for w in range(B):
            ro = ow.step_contact(so[w], a[w].astype(np.float64), xo[w])
            mo = ro["m"]
            assert r["nc"][w] == ro["nc"] and r["m"][w] == mo
            assert np.array_equal(r["labels"][w][:mo], ro["mapping"])          # bit-exact contact set / classification
            assert r["status"][w] == ro["status"]                               # same solver branch
            assert np.array_equal(r["cinfo"][w][: ro["nc"], 7:9].astype(int), ro["bodies"])
            assert np.array_equal(r["cinfo"][w][: ro["nc"], 9].astype(int), ro["type"])
            if mo:
                assert np.abs(r["x"][w][:mo] - ro["x"]).max() < 1e-5 * max(1.0, np.abs(ro["x"]).max())
            assert rel_err(r["next"][w], ro["next_state"]) < 1e-4
            so[w] = ro["next_state"]
            xo[w] = ro["x"] if mo else None
            seen_status.add(int(ro["status"]))
        # feed the oracle's fp64 state to both sides so rounding of the fp32 rows does not accumulate into tie flips
        se = so.astype(np.float32)
        so = se.astype(np.float64)
        xe, me = r["x"], r["m"]
    assert len(seen_status) >= 1


def test_oracle_contact_step_invariants(oracle_mod):
    """Properties the reference's own tests rely on (unittests/GradientTestUtils.hpp verifyRecoveredLCPConstraints,
    LCPUtils::isLCPSolutionValid): a solution that did not need the friction-drop fallback is a valid boxed-LCP
    solution; clamping normal rows end with ~zero relative normal velocity; positions advance with the pre-step velocity."""
    for name in ("half_cheetah", "atlas_ground"):
        raw = load_raw(name)
        ow = ob.OracleContactWorld(raw)
        s, a = contact_inputs(raw, name, 6, seed=9)
        total_contacts = 0
        for w in range(6):
            s64, a64 = s[w].astype(np.float64), a[w].astype(np.float64)
            r = ow.step_contact(s64, a64)
            n = raw.ndof
            total_contacts += r["nc"]
            if r["m"] == 0:
                continue
            if not (r["status"] & 16):
                cfm = 1e-4 if (r["status"] & 8) else 0.0
                assert ob.lcp_valid(r["A"] + cfm * np.eye(r["m"]), r["x"], r["b"], r["hi"], r["lo"], r["findex"])
            assert np.allclose(r["A"], r["A"].T, atol=1e-9)
            w_ = r["A"] @ r["x"] - r["b"]
            for j in range(r["m"]):
                if r["mapping"][j] == -2 and r["findex"][j] == -1 and not (r["status"] & 24):
                    assert abs(w_[j]) < 1e-6
            # generic joints: q+ = q + dt * v_t (World.cpp:307-322)
            q, v = s64[:n], s64[n:]
            idx = [d for d in range(n) if not (name == "atlas_ground" and d < 6)]
            assert np.allclose(r["next_state"][:n][idx], (q + raw.dt * v)[idx], atol=1e-12)
        assert total_contacts > 0
