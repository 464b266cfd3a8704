"""LCP stage of the oracle, pinned against the real reference code and the reference's literal instances.
  * nb2_dantzig.cuh (the Dantzig restatement shared by the CUDA library and the oracle) vs the reference's own
    dSolveLCP compiled from /root/reference/dart/external/odelcpsolver (oracle/_ref/libodelcp.so);
  * the literal (A, x, lo, hi, b, fIndex) instances of unittests/unit/test_LCPUtils.cpp (tests/golden/lcp_fixtures.json):
    the chain's answer must satisfy LCPUtils::isLCPSolutionValid, as the reference's tests assert;
  * pinv_solve vs numpy.linalg.lstsq/pinv (Eigen completeOrthogonalDecomposition semantics: min-norm least squares)."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from tests.util import ROOT


def _contact_lcp(rng, nc, reg):
    n = 3 * nc
    J = rng.normal(size=(n, rng.integers(4, 14)))
    A = J @ J.T + reg * np.eye(n)
    b = rng.normal(size=n)
    lo, hi, fi = np.zeros(n), np.zeros(n), -np.ones(n, int)
    for c in range(nc):
        hi[3 * c] = np.inf
        mu = rng.uniform(0.2, 1.2)
        for k in (1, 2):
            lo[3 * c + k], hi[3 * c + k], fi[3 * c + k] = -mu, mu, 3 * c
    return A, b, lo, hi, fi


@pytest.mark.skipif(ob.ref_ode() is None, reason="oracle/_ref/libodelcp.so not built (needs /root/reference)")
@pytest.mark.parametrize("early", [True, False])
def test_dantzig_matches_reference_dSolveLCP(oracle_mod, early):
    rng = np.random.default_rng(1)
    for trial in range(300):
        A, b, lo, hi, fi = _contact_lcp(rng, int(rng.integers(1, 9)), 1e-6)
        xr, okr = ob.ref_dsolve_lcp(A, b, lo, hi, fi, early)
        xo, oko = ob.dantzig(A, b, lo, hi, fi, early)
        assert okr == oko
        if okr:
            assert np.abs(xr - xo).max() <= 1e-7 * max(1.0, np.abs(xr).max())


@pytest.mark.skipif(ob.ref_ode() is None, reason="oracle/_ref/libodelcp.so not built (needs /root/reference)")
def test_dantzig_on_exactly_singular_problems_mostly_matches(oracle_mod, capfd):
    """A = J J^T with redundant rows: pivoting on a singular factor is decided by rounding; the two implementations
    use different (mathematically equal) factor updates, so a few percent of such instances may differ."""
    rng = np.random.default_rng(2)
    agree = total = 0
    for trial in range(300):
        A, b, lo, hi, fi = _contact_lcp(rng, int(rng.integers(1, 9)), 0.0)
        xr, okr = ob.ref_dsolve_lcp(A, b, lo, hi, fi, True)
        xo, oko = ob.dantzig(A, b, lo, hi, fi, True)
        total += 1
        if okr == oko and (not okr or np.abs(xr - xo).max() <= 1e-6 * max(1.0, np.abs(xr).max())):
            agree += 1
    capfd.readouterr()
    assert agree >= 0.95 * total, (agree, total)


def test_reference_literal_instances(oracle_mod):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "lcp_fixtures.json")))
    assert len(fx) >= 6
    for f in fx:
        n = f["n"]
        A = np.array(f["A"]).reshape(n, n)
        b, lo, hi = np.array(f["b"]), np.array(f["lo"]), np.array(f["hi"])
        fi = np.array(f["findex"])
        x, mapping, status = ob.solve_chain(A, b, lo, hi, fi)
        friction_dropped = bool(status & 16)
        if not friction_dropped:  # the final fallback is allowed to violate friction rows (BoxedLcpConstraintSolver.cpp:657)
            assert ob.lcp_valid(A + (1e-4 * np.eye(n) if status & 8 else 0), x, b, hi, lo, fi), (f["name"], status, x)
        assert np.isfinite(x).all()
        assert set(np.unique(mapping)) <= set(range(-4, n))


def test_pinv_solve_is_min_norm_least_squares(oracle_mod):
    rng = np.random.default_rng(3)
    for m, n, r in [(6, 6, 6), (12, 12, 6), (9, 9, 3), (5, 8, 4), (8, 5, 5), (24, 24, 12)]:
        Q = rng.normal(size=(m, r)) @ rng.normal(size=(r, n))
        b = Q @ rng.normal(size=n) if rng.random() < 0.7 else rng.normal(size=m)
        x = ob.pinv_solve(Q, b)
        ref = np.linalg.pinv(Q) @ b
        assert np.abs(x - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
