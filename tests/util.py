import os

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
MODELS = os.path.join(ROOT, "tests", "golden", "models")


def load_raw(name):
    import nimblephysics_b200 as nb

    return nb.RawModel.load(os.path.join(MODELS, f"{name}.json"))


def sample_inputs(raw, B, seed=1234, q_scale=0.4, v_scale=1.0, tau_scale=5.0):
    """Seeded synthetic (q, qdot, tau) batches; fp32 values, shared by oracle (cast to fp64) and GPU."""
    rng = np.random.default_rng(seed)
    n, na = raw.ndof, len(raw.action_map)
    q = rng.uniform(-q_scale, q_scale, (B, n))
    v = rng.uniform(-v_scale, v_scale, (B, n))
    a = rng.uniform(-tau_scale, tau_scale, (B, na))
    s = np.concatenate([q, v], 1).astype(np.float32)
    return s, a.astype(np.float32), rng.normal(size=(B, 2 * n)).astype(np.float32)


def rel_err(x, ref):
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


