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




def contact_inputs(raw, name, B, seed=0):
    """Seeded contact-rich inputs: half-cheetah bench pose q[1]=-0.1, q[2]=0.03 + noise
    (python/nimblephysics_benchmarks/half_cheetah_bench.py:19-20); Atlas q[0]=-pi/2, q[4] so the feet penetrate 6-10 mm
    (unittests/unit/test_AtlasGradients.cpp:235-236) + joint noise."""
    rng = np.random.default_rng(seed)
    n, na = raw.ndof, len(raw.action_map)
    s = np.zeros((B, 2 * n))
    for w in range(B):
        q = np.zeros(n)
        if name == "half_cheetah":
            q[1] = -0.1 + rng.normal(0, 0.01)
            q[2] = 0.03
            q[3:] = rng.normal(0, 0.02, n - 3)
            v = rng.normal(0, 0.1, n)
        else:
            q[0] = -0.5 * np.pi
            q[4] = -0.01 + rng.uniform(-0.004, 0.0)
            q[6:] = rng.normal(0, 0.01, n - 6)
            v = rng.normal(0, 0.05, n)
        s[w, :n], s[w, n:] = q, v
    a = rng.uniform(-2, 2, (B, na))
    if name != "half_cheetah":
        a[:, :6] = 0.0
    return s.astype(np.float32), a.astype(np.float32)


def multiarm_world(nlinks=5, length=0.2, gravity=(0.0, 0.0, 0.0)):
    """createMultiarmRobot(nlinks, length) (unittests/TestHelpers.hpp): a chain of revolute joints with alternating axes."""
    import nimblephysics_b200 as nb

    w = nb.World()
    w.setGravity(list(gravity))
    w.setTimeStep(1e-3)
    sk = nb.Skeleton("arm")
    parent = None
    axes = [[0, 0, 1], [0, 1, 0], [1, 0, 0]]
    for k in range(nlinks):
        j, b = sk.createRevoluteJointAndBodyNodePair(parent)
        j.setAxis(axes[k % 3])
        T = nb.Isometry3()
        T.set_translation([0, 0, length if k else 0.0])
        j.setTransformFromParentBodyNode(T)
        b.setMass(1.0 + 0.1 * k)
        b.setLocalCOM([0.01 * k, 0.0, length / 2])
        b.setMomentOfInertia(0.02, 0.03, 0.01, 0.001, 0.0, 0.002)
        parent = b
    w.addSkeleton(sk)
    return w
