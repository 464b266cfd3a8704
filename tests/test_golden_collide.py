"""Golden vectors the reference's own unit tests hold for the contact generator of this path, applied to BOTH the oracle
(oracle/contact.hpp) and the device code compiled for the host (csrc/nb2_geom.cuh through tests/host_emul):

  unittests/unit/test_DARTCollide.cpp:554-592  BOX_BOX_FACE_FACE_COLLISION_ANNOTATION — the analytic collideBoxBox on a unit box at z = -0.5 and
      a 0.5-box at (0, 0.5, 0.25): 4 contacts; (+-0.25, 0.5, 0) typed EDGE_EDGE, (+-0.25, 0.25, 0) typed FACE_VERTEX.
  unittests/unit/test_DARTCollide.cpp:1639-1727 / 1818-1900 / 2572-2672: the sphere-vs-box cases (vertex, face; a capsule end acting as a sphere):
      expected point, normal and depth of the single contact (the reference exercises its mesh path there; the analytic box-sphere routine of
      this path must give the same geometry).
"""
import numpy as np
import pytest

import nimblephysics_b200 as nb
from tests.host_emul.binding import EmulWorld


@pytest.fixture(scope="module")
def oracle_mod():
    from oracle import binding
    binding.build()
    return binding


def _two_body_world(static_shape, static_pos, moving_shape):
    w = nb.World()
    w.setGravity([0, 0, 0])
    g = nb.Skeleton("fixed")
    g.setMobile(False)
    j, b = g.createWeldJointAndBodyNodePair()
    b.createShapeNode(static_shape).createCollisionAspect()
    T = nb.Isometry3(); T.set_translation(static_pos)
    j.setTransformFromParentBodyNode(T)
    w.addSkeleton(g)
    s = nb.Skeleton("moving")
    j, b = s.createFreeJointAndBodyNodePair()
    b.setMass(1.0)
    b.createShapeNode(moving_shape).createCollisionAspect()
    w.addSkeleton(s)
    return w


def _contacts(oracle_mod, world, pos, euler=(0, 0, 0)):
    """contacts of the two-body world with the moving body at `pos`: (oracle, device code on the host) as lists of (point, normal, depth, type)."""
    raw = nb.flatten_world(world)
    n = raw.ndof
    s = np.zeros(2 * n); s[0:3] = euler; s[3:6] = pos
    ro = oracle_mod.OracleContactWorld(raw).step_contact(s, np.zeros(len(raw.action_map)))
    orc = [(ro["point"][k], ro["normal"][k], ro["depth"][k], int(ro["type"][k])) for k in range(ro["nc"])]
    r = EmulWorld(nb.compile_model(raw)).forward_contact(s[None].astype(np.float32), np.zeros((1, len(raw.action_map)), np.float32))
    ci = r["cinfo"][0][: r["nc"][0]]
    dev = [(c[0:3].astype(np.float64), c[3:6].astype(np.float64), float(c[6]), int(c[9])) for c in ci]
    return orc, dev


def test_box_box_face_face_annotation_golden(oracle_mod):
    world = _two_body_world(nb.BoxShape([1.0, 1.0, 1.0]), [0.0, 0.0, -0.5], nb.BoxShape([0.5, 0.5, 0.5]))
    golden = {(0.25, 0.5): 3, (-0.25, 0.5): 3, (0.25, 0.25): 2, (-0.25, 0.25): 2}  # ContactType: EDGE_EDGE = 3, FACE_VERTEX = 2 (Contact.hpp:45-70)
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, [0.0, 0.5, 0.25])):
        assert len(cs) == 4, (name, cs)
        seen = {}
        for p, nrm, depth, typ in cs:
            assert abs(p[2]) < 1e-6 and abs(depth) < 1e-6, (name, p, depth)
            assert abs(abs(nrm[2]) - 1.0) < 1e-9, (name, nrm)
            seen[(round(float(p[0]), 6), round(float(p[1]), 6))] = typ
        assert seen == golden, (name, seen)


def test_sphere_box_golden_cases(oracle_mod):
    r = 0.5
    box = nb.BoxShape([1.0, 1.0, 1.0])
    # VERTEX_SPHERE_COLLISION (:1639-1727): sphere centre on the box diagonal, 0.01 inside the touching distance per axis
    c = 0.5 + np.sqrt(0.25 / 3) - 0.01
    world = _two_body_world(box, [0, 0, 0], nb.SphereShape(r))
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, [c, c, c])):
        assert len(cs) == 1, (name, cs)
        p, nrm, depth, typ = cs[0]
        assert np.allclose(p, [0.5, 0.5, 0.5], atol=1e-6), (name, p)
        assert np.allclose(np.abs(nrm), np.ones(3) / np.sqrt(3), atol=1e-6) and len({np.sign(v) for v in nrm}) == 1, (name, nrm)
        assert abs(depth - np.sqrt(3 * 0.01 * 0.01)) < 1e-6, (name, depth)
    # FACE_SPHERE_COLLISION (:1818-1900): sphere above the +x face, 0.01 deep
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, [0.5 + r - 0.01, 0.0, 0.0])):
        assert len(cs) == 1, (name, cs)
        p, nrm, depth, typ = cs[0]
        assert np.allclose(np.abs(nrm), [1, 0, 0], atol=1e-9) and abs(depth - 0.01) < 1e-6, (name, nrm, depth)
        assert abs(p[0] - 0.5) < 0.011 and np.allclose(p[1:], 0, atol=1e-9), (name, p)
    # CAPSULE_BOX_AS_SPHERE_COLLISION (:2572-2672): a capsule standing on its end above the +z face touches with its end sphere
    h, rc = 1.0, 0.25
    world = _two_body_world(box, [0, 0, 0], nb.CapsuleShape(rc, h))
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, [0.1, -0.1, 0.5 + h / 2 + rc - 0.01])):
        assert len(cs) == 1, (name, cs)
        p, nrm, depth, typ = cs[0]
        assert np.allclose(np.abs(nrm), [0, 0, 1], atol=1e-9) and abs(depth - 0.01) < 1e-6, (name, nrm, depth)
        assert np.allclose(p[:2], [0.1, -0.1], atol=1e-6) and abs(p[2] - 0.5) < 0.011, (name, p)


def _one(cs, name):
    assert len(cs) == 1, (name, cs)
    return cs[0]


def test_capsule_capsule_and_capsule_sphere_golden_cases(oracle_mod):
    """unittests/unit/test_DARTCollide.cpp: CAPSULE_CAPSULE_T_SHAPED (:2167-2245), X_SHAPED (:2248-2326), CAPSULE_SPHERE_END (:2424-2495) and
    CAPSULE_SPHERE_SIDE (:2498-2570), forward and "backwards" (objects swapped): expected point, normal (object 2 -> object 1), depth 0.01, type."""
    h, r1, r2 = 1.0, 0.4, 0.3
    px = r1 - 0.01 * r1 / (r1 + r2)
    cap1, cap2, sph2 = nb.CapsuleShape(r1, h), nb.CapsuleShape(r2, h), nb.SphereShape(r2)
    cases = [
        # (static shape, moving shape, moving position, moving rotation vector, expected point, expected normal, expected type)
        ("T", cap1, cap2, [r1 + r2 + h / 2 - 0.01, 0, 0], [0, np.pi / 2, 0], [px, 0, 0], [-1, 0, 0], 13),          # PIPE_SPHERE
        ("X", cap1, cap2, [0, r1 + r2 - 0.01, 0], [0, np.pi / 2, 0], [0, px, 0], [0, -1, 0], 15),                    # PIPE_PIPE
        ("sphere at the end", cap1, sph2, [0, 0, h / 2 + r1 + r2 - 0.01], [0, 0, 0], [0, 0, h / 2 + px], [0, 0, -1], 6),   # SPHERE_SPHERE
        ("sphere at the side", cap1, sph2, [r1 + r2 - 0.01, 0, 0], [0, 0, 0], [px, 0, 0], [-1, 0, 0], 13),           # PIPE_SPHERE
    ]
    for label, s_static, s_moving, pos, rot, ep, en, et in cases:
        world = _two_body_world(s_static, [0, 0, 0], s_moving)
        for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, pos, rot)):
            p, nrm, depth, typ = _one(cs, (label, name))
            assert np.allclose(p, ep, atol=1e-6) and np.allclose(nrm, en, atol=1e-6) and abs(depth - 0.01) < 1e-6 and typ == et, (label, name, p, nrm, depth, typ)
    # backwards: the small shape is object 1 (static, placed where it was), the big capsule moves but sits at the origin
    back = [
        ("T backwards", cap2, [r1 + r2 + h / 2 - 0.01, 0, 0], [0, np.pi / 2, 0], [px, 0, 0], [1, 0, 0], 14),        # SPHERE_PIPE
        ("sphere at the side backwards", sph2, [r1 + r2 - 0.01, 0, 0], [0, 0, 0], [px, 0, 0], [1, 0, 0], 14),       # SPHERE_PIPE
        ("sphere at the end backwards", sph2, [0, 0, h / 2 + r1 + r2 - 0.01], [0, 0, 0], [0, 0, h / 2 + px], [0, 0, 1], 6),
    ]
    for label, s_static, pos, rot, ep, en, et in back:
        w = nb.World(); w.setGravity([0, 0, 0])
        g = nb.Skeleton("fixed"); g.setMobile(False)
        j, b = g.createWeldJointAndBodyNodePair()
        b.createShapeNode(s_static).createCollisionAspect()
        from scipy.spatial.transform import Rotation
        T = nb.Isometry3(); T.set_translation(pos); T.set_rotation(Rotation.from_rotvec(rot).as_matrix())
        j.setTransformFromParentBodyNode(T)
        w.addSkeleton(g)
        s = nb.Skeleton("moving")
        j, b = s.createFreeJointAndBodyNodePair(); b.setMass(1.0)
        b.createShapeNode(cap1).createCollisionAspect()
        w.addSkeleton(s)
        for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, w, [0, 0, 0])):
            p, nrm, depth, typ = _one(cs, (label, name))
            assert np.allclose(p, ep, atol=1e-6) and np.allclose(nrm, en, atol=1e-6) and abs(depth - 0.01) < 1e-6 and typ == et, (label, name, p, nrm, depth, typ)


def test_sphere_sphere_contact(oracle_mod):
    """collideSphereSphere (DARTCollide.cpp:1812-1882): point at the radius-weighted position between the centres, normal from object 2 to 1."""
    r0, r1 = 0.4, 0.3
    world = _two_body_world(nb.SphereShape(r0), [0, 0, 0], nb.SphereShape(r1))
    d = np.array([1.0, 2.0, -2.0]) / 3.0
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, list(d * (r0 + r1 - 0.02)))):
        p, nrm, depth, typ = _one(cs, name)
        assert np.allclose(nrm, -d, atol=1e-6) and abs(depth - 0.02) < 1e-6 and typ == 6, (name, nrm, depth, typ)
        assert np.allclose(p, d * (r0 + r1 - 0.02) * r0 / (r0 + r1), atol=1e-6), (name, p)
    for name, cs in zip(("oracle", "device code"), _contacts(oracle_mod, world, list(d * (r0 + r1 + 0.01)))):
        assert len(cs) == 0, (name, cs)
