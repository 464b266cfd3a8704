"""Minimal .urdf / .skel -> Skeleton/World loaders.

Keeps the reference's loader *surface* (``loadWorld(path)``,
``World.loadSkeleton(path)``; reference python/nimblephysics/loader.py:12-15,
dart/utils/UniversalLoader.cpp:39-90) with a new implementation that only
understands what the hot path needs: revolute / continuous / prismatic /
fixed / floating joints, inertials, and box / sphere / capsule collision
primitives.  Mesh colliders are skipped (reference needs assimp for them;
SURVEY §7 "mesh colliders").

Semantics restated from the reference (not its code):
  * URDF root link gets a FreeJoint unless it is called "world"
    (dart/utils/urdf/DartLoader.cpp:206-239); fixed -> WeldJoint (:483);
    limits/damping (:402-436); inertia rotated by the inertial rpy (:519-538).
    Child links are visited in joint-name order (urdfdom keeps joints in a
    name-sorted map).
  * .skel: body <transformation> is the body's world pose at q=0, the joint
    <transformation> is child-body->joint, parent->joint is derived
    (dart/utils/SkelParser.cpp:1539-1561); a body with <inertia> but no
    <moment_of_inertia> takes the moment of its first shape, unrotated
    (:619-645); euler XYZ angles (dart/math/Geometry.cpp:1767-1797).
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import warnings

import numpy as np

from .world import (FREE, PRISMATIC, REVOLUTE, WELD, BodyNode, BoxShape, CapsuleShape, Joint, ShapeNode,
                    Skeleton, SphereShape, World)


def _floats(s: str) -> List[float]:
    return [float(x) for x in s.split()]


def _rpy_to_R(rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _euler_xyz_to_R(a) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = math.cos(a[0]), math.sin(a[0]), math.cos(a[1]), math.sin(a[1]), math.cos(a[2]), math.sin(a[2])
    return np.array([
        [cy * cz, -cy * sz, sy],
        [cx * sz + cz * sx * sy, cx * cz - sx * sy * sz, -cy * sx],
        [sx * sz - cx * cz * sy, cz * sx + cx * sy * sz, cx * cy],
    ])


def _T(R=None, p=None) -> np.ndarray:
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if p is not None:
        T[:3, 3] = p
    return T


_SKIPPED_WARNED = set()


def _warn_skipped_collider(geom_el, where: str):
    """A collision geometry this path has no generator for (mesh, cylinder, ...): say so ONCE per kind and file instead of dropping it silently —
    a robot standing on mesh feet in the reference would fall through the floor here."""
    kinds = [c.tag for c in list(geom_el)] if geom_el is not None else []
    key = (where, tuple(kinds))
    if kinds and key not in _SKIPPED_WARNED:
        _SKIPPED_WARNED.add(key)
        warnings.warn(f"nimblephysics_b200 loader: collision geometry {kinds} in {where} has no contact generator on the GPU path and is "
                      "SKIPPED (supported: box, sphere, capsule)", stacklevel=3)


def _urdf_origin(el) -> np.ndarray:
    if el is None:
        return np.eye(4)
    xyz = _floats(el.get("xyz", "0 0 0"))
    rpy = _floats(el.get("rpy", "0 0 0"))
    return _T(_rpy_to_R(rpy), xyz)


def _urdf_geometry(geom_el):
    if geom_el is None:
        return None
    box = geom_el.find("box")
    if box is not None:
        return BoxShape(_floats(box.get("size")))
    sph = geom_el.find("sphere")
    if sph is not None:
        return SphereShape(float(sph.get("radius")))
    cap = geom_el.find("capsule")
    if cap is not None:
        return CapsuleShape(float(cap.get("radius")), float(cap.get("length", cap.get("height", "0"))))
    _warn_skipped_collider(geom_el, "a .urdf collision element")
    return None  # cylinder / mesh: not a supported collider


def load_urdf_skeleton(path: str) -> Skeleton:
    root = ET.parse(path).getroot()
    skel = Skeleton(root.get("name", "skeleton"))
    links: Dict[str, ET.Element] = {l.get("name"): l for l in root.findall("link")}
    joints = sorted(root.findall("joint"), key=lambda j: j.get("name"))
    children: Dict[str, List[ET.Element]] = {name: [] for name in links}
    is_child = set()
    for j in joints:
        p, c = j.find("parent").get("link"), j.find("child").get("link")
        if p not in links or c not in links:
            raise ValueError(f"URDF joint {j.get('name')} references a missing link")
        children[p].append(j)
        is_child.add(c)
    roots = [n for n in links if n not in is_child]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")

    def fill_body(body: BodyNode, link_el):
        inertial = link_el.find("inertial")
        if inertial is not None:
            T = _urdf_origin(inertial.find("origin"))
            body.com = T[:3, 3].copy()
            m = inertial.find("mass")
            body.mass = float(m.get("value")) if m is not None else 0.0
            ine = inertial.find("inertia")
            if ine is not None:
                g = lambda k: float(ine.get(k, "0"))
                J = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
                R = T[:3, :3]
                body.moment = R @ J @ R.T
        for col in link_el.findall("collision"):
            shape = _urdf_geometry(col.find("geometry"))
            if shape is None:
                continue
            sn = ShapeNode(shape, _urdf_origin(col.find("origin")))
            sn.has_collision = True
            body.shapes.append(sn)

    def make_joint(jel, parent_body: Optional[BodyNode], child_name: str) -> BodyNode:
        jt = jel.get("type")
        kind = {"revolute": REVOLUTE, "continuous": REVOLUTE, "prismatic": PRISMATIC, "fixed": WELD,
                "floating": FREE}.get(jt)
        if kind is None:
            raise NotImplementedError(f"URDF joint type '{jt}' is outside the hot-path scope")
        joint, body = skel._create(kind, parent_body, jel.get("name"), child_name)
        joint.T_pj = _urdf_origin(jel.find("origin"))
        if kind in (REVOLUTE, PRISMATIC):
            ax = jel.find("axis")
            joint.setAxis(_floats(ax.get("xyz")) if ax is not None else [1.0, 0.0, 0.0])
            lim = jel.find("limit")
            if lim is not None:
                lo, hi = float(lim.get("lower", "0")), float(lim.get("upper", "0"))
                vel, eff = float(lim.get("velocity", "0")), float(lim.get("effort", "0"))
                joint.pos_lo[0], joint.pos_hi[0] = lo, hi
                joint.vel_lo[0], joint.vel_hi[0] = -vel, vel
                joint.force_lo[0], joint.force_hi[0] = -eff, eff
                if lo > 0 or hi < 0:  # zero outside the limits -> start in the middle (DartLoader.cpp:413-431)
                    joint.init_pos[0] = 0.5 * (lo + hi)
                    joint.rest[0] = joint.init_pos[0]
            if jt == "continuous":
                joint.pos_lo[0], joint.pos_hi[0] = -math.inf, math.inf
            dyn = jel.find("dynamics")
            if dyn is not None:
                joint.damping[0] = float(dyn.get("damping", "0"))
        fill_body(body, links[child_name])
        return body

    def recurse(parent_body: Optional[BodyNode], link_name: str):
        for jel in children[link_name]:
            cname = jel.find("child").get("link")
            body = make_joint(jel, parent_body, cname)
            recurse(body, cname)

    rname = roots[0]
    if rname == "world":
        recurse(None, rname)
    else:
        joint, body = skel._create(FREE, None, "rootJoint", rname)
        fill_body(body, links[rname])
        recurse(body, rname)
    return skel


# ---------------------------------------------------------------- .skel ----
def _skel_T(el) -> np.ndarray:
    if el is None:
        return np.eye(4)
    v = _floats(el.text)
    return _T(_euler_xyz_to_R(v[3:6]), v[0:3])


def _skel_shape(geom_el):
    if geom_el is None:
        return None
    box = geom_el.find("box")
    if box is not None:
        return BoxShape(_floats(box.find("size").text))
    sph = geom_el.find("sphere")
    if sph is not None:
        return SphereShape(float(sph.find("radius").text))
    cap = geom_el.find("capsule")
    if cap is not None:
        return CapsuleShape(float(cap.find("radius").text), float(cap.find("height").text))
    ell = geom_el.find("ellipsoid")
    if ell is not None:
        s = _floats(ell.find("size").text)
        if abs(s[0] - s[1]) < 1e-12 and abs(s[1] - s[2]) < 1e-12:
            return SphereShape(0.5 * s[0])
    return None


def _read_skel_skeleton(sk_el) -> Skeleton:
    skel = Skeleton(sk_el.get("name", "skeleton"))
    mob = sk_el.find("mobile")
    if mob is not None:
        skel.mobile = mob.text.strip().lower() in ("1", "true")
    frame = _skel_T(sk_el.find("transformation"))
    body_info = {}
    for b_el in sk_el.findall("body"):
        name = b_el.get("name")
        W = frame @ _skel_T(b_el.find("transformation"))
        body_info[name] = (b_el, W)
    joint_els = sk_el.findall("joint")
    by_child = {}
    for j_el in joint_els:
        c = j_el.find("child").text.strip()
        if c in by_child:
            continue  # a body keeps its first parent joint (SkelParser.cpp:1563-1573)
        by_child[c] = j_el
    created: Dict[str, BodyNode] = {}

    def create(child_name: str):
        if child_name in created:
            return created[child_name]
        j_el = by_child[child_name]
        pname = j_el.find("parent").text.strip()
        parent_body = None
        parentW = np.eye(4)
        if pname != "world" or "world" in body_info:
            if pname not in created:
                if pname in by_child:
                    create(pname)
                else:
                    # parent body without a joint: reference inserts a root FreeJoint (SkelParser.cpp:1010-1025)
                    _, W = body_info[pname]
                    j, b = skel._create(FREE, None, "root", pname)
                    j.T_pj = W.copy()
                    _fill_skel_body(b, body_info[pname][0])
                    created[pname] = b
            parent_body = created[pname]
            parentW = body_info[pname][1]
        jt = j_el.get("type")
        kind = {"weld": WELD, "revolute": REVOLUTE, "prismatic": PRISMATIC, "free": FREE}.get(jt)
        if kind is None:
            raise NotImplementedError(f".skel joint type '{jt}' is outside the hot-path scope")
        joint, body = skel._create(kind, parent_body, j_el.get("name"), child_name)
        b_el, childW = body_info[child_name]
        T_cj = _skel_T(j_el.find("transformation"))
        joint.T_cj = T_cj
        joint.T_pj = np.linalg.inv(parentW) @ childW @ T_cj
        if kind in (REVOLUTE, PRISMATIC):
            ax = j_el.find("axis")
            joint.setAxis(_floats(ax.find("xyz").text))
            d = ax.find("damping")
            if d is not None:
                joint.damping[0] = float(d.text)
            dyn = ax.find("dynamics")
            if dyn is not None:
                for tag, arr in (("damping", joint.damping), ("spring_rest_position", joint.rest),
                                 ("spring_stiffness", joint.spring)):
                    e = dyn.find(tag)
                    if e is not None:
                        arr[0] = float(e.text)
            lim = ax.find("limit")
            if lim is not None:
                lo, hi = lim.find("lower"), lim.find("upper")
                if lo is not None:
                    joint.pos_lo[0] = float(lo.text)
                if hi is not None:
                    joint.pos_hi[0] = float(hi.text)
            ip = j_el.find("init_pos")
            if ip is not None and ip.text.strip():
                joint.init_pos[0] = _floats(ip.text)[0]
        elif kind == FREE:
            ip = j_el.find("init_pos")
            if ip is not None and ip.text.strip():
                joint.init_pos[:] = _floats(ip.text)[:6]
        _fill_skel_body(body, b_el)
        created[child_name] = body
        return body

    for j_el in joint_els:
        create(j_el.find("child").text.strip())
    return skel


def _fill_skel_body(body: BodyNode, b_el):
    g = b_el.find("gravity")
    if g is not None:
        body.gravity_mode = g.text.strip().lower() in ("1", "true")
    # shapes first (visualisation, then collision) — the first one may define the moment
    first_shape = None
    for tag in ("visualization_shape", "collision_shape"):
        for s_el in b_el.findall(tag):
            shape = _skel_shape(s_el.find("geometry"))
            if shape is None:
                continue
            if first_shape is None:
                first_shape = shape
            if tag == "collision_shape":
                sn = ShapeNode(shape, _skel_T(s_el.find("transformation")))
                sn.has_collision = True
                c = s_el.find("collidable")
                if c is not None:
                    sn.collidable = bool(float(c.text))
                body.shapes.append(sn)
    ine = b_el.find("inertia")
    if ine is not None:
        body.mass = float(ine.find("mass").text)
        off = ine.find("offset")
        if off is not None:
            body.com = np.array(_floats(off.text))
        moi = ine.find("moment_of_inertia")
        if moi is not None:
            g = lambda k: float(moi.find(k).text) if moi.find(k) is not None else 0.0
            body.setMomentOfInertia(g("ixx"), g("iyy"), g("izz"), g("ixy"), g("ixz"), g("iyz"))
        elif first_shape is not None:
            body.moment = first_shape.compute_inertia(body.mass)


def load_skel_world(path: str) -> World:
    root = ET.parse(path).getroot()
    w_el = root.find("world")
    world = World()
    world.setGravity([0.0, -9.81, 0.0])  # SkelParser default; overwritten below when present
    phys = w_el.find("physics")
    if phys is not None:
        ts = phys.find("time_step")
        if ts is not None:
            world.setTimeStep(float(ts.text))
        g = phys.find("gravity")
        if g is not None:
            world.setGravity(_floats(g.text))
    for sk_el in w_el.findall("skeleton"):
        world.addSkeleton(_read_skel_skeleton(sk_el))
    return world


def _sdf_pose(el) -> np.ndarray:
    """"x y z roll pitch yaw", extrinsic rotation = Rz(yaw) Ry(pitch) Rx(roll) (dart/utils/XmlHelpers.cpp:382-420)."""
    if el is None or el.text is None:
        return np.eye(4)
    v = _floats(el.text)
    return _T(_rpy_to_R(v[3:6]), v[0:3])


def _sdf_shape(geom_el):
    if geom_el is None:
        return None
    b = geom_el.find("box")
    if b is not None:
        return BoxShape(_floats(b.find("size").text))
    sp = geom_el.find("sphere")
    if sp is not None:
        return SphereShape(float(sp.find("radius").text))
    _warn_skipped_collider(geom_el, "a .sdf collision element")
    return None  # cylinders / meshes: no collider on the hot path (the reference meshes need assimp)


def load_sdf_skeleton(path: str) -> Skeleton:
    """dart/utils/sdf/SdfParser.cpp: link poses live in the model frame (:1117-1127); a joint's <pose> is child-link -> joint,
    parent-to-joint = parentWorld^-1 childWorld childToJoint (:1608-1624); <axis><xyz> is in the joint frame unless
    <use_parent_model_frame> (:1670-1690); links without a parent joint hang on a root FreeJoint placed at the link pose
    (:858-871).  Bodies are created in the reference's order — alphabetical by link name, a missing parent first
    (:843-880, std::map iteration) — because that order is the DoF order of the Skeleton."""
    root = ET.parse(path).getroot()
    model = root.find("model") if root.tag == "sdf" else root
    if model is None:
        w = root.find("world")
        model = w.find("model") if w is not None else None
    if model is None:
        raise ValueError("no <model> in the SDF file")
    skel = Skeleton(model.get("name", "skeleton"))
    st = model.find("static")
    if st is not None and st.text.strip().lower() in ("1", "true"):
        skel.setMobile(False)
    skel_frame = _sdf_pose(model.find("pose"))
    links = {l.get("name"): l for l in model.findall("link")}
    init_T = {name: skel_frame @ _sdf_pose(l.find("pose")) for name, l in links.items()}
    joints = {}  # keyed by child link name (SdfParser.cpp:1498-1530)
    for j in model.findall("joint"):
        child = j.find("child").text.strip()
        if child in joints:
            continue
        joints[child] = j
    created: Dict[str, BodyNode] = {}

    def fill_body(body: BodyNode, l):
        g = l.find("gravity")
        if g is not None:
            body.gravity_mode = g.text.strip().lower() in ("1", "true")
        ine = l.find("inertial")
        if ine is not None:
            m = ine.find("mass")
            if m is not None:
                body.setMass(float(m.text))          # Inertia::setMass first (scales the default moment) ...
            if ine.find("pose") is not None:
                body.com = _sdf_pose(ine.find("pose"))[:3, 3].copy()  # only the translation is used (:1143-1149)
            moi = ine.find("inertia")
            if moi is not None:                       # ... then the moment is set explicitly
                gv = lambda k: float(moi.find(k).text) if moi.find(k) is not None else 0.0
                body.setMomentOfInertia(gv("ixx"), gv("iyy"), gv("izz"), gv("ixy"), gv("ixz"), gv("iyz"))
        for col in l.findall("collision"):
            shape = _sdf_shape(col.find("geometry"))
            if shape is None:
                continue
            sn = ShapeNode(shape, _sdf_pose(col.find("pose")))
            sn.has_collision = True
            body.shapes.append(sn)

    def create(name: str):
        l = links[name]
        jel = joints.get(name)
        if jel is None:  # root: FreeJoint at the link's pose
            joint, body = skel._create(FREE, None, "root", name)
            joint.T_pj = init_T[name].copy()
            fill_body(body, l)
            created[name] = body
            return
        pname = jel.find("parent").text.strip()
        parent_body = created.get(pname)
        jt = jel.get("type")
        kind = {"revolute": REVOLUTE, "prismatic": PRISMATIC, "fixed": WELD}.get(jt)
        if kind is None:
            raise NotImplementedError(f"SDF joint type '{jt}' is outside the hot-path scope")
        joint, body = skel._create(kind, parent_body, jel.get("name"), name)
        child_world = init_T[name]
        parent_world = init_T[pname] if pname in init_T else np.eye(4)
        child_to_joint = _sdf_pose(jel.find("pose"))
        joint.T_cj = child_to_joint
        joint.T_pj = np.linalg.inv(parent_world) @ child_world @ child_to_joint
        if kind in (REVOLUTE, PRISMATIC):
            ax = jel.find("axis")
            if ax is None:
                raise ValueError(f"SDF joint {jel.get('name')} has no <axis>")
            xyz = np.array(_floats(ax.find("xyz").text))
            upf = ax.find("use_parent_model_frame")
            if upf is not None and upf.text.strip().lower() in ("1", "true"):
                parent_model_frame = np.linalg.inv(child_world @ child_to_joint) @ skel_frame
                xyz = parent_model_frame[:3, :3] @ xyz
            joint.setAxis(xyz)
            dyn = ax.find("dynamics")
            if dyn is not None and dyn.find("damping") is not None:
                joint.damping[0] = float(dyn.find("damping").text)
            lim = ax.find("limit")
            lo, hi = -math.inf, math.inf
            if lim is not None:
                if lim.find("lower") is not None:
                    lo = float(lim.find("lower").text)
                if lim.find("upper") is not None:
                    hi = float(lim.find("upper").text)
            joint.pos_lo[0], joint.pos_hi[0] = lo, hi
            if 0.0 < lo or hi < 0.0:  # zero outside the limits (:1722-1737)
                init = 0.5 * (lo + hi) if (math.isfinite(lo) and math.isfinite(hi)) else (lo if math.isfinite(lo) else hi)
                if math.isfinite(init):
                    joint.init_pos[0] = init
                    joint.rest[0] = init
        fill_body(body, l)
        created[name] = body

    remaining = sorted(links)
    cur = remaining[0] if remaining else None
    while remaining:
        name = cur if cur in remaining else remaining[0]
        jel = joints.get(name)
        if jel is not None:
            pname = jel.find("parent").text.strip()
            if pname not in created and pname != "world" and pname != "":
                if pname not in links:
                    raise ValueError(f"SDF joint {jel.get('name')} references the missing link {pname}")
                cur = pname  # create the parent before the current joint
                continue
        create(name)
        remaining.remove(name)
        cur = remaining[0] if remaining else None
    return skel


def load_skeleton(path: str) -> Skeleton:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".urdf":
        return load_urdf_skeleton(path)
    if ext == ".skel":
        w = load_skel_world(path)
        if len(w.skeletons) != 1:
            raise ValueError(".skel file holds several skeletons; use loadWorld()")
        return w.skeletons[0]
    if ext in (".sdf", ".world"):
        return load_sdf_skeleton(path)
    raise NotImplementedError(f"unsupported model file type '{ext}' (urdf, sdf and skel only)")


def loadWorld(path: str) -> World:
    """reference: python/nimblephysics/loader.py:12-15 -> UniversalLoader::loadWorld."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".skel":
        return load_skel_world(path)
    if ext == ".urdf":
        w = World()
        w.addSkeleton(load_urdf_skeleton(path))
        return w
    if ext in (".sdf", ".world"):
        w = World()
        w.addSkeleton(load_sdf_skeleton(path))
        return w
    raise NotImplementedError(f"unsupported world file type '{ext}'")
