"""Aircraft description: aero/motor table + rigid-body data read from a URDF.

The aero numbers are those of the reference's ``my_models/fixedwing/fixewing.yaml:1-71``.  The rigid body
(``fixedwing.urdf`` upstream) is NOT in the reference; ``params/fixedwing_placeholder.urdf`` is a stand-in
flagged PLACEHOLDER_NOT_UPSTREAM and any PyFlyt-style URDF (base link + fixed-joint links) can be loaded
instead.

Bullet loads such a URDF as a btMultiBody whose links hang off fixed joints; its articulated-body pass is then
exactly one rigid body expressed at the *base-link CoM* (point O).  ``composite_body`` reduces the link tree
to what the kernels need: total mass, composite CoM offset ``com`` from O, the inertia about O, the 6x6
spatial inertia inverse, and the link-CoM offsets at which PyFlyt applies surface/motor forces in LINK_FRAME
(reference call sites: envs/fixedwing_envs/fixedwing_base_env.py:230-237).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np
import yaml

_PARAMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "params")
DEFAULT_URDF = os.path.join(_PARAMS, "fixedwing_placeholder.urdf")
DEFAULT_AERO = os.path.join(_PARAMS, "fixedwing_aero.yaml")

SURFACE_ORDER = ("left_wing_flapped", "right_wing_flapped", "horizontal_tail", "vertical_tail", "main_wing")


def _vec(s: str | None, n: int = 3) -> np.ndarray:
    if s is None:
        return np.zeros(n)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if v.shape != (n,):
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def rpy_matrix(rpy: np.ndarray) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw -> rotation matrix (child frame -> parent frame)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


@dataclass
class Link:
    name: str
    mass: float
    com: np.ndarray          # CoM in the base-link frame
    inertia: np.ndarray      # 3x3 about the link CoM, base-link axes
    rot: np.ndarray          # link inertial frame -> base-link frame


@dataclass
class RigidBody:
    """Composite of a base link and its fixed-joint children, expressed at the base-link CoM."""
    mass: float
    com: np.ndarray                     # composite CoM relative to O (base-link CoM)
    inertia_o: np.ndarray               # 3x3 about O
    link_offsets: dict[str, np.ndarray]  # link CoM relative to O
    collision_points: np.ndarray        # [P,3] relative to O
    placeholder: bool = False
    links: list[Link] = field(default_factory=list)

    def spatial_inertia(self) -> np.ndarray:
        """6x6 [[I_O, M[c]x], [-M[c]x, M*1]] acting on (angular acc, linear acc of O)."""
        cx = skew(self.com)
        top = np.hstack([self.inertia_o, self.mass * cx])
        bot = np.hstack([-self.mass * cx, self.mass * np.eye(3)])
        return np.vstack([top, bot])

    def spatial_inertia_inv(self) -> np.ndarray:
        return np.linalg.inv(self.spatial_inertia())


def skew(v: np.ndarray) -> np.ndarray:
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=np.float64)


def parallel_axis(inertia_com: np.ndarray, mass: float, d: np.ndarray) -> np.ndarray:
    """Inertia about a point displaced by ``d`` from the CoM."""
    return inertia_com + mass * (float(d @ d) * np.eye(3) - np.outer(d, d))


def load_urdf(path: str = DEFAULT_URDF) -> RigidBody:
    tree = ET.parse(path)
    root = tree.getroot()
    with open(path, "r", encoding="utf-8") as f:
        placeholder = "PLACEHOLDER_NOT_UPSTREAM" in f.read()

    raw_links = {}
    for l in root.findall("link"):
        inertial = l.find("inertial")
        if inertial is None:
            raw_links[l.get("name")] = (0.0, np.zeros(3), np.eye(3), np.zeros((3, 3)))
            continue
        org = inertial.find("origin")
        xyz = _vec(org.get("xyz") if org is not None else None)
        rpy = _vec(org.get("rpy") if org is not None else None)
        mass = float(inertial.find("mass").get("value"))
        it = inertial.find("inertia")
        ixx, ixy, ixz = float(it.get("ixx")), float(it.get("ixy", 0)), float(it.get("ixz", 0))
        iyy, iyz, izz = float(it.get("iyy")), float(it.get("iyz", 0)), float(it.get("izz"))
        I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=np.float64)
        raw_links[l.get("name")] = (mass, xyz, rpy_matrix(rpy), I)

    parent_of, joint_tf = {}, {}
    for j in root.findall("joint"):
        if j.get("type") != "fixed":
            raise ValueError(f"joint {j.get('name')}: only fixed joints are supported (got {j.get('type')})")
        child, parent = j.find("child").get("link"), j.find("parent").get("link")
        org = j.find("origin")
        xyz = _vec(org.get("xyz") if org is not None else None)
        rpy = _vec(org.get("rpy") if org is not None else None)
        parent_of[child] = parent
        joint_tf[child] = (rpy_matrix(rpy), xyz)

    roots = [n for n in raw_links if n not in parent_of]
    if len(roots) != 1:
        raise ValueError(f"expected exactly one base link, found {roots}")
    base = roots[0]

    def to_base(name: str) -> tuple[np.ndarray, np.ndarray]:
        R, t = np.eye(3), np.zeros(3)
        while name != base:
            Rj, tj = joint_tf[name]
            R, t = Rj @ R, Rj @ t + tj
            name = parent_of[name]
        return R, t

    links: list[Link] = []
    for name, (mass, xyz, Rin, I) in raw_links.items():
        R, t = to_base(name)
        com = R @ xyz + t
        Rl = R @ Rin
        links.append(Link(name, mass, com, Rl @ I @ Rl.T, Rl))

    base_com = next(l.com for l in links if l.name == base)
    total = sum(l.mass for l in links)
    if total <= 0:
        raise ValueError("URDF has no mass")
    com = sum(l.mass * (l.com - base_com) for l in links) / total
    inertia_o = sum(parallel_axis(l.inertia, l.mass, l.com - base_com) for l in links)
    offsets = {l.name: l.com - base_com for l in links}

    pts = []
    cp = root.find("fwsim_collision_points")
    if cp is not None:
        pts = [_vec(p.get("xyz")) - base_com for p in cp.findall("point")]
    else:  # derive probe points from collision boxes (8 corners each)
        for l in root.findall("link"):
            R, t = to_base(l.get("name"))
            for c in l.findall("collision"):
                box = c.find("geometry/box")
                if box is None:
                    continue
                half = _vec(box.get("size")) / 2
                org = c.find("origin")
                o = _vec(org.get("xyz") if org is not None else None)
                Ro = rpy_matrix(_vec(org.get("rpy") if org is not None else None))
                for sx in (-1, 1):
                    for sy in (-1, 1):
                        for sz in (-1, 1):
                            pts.append(R @ (Ro @ (half * np.array([sx, sy, sz])) + o) + t - base_com)
    if not pts:
        pts = [np.zeros(3)]
    return RigidBody(float(total), com, inertia_o, offsets, np.array(pts, dtype=np.float64), placeholder, links)


def quat_matrix(q) -> np.ndarray:
    """pybullet (x, y, z, w) quaternion -> rotation matrix."""
    x, y, z, w = (float(v) for v in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def body_from_bullet(links: list[dict], collision_points=None) -> RigidBody:
    """Composite rigid body from what pybullet reports for a loaded multibody (scripts/record_pyflyt_golden.py dumps it):
    one dict per link with ``name``, ``mass``, ``inertia_diag`` (getDynamicsInfo local inertia diagonal, in the link's
    inertial frame), ``com`` (link CoM relative to the base-link CoM, base axes, body at the identity orientation) and
    ``inertial_quat`` (orientation of the inertial frame in base axes, xyzw).  The first link is the base.  This is the
    path by which a recording made with the REAL fixedwing.urdf replaces params/fixedwing_placeholder.urdf."""
    out: list[Link] = []
    for d in links:
        R = quat_matrix(d.get("inertial_quat", (0.0, 0.0, 0.0, 1.0)))
        I = R @ np.diag(np.asarray(d["inertia_diag"], dtype=np.float64)) @ R.T
        out.append(Link(str(d["name"]), float(d["mass"]), np.asarray(d["com"], dtype=np.float64), I, R))
    total = sum(l.mass for l in out)
    if total <= 0:
        raise ValueError("body has no mass")
    com = sum(l.mass * l.com for l in out) / total
    inertia_o = sum(parallel_axis(l.inertia, l.mass, l.com) for l in out)
    offsets = {l.name: l.com for l in out}
    pts = np.asarray(collision_points if collision_points is not None and len(collision_points) else [[0.0, 0.0, 0.0]],
                     dtype=np.float64)
    return RigidBody(float(total), com, inertia_o, offsets, pts, False, out)


@dataclass
class AeroTable:
    names: list[str]
    links: list[str]
    lift_unit: np.ndarray     # [5,3]
    fwd_unit: np.ndarray      # [5,3]
    cols: dict[str, np.ndarray]   # per-surface scalar columns
    motor: dict


def load_aero(path: str = DEFAULT_AERO) -> AeroTable:
    with open(path, "r", encoding="utf-8") as f:
        doc = yaml.safe_load(f)
    columns = doc["columns"]
    rows = {r[0]: dict(zip(columns, r)) for r in doc["surfaces"]}
    missing = [n for n in SURFACE_ORDER if n not in rows]
    if missing:
        raise ValueError(f"aero table lacks surfaces {missing}")
    ordered = [rows[n] for n in SURFACE_ORDER]
    scalar_cols = [c for c in columns if c not in ("name", "link", "lift_unit")]
    cols = {c: np.array([float(r[c]) for r in ordered], dtype=np.float64) for c in scalar_cols}
    fwd = np.tile(np.asarray(doc.get("forward_unit", [1, 0, 0]), dtype=np.float64), (len(ordered), 1))
    lift = np.array([r["lift_unit"] for r in ordered], dtype=np.float64)
    return AeroTable(list(SURFACE_ORDER), [r["link"] for r in ordered], lift, fwd, cols, dict(doc["motor"]))
