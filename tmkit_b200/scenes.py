"""Neutral scene descriptions used on both sides of the parity check.

A scene here is plain data (frames, geometry, allowed pairs) - no arithmetic beyond
constant folding (one exception: ``clutter_scene`` rejects random obstacles that touch the robot at the start
configuration with a conservative bounding test of its own, ``world_poses``).  The same description is pushed (a) through the product's
scene-ingest C ABI (``aa_rx_sg_add_frame_*`` / ``aa_rx_geom_*``, see
``include/tmkit_b200/amino_rx.h``) and (b) into flat numpy arrays for the CPU oracle.

Sources (reference = /root/reference, read at authoring time only):

* Sussman block scene, tables, classes : ``demo/baxter-sussman/table.robray:1-81``,
  ``class.robray:1-65``, ``sussman-0.robray:1-34``, ``sussman-1.robray:1-29``.
  Values are transcribed as data; the robray expressions (``(block_size +
  table_size)/2 + .1*mm``) are evaluated here.
* Allowed pairs                         : ``demo/baxter-sussman/allowed-collision.robray:1-31``.
* Start configuration                   : ``demo/baxter-sussman/q0.tmp:1-2``.
* Baxter kinematics / collision shapes  : **RESTATED FIXTURE**.  ``baxter_description``
  is not in the reference tree (``doc/md/tutorial.md:34-47`` tells the user to fetch
  it).  Joint origins, limits and collision cylinders are restated from memory of
  ``baxter.urdf``; the torso / pedestal meshes are procedural
  triangle soups.  Frame names are the ones the reference's allowed-collision list
  uses, so that list applies verbatim.

Conventions:  transform = quaternion x,y,z,w + translation x,y,z;
``rpy`` is URDF roll-pitch-yaw (R = Rz(y) Ry(p) Rx(r)); box ``dimension`` = full
extents, centred; sphere centred; **cylinder: axis +z, base at the frame origin**
(Amino convention, inferred from the reference's own frame names: the URDF
``right_upper_shoulder`` cylinder of length 0.2722 has origin z=0.1361, and the
reference calls its collision frame plain ``right_s0`` - i.e. the URDF importer's
-h/2 shift cancelled the origin exactly, which only happens for a base-at-origin
cylinder).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
SPHERE, BOX, CYLINDER, MESH = 0, 1, 2, 3


def quat_from_rpy(r: float, p: float, y: float) -> Tuple[float, float, float, float]:
    """URDF roll-pitch-yaw to quaternion (x,y,z,w)."""
    cr, sr = math.cos(r / 2), math.sin(r / 2)
    cp, sp = math.cos(p / 2), math.sin(p / 2)
    cy, sy = math.cos(y / 2), math.sin(y / 2)
    return (
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
        cr * cp * cy + sr * sp * sy,
    )


@dataclass
class Geom:
    shape: int
    dims: Tuple[float, float, float]  # sphere: (r,0,0); box: full extents; cylinder: (height, radius, 0)
    verts: Optional[np.ndarray] = None  # mesh: (nv,3) float64
    tris: Optional[np.ndarray] = None  # mesh: (nt,3) int32


@dataclass
class Frame:
    name: str
    parent: str  # "" = root
    type: int = FIXED
    quat: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)
    xyz: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    axis: Tuple[float, float, float] = (0.0, 0.0, 1.0)
    config: str = ""
    offset: float = 0.0
    limits: Optional[Tuple[float, float]] = None
    geoms: List[Geom] = field(default_factory=list)


@dataclass
class Scene:
    frames: List[Frame] = field(default_factory=list)
    allowed: List[Tuple[str, str]] = field(default_factory=list)

    def add(self, fr: Frame) -> Frame:
        self.frames.append(fr)
        return fr

    def frame(self, name: str) -> Frame:
        for f in self.frames:
            if f.name == name:
                return f
        raise KeyError(name)

    def extend(self, other: "Scene") -> "Scene":
        self.frames.extend(other.frames)
        self.allowed.extend(other.allowed)
        return self

    def config_names(self) -> List[str]:
        out: List[str] = []
        for f in self.frames:
            if f.type != FIXED and f.config not in out:
                out.append(f.config)
        return out


# --------------------------------------------------------------------------------------
# Procedural meshes (stand-ins for the absent torso / pedestal DAE files)
# --------------------------------------------------------------------------------------

def _grid_box_mesh(half: Sequence[float], centre: Sequence[float], n: int, taper: float = 1.0):
    """Closed box surface, each face tessellated n x n, optionally tapered in x/y with z."""
    hx, hy, hz = half
    verts: List[Tuple[float, float, float]] = []
    tris: List[Tuple[int, int, int]] = []
    index: Dict[Tuple[int, int, int], int] = {}

    def vid(i: int, j: int, k: int) -> int:
        key = (i, j, k)
        if key not in index:
            u, v, w = (2.0 * i / n - 1.0), (2.0 * j / n - 1.0), (2.0 * k / n - 1.0)
            s = 1.0 + (taper - 1.0) * (w + 1.0) / 2.0
            index[key] = len(verts)
            verts.append((centre[0] + u * hx * s, centre[1] + v * hy * s, centre[2] + w * hz))
        return index[key]

    def quad(a, b, c, d):
        tris.append((a, b, c))
        tris.append((a, c, d))

    for a in range(n):
        for b in range(n):
            quad(vid(a, b, 0), vid(a, b + 1, 0), vid(a + 1, b + 1, 0), vid(a + 1, b, 0))
            quad(vid(a, b, n), vid(a + 1, b, n), vid(a + 1, b + 1, n), vid(a, b + 1, n))
            quad(vid(a, 0, b), vid(a + 1, 0, b), vid(a + 1, 0, b + 1), vid(a, 0, b + 1))
            quad(vid(a, n, b), vid(a, n, b + 1), vid(a + 1, n, b + 1), vid(a + 1, n, b))
            quad(vid(0, a, b), vid(0, a, b + 1), vid(0, a + 1, b + 1), vid(0, a + 1, b))
            quad(vid(n, a, b), vid(n, a + 1, b), vid(n, a + 1, b + 1), vid(n, a, b + 1))
    return np.asarray(verts, dtype=np.float64), np.asarray(tris, dtype=np.int32)


def _tube_mesh(r0: float, r1: float, z0: float, z1: float, nseg: int, nring: int):
    """Closed surface of revolution about z (capped frustum)."""
    verts: List[Tuple[float, float, float]] = []
    tris: List[Tuple[int, int, int]] = []
    for k in range(nring + 1):
        t = k / nring
        r = r0 + (r1 - r0) * t
        z = z0 + (z1 - z0) * t
        for s in range(nseg):
            a = 2.0 * math.pi * s / nseg
            verts.append((r * math.cos(a), r * math.sin(a), z))
    for k in range(nring):
        for s in range(nseg):
            a = k * nseg + s
            b = k * nseg + (s + 1) % nseg
            c = (k + 1) * nseg + (s + 1) % nseg
            d = (k + 1) * nseg + s
            tris.append((a, b, c))
            tris.append((a, c, d))
    cb = len(verts)
    verts.append((0.0, 0.0, z0))
    ct = len(verts)
    verts.append((0.0, 0.0, z1))
    for s in range(nseg):
        tris.append((cb, (s + 1) % nseg, s))
        tris.append((ct, nring * nseg + s, nring * nseg + (s + 1) % nseg))
    return np.asarray(verts, dtype=np.float64), np.asarray(tris, dtype=np.int32)


# --------------------------------------------------------------------------------------
# Baxter (restated fixture)
# --------------------------------------------------------------------------------------

HALF_PI = math.pi / 2

# joint name suffix -> (xyz, rpy, (lo, hi)) - SURVEY.md Appendix A.6 (restated)
_ARM_CHAIN = [
    ("s0", (0.055695, 0.0, 0.011038), (0.0, 0.0, 0.0), (-1.70168, 1.70168)),
    ("s1", (0.069, 0.0, 0.27035), (-HALF_PI, 0.0, 0.0), (-2.147, 1.047)),
    ("e0", (0.102, 0.0, 0.0), (HALF_PI, 0.0, HALF_PI), (-3.05418, 3.05418)),
    ("e1", (0.069, 0.0, 0.26242), (-HALF_PI, -HALF_PI, 0.0), (-0.05, 2.618)),
    ("w0", (0.10359, 0.0, 0.0), (HALF_PI, 0.0, HALF_PI), (-3.059, 3.059)),
    ("w1", (0.01, 0.0, 0.2707), (-HALF_PI, -HALF_PI, 0.0), (-1.5708, 2.094)),
    ("w2", (0.115975, 0.0, 0.0), (HALF_PI, 0.0, HALF_PI), (-3.059, 3.059)),
]

# link collision cylinders: joint suffix -> (link name, length, radius, URDF origin z)
# The Amino-style importer shifts a URDF (centred) cylinder by -length/2, so the
# collision child frame sits at z = origin_z - length/2; a zero shift means the geometry
# hangs directly on the joint frame (this is why the reference names `right_s0`).
_ARM_CYLS = [
    ("s0", "upper_shoulder", 0.2722, 0.06, 0.1361),
    ("s1", "lower_shoulder", 0.12, 0.06, 0.0),
    ("e0", "upper_elbow", 0.107, 0.06, -0.0535),
    ("e1", "lower_elbow", 0.10, 0.06, 0.0),
    ("w0", "upper_forearm", 0.088, 0.06, -0.044),
    ("w1", "lower_forearm", 0.10, 0.06, 0.0),
    ("w2", "wrist", 0.165, 0.06, -0.0825),
]

RIGHT_ARM = ["right_s0", "right_s1", "right_e0", "right_e1", "right_w0", "right_w1", "right_w2"]
LEFT_ARM = ["left_s0", "left_s1", "left_e0", "left_e1", "left_w0", "left_w1", "left_w2"]


def arm_limits(side: str = "right") -> np.ndarray:
    return np.asarray([lim for (_, _, _, lim) in _ARM_CHAIN], dtype=np.float64)


def _add_arm(sc: Scene, side: str, sign: float) -> None:
    mount = f"{side}_torso_arm_mount"
    sc.add(Frame(mount, "torso_t0", FIXED, quat_from_rpy(0, 0, sign * 0.7854),
                 (0.024645, sign * 0.219645, 0.118588)))
    parent = mount
    for suffix, xyz, rpy, lim in _ARM_CHAIN:
        name = f"{side}_{suffix}"
        sc.add(Frame(name, parent, REVOLUTE, quat_from_rpy(*rpy), xyz, (0.0, 0.0, 1.0), name, 0.0, lim))
        parent = name
    for suffix, link, length, radius, oz in _ARM_CYLS:
        jf = f"{side}_{suffix}"
        shift = oz - length / 2
        g = Geom(CYLINDER, (length, radius, 0.0))
        if abs(shift) < 1e-12:
            sc.frame(jf).geoms.append(g)
        else:
            sc.add(Frame(f"{side}_{link}-collision", jf, FIXED, (0, 0, 0, 1), (0.0, 0.0, shift), geoms=[g]))
    # dummy 1 mm spheres on the *_visual links (frames named after their fixed joints)
    sc.add(Frame(f"{side}_e0_fixed", f"{side}_e0", FIXED, (0, 0, 0, 1), (0.107, 0.0, 0.0),
                 geoms=[Geom(SPHERE, (0.001, 0, 0))]))
    sc.add(Frame(f"{side}_w0_fixed", f"{side}_w0", FIXED, (0, 0, 0, 1), (0.088, 0.0, 0.0),
                 geoms=[Geom(SPHERE, (0.001, 0, 0))]))
    sc.add(Frame(f"{side}_hand", f"{side}_w2", FIXED, (0, 0, 0, 1), (0.0, 0.0, 0.11355)))
    sc.add(Frame(f"{side}_hand-collision", f"{side}_hand", FIXED, (0, 0, 0, 1), (0.0, 0.0, -0.0464),
                 geoms=[Geom(CYLINDER, (0.0464, 0.04, 0.0))]))
    sc.add(Frame(f"{side}_endpoint", f"{side}_hand", FIXED, (0, 0, 0, 1), (0.0, 0.0, 0.025)))


def baxter_fixture(mesh_res: int = 6) -> Scene:
    """Baxter, restated.  27 collision geometries: 10 per arm, torso + pedestal meshes,
    3 dummy head spheres, 2 head bounding spheres.  ``mesh_res`` scales the procedural
    torso / pedestal tessellation (6 -> 432 + 352 triangles)."""
    sc = Scene()
    tv, tt = _grid_box_mesh((0.13, 0.17, 0.33), (-0.03, 0.0, 0.22), mesh_res, taper=0.8)
    sc.add(Frame("torso_t0", "", FIXED, geoms=[Geom(MESH, (0, 0, 0), tv, tt)]))
    pv, pt = _tube_mesh(0.42, 0.16, -0.93, -0.12, 4 * mesh_res - 2, mesh_res)
    sc.add(Frame("pedestal_fixed", "torso_t0", FIXED, geoms=[Geom(MESH, (0, 0, 0), pv, pt)]))
    sc.add(Frame("head_pan", "torso_t0", REVOLUTE, (0, 0, 0, 1), (0.06, 0.0, 0.686), (0, 0, 1), "head_pan",
                 0.0, (-HALF_PI, HALF_PI), geoms=[Geom(SPHERE, (0.001, 0, 0))]))
    sc.add(Frame("head_nod", "head_pan", FIXED, (0, 0, 0, 1), (0.1227, 0.0, 0.0),
                 geoms=[Geom(SPHERE, (0.001, 0, 0))]))
    sc.add(Frame("sonar_s0", "torso_t0", FIXED, (0, 0, 0, 1), (0.0947, 0.0, 0.817),
                 geoms=[Geom(SPHERE, (0.001, 0, 0))]))
    for k, y in ((1, -0.04), (2, 0.04)):
        sc.add(Frame(f"collision_head_{k}", "torso_t0", FIXED, (0, 0, 0, 1), (0.11, 0.0, 0.75)))
        sc.add(Frame(f"collision_head_link_{k}-collision", f"collision_head_{k}", FIXED, (0, 0, 0, 1),
                     (-0.07, y, 0.0), geoms=[Geom(SPHERE, (0.22, 0, 0))]))
    _add_arm(sc, "right", -1.0)
    _add_arm(sc, "left", +1.0)
    return sc


# reference: demo/baxter-sussman/allowed-collision.robray:1-31
ALLOWED_BAXTER = [
    ("collision_head_link_1-collision", "collision_head_link_2-collision"),
    ("collision_head_link_1-collision", "head_nod"),
    ("collision_head_link_1-collision", "head_pan"),
    ("collision_head_link_1-collision", "sonar_s0"),
    ("collision_head_link_1-collision", "torso_t0"),
    ("collision_head_link_2-collision", "head_nod"),
    ("collision_head_link_2-collision", "head_pan"),
    ("collision_head_link_2-collision", "sonar_s0"),
    ("collision_head_link_2-collision", "torso_t0"),
    ("left_e0_fixed", "left_lower_elbow-collision"),
    ("left_hand-collision", "left_wrist-collision"),
    ("left_lower_elbow-collision", "left_upper_forearm-collision"),
    ("left_lower_forearm-collision", "left_w0_fixed"),
    ("left_lower_forearm-collision", "left_wrist-collision"),
    ("left_lower_shoulder-collision", "left_s0"),
    ("left_lower_shoulder-collision", "left_upper_elbow-collision"),
    ("left_s0", "torso_t0"),
    ("left_upper_forearm-collision", "left_w0_fixed"),
    ("pedestal_fixed", "torso_t0"),
    ("right_e0_fixed", "right_lower_elbow-collision"),
    ("right_e0_fixed", "right_upper_forearm-collision"),
    ("right_hand-collision", "right_wrist-collision"),
    ("right_lower_elbow-collision", "right_upper_forearm-collision"),
    ("right_lower_forearm-collision", "right_w0_fixed"),
    ("right_lower_forearm-collision", "right_wrist-collision"),
    ("right_lower_shoulder-collision", "right_s0"),
    ("right_lower_shoulder-collision", "right_upper_elbow-collision"),
    ("right_s0", "right_upper_elbow-collision"),
    ("right_s0", "torso_t0"),
    ("right_upper_forearm-collision", "right_w0_fixed"),
    ("right_w0_fixed", "right_wrist-collision"),
]

# reference: demo/baxter-sussman/q0.tmp:1-2 (map by NAME; the file's order is not chain order)
Q0 = {
    "right_w2": 0.0, "right_w1": 1.5707963267948966, "right_w0": 0.0,
    "right_s1": -0.7853981633974483, "right_s0": 0.1570796350201586,
    "right_e1": 0.7853981633974483, "right_e0": 0.0,
    "left_w2": 0.0, "left_w1": 0.7853981633974483, "left_w0": 0.0,
    "left_s1": -1.5707963267948966, "left_s0": 2.04203514993196,
    "left_e1": 2.356194490192345, "left_e0": 0.0,
}

# --------------------------------------------------------------------------------------
# Sussman scene (reference demo/baxter-sussman/*.robray)
# --------------------------------------------------------------------------------------
CM, MM = 1e-2, 1e-3
BLOCK_SIZE = 10 * CM  # class.robray:16
TABLE_SIZE = 1 * CM  # class.robray:17
TABLE_STACK = (BLOCK_SIZE + TABLE_SIZE) / 2 + 0.1 * MM  # class.robray:54
BLOCK_STACK = BLOCK_SIZE + 0.1 * MM  # class.robray:55
TABLE_HEIGHT = 0.05  # table.robray:5


def _box(dx, dy, dz) -> Geom:
    return Geom(BOX, (dx, dy, dz))


def tables_scene() -> Scene:
    """table.robray:7-81."""
    sc = Scene()
    sc.add(Frame("table_base", "", FIXED, quat_from_rpy(0, 0, math.radians(-45)), (0.1, -0.3, TABLE_HEIGHT)))
    sc.add(Frame("front_table", "table_base", FIXED, (0, 0, 0, 1), (0.6, 0, 0), geoms=[_box(0.75, 0.75, TABLE_SIZE)]))
    sc.add(Frame("curtain", "", FIXED, quat_from_rpy(0, 0, 0), (-0.35, 0, 0), geoms=[_box(0.01, 5, 5)]))
    sc.add(Frame("bookshelf", "", FIXED, quat_from_rpy(0, 0, 0), (0, -1.15, 0), geoms=[_box(1, 0.01, 5)]))
    depth, length1, thick = 0.5, 1.5, 0.01
    sc.add(Frame("lab_front_table", "", FIXED, (0, 0, 0, 1), (0.8, -0.2, TABLE_HEIGHT), geoms=[_box(depth, length1, thick)]))
    sc.add(Frame("left_table", "lab_front_table", FIXED, quat_from_rpy(0, 0, HALF_PI),
                 (-length1 / 2 + depth / 2, length1 / 2 + depth / 2, 0), geoms=[_box(depth, length1, thick)]))
    sc.add(Frame("right_table", "lab_front_table", FIXED, quat_from_rpy(0, 0, -HALF_PI),
                 (-length1 / 2 - depth / 2, -length1 / 2 + depth / 2, 0), geoms=[_box(depth, length1, thick)]))
    return sc


def sussman_scene(which: int = 0) -> Scene:
    """sussman-0.robray (start) / sussman-1.robray (goal) + class.robray grasp frame."""
    sc = tables_scene()
    blk = lambda: _box(BLOCK_SIZE, BLOCK_SIZE, BLOCK_SIZE)
    if which == 0:
        sc.add(Frame("block_a", "front_table", FIXED, (0, 0, 0, 1), (0, 0, TABLE_STACK), geoms=[blk()]))
        sc.add(Frame("block_b", "front_table", FIXED, (0, 0, 0, 1), (0, -0.25, TABLE_STACK), geoms=[blk()]))
        sc.add(Frame("block_c", "block_a", FIXED, (0, 0, 0, 1), (0, 0, BLOCK_STACK), geoms=[blk()]))
    else:
        raise ValueError("goal scene sussman-1 only names relations; it has no root placement")
    return sc


def baxter_sussman(mesh_res: int = 6) -> Scene:
    """Config C1/C2 scene: Baxter fixture + Sussman start scene + the 31 allowed pairs
    + the grasp frame (class.robray:61-65)."""
    sc = baxter_fixture(mesh_res)
    sc.extend(sussman_scene(0))
    sc.add(Frame("end_effector_grasp", "right_endpoint", FIXED, (0, 1, 0, 0), (0, 0, 0.075)))
    sc.allowed = list(ALLOWED_BAXTER)
    return sc


def q0_vector(sc: Scene) -> np.ndarray:
    names = sc.config_names()
    return np.asarray([Q0.get(n, 0.0) for n in names], dtype=np.float64)


# --------------------------------------------------------------------------------------
# Scaling scenes
# --------------------------------------------------------------------------------------

def blocksworld_scene(n_blocks: int, seed: int = 0, mesh_res: int = 6, grasped: bool = True) -> Scene:
    """C4: front table + n 0.1 m cubes on the 0.1 m grid (tm-blocks.py:6 RESOLUTION,
    genscene.lisp:4-30 draws positions without replacement); stacked when the grid is
    exhausted.  With ``grasped`` the first block hangs on the right grasp frame, so
    moving-box vs static-box and moving-box vs mesh pairs are exercised."""
    rng = np.random.default_rng(seed)
    sc = baxter_fixture(mesh_res)
    sc.extend(tables_scene())
    sc.add(Frame("end_effector_grasp", "right_endpoint", FIXED, (0, 1, 0, 0), (0, 0, 0.075)))
    cells = [(i, j) for i in range(-3, 4) for j in range(-3, 4)]
    order = rng.permutation(len(cells))
    height: Dict[Tuple[int, int], str] = {}
    for b in range(n_blocks):
        name = f"block_{b}"
        if grasped and b == 0:
            sc.add(Frame(name, "end_effector_grasp", FIXED, (0, 0, 0, 1), (0, 0, 0), geoms=[_box(0.1, 0.1, 0.1)]))
            continue
        cell = cells[order[b % len(cells)]]
        if cell in height:
            sc.add(Frame(name, height[cell], FIXED, (0, 0, 0, 1), (0, 0, BLOCK_STACK), geoms=[_box(0.1, 0.1, 0.1)]))
        else:
            sc.add(Frame(name, "front_table", FIXED, (0, 0, 0, 1), (0.1 * cell[0], 0.1 * cell[1], TABLE_STACK),
                         geoms=[_box(0.1, 0.1, 0.1)]))
        height[cell] = name
    sc.allowed = list(ALLOWED_BAXTER)
    return sc


def _rand_quat(rng) -> Tuple[float, float, float, float]:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(x) for x in q)


def _q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _q_rot(q, v):
    u = np.asarray(q[:3], dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    t = 2.0 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


def world_poses(sc: Scene, q: Dict[str, float]) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """World pose (quaternion xyzw, translation) of every frame at the configuration ``q`` (variables not
    listed are 0).  Scene-authoring aid only (the rejection test of ``clutter_scene``): the product's FK is
    aa_rx_sg_tf; the CPU checker has its own."""
    out: Dict[str, Tuple[np.ndarray, np.ndarray]] = {"": (np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3))}
    pending = list(sc.frames)
    while pending:
        rest = []
        for f in pending:
            if f.parent not in out:
                rest.append(f)
                continue
            r, v = np.asarray(f.quat, dtype=np.float64), np.asarray(f.xyz, dtype=np.float64)
            if f.type == REVOLUTE:
                ang = q.get(f.config, 0.0) + f.offset
                ax = np.asarray(f.axis, dtype=np.float64)
                r = _q_mul(r, np.concatenate([ax * math.sin(ang / 2), [math.cos(ang / 2)]]))
            elif f.type == PRISMATIC:
                v = v + _q_rot(r, np.asarray(f.axis, dtype=np.float64) * (q.get(f.config, 0.0) + f.offset))
            pr, pv = out[f.parent]
            out[f.name] = (_q_mul(pr, r), _q_rot(pr, v) + pv)
        if len(rest) == len(pending):
            raise ValueError("cycle or unknown parent in scene graph")
        pending = rest
    return out


def _bounds_at(sc: Scene, q: Dict[str, float]):
    """Conservative world-frame bounds of every geometry at q: spheres, capsules (cylinders) and oriented
    boxes (boxes; meshes by the box of their vertices)."""
    tf = world_poses(sc, q)
    out = []
    for f in sc.frames:
        r, v = tf[f.name]
        for g in f.geoms:
            if g.shape == SPHERE:
                out.append(("sphere", v, float(g.dims[0])))
            elif g.shape == CYLINDER:  # base at the frame origin, axis +z
                out.append(("capsule", v, v + _q_rot(r, (0.0, 0.0, float(g.dims[0]))), float(g.dims[1])))
            elif g.shape == BOX:
                out.append(("obb", r, v, 0.5 * np.asarray(g.dims, dtype=np.float64)))
            else:
                vv = np.asarray(g.verts, dtype=np.float64)
                lo, hi = vv.min(axis=0), vv.max(axis=0)
                out.append(("obb", r, v + _q_rot(r, 0.5 * (lo + hi)), 0.5 * (hi - lo)))
    return out


def _ball_touches(bounds, c, rho: float) -> bool:
    c = np.asarray(c, dtype=np.float64)
    for b in bounds:
        if b[0] == "sphere":
            d = np.linalg.norm(c - b[1]) - b[2]
        elif b[0] == "capsule":
            p0, p1 = b[1], b[2]
            e = p1 - p0
            t = min(1.0, max(0.0, float(np.dot(c - p0, e) / max(np.dot(e, e), 1e-30))))
            d = np.linalg.norm(c - (p0 + t * e)) - b[3]
        else:
            r, centre, half = b[1], b[2], b[3]
            local = _q_rot(np.array([-r[0], -r[1], -r[2], r[3]]), c - centre)
            d = np.linalg.norm(np.maximum(np.abs(local) - half, 0.0))
        if d <= rho:
            return True
    return False


def clutter_scene(n_prims: int = 512, seed: int = 0, mesh_res: int = 6, reject_at_q0: bool = True) -> Scene:
    """C5: Baxter + n random primitives: 50 % boxes (extents U(0.02,0.3)^3), 30 %
    cylinders (r U(0.01,0.1), h U(0.05,0.4)), 20 % spheres (r U(0.02,0.15)); centres
    U([-1.5,1.5]^2 x [-0.5,1.5]); uniform orientations; candidates that intersect the robot at the start
    configuration ``Q0`` are rejected and redrawn.  The rejection test is conservative and device-free: the
    candidate's bounding ball against capsule / oriented-box / sphere bounds of the robot geometry at Q0 (it
    rejects every intersecting candidate and a few near misses).  ``reject_at_q0=False`` gives the round-1 scene,
    where such obstacles stay and are handled by the start-state allowance (kept as a second test: 98 % of its
    configurations are invalid)."""
    rng = np.random.default_rng(seed)
    sc = baxter_fixture(mesh_res)
    bounds = _bounds_at(sc, Q0) if reject_at_q0 else None
    i = 0
    while i < n_prims:
        u = rng.random()
        c = (float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-0.5, 1.5)))
        if u < 0.5:
            g = Geom(BOX, tuple(float(x) for x in rng.uniform(0.02, 0.3, size=3)))
            centre_local, rho = (0.0, 0.0, 0.0), 0.5 * float(np.linalg.norm(g.dims))
        elif u < 0.8:
            g = Geom(CYLINDER, (float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.01, 0.1)), 0.0))
            centre_local, rho = (0.0, 0.0, 0.5 * g.dims[0]), math.hypot(0.5 * g.dims[0], g.dims[1])
        else:
            g = Geom(SPHERE, (float(rng.uniform(0.02, 0.15)), 0.0, 0.0))
            centre_local, rho = (0.0, 0.0, 0.0), g.dims[0]
        quat = _rand_quat(rng)
        if bounds is not None and _ball_touches(bounds, np.asarray(c) + _q_rot(np.asarray(quat), centre_local), rho + 1e-6):
            continue
        sc.add(Frame(f"obstacle_{i}", "", FIXED, quat, c, geoms=[g]))
        i += 1
    sc.allowed = list(ALLOWED_BAXTER)
    return sc


# --------------------------------------------------------------------------------------
# Flat arrays (oracle input; also what tests use to cross-check the product's export)
# --------------------------------------------------------------------------------------

def flatten(sc: Scene) -> Dict[str, np.ndarray]:
    """Topologically sort frames (stable: parent before child, otherwise insertion order -
    the same rule the product's aa_rx_sg_init uses) and emit flat arrays."""
    by_name = {f.name: f for f in sc.frames}
    if len(by_name) != len(sc.frames):
        raise ValueError("duplicate frame name")
    order: List[Frame] = []
    placed: Dict[str, int] = {}
    pending = list(sc.frames)
    while pending:
        rest = []
        for f in pending:
            if f.parent == "" or f.parent in placed:
                placed[f.name] = len(order)
                order.append(f)
            else:
                if f.parent not in by_name:
                    raise KeyError(f"frame {f.name}: unknown parent {f.parent}")
                rest.append(f)
        if len(rest) == len(pending):
            raise ValueError("cycle in scene graph")
        pending = rest
    cfg = []
    for f in order:
        if f.type != FIXED and f.config not in cfg:
            cfg.append(f.config)
    n_f = len(order)
    out = {
        "names": [f.name for f in order],
        "config_names": cfg,
        "parent": np.asarray([placed[f.parent] if f.parent else -1 for f in order], dtype=np.int32),
        "type": np.asarray([f.type for f in order], dtype=np.int32),
        "E": np.asarray([list(f.quat) + list(f.xyz) for f in order], dtype=np.float64).reshape(n_f, 7),
        "axis": np.asarray([f.axis for f in order], dtype=np.float64).reshape(n_f, 3),
        "offset": np.asarray([f.offset for f in order], dtype=np.float64),
        "qidx": np.asarray([cfg.index(f.config) if f.type != FIXED else -1 for f in order], dtype=np.int32),
    }
    lim = np.zeros((len(cfg), 2))
    lim[:, 0], lim[:, 1] = -math.pi, math.pi
    for f in order:
        if f.type != FIXED and f.limits is not None:
            lim[cfg.index(f.config)] = f.limits
    out["limits"] = lim
    g_frame, g_shape, g_dim, g_mesh = [], [], [], []
    verts, tris, vstart, tstart = [], [], [0], [0]
    for fi, f in enumerate(order):
        for g in f.geoms:
            g_frame.append(fi)
            g_shape.append(g.shape)
            g_dim.append(g.dims)
            if g.shape == MESH:
                g_mesh.append(len(vstart) - 1)
                verts.append(np.asarray(g.verts, dtype=np.float64))
                tris.append(np.asarray(g.tris, dtype=np.int32))
                vstart.append(vstart[-1] + len(g.verts))
                tstart.append(tstart[-1] + len(g.tris))
            else:
                g_mesh.append(-1)
    out["g_frame"] = np.asarray(g_frame, dtype=np.int32)
    out["g_shape"] = np.asarray(g_shape, dtype=np.int32)
    out["g_dim"] = np.asarray(g_dim, dtype=np.float64).reshape(len(g_frame), 3)
    out["g_mesh"] = np.asarray(g_mesh, dtype=np.int32)
    out["mesh_vstart"] = np.asarray(vstart, dtype=np.int32)
    out["mesh_tstart"] = np.asarray(tstart, dtype=np.int32)
    out["verts"] = np.concatenate(verts) if verts else np.zeros((0, 3))
    out["tris"] = np.concatenate(tris) if tris else np.zeros((0, 3), dtype=np.int32)
    allowed = np.zeros((n_f, n_f), dtype=np.uint8)
    for a, b in sc.allowed:
        i, j = placed[a], placed[b]
        allowed[i, j] = allowed[j, i] = 1
    out["allowed"] = allowed
    return out


# --------------------------------------------------------------------------------------
# Scene plugin source (what the reference's `aarxc` emits: Makefile.am:213-231)
# --------------------------------------------------------------------------------------

def to_plugin_source(sc: Scene, scene_name: str) -> str:
    """C source of a scene plugin: one constructor ``aa_rx_dl_sg__<scene_name>`` that feeds the scene
    through the ingest ABI (aa_rx_sg_add_frame_*, aa_rx_geom_*, aa_rx_geom_attach,
    aa_rx_sg_allow_collision_name).  Compile with ``gcc -shared -fPIC -I include`` and load it with
    aa_rx_dl_sg (demo/tmpexec.c:172)."""
    def arr(xs, fmt="%.17g", suffix=""):
        def lit(x):
            t = fmt % float(x)
            if "." not in t and "e" not in t and "n" not in t:
                t += ".0"  # "0f" is not a C literal
            return t + suffix
        return "{" + ", ".join(lit(x) for x in xs) + "}"

    out = ['#include "tmkit_b200/amino_rx.h"', ""]
    mesh_id = 0
    body = []
    for f in sc.frames:
        body.append("    {")
        body.append(f"        static const double q[4] = {arr(f.quat)}, v[3] = {arr(f.xyz)}, ax[3] = {arr(f.axis)};")
        if f.type == FIXED:
            body.append(f'        aa_rx_sg_add_frame_fixed(sg, "{f.parent}", "{f.name}", q, v);')
            body.append("        (void)ax;")
        else:
            fn = "aa_rx_sg_add_frame_revolute" if f.type == REVOLUTE else "aa_rx_sg_add_frame_prismatic"
            body.append(f'        {fn}(sg, "{f.parent}", "{f.name}", q, v, "{f.config}", ax, {float(f.offset):.17g});')
            if f.limits is not None:
                body.append(f'        aa_rx_sg_set_limit_pos(sg, "{f.config}", {float(f.limits[0]):.17g}, {float(f.limits[1]):.17g});')
        for g in f.geoms:
            if g.shape == BOX:
                body.append(f"        {{ static const double d[3] = {arr(g.dims)}; aa_rx_geom_attach(sg, \"{f.name}\", aa_rx_geom_box(0, d)); }}")
            elif g.shape == SPHERE:
                body.append(f'        aa_rx_geom_attach(sg, "{f.name}", aa_rx_geom_sphere(0, {float(g.dims[0]):.17g}));')
            elif g.shape == CYLINDER:
                body.append(f'        aa_rx_geom_attach(sg, "{f.name}", aa_rx_geom_cylinder(0, {float(g.dims[0]):.17g}, {float(g.dims[1]):.17g}));')
            else:
                vv = np.asarray(g.verts, dtype=np.float32).reshape(-1)
                ii = np.asarray(g.tris, dtype=np.uint32).reshape(-1)
                out.append(f"static const float mesh{mesh_id}_v[] = {arr(vv, '%.9g', 'f')};")
                out.append(f"static const unsigned mesh{mesh_id}_i[] = {{{', '.join(str(int(x)) for x in ii)}}};")
                body.append("        {")
                body.append("            struct aa_rx_mesh *m = aa_rx_mesh_create();")
                body.append(f"            aa_rx_mesh_set_vertices(m, {len(vv) // 3}, mesh{mesh_id}_v, 0);")
                body.append(f"            aa_rx_mesh_set_indices(m, {len(ii) // 3}, mesh{mesh_id}_i, 0);")
                body.append(f'            aa_rx_geom_attach(sg, "{f.name}", aa_rx_geom_mesh(0, m));')
                body.append("            aa_rx_mesh_destroy(m);")
                body.append("        }")
                mesh_id += 1
        body.append("    }")
    for a, b in sc.allowed:
        body.append(f'    aa_rx_sg_allow_collision_name(sg, "{a}", "{b}", 1);')
    out.append("")
    out.append(f"struct aa_rx_sg *aa_rx_dl_sg__{scene_name}(struct aa_rx_sg *sg) {{")
    out.extend(body)
    out.append("    return sg;")
    out.append("}")
    return "\n".join(out) + "\n"
