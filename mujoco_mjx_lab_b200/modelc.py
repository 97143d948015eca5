"""MJCF -> constants compiler for the humanoid model family (numpy only).

The reference gets these numbers from the MuJoCo C compiler
(`mujoco.MjModel.from_xml_path`, reference src/training_utils.py:80) followed by
`mjx.put_model` (reference src/training_utils.py:105).  Neither `mujoco` nor `mujoco.mjx`
is installed here, so this module restates the part of the MJCF compiler that
`models/humanoid_mjx.xml` / `models/humanoid.xml` exercise:

  * nested <default class> inheritance and `childclass`,
  * capsule `fromto` -> (pos, quat, half-length), angles in degrees,
  * geom mass / inertia at density 1000 and the per-body inertial frame,
  * joint / dof / actuator / fixed-tendon / site / touch-sensor tables,
  * the contype/conaffinity + weld-parent + <exclude> collision filter and the static
    contact-slot / constraint-row layout MJX uses,
  * geom-pair contact parameter mixing (friction max, solref/solimp solmix average),
  * the `mj_setConst` quantities evaluated at qpos0: dof_invweight0, body_invweight0,
    tendon_invweight0, stat.meaninertia, body_subtreemass.

Everything is computed in float64 and cast to float32 when packed (mjx.put_model does the same).
Output: a plain dict of numpy arrays (`compile_mjcf`) and a packed POD blob (`pack_blob`) whose layout
is `struct mjxb_model_blob` in csrc/mjxb_model.h.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from typing import Any, Dict, List, Optional

import numpy as np

# geom type enum follows MuJoCo (mjtGeom): plane 0, sphere 2, capsule 3, box 6
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX}
JNT_FREE, JNT_HINGE = 0, 3
# pair kinds used by the collision kernels (ordered by (type1,type2))
PAIR_PLANE_SPHERE, PAIR_PLANE_CAPSULE, PAIR_SPHERE_SPHERE, PAIR_SPHERE_CAPSULE, PAIR_CAPSULE_CAPSULE = 0, 1, 2, 3, 4
_PAIR_KIND = {
    (GEOM_PLANE, GEOM_SPHERE): PAIR_PLANE_SPHERE,
    (GEOM_PLANE, GEOM_CAPSULE): PAIR_PLANE_CAPSULE,
    (GEOM_SPHERE, GEOM_SPHERE): PAIR_SPHERE_SPHERE,
    (GEOM_SPHERE, GEOM_CAPSULE): PAIR_SPHERE_CAPSULE,
    (GEOM_CAPSULE, GEOM_CAPSULE): PAIR_CAPSULE_CAPSULE,
}
_PAIR_NCON = {PAIR_PLANE_SPHERE: 1, PAIR_PLANE_CAPSULE: 2, PAIR_SPHERE_SPHERE: 1, PAIR_SPHERE_CAPSULE: 1,
              PAIR_CAPSULE_CAPSULE: 1}
SOLVER_CG, SOLVER_NEWTON = 1, 2
INT_EULER, INT_IMPLICITFAST = 0, 3
MJ_MINVAL = 1e-15


# ----------------------------------------------------------------------------- small math (f64)
def _vec(s: Optional[str], n: Optional[int] = None, default=None) -> np.ndarray:
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size < n and default is not None:
        out = np.array(default, dtype=np.float64)
        out[: v.size] = v
        return out
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def z_to_quat(vec):
    """Quaternion rotating +z onto `vec` (MuJoCo user-side z2quat semantics)."""
    v = np.asarray(vec, dtype=np.float64)
    v = v / np.linalg.norm(v)
    axis = np.cross([0.0, 0.0, 1.0], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = math.atan2(s, v[2])
    return np.concatenate([[math.cos(ang / 2)], axis * math.sin(ang / 2)])


def axis_angle_quat(axis, angle):
    return np.concatenate([[math.cos(angle / 2)], np.asarray(axis) * math.sin(angle / 2)])


# ----------------------------------------------------------------------------- defaults handling
_BUILTIN = {
    "geom": dict(type="sphere", contype="1", conaffinity="1", condim="3", friction="1 0.005 0.0001",
                 solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", density="1000", margin="0", gap="0",
                 solmix="1", priority="0"),
    "joint": dict(type="hinge", pos="0 0 0", axis="0 0 1", stiffness="0", damping="0", armature="0",
                  springref="0", ref="0", solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
                  margin="0"),
    "site": dict(type="sphere", pos="0 0 0", size="0.005 0.005 0.005"),
    "motor": dict(gear="1", ctrllimited="auto"),
    "tendon": dict(solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2", margin="0"),
}


class _Defaults:
    """Nested <default class=...> tree: class name -> {tag -> attrib dict} with parent inheritance."""

    def __init__(self, root: Optional[ET.Element]):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {}
        base = {k: dict(v) for k, v in _BUILTIN.items()}
        self.classes["main"] = base
        if root is not None:
            self._walk(root, "main", top=True)

    def _walk(self, node: ET.Element, parent: str, top: bool = False):
        name = "main" if top else node.get("class")
        if not top:
            self.classes[name] = {k: dict(v) for k, v in self.classes[parent].items()}
        cur = self.classes[name]
        for child in node:
            if child.tag == "default":
                continue
            cur.setdefault(child.tag, {})
            cur[child.tag].update(child.attrib)
        for child in node:
            if child.tag == "default":
                self._walk(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.get("class") or childclass or "main"
        out = dict(self.classes[cls].get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


# ----------------------------------------------------------------------------- geom mass properties
def _geom_mass_inertia(gtype: int, size: np.ndarray, density: float):
    """Mass and diagonal inertia (about the geom centre, geom axes). SURVEY.md A.6."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        m = 4.0 / 3.0 * math.pi * r ** 3 * density
        i = 0.4 * m * r * r
        return m, np.array([i, i, i])
    if gtype == GEOM_CAPSULE:
        r, h = size[0], 2.0 * size[1]
        m_cyl = math.pi * r * r * h * density
        m_sph = 4.0 / 3.0 * math.pi * r ** 3 * density
        m = m_cyl + m_sph
        ixx = m_cyl * (3 * r * r + h * h) / 12.0 + m_sph * (0.4 * r * r + 0.375 * r * h + 0.25 * h * h)
        izz = m_cyl * r * r / 2.0 + m_sph * 0.4 * r * r
        return m, np.array([ixx, ixx, izz])
    if gtype == GEOM_PLANE:
        return 0.0, np.zeros(3)
    raise NotImplementedError(f"geom type {gtype}")


# ----------------------------------------------------------------------------- compiler
def compile_mjcf(xml_path: str, solver_overrides: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    root = ET.parse(xml_path).getroot()
    comp = root.find("compiler")
    angle_deg = True if comp is None else comp.get("angle", "degree") == "degree"
    ang = math.pi / 180.0 if angle_deg else 1.0

    defaults = _Defaults(root.find("default"))

    # ---- options
    opt = dict(timestep=0.002, solver=SOLVER_NEWTON, integrator=INT_EULER, iterations=100, ls_iterations=50,
               tolerance=1e-8, ls_tolerance=0.01, impratio=1.0, gravity=np.array([0.0, 0.0, -9.81]),
               cone=0, eulerdamp=1, jacobian="auto")
    o = root.find("option")
    if o is not None:
        if o.get("timestep"): opt["timestep"] = float(o.get("timestep"))
        if o.get("solver"): opt["solver"] = {"CG": SOLVER_CG, "Newton": SOLVER_NEWTON}[o.get("solver")]
        if o.get("integrator"):
            opt["integrator"] = {"Euler": INT_EULER, "implicitfast": INT_IMPLICITFAST}[o.get("integrator")]
        for k in ("iterations", "ls_iterations"):
            if o.get(k): opt[k] = int(o.get(k))
        for k in ("tolerance", "ls_tolerance", "impratio"):
            if o.get(k): opt[k] = float(o.get(k))
        if o.get("gravity"): opt["gravity"] = _vec(o.get("gravity"))
        if o.get("cone"): opt["cone"] = {"pyramidal": 0, "elliptic": 1}[o.get("cone")]
        if o.get("jacobian"): opt["jacobian"] = o.get("jacobian")
        fl = o.find("flag")
        if fl is not None and fl.get("eulerdamp") == "disable":
            opt["eulerdamp"] = 0
    if solver_overrides:
        opt.update(solver_overrides)
    if opt["cone"] != 0:
        raise NotImplementedError("only the pyramidal cone is supported")

    bodies: List[dict] = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                               jnts=[], geoms=[], sites=[])]
    joints: List[dict] = []
    geoms: List[dict] = []
    sites: List[dict] = []

    def add_geom(e: ET.Element, bid: int, childclass: Optional[str]):
        a = defaults.resolve("geom", e, childclass)
        gtype = _GEOM_TYPES[a.get("type", "sphere")]
        size = _vec(a.get("size"), 3, [0.0, 0.0, 0.0])
        if size is None:
            size = np.zeros(3)
        pos = _vec(a.get("pos"), 3, [0, 0, 0]) if a.get("pos") else np.zeros(3)
        quat = _vec(a.get("quat")) if a.get("quat") else np.array([1.0, 0, 0, 0])
        if a.get("fromto") is not None:
            ft = _vec(a.get("fromto"))
            p_from, p_to = ft[:3], ft[3:]
            pos = 0.5 * (p_from + p_to)
            v = p_from - p_to  # MuJoCo: z axis points from "to" towards "from"
            r = size[0]
            size = np.array([r, 0.5 * np.linalg.norm(v), 0.0])
            quat = z_to_quat(v)
        elif a.get("zaxis") is not None:
            quat = z_to_quat(_vec(a.get("zaxis")))
        quat = quat / np.linalg.norm(quat)
        fr = _vec(a.get("friction"), 3, [1.0, 0.005, 0.0001])
        g = dict(name=a.get("name", f"geom{len(geoms)}"), type=gtype, body=bid, size=size, pos=pos, quat=quat,
                 contype=int(a["contype"]), conaffinity=int(a["conaffinity"]), condim=int(a["condim"]),
                 friction=fr, solref=_vec(a["solref"]), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                 density=float(a["density"]), margin=float(a["margin"]), gap=float(a["gap"]),
                 solmix=float(a["solmix"]), priority=int(a["priority"]))
        if g["margin"] != 0.0 or g["gap"] != 0.0:
            raise NotImplementedError("geom margin/gap")
        geoms.append(g)
        bodies[bid]["geoms"].append(len(geoms) - 1)

    def add_site(e: ET.Element, bid: int, childclass: Optional[str]):
        a = defaults.resolve("site", e, childclass)
        s = dict(name=a.get("name"), body=bid, pos=_vec(a.get("pos"), 3, [0, 0, 0]),
                 quat=_vec(a.get("quat")) if a.get("quat") else np.array([1.0, 0, 0, 0]),
                 size=_vec(a.get("size"), 3, [0.005, 0.005, 0.005]), type=_GEOM_TYPES.get(a.get("type"), GEOM_SPHERE))
        sites.append(s)
        bodies[bid]["sites"].append(len(sites) - 1)

    def add_joint(e: ET.Element, bid: int, childclass: Optional[str], free: bool):
        if free:
            j = dict(name=e.get("name", "free"), type=JNT_FREE, body=bid, pos=np.zeros(3), axis=np.array([0, 0, 1.0]),
                     limited=0, range=np.zeros(2), stiffness=0.0, damping=0.0, armature=0.0, springref=0.0, ref=0.0,
                     solref=np.array([0.02, 1.0]), solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]), margin=0.0)
        else:
            a = defaults.resolve("joint", e, childclass)
            jt = a.get("type", "hinge")
            if jt not in ("hinge", "free"):
                raise NotImplementedError(f"joint type {jt}")
            axis = _vec(a["axis"])
            axis = axis / np.linalg.norm(axis)
            rng = _vec(a.get("range")) * ang if a.get("range") else np.zeros(2)
            lim = a.get("limited", "auto")
            limited = 1 if (lim == "true" or (lim == "auto" and a.get("range"))) else 0
            j = dict(name=a.get("name"), type=JNT_HINGE if jt == "hinge" else JNT_FREE, body=bid, pos=_vec(a["pos"]),
                     axis=axis, limited=limited, range=rng, stiffness=float(a["stiffness"]), damping=float(a["damping"]),
                     armature=float(a["armature"]), springref=float(a["springref"]) * ang, ref=float(a["ref"]) * ang,
                     solref=_vec(a["solreflimit"]), solimp=_vec(a["solimplimit"], 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                     margin=float(a["margin"]))
            if j["margin"] != 0.0:
                raise NotImplementedError("joint margin")
        joints.append(j)
        bodies[bid]["jnts"].append(len(joints) - 1)

    def walk_body(e: ET.Element, parent: int, childclass: Optional[str]):
        cc = e.get("childclass") or childclass
        pos = _vec(e.get("pos"), 3, [0, 0, 0]) if e.get("pos") else np.zeros(3)
        quat = _vec(e.get("quat")) if e.get("quat") else np.array([1.0, 0, 0, 0])
        bodies.append(dict(name=e.get("name"), parent=parent, pos=pos, quat=quat / np.linalg.norm(quat),
                           jnts=[], geoms=[], sites=[]))
        bid = len(bodies) - 1
        for c in e:
            if c.tag == "freejoint":
                add_joint(c, bid, cc, free=True)
            elif c.tag == "joint":
                add_joint(c, bid, cc, free=False)
            elif c.tag == "geom":
                add_geom(c, bid, cc)
            elif c.tag == "site":
                add_site(c, bid, cc)
            elif c.tag == "inertial":
                raise NotImplementedError("<inertial>")
        for c in e:
            if c.tag == "body":
                walk_body(c, bid, cc)

    wb = root.find("worldbody")
    for c in wb:
        if c.tag == "geom":
            add_geom(c, 0, None)
        elif c.tag == "site":
            add_site(c, 0, None)
    for c in wb:
        if c.tag == "body":
            walk_body(c, 0, None)

    nbody, njnt, ngeom, nsite = len(bodies), len(joints), len(geoms), len(sites)

    # ---- qpos / dof addressing
    nq = nv = 0
    for j in joints:
        j["qposadr"], j["dofadr"] = nq, nv
        if j["type"] == JNT_FREE:
            nq, nv = nq + 7, nv + 6
        else:
            nq, nv = nq + 1, nv + 1
    qpos0 = np.zeros(nq)
    qpos_spring = np.zeros(nq)
    dof_body = np.zeros(nv, np.int32)
    dof_jnt = np.zeros(nv, np.int32)
    dof_parent = -np.ones(nv, np.int32)
    dof_armature = np.zeros(nv)
    dof_damping = np.zeros(nv)
    dof_stiffness = np.zeros(nv)
    body_dofadr = -np.ones(nbody, np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_lastdof = -np.ones(nbody, np.int32)  # last dof of the closest ancestor-or-self with dofs
    for b in range(1, nbody):
        bd = bodies[b]
        last = body_lastdof[bd["parent"]]
        for ji in bd["jnts"]:
            j = joints[ji]
            nd = 6 if j["type"] == JNT_FREE else 1
            if body_dofadr[b] < 0:
                body_dofadr[b] = j["dofadr"]
            body_dofnum[b] += nd
            for k in range(nd):
                d = j["dofadr"] + k
                dof_body[d], dof_jnt[d], dof_parent[d] = b, ji, last
                dof_armature[d], dof_damping[d] = j["armature"], j["damping"]
                dof_stiffness[d] = j["stiffness"]
                last = d
            if j["type"] == JNT_FREE:
                qpos0[j["qposadr"]: j["qposadr"] + 3] = bd["pos"]
                qpos0[j["qposadr"] + 3: j["qposadr"] + 7] = bd["quat"]
                qpos_spring[j["qposadr"]: j["qposadr"] + 7] = qpos0[j["qposadr"]: j["qposadr"] + 7]
            else:
                qpos0[j["qposadr"]] = j["ref"]
                qpos_spring[j["qposadr"]] = j["springref"]
        body_lastdof[b] = last

    # ---- body tree helpers (bodies are in DFS pre-order => each subtree is a contiguous id range)
    parent = np.array([b["parent"] for b in bodies], np.int32)
    depth = np.zeros(nbody, np.int32)
    for b in range(1, nbody):
        depth[b] = depth[parent[b]] + 1
    subtree_end = np.arange(1, nbody + 1, dtype=np.int32)
    for b in range(nbody - 1, 0, -1):
        subtree_end[parent[b]] = max(subtree_end[parent[b]], subtree_end[b])
    rootid = np.zeros(nbody, np.int32)
    weldid = np.zeros(nbody, np.int32)
    for b in range(1, nbody):
        rootid[b] = b if parent[b] == 0 else rootid[parent[b]]
        weldid[b] = b if bodies[b]["jnts"] else weldid[parent[b]]

    # ---- body inertial properties from geoms
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_inertia = np.zeros((nbody, 3, 3))  # full tensor about ipos, in body axes
    for g in geoms:
        g["mass"], g["inertia_diag"] = _geom_mass_inertia(g["type"], g["size"], g["density"])
    for b in range(1, nbody):
        gs = [geoms[i] for i in bodies[b]["geoms"]]
        m = sum(g["mass"] for g in gs)
        body_mass[b] = m
        if m <= 0:
            continue
        com = sum(g["mass"] * g["pos"] for g in gs) / m
        body_ipos[b] = com
        inert = np.zeros((3, 3))
        for g in gs:
            r = quat_to_mat(g["quat"])
            d = g["pos"] - com
            inert += r @ np.diag(g["inertia_diag"]) @ r.T + g["mass"] * (d @ d * np.eye(3) - np.outer(d, d))
        body_inertia[b] = inert
    subtreemass = np.array([body_mass[b: subtree_end[b]].sum() for b in range(nbody)])

    # ---- actuators (motors on hinge joints)
    jname = {j["name"]: i for i, j in enumerate(joints)}
    act = []
    a_root = root.find("actuator")
    if a_root is not None:
        for e in a_root:
            if e.tag != "motor":
                raise NotImplementedError(e.tag)
            a = defaults.resolve("motor", e, None)
            cr = _vec(a.get("ctrlrange")) if a.get("ctrlrange") else np.zeros(2)
            cl = a.get("ctrllimited", "auto")
            limited = 1 if (cl == "true" or (cl == "auto" and a.get("ctrlrange"))) else 0
            gear = _vec(a.get("gear"), 6, [1.0, 0, 0, 0, 0, 0])
            ji = jname[a["joint"]]
            act.append(dict(name=a.get("name"), dof=joints[ji]["dofadr"], gear=gear[0], ctrlrange=cr, ctrllimited=limited))
    nu = len(act)

    # ---- fixed tendons
    tendons = []
    t_root = root.find("tendon")
    if t_root is not None:
        for e in t_root:
            if e.tag != "fixed":
                raise NotImplementedError(e.tag)
            a = dict(defaults.classes["main"].get("tendon", {}))
            a.update(e.attrib)
            wraps = [(joints[jname[w.get("joint")]]["dofadr"], joints[jname[w.get("joint")]]["qposadr"], float(w.get("coef")))
                     for w in e if w.tag == "joint"]
            rng = _vec(a.get("range")) if a.get("range") else np.zeros(2)
            lim = a.get("limited", "auto")
            tendons.append(dict(name=a.get("name"), wraps=wraps, range=rng,
                                limited=1 if (lim == "true" or (lim == "auto" and a.get("range"))) else 0,
                                solref=_vec(a["solreflimit"]), solimp=_vec(a["solimplimit"], 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                                margin=float(a["margin"])))
    ntendon = len(tendons)

    # ---- sensors (touch only)
    sname = {s["name"]: i for i, s in enumerate(sites)}
    sensors = []
    s_root = root.find("sensor")
    if s_root is not None:
        for e in s_root:
            if e.tag != "touch":
                raise NotImplementedError(e.tag)
            sensors.append(dict(name=e.get("name"), site=sname[e.get("site")]))

    # ---- keyframes
    keys = {}
    k_root = root.find("keyframe")
    if k_root is not None:
        for e in k_root:
            q = _vec(e.get("qpos")) if e.get("qpos") else qpos0.copy()
            keys[e.get("name")] = q

    # ---- collision pairs (MJX collision_driver._geom_pairs semantics)
    bname = {b["name"]: i for i, b in enumerate(bodies)}
    excl = set()
    c_root = root.find("contact")
    if c_root is not None:
        for e in c_root:
            if e.tag == "exclude":
                b1, b2 = bname[e.get("body1")], bname[e.get("body2")]
                excl.add((min(b1, b2), max(b1, b2)))
            else:
                raise NotImplementedError(e.tag)
    raw_pairs = []
    for b1 in range(nbody):
        for b2 in range(b1, nbody):
            if b1 == b2 or (b1, b2) in excl:
                continue
            w1, w2 = weldid[b1], weldid[b2]
            if w1 == w2:
                continue
            w1p, w2p = weldid[parent[w1]], weldid[parent[w2]]
            if w1 != 0 and w2 != 0 and (w1 == w2p or w2 == w1p):
                continue
            for g1 in bodies[b1]["geoms"]:
                for g2 in bodies[b2]["geoms"]:
                    ga, gb = geoms[g1], geoms[g2]
                    if not ((ga["contype"] & gb["conaffinity"]) or (gb["contype"] & ga["conaffinity"])):
                        continue
                    i1, i2 = (g1, g2) if ga["type"] <= gb["type"] else (g2, g1)
                    raw_pairs.append((i1, i2))
    groups: Dict[tuple, list] = {}
    for (g1, g2) in raw_pairs:
        ga, gb = geoms[g1], geoms[g2]
        if ga["priority"] != gb["priority"]:
            raise NotImplementedError("geom priority")
        condim = max(ga["condim"], gb["condim"])
        groups.setdefault((condim, ga["type"], gb["type"]), []).append((g1, g2))
    # contact slots are ordered by condim (frictionless first), groups keep first-encounter order inside a condim
    order = sorted(groups.keys(), key=lambda k: k[0])
    pairs = []
    con_adr = 0
    for key in order:
        condim, t1, t2 = key
        if condim not in (1, 3):
            raise NotImplementedError(f"condim {condim}")
        kind = _PAIR_KIND[(t1, t2)]
        for (g1, g2) in groups[key]:
            ga, gb = geoms[g1], geoms[g2]
            mix = ga["solmix"] / (ga["solmix"] + gb["solmix"])
            fr = np.maximum(ga["friction"], gb["friction"])
            if ga["solref"][0] > 0 and gb["solref"][0] > 0:
                solref = mix * ga["solref"] + (1 - mix) * gb["solref"]
            else:
                solref = np.minimum(ga["solref"], gb["solref"])
            solimp = mix * ga["solimp"] + (1 - mix) * gb["solimp"]
            pairs.append(dict(g1=g1, g2=g2, kind=kind, condim=condim, ncon=_PAIR_NCON[kind], con_adr=con_adr,
                              mu=fr[0], friction=np.array([fr[0], fr[0], fr[1], fr[2], fr[2]]),
                              solref=solref, solimp=solimp))
            con_adr += _PAIR_NCON[kind]
    ncon = con_adr
    npair = len(pairs)

    # ---- constraint row layout: joint limits, tendon limits, condim-1 contacts, condim-3 contacts x4
    lim_jnts = [i for i, j in enumerate(joints) if j["limited"] and j["type"] == JNT_HINGE]
    lim_tendons = [i for i, t in enumerate(tendons) if t["limited"]]
    nlimit, ntlimit = len(lim_jnts), len(lim_tendons)
    row = nlimit + ntlimit
    ncon1 = sum(p["ncon"] for p in pairs if p["condim"] == 1)
    for p in pairs:
        if p["condim"] == 1:
            p["efc_adr"] = row + p["con_adr"]
        else:
            p["efc_adr"] = row + ncon1 + 4 * (p["con_adr"] - ncon1)
    nefc = row + ncon1 + 4 * (ncon - ncon1)

    model: Dict[str, Any] = dict(
        opt=opt, nq=nq, nv=nv, nu=nu, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite, ntendon=ntendon,
        nsensor=len(sensors), npair=npair, ncon=ncon, nefc=nefc, nlimit=nlimit, ntlimit=ntlimit, ncon1=ncon1,
        body_name=[b["name"] for b in bodies], body_parent=parent, body_depth=depth, body_subtree_end=subtree_end,
        body_rootid=rootid, body_weldid=weldid, body_pos=np.array([b["pos"] for b in bodies]),
        body_quat=np.array([b["quat"] for b in bodies]), body_ipos=body_ipos, body_inertia=body_inertia,
        body_mass=body_mass, body_subtreemass=subtreemass, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        body_jntadr=np.array([b["jnts"][0] if b["jnts"] else -1 for b in bodies], np.int32),
        body_jntnum=np.array([len(b["jnts"]) for b in bodies], np.int32),
        jnt_name=[j["name"] for j in joints], jnt_type=np.array([j["type"] for j in joints], np.int32),
        jnt_body=np.array([j["body"] for j in joints], np.int32),
        jnt_qposadr=np.array([j["qposadr"] for j in joints], np.int32),
        jnt_dofadr=np.array([j["dofadr"] for j in joints], np.int32),
        jnt_pos=np.array([j["pos"] for j in joints]), jnt_axis=np.array([j["axis"] for j in joints]),
        jnt_range=np.array([j["range"] for j in joints]), jnt_limited=np.array([j["limited"] for j in joints], np.int32),
        jnt_stiffness=np.array([j["stiffness"] for j in joints]),
        jnt_solref=np.array([j["solref"] for j in joints]), jnt_solimp=np.array([j["solimp"] for j in joints]),
        lim_jnts=np.array(lim_jnts, np.int32), lim_tendons=np.array(lim_tendons, np.int32),
        dof_body=dof_body, dof_jnt=dof_jnt, dof_parent=dof_parent, dof_armature=dof_armature,
        dof_damping=dof_damping, dof_stiffness=dof_stiffness,
        qpos0=qpos0, qpos_spring=qpos_spring,
        geom_name=[g["name"] for g in geoms], geom_type=np.array([g["type"] for g in geoms], np.int32),
        geom_body=np.array([g["body"] for g in geoms], np.int32), geom_size=np.array([g["size"] for g in geoms]),
        geom_pos=np.array([g["pos"] for g in geoms]), geom_quat=np.array([g["quat"] for g in geoms]),
        geom_mass=np.array([g["mass"] for g in geoms]),
        site_name=[s["name"] for s in sites], site_body=np.array([s["body"] for s in sites], np.int32),
        site_pos=np.array([s["pos"] for s in sites]).reshape(-1, 3), site_quat=np.array([s["quat"] for s in sites]).reshape(-1, 4),
        site_size=np.array([s["size"] for s in sites]).reshape(-1, 3), site_type=np.array([s["type"] for s in sites], np.int32),
        sensor_name=[s["name"] for s in sensors], sensor_site=np.array([s["site"] for s in sensors], np.int32),
        act_name=[a["name"] for a in act], act_dof=np.array([a["dof"] for a in act], np.int32),
        act_gear=np.array([a["gear"] for a in act]), act_ctrlrange=np.array([a["ctrlrange"] for a in act]).reshape(-1, 2),
        act_ctrllimited=np.array([a["ctrllimited"] for a in act], np.int32),
        tendons=tendons, pairs=pairs, keyframes=keys,
    )
    _set_const(model)
    return model


# ----------------------------------------------------------------------------- numpy forward pieces (f64)
def np_kinematics(model: Dict[str, Any], qpos: np.ndarray):
    """Body frames, inertial frames, joint anchors/axes at `qpos`."""
    nb = model["nbody"]
    xpos, xquat = np.zeros((nb, 3)), np.zeros((nb, 4))
    xquat[0, 0] = 1.0
    xanchor, xaxis = np.zeros((model["njnt"], 3)), np.zeros((model["njnt"], 3))
    for b in range(1, nb):
        p = model["body_parent"][b]
        pos = xpos[p] + quat_to_mat(xquat[p]) @ model["body_pos"][b]
        quat = quat_mul(xquat[p], model["body_quat"][b])
        ja, jn = model["body_jntadr"][b], model["body_jntnum"][b]
        for j in range(ja, ja + jn):
            qa = model["jnt_qposadr"][j]
            if model["jnt_type"][j] == JNT_FREE:
                xanchor[j], xaxis[j] = qpos[qa:qa + 3], [0, 0, 1.0]
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
            else:
                r = quat_to_mat(quat)
                xanchor[j] = r @ model["jnt_pos"][j] + pos
                xaxis[j] = r @ model["jnt_axis"][j]
                quat = quat_mul(quat, axis_angle_quat(model["jnt_axis"][j], qpos[qa] - model["qpos0"][qa]))
                pos = xanchor[j] - quat_to_mat(quat) @ model["jnt_pos"][j]
        xpos[b], xquat[b] = pos, quat
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, model["body_ipos"])
    return xpos, xquat, xmat, xipos, xanchor, xaxis


def np_mass_matrix(model: Dict[str, Any], qpos: np.ndarray):
    """Dense joint-space inertia via per-body point Jacobians (independent of the CRB recursion the
    CPU restatement and the kernels use): M = sum_b m_b Jp_b^T Jp_b + Jr_b^T I_b Jr_b + diag(armature)."""
    nb, nv = model["nbody"], model["nv"]
    xpos, xquat, xmat, xipos, xanchor, xaxis = np_kinematics(model, qpos)
    jacp, jacr = np_body_jacobians(model, xmat, xipos, xanchor, xaxis)
    m = np.diag(model["dof_armature"]).astype(np.float64)
    for b in range(1, nb):
        iw = xmat[b] @ model["body_inertia"][b] @ xmat[b].T
        m += model["body_mass"][b] * jacp[b].T @ jacp[b] + jacr[b].T @ iw @ jacr[b]
    return m, (xpos, xquat, xmat, xipos, xanchor, xaxis, jacp, jacr)


def np_body_jacobians(model, xmat, xipos, xanchor, xaxis):
    """World-frame Jacobians (3 x nv each) of every body's inertial-frame origin and orientation."""
    nb, nv = model["nbody"], model["nv"]
    jacp, jacr = np.zeros((nb, 3, nv)), np.zeros((nb, 3, nv))
    for b in range(1, nb):
        a = b
        while a > 0:
            ja, jn = model["body_jntadr"][a], model["body_jntnum"][a]
            for j in range(ja, ja + jn):
                d = model["jnt_dofadr"][j]
                if model["jnt_type"][j] == JNT_FREE:
                    jacp[b][:, d:d + 3] = np.eye(3)
                    for k in range(3):
                        ax = xmat[a][:, k]  # rotational free dofs are body-frame
                        jacr[b][:, d + 3 + k] = ax
                        jacp[b][:, d + 3 + k] = np.cross(ax, xipos[b] - xanchor[j])
                else:
                    jacr[b][:, d] = xaxis[j]
                    jacp[b][:, d] = np.cross(xaxis[j], xipos[b] - xanchor[j])
            a = model["body_parent"][a]
    return jacp, jacr


def _set_const(model: Dict[str, Any]):
    """mj_setConst quantities at qpos0 (last bullets)."""
    nv, nb = model["nv"], model["nbody"]
    m, (xpos, xquat, xmat, xipos, xanchor, xaxis, jacp, jacr) = np_mass_matrix(model, model["qpos0"])
    minv = np.linalg.inv(m)
    dof_inv = np.diag(minv).copy()
    for j in range(model["njnt"]):
        if model["jnt_type"][j] == JNT_FREE:
            d = model["jnt_dofadr"][j]
            dof_inv[d:d + 3] = dof_inv[d:d + 3].mean()
            dof_inv[d + 3:d + 6] = dof_inv[d + 3:d + 6].mean()
    body_inv = np.zeros((nb, 2))
    for b in range(1, nb):
        if model["body_weldid"][b] == 0:
            continue
        ap = jacp[b] @ minv @ jacp[b].T
        ar = jacr[b] @ minv @ jacr[b].T
        body_inv[b] = [max(MJ_MINVAL, np.trace(ap) / 3.0), max(MJ_MINVAL, np.trace(ar) / 3.0)]
    for t in model["tendons"]:
        jt = np.zeros(nv)
        for (dof, _, coef) in t["wraps"]:
            jt[dof] = coef
        t["J"] = jt
        t["invweight0"] = float(jt @ minv @ jt)
    model["dof_invweight0"] = dof_inv
    model["body_invweight0"] = body_inv
    model["meaninertia"] = float(np.trace(m) / nv)
    model["qM0"] = m
    model["xpos0"] = xpos
    # contact-row invweights
    for p in model["pairs"]:
        b1, b2 = model["geom_body"][p["g1"]], model["geom_body"][p["g2"]]
        iw = body_inv[b1, 0] + body_inv[b2, 0]
        p["b1"], p["b2"] = int(b1), int(b2)
        if p["condim"] == 1:
            p["invweight"] = iw
        else:
            mu = p["mu"]
            p["invweight"] = (iw + mu * mu * iw) * 2.0 * mu * mu / model["opt"]["impratio"]


# ----------------------------------------------------------------------------- blob packing
MAXBODY, MAXJNT, MAXDOF, MAXQ, MAXGEOM, MAXPAIR, MAXU, MAXTENDON, MAXWRAP, MAXSITE, MAXSENSOR = \
    20, 24, 32, 32, 24, 192, 24, 4, 4, 4, 4
BLOB_MAGIC, BLOB_VERSION = 0x4D4A5842, 3  # "MJXB"

_I, _F = np.int32, np.float32
BLOB_DTYPE = np.dtype([
    ("magic", _I), ("version", _I),
    ("nq", _I), ("nv", _I), ("nu", _I), ("nbody", _I), ("njnt", _I), ("ngeom", _I), ("nsite", _I), ("ntendon", _I),
    ("nsensor", _I), ("npair", _I), ("ncon", _I), ("nefc", _I), ("nlimit", _I), ("ntlimit", _I), ("ncon1", _I),
    ("solver", _I), ("iterations", _I), ("ls_iterations", _I), ("integrator", _I), ("eulerdamp", _I), ("maxdepth", _I),
    ("timestep", _F), ("gravity", _F, 3), ("tolerance", _F), ("ls_tolerance", _F), ("impratio", _F), ("meaninertia", _F),
    # bodies
    ("body_parent", _I, MAXBODY), ("body_depth", _I, MAXBODY), ("body_subtree_end", _I, MAXBODY),
    ("body_jntadr", _I, MAXBODY), ("body_jntnum", _I, MAXBODY), ("body_dofadr", _I, MAXBODY), ("body_dofnum", _I, MAXBODY),
    ("body_pos", _F, (MAXBODY, 3)), ("body_quat", _F, (MAXBODY, 4)), ("body_ipos", _F, (MAXBODY, 3)),
    ("body_inertia", _F, (MAXBODY, 6)),  # xx yy zz xy xz yz, about ipos, body axes
    ("body_mass", _F, MAXBODY), ("body_invweight0", _F, (MAXBODY, 2)),
    # joints
    ("jnt_type", _I, MAXJNT), ("jnt_body", _I, MAXJNT), ("jnt_qposadr", _I, MAXJNT), ("jnt_dofadr", _I, MAXJNT),
    ("jnt_limited", _I, MAXJNT),
    ("jnt_pos", _F, (MAXJNT, 3)), ("jnt_axis", _F, (MAXJNT, 3)), ("jnt_range", _F, (MAXJNT, 2)),
    ("jnt_solref", _F, (MAXJNT, 2)), ("jnt_solimp", _F, (MAXJNT, 5)),
    ("lim_jnt", _I, MAXJNT),
    # dofs
    ("dof_body", _I, MAXDOF), ("dof_jnt", _I, MAXDOF), ("dof_parent", _I, MAXDOF),
    ("dof_armature", _F, MAXDOF), ("dof_damping", _F, MAXDOF), ("dof_stiffness", _F, MAXDOF),
    ("dof_invweight0", _F, MAXDOF),
    ("qpos0", _F, MAXQ), ("qpos_spring", _F, MAXQ),
    # geoms
    ("geom_type", _I, MAXGEOM), ("geom_body", _I, MAXGEOM),
    ("geom_size", _F, (MAXGEOM, 3)), ("geom_pos", _F, (MAXGEOM, 3)), ("geom_quat", _F, (MAXGEOM, 4)),
    # pairs
    ("pair_g1", _I, MAXPAIR), ("pair_g2", _I, MAXPAIR), ("pair_kind", _I, MAXPAIR), ("pair_condim", _I, MAXPAIR),
    ("pair_conadr", _I, MAXPAIR), ("pair_efcadr", _I, MAXPAIR),
    ("pair_mu", _F, MAXPAIR), ("pair_invweight", _F, MAXPAIR), ("pair_solref", _F, (MAXPAIR, 2)),
    ("pair_solimp", _F, (MAXPAIR, 5)),
    # actuators
    ("act_dof", _I, MAXU), ("act_ctrllimited", _I, MAXU), ("act_gear", _F, MAXU), ("act_ctrlrange", _F, (MAXU, 2)),
    # tendons
    ("ten_limited", _I, MAXTENDON), ("ten_nwrap", _I, MAXTENDON), ("ten_dof", _I, (MAXTENDON, MAXWRAP)),
    ("ten_qpos", _I, (MAXTENDON, MAXWRAP)), ("ten_coef", _F, (MAXTENDON, MAXWRAP)), ("ten_range", _F, (MAXTENDON, 2)),
    ("ten_solref", _F, (MAXTENDON, 2)), ("ten_solimp", _F, (MAXTENDON, 5)), ("ten_invweight0", _F, MAXTENDON),
    ("lim_ten", _I, MAXTENDON),
    # sites / sensors
    ("site_body", _I, MAXSITE), ("site_type", _I, MAXSITE), ("site_pos", _F, (MAXSITE, 3)),
    ("site_quat", _F, (MAXSITE, 4)), ("site_size", _F, (MAXSITE, 3)),
    ("sensor_site", _I, MAXSENSOR),
])


def pack_blob(model: Dict[str, Any]) -> np.ndarray:
    """Pack `model` into the POD layout of `struct mjxb_model_blob` (csrc/mjxb_model.h)."""
    b = np.zeros((), dtype=BLOB_DTYPE)
    caps = dict(nbody=MAXBODY, njnt=MAXJNT, nv=MAXDOF, nq=MAXQ, ngeom=MAXGEOM, npair=MAXPAIR, nu=MAXU,
                ntendon=MAXTENDON, nsite=MAXSITE, nsensor=MAXSENSOR)
    for k, cap in caps.items():
        if model[k] > cap:
            raise ValueError(f"model {k}={model[k]} exceeds compiled capacity {cap}")
    b["magic"], b["version"] = BLOB_MAGIC, BLOB_VERSION
    for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "ntendon", "nsensor", "npair", "ncon", "nefc",
              "nlimit", "ntlimit", "ncon1"):
        b[k] = model[k]
    o = model["opt"]
    b["solver"], b["iterations"], b["ls_iterations"] = o["solver"], o["iterations"], o["ls_iterations"]
    b["integrator"], b["eulerdamp"] = o["integrator"], o["eulerdamp"]
    b["maxdepth"] = int(model["body_depth"].max())
    b["timestep"], b["gravity"], b["tolerance"], b["ls_tolerance"] = o["timestep"], o["gravity"], o["tolerance"], o["ls_tolerance"]
    b["impratio"], b["meaninertia"] = o["impratio"], model["meaninertia"]

    def put(name, arr):
        arr = np.asarray(arr)
        n = arr.shape[0]
        if n:
            b[name][:n] = arr

    for k in ("body_parent", "body_depth", "body_subtree_end", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum",
              "body_pos", "body_quat", "body_ipos", "body_mass", "body_invweight0"):
        put(k, model[k])
    ii = model["body_inertia"]
    put("body_inertia", np.stack([ii[:, 0, 0], ii[:, 1, 1], ii[:, 2, 2], ii[:, 0, 1], ii[:, 0, 2], ii[:, 1, 2]], axis=1))
    for k in ("jnt_type", "jnt_body", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "jnt_pos", "jnt_axis", "jnt_range",
              "jnt_solref", "jnt_solimp"):
        put(k, model[k])
    put("lim_jnt", model["lim_jnts"])
    for k in ("dof_body", "dof_jnt", "dof_parent", "dof_armature", "dof_damping", "dof_stiffness", "dof_invweight0",
              "qpos0", "qpos_spring", "geom_type", "geom_body", "geom_size", "geom_pos", "geom_quat"):
        put(k, model[k])
    ps = model["pairs"]
    put("pair_g1", [p["g1"] for p in ps]); put("pair_g2", [p["g2"] for p in ps])
    put("pair_kind", [p["kind"] for p in ps]); put("pair_condim", [p["condim"] for p in ps])
    put("pair_conadr", [p["con_adr"] for p in ps]); put("pair_efcadr", [p["efc_adr"] for p in ps])
    put("pair_mu", [p["mu"] for p in ps]); put("pair_invweight", [p["invweight"] for p in ps])
    put("pair_solref", np.array([p["solref"] for p in ps]).reshape(-1, 2))
    put("pair_solimp", np.array([p["solimp"] for p in ps]).reshape(-1, 5))
    put("act_dof", model["act_dof"]); put("act_ctrllimited", model["act_ctrllimited"])
    put("act_gear", model["act_gear"]); put("act_ctrlrange", model["act_ctrlrange"])
    for i, t in enumerate(model["tendons"]):
        if len(t["wraps"]) > MAXWRAP:
            raise ValueError("too many tendon wraps")
        b["ten_limited"][i], b["ten_nwrap"][i] = t["limited"], len(t["wraps"])
        for w, (dof, qadr, coef) in enumerate(t["wraps"]):
            b["ten_dof"][i, w], b["ten_qpos"][i, w], b["ten_coef"][i, w] = dof, qadr, coef
        b["ten_range"][i], b["ten_solref"][i], b["ten_solimp"][i] = t["range"], t["solref"], t["solimp"]
        b["ten_invweight0"][i] = t["invweight0"]
    put("lim_ten", model["lim_tendons"])
    for k in ("site_body", "site_type", "site_pos", "site_quat", "site_size", "sensor_site"):
        put(k, model[k])
    return b


def blob_to_json(blob: np.ndarray) -> Dict[str, Any]:
    return {name: np.asarray(blob[name]).tolist() for name in BLOB_DTYPE.names}


def blob_from_json(d: Dict[str, Any]) -> np.ndarray:
    b = np.zeros((), dtype=BLOB_DTYPE)
    for name in BLOB_DTYPE.names:
        b[name] = np.asarray(d[name], dtype=BLOB_DTYPE[name].base)
    return b


def emit_c_header() -> str:
    """C declaration of the blob (kept in sync with BLOB_DTYPE; csrc/mjxb_model.h is generated from this)."""
    lines = ["// GENERATED by mujoco_mjx_lab_b200/modelc.py:emit_c_header() -- do not edit by hand.",
             "#pragma once", "#include <stdint.h>", ""]
    for k, v in dict(MJXB_MAXBODY=MAXBODY, MJXB_MAXJNT=MAXJNT, MJXB_MAXDOF=MAXDOF, MJXB_MAXQ=MAXQ, MJXB_MAXGEOM=MAXGEOM,
                     MJXB_MAXPAIR=MAXPAIR, MJXB_MAXU=MAXU, MJXB_MAXTENDON=MAXTENDON, MJXB_MAXWRAP=MAXWRAP,
                     MJXB_MAXSITE=MAXSITE, MJXB_MAXSENSOR=MAXSENSOR, MJXB_BLOB_MAGIC=BLOB_MAGIC,
                     MJXB_BLOB_VERSION=BLOB_VERSION).items():
        lines.append(f"#define {k} {v}")
    lines += ["", "typedef struct mjxb_model_blob {"]
    for name in BLOB_DTYPE.names:
        ft = BLOB_DTYPE[name]
        ctype = "int32_t" if ft.base == np.int32 else "float"
        dims = "".join(f"[{d}]" for d in ft.shape)
        lines.append(f"  {ctype} {name}{dims};")
    lines += ["} mjxb_model_blob;", ""]
    return "\n".join(lines)


# ----------------------------------------------------------------------------- (de)serialisation of the compiled dict
def _to_jsonable(x):
    if isinstance(x, np.ndarray):
        return {"__nd__": x.tolist(), "dtype": str(x.dtype)}
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, dict):
        return {k: _to_jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_to_jsonable(v) for v in x]
    return x


def _from_jsonable(x):
    if isinstance(x, dict):
        if "__nd__" in x:
            return np.array(x["__nd__"], dtype=np.dtype(x["dtype"]))
        return {k: _from_jsonable(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_from_jsonable(v) for v in x]
    return x


def save_model(model: Dict[str, Any], path: str):
    import json
    with open(path, "w") as fh:
        json.dump(_to_jsonable(model), fh)


def load_model(path: str) -> Dict[str, Any]:
    import json
    with open(path) as fh:
        m = _from_jsonable(json.load(fh))
    for t in m["tendons"]:
        t["wraps"] = [tuple(w) for w in t["wraps"]]
    return m


def builtin_model(name: str = "humanoid_mjx") -> Dict[str, Any]:
    """Compiled constants shipped with the package (generated by tools/compile_models.py from the reference's XML)."""
    import os
    return load_model(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", f"{name}.json"))
