"""MJCF-subset model compiler (host side, init only).

The reference builds its physics model by handing `assets/rodent/rodent.xml` to the
MuJoCo C library (`mujoco.MjSpec.from_file(...)`, edits, `.compile()`;
reference `track_mjx/environment/walker/rodent.py:51-87`, `spec_utils.py:19-52`)
and then overriding solver options (`environment/task/single_clip_tracking.py:64-72`).
MuJoCo is not available to this build (not vendored, not installed), so this module
restates the part of the MuJoCo model compiler that the rodent model exercises:

* `<default>` class inheritance for joint / geom / general / tendon,
* body tree in depth-first order, joints in order of appearance inside each body,
* `angle="radian"`, `euler` (sequence "xyz", intrinsic) and `quat` geom/body frames,
* mass and inertia from geom density for sphere / capsule / ellipsoid / box,
  composed into a body inertial frame (principal axes),
* fixed tendons, `general` actuators (gain/bias/dyn parameters, gear),
* the constants MuJoCo derives at `qpos0`: `dof_invweight0`, `body_invweight0`,
  `stat.meaninertia`.

and the two spec edits the reference walker applies before compiling:
torque-actuator rewrite (`rodent.py:70-78`) and the uniform rescale
(`spec_utils.dm_scale_spec`, `spec_utils.py:19-52`).

Everything is float64 numpy; the result is a plain dict of arrays named after the
`mjModel` fields they stand for.  `model_blob.pack` turns it into the flat fp32
constant table the CUDA kernels and the CPU oracle consume.
"""

from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from typing import Any

import numpy as np

MJ_MINVAL = 1e-15

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_BOX = 0, 2, 3, 4, 6
_GEOM_TYPES = {
    "plane": GEOM_PLANE,
    "sphere": GEOM_SPHERE,
    "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID,
    "box": GEOM_BOX,
}
JNT_FREE, JNT_HINGE = 0, 3


# --------------------------------------------------------------------------- #
# small quaternion helpers (w, x, y, z)
# --------------------------------------------------------------------------- #
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def mat_to_quat(m):
    """Rotation matrix -> unit quaternion (w >= 0 branch selection on the largest term)."""
    t = np.trace(m)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def euler_xyz_to_quat(e):
    """MuJoCo default eulerseq "xyz": intrinsic rotations, q = qx * qy * qz."""
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for axis, a in enumerate(e):
        t = np.zeros(4)
        t[0] = math.cos(a / 2)
        t[axis + 1] = math.sin(a / 2)
        q = quat_mul(q, t)
    return q


def rotate(v, q):
    return quat_to_mat(q) @ np.asarray(v, dtype=np.float64)


# --------------------------------------------------------------------------- #
# XML parsing with default classes
# --------------------------------------------------------------------------- #
_DEFAULTABLE = ("joint", "geom", "general", "tendon", "site")


def _parse_defaults(root: ET.Element) -> dict[str, dict[str, dict[str, str]]]:
    """class name -> element tag -> attribute dict (already merged with ancestors)."""
    classes: dict[str, dict[str, dict[str, str]]] = {}

    def visit(node: ET.Element, parent: dict[str, dict[str, str]], name: str):
        cur = {tag: dict(parent.get(tag, {})) for tag in _DEFAULTABLE}
        for child in node:
            if child.tag in _DEFAULTABLE:
                cur[child.tag].update(child.attrib)
        classes[name] = cur
        for child in node:
            if child.tag == "default":
                visit(child, cur, child.attrib["class"])

    empty = {tag: {} for tag in _DEFAULTABLE}
    tops = [d for d in root.findall("default")]
    if not tops:
        classes["main"] = empty
    for top in tops:
        visit(top, empty, top.attrib.get("class", "main"))
    return classes


def _fvec(s: str | None, n: int | None = None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) < n and default is not None:
        full = np.array(default, dtype=np.float64)
        full[: len(v)] = v
        v = full
    return v


def _resolve(el: ET.Element, classes, tag: str | None = None) -> dict[str, str]:
    tag = tag or el.tag
    cls = el.attrib.get("class", "main")
    out = dict(classes[cls].get(tag, {}))
    out.update(el.attrib)
    return out


def _frame_quat(a: dict[str, str]) -> np.ndarray:
    if "quat" in a:
        q = _fvec(a["quat"])
        return q / np.linalg.norm(q)
    if "euler" in a:
        return euler_xyz_to_quat(_fvec(a["euler"]))
    return np.array([1.0, 0.0, 0.0, 0.0])


# --------------------------------------------------------------------------- #
# geom mass properties (MuJoCo formulas for primitive shapes)
# --------------------------------------------------------------------------- #
def _geom_volume_inertia(gtype: int, size: np.ndarray) -> tuple[float, np.ndarray]:
    """Returns (volume, unit-density diagonal inertia in the geom frame)."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        vol = 4.0 / 3.0 * math.pi * r**3
        i = 2.0 / 5.0 * vol * r * r
        return vol, np.array([i, i, i])
    if gtype == GEOM_CAPSULE:
        r, height = size[0], 2 * size[1]
        vol_s = 4.0 / 3.0 * math.pi * r**3
        vol_c = math.pi * r * r * height
        ixx = vol_c * (3 * r * r + height * height) / 12.0
        izz = vol_c * r * r / 2.0
        sph = 2.0 * vol_s * r * r / 5.0
        ixx += sph + vol_s * height * (3 * r + 2 * height) / 8.0
        izz += sph
        return vol_s + vol_c, np.array([ixx, ixx, izz])
    if gtype == GEOM_ELLIPSOID:
        a, b, c = size[:3]
        vol = 4.0 / 3.0 * math.pi * a * b * c
        return vol, vol / 5.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    if gtype == GEOM_BOX:
        a, b, c = size[:3]
        vol = 8 * a * b * c
        return vol, vol / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    if gtype == GEOM_PLANE:
        return 0.0, np.zeros(3)
    raise ValueError(f"unsupported geom type {gtype}")


def _principal_axes(full: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Symmetric 3x3 inertia -> (iquat, principal inertias sorted descending)."""
    w, v = np.linalg.eigh(full)
    order = np.argsort(-w)
    w, v = w[order], v[:, order]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    return mat_to_quat(v), w


# --------------------------------------------------------------------------- #
# the compiler
# --------------------------------------------------------------------------- #
def compile_mjcf(
    xml_path: str,
    torque_actuators: bool = False,
    rescale_factor: float = 1.0,
    scale_root_body: str = "walker",
) -> dict[str, Any]:
    """Parse + edit + compile, mirroring `Rodent._build_spec` then `MjSpec.compile()`.

    Args:
      xml_path: MJCF file.
      torque_actuators: reference `rodent.py:70-78` — gain := forcerange_hi, bias removed.
      rescale_factor: reference `spec_utils.dm_scale_spec` — scales body pos and geom
        size/pos of every body *below* `scale_root_body`, and actuator gear by s^2.
    """
    root = ET.parse(xml_path).getroot()
    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    if comp.get("angle", "degree") != "radian":
        raise ValueError("only angle=radian models are supported")
    classes = _parse_defaults(root)

    opt_el = root.find("option")
    opt = dict(opt_el.attrib) if opt_el is not None else {}

    bodies: list[dict[str, Any]] = []
    joints: list[dict[str, Any]] = []
    geoms: list[dict[str, Any]] = []

    def add_body(el: ET.Element | None, parent: int, scaled: bool):
        bid = len(bodies)
        if el is None:  # world
            b = dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]))
            children = root.find("worldbody")
        else:
            a = el.attrib
            pos = _fvec(a.get("pos"), default=[0, 0, 0])
            if scaled:
                pos = pos * rescale_factor
            b = dict(name=a.get("name", f"body{bid}"), parent=parent, pos=pos, quat=_frame_quat(a))
            children = el
        b["jnt"], b["geom"] = [], []
        bodies.append(b)
        # children of `scale_root_body` (and below) are scaled; the root body itself is not
        child_scaled = scaled or (b["name"] == scale_root_body)
        for ch in children:
            if ch.tag in ("joint", "freejoint"):
                jid = len(joints)
                if ch.tag == "freejoint" or ch.attrib.get("type") == "free":
                    j = dict(name=ch.attrib.get("name", ""), type=JNT_FREE, body=bid, pos=np.zeros(3),
                             axis=np.array([0.0, 0, 1]), limited=False, range=np.zeros(2), stiffness=0.0,
                             damping=0.0, armature=0.0, springref=0.0, ref=0.0, margin=0.0,
                             solref=np.array([0.02, 1.0]), solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]))
                else:
                    a = _resolve(ch, classes, "joint")
                    if a.get("type", "hinge") != "hinge":
                        raise ValueError("only free and hinge joints are supported")
                    axis = _fvec(a.get("axis"), default=[0, 0, 1])
                    axis = axis / np.linalg.norm(axis)
                    rng = _fvec(a.get("range"), default=[0, 0])
                    lim = a.get("limited", "auto")
                    limited = (lim == "true") or (lim == "auto" and "range" in a)
                    j = dict(name=a.get("name", ""), type=JNT_HINGE, body=bid,
                             pos=_fvec(a.get("pos"), default=[0, 0, 0]), axis=axis, limited=limited, range=rng,
                             stiffness=float(a.get("stiffness", 0)), damping=float(a.get("damping", 0)),
                             armature=float(a.get("armature", 0)), springref=float(a.get("springref", 0)),
                             ref=float(a.get("ref", 0)), margin=float(a.get("margin", 0)),
                             solref=_fvec(a.get("solreflimit"), 2, [0.02, 1.0]),
                             solimp=_fvec(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]))
                joints.append(j)
                b["jnt"].append(jid)
            elif ch.tag == "geom":
                a = _resolve(ch, classes, "geom")
                gtype = _GEOM_TYPES[a.get("type", "sphere")]
                size = _fvec(a.get("size"), 3, [0, 0, 0])
                gpos = _fvec(a.get("pos"), default=[0, 0, 0])
                if "fromto" in a:
                    raise ValueError("fromto geoms are not supported")
                if scaled:
                    size, gpos = size * rescale_factor, gpos * rescale_factor
                g = dict(name=a.get("name", ""), type=gtype, body=bid, size=size, pos=gpos, quat=_frame_quat(a),
                         density=float(a.get("density", 1000.0)), contype=int(a.get("contype", 1)),
                         conaffinity=int(a.get("conaffinity", 1)), condim=int(a.get("condim", 3)),
                         priority=int(a.get("priority", 0)),
                         friction=_fvec(a.get("friction"), 3, [1.0, 0.005, 0.0001]),
                         solref=_fvec(a.get("solref"), 2, [0.02, 1.0]),
                         solimp=_fvec(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                         margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)))
                if "mass" in a:
                    raise ValueError("explicit geom mass is not supported")
                geoms.append(g)
                b["geom"].append(len(geoms) - 1)
            elif ch.tag == "inertial":
                raise ValueError("explicit <inertial> is not supported")
        for ch in children:
            if ch.tag == "body":
                add_body(ch, bid, child_scaled)

    # MuJoCo numbers bodies depth-first; joints/geoms are numbered body by body.  The
    # recursion above appends a body's joints/geoms before descending, which yields the
    # same ids because all of a body's own elements precede those of later bodies.
    add_body(None, 0, False)
    # renumber joints and geoms by body id (they were appended in visit order == body order)
    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)

    # ---------------- bodies: inertial frames from geoms ----------------
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for bid, b in enumerate(bodies):
        gl = [geoms[g] for g in b["geom"]]
        props = []
        for g in gl:
            vol, unit_i = _geom_volume_inertia(g["type"], g["size"])
            props.append((g["density"] * vol, g["density"] * unit_i, g))
        props = [p for p in props if p[0] > 0]
        if not props:
            continue
        if len(props) == 1:
            m, i, g = props[0]
            body_mass[bid], body_ipos[bid], body_iquat[bid], body_inertia[bid] = m, g["pos"], g["quat"], i
            continue
        mtot = sum(p[0] for p in props)
        com = sum(p[0] * p[2]["pos"] for p in props) / mtot
        full = np.zeros((3, 3))
        for m, i, g in props:
            r = quat_to_mat(g["quat"])
            d = g["pos"] - com
            full += r @ np.diag(i) @ r.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        q, w = _principal_axes(full)
        body_mass[bid], body_ipos[bid], body_iquat[bid], body_inertia[bid] = mtot, com, q, w

    # ---------------- joints / dofs ----------------
    jnt_qposadr, jnt_dofadr = np.zeros(njnt, int), np.zeros(njnt, int)
    nq = nv = 0
    for jid, j in enumerate(joints):
        jnt_qposadr[jid], jnt_dofadr[jid] = nq, nv
        nq += 7 if j["type"] == JNT_FREE else 1
        nv += 6 if j["type"] == JNT_FREE else 1
    dof_bodyid, dof_jntid = np.zeros(nv, int), np.zeros(nv, int)
    dof_armature, dof_damping = np.zeros(nv), np.zeros(nv)
    qpos0, qpos_spring = np.zeros(nq), np.zeros(nq)
    for jid, j in enumerate(joints):
        w = 6 if j["type"] == JNT_FREE else 1
        d0 = jnt_dofadr[jid]
        dof_bodyid[d0 : d0 + w] = j["body"]
        dof_jntid[d0 : d0 + w] = jid
        dof_armature[d0 : d0 + w] = j["armature"]
        dof_damping[d0 : d0 + w] = j["damping"]
        q0 = jnt_qposadr[jid]
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[q0 : q0 + 3], qpos0[q0 + 3 : q0 + 7] = b["pos"], b["quat"]
            qpos_spring[q0 : q0 + 7] = qpos0[q0 : q0 + 7]
        else:
            qpos0[q0] = j["ref"]
            qpos_spring[q0] = j["springref"]
    body_parentid = np.array([b["parent"] for b in bodies])
    body_jntnum = np.array([len(b["jnt"]) for b in bodies])
    body_jntadr = np.array([b["jnt"][0] if b["jnt"] else -1 for b in bodies])
    body_dofnum = np.array([sum(6 if joints[j]["type"] == JNT_FREE else 1 for j in b["jnt"]) for b in bodies])
    body_dofadr = np.array([jnt_dofadr[b["jnt"][0]] if b["jnt"] else -1 for b in bodies])
    body_rootid = np.zeros(nbody, int)
    for bid in range(1, nbody):
        p = body_parentid[bid]
        body_rootid[bid] = bid if p == 0 else body_rootid[p]
    # dof_parentid: previous dof in the same body, else last dof of the nearest ancestor with dofs
    dof_parentid = np.full(nv, -1)
    for d in range(nv):
        bid = dof_bodyid[d]
        if d > body_dofadr[bid]:
            dof_parentid[d] = d - 1
        else:
            p = body_parentid[bid]
            while p > 0 and body_dofnum[p] == 0:
                p = body_parentid[p]
            if p > 0:
                dof_parentid[d] = body_dofadr[p] + body_dofnum[p] - 1

    # ---------------- tendons (fixed only) ----------------
    jname = {j["name"]: i for i, j in enumerate(joints)}
    tendons = []
    ten_el = root.find("tendon")
    if ten_el is not None:
        for t in ten_el:
            if t.tag != "fixed":
                raise ValueError("only fixed tendons are supported")
            a = _resolve(t, classes, "tendon")
            lim = a.get("limited", "auto")
            if lim == "true" or (lim == "auto" and "range" in a):
                raise ValueError("tendon limits are not supported")
            if float(a.get("stiffness", 0)) or float(a.get("damping", 0)) or float(a.get("frictionloss", 0)):
                raise ValueError("tendon spring/damper/frictionloss are not supported")
            row = np.zeros(nv)
            for w in t.findall("joint"):
                row[jnt_dofadr[jname[w.attrib["joint"]]]] += float(w.attrib["coef"])
            tendons.append(dict(name=a.get("name", ""), row=row))
    tname = {t["name"]: i for i, t in enumerate(tendons)}
    ntendon = len(tendons)
    ten_J = np.array([t["row"] for t in tendons]).reshape(ntendon, nv)

    # ---------------- actuators (general) ----------------
    acts = []
    act_el = root.find("actuator")
    if act_el is not None:
        for el in act_el:
            if el.tag != "general":
                raise ValueError("only <general> actuators are supported")
            a = _resolve(el, classes, "general")
            gear = _fvec(a.get("gear"), 6, [1, 0, 0, 0, 0, 0])
            gainprm = _fvec(a.get("gainprm"), 10, [1] + [0] * 9)
            biasprm = _fvec(a.get("biasprm"), 10, [0] * 10)
            dynprm = _fvec(a.get("dynprm"), 10, [1] + [0] * 9)
            forcerange = _fvec(a.get("forcerange"), 2, [0, 0])
            ctrlrange = _fvec(a.get("ctrlrange"), 2, [0, 0])
            biastype = a.get("biastype", "none")
            if torque_actuators:  # reference rodent.py:70-78
                if forcerange.size >= 2:
                    gainprm[0] = forcerange[1]
                biastype = "none"
                biasprm = np.zeros(10)
            gear = gear * rescale_factor * rescale_factor if rescale_factor != 1.0 else gear  # spec_utils.py:38-42
            if a.get("gaintype", "fixed") != "fixed":
                raise ValueError("only gaintype=fixed is supported")
            cl = a.get("ctrllimited", "auto")
            fl = a.get("forcelimited", "auto")
            act = dict(name=a.get("name", ""), gear=gear[0], gain=gainprm[0], biastype=biastype,
                       biasprm=biasprm[:3], dyntype=a.get("dyntype", "none"), dynprm=dynprm[0],
                       ctrllimited=(cl == "true") or (cl == "auto" and "ctrlrange" in a), ctrlrange=ctrlrange,
                       forcelimited=(fl == "true") or (fl == "auto" and "forcerange" in a), forcerange=forcerange)
            if "joint" in a:
                act["moment"] = np.zeros(nv)
                act["moment"][jnt_dofadr[jname[a["joint"]]]] = act["gear"]
                act["trn"] = ("joint", jname[a["joint"]])
            elif "tendon" in a:
                act["moment"] = act["gear"] * ten_J[tname[a["tendon"]]]
                act["trn"] = ("tendon", tname[a["tendon"]])
            else:
                raise ValueError("only joint and tendon transmissions are supported")
            if act["dyntype"] not in ("none", "filter"):
                raise ValueError("only dyntype none/filter are supported")
            acts.append(act)
    nu = len(acts)
    na = sum(1 for a in acts if a["dyntype"] != "none")
    if na not in (0, nu):
        raise ValueError("mixed stateful/stateless actuators are not supported")

    model: dict[str, Any] = dict(
        nq=nq, nv=nv, nu=nu, na=na, nbody=nbody, njnt=njnt, ntendon=ntendon,
        body_names=[b["name"] for b in bodies], jnt_names=[j["name"] for j in joints],
        geom_names=[g["name"] for g in geoms], actuator_names=[a["name"] for a in acts],
        body_parentid=body_parentid, body_rootid=body_rootid,
        body_pos=np.array([b["pos"] for b in bodies]), body_quat=np.array([b["quat"] for b in bodies]),
        body_ipos=body_ipos, body_iquat=body_iquat, body_mass=body_mass, body_inertia=body_inertia,
        body_jntadr=body_jntadr, body_jntnum=body_jntnum, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        jnt_type=np.array([j["type"] for j in joints]), jnt_bodyid=np.array([j["body"] for j in joints]),
        jnt_qposadr=jnt_qposadr, jnt_dofadr=jnt_dofadr,
        jnt_pos=np.array([j["pos"] for j in joints]), jnt_axis=np.array([j["axis"] for j in joints]),
        jnt_limited=np.array([j["limited"] for j in joints]), jnt_range=np.array([j["range"] for j in joints]),
        jnt_stiffness=np.array([j["stiffness"] for j in joints]), jnt_margin=np.array([j["margin"] for j in joints]),
        jnt_solref=np.array([j["solref"] for j in joints]), jnt_solimp=np.array([j["solimp"] for j in joints]),
        qpos0=qpos0, qpos_spring=qpos_spring,
        dof_bodyid=dof_bodyid, dof_jntid=dof_jntid, dof_parentid=dof_parentid,
        dof_armature=dof_armature, dof_damping=dof_damping,
        ten_J=ten_J,
        actuator_moment=np.array([a["moment"] for a in acts]).reshape(nu, nv),
        actuator_gain=np.array([a["gain"] for a in acts]),
        actuator_biastype=[a["biastype"] for a in acts],
        actuator_biasprm=np.array([a["biasprm"] for a in acts]).reshape(nu, 3),
        actuator_dyntype=[a["dyntype"] for a in acts],
        actuator_dynprm=np.array([a["dynprm"] for a in acts]),
        actuator_ctrllimited=np.array([a["ctrllimited"] for a in acts]),
        actuator_ctrlrange=np.array([a["ctrlrange"] for a in acts]).reshape(nu, 2),
        actuator_forcelimited=np.array([a["forcelimited"] for a in acts]),
        actuator_forcerange=np.array([a["forcerange"] for a in acts]).reshape(nu, 2),
        actuator_trn=[a["trn"] for a in acts],
        geoms=geoms,
        opt=dict(
            timestep=float(opt.get("timestep", 0.002)), gravity=np.array([0.0, 0.0, -9.81]),
            solver=opt.get("solver", "Newton").lower(), iterations=int(opt.get("iterations", 100)),
            ls_iterations=int(opt.get("ls_iterations", 50)), tolerance=float(opt.get("tolerance", 1e-8)),
            ls_tolerance=float(opt.get("ls_tolerance", 0.01)), impratio=float(opt.get("impratio", 1.0)),
            cone=opt.get("cone", "pyramidal"),
        ),
    )
    _collision_pairs(model)
    _set_const(model)
    return model


def _collision_pairs(model: dict[str, Any]) -> None:
    """Static plane-vs-primitive contact set, as MJX enumerates it at trace time.

    MJX keeps every geom pair that passes the contype/conaffinity filter
    ((contype1 & conaffinity2) | (contype2 & conaffinity1)) and is not excluded; here only
    pairs whose first geom is a plane are supported (the rodent model has no others).
    Contact parameters follow MuJoCo's mixing rule with geom `priority`.
    """
    geoms = model["geoms"]
    planes = [i for i, g in enumerate(geoms) if g["type"] == GEOM_PLANE]
    others = [i for i, g in enumerate(geoms) if g["type"] != GEOM_PLANE]
    pairs = []
    for i in others:
        for k in others:
            if k <= i or geoms[i]["body"] == geoms[k]["body"]:
                continue
            a, b = geoms[i], geoms[k]
            if (a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"]):
                raise ValueError("non-plane collision pairs are not supported by this build")
    for p in planes:
        for i in others:
            a, b = geoms[p], geoms[i]
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            if b["type"] not in (GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID):
                raise ValueError("only plane-sphere/capsule/ellipsoid collisions are supported")
            if a["priority"] != b["priority"]:
                hi = a if a["priority"] > b["priority"] else b
                friction, solref, solimp, condim = hi["friction"], hi["solref"], hi["solimp"], hi["condim"]
            else:
                friction = np.maximum(a["friction"], b["friction"])
                mix = 0.5  # solmix equal
                solref = mix * a["solref"] + (1 - mix) * b["solref"]
                solimp = mix * a["solimp"] + (1 - mix) * b["solimp"]
                condim = max(a["condim"], b["condim"])
            if condim != 3:
                raise ValueError("only condim=3 contacts are supported")
            margin = max(a["margin"], b["margin"])
            gap = max(a["gap"], b["gap"])
            pairs.append(dict(plane=p, geom=i, friction=friction, solref=solref, solimp=solimp,
                              includemargin=margin - gap))
    # MJX groups contacts by collision function: plane-sphere, plane-capsule, plane-ellipsoid
    order = {GEOM_SPHERE: 0, GEOM_CAPSULE: 1, GEOM_ELLIPSOID: 2}
    pairs.sort(key=lambda p: (order[geoms[p["geom"]]["type"]], p["geom"]))
    model["contact_pairs"] = pairs
    model["ncon"] = sum(2 if geoms[p["geom"]]["type"] == GEOM_CAPSULE else 1 for p in pairs)


# --------------------------------------------------------------------------- #
# fp64 numpy kinematics + mass matrix (used for the qpos0 constants and by tests)
# --------------------------------------------------------------------------- #
def kinematics(model: dict[str, Any], qpos: np.ndarray):
    """Body frames, joint anchors/axes for `qpos` (MuJoCo mj_kinematics semantics)."""
    nbody = model["nbody"]
    xpos, xquat = np.zeros((nbody, 3)), np.zeros((nbody, 4))
    xquat[0] = [1, 0, 0, 0]
    xanchor, xaxis = np.zeros((model["njnt"], 3)), np.zeros((model["njnt"], 3))
    for b in range(1, nbody):
        p = model["body_parentid"][b]
        pos = xpos[p] + rotate(model["body_pos"][b], xquat[p])
        quat = quat_mul(xquat[p], model["body_quat"][b])
        for k in range(model["body_jntnum"][b]):
            j = model["body_jntadr"][b] + k
            qa = model["jnt_qposadr"][j]
            if model["jnt_type"][j] == JNT_FREE:
                xanchor[j], xaxis[j] = qpos[qa : qa + 3], [0, 0, 1]
                pos = qpos[qa : qa + 3].copy()
                quat = qpos[qa + 3 : qa + 7] / np.linalg.norm(qpos[qa + 3 : qa + 7])
            else:
                xanchor[j] = rotate(model["jnt_pos"][j], quat) + pos
                xaxis[j] = rotate(model["jnt_axis"][j], quat)
                ang = qpos[qa] - model["qpos0"][qa]
                qloc = np.concatenate([[math.cos(ang / 2)], model["jnt_axis"][j] * math.sin(ang / 2)])
                quat = quat_mul(quat, qloc)
                pos = xanchor[j] - rotate(model["jnt_pos"][j], quat)
        xpos[b], xquat[b] = pos, quat
    return xpos, xquat, xanchor, xaxis


def body_jacobians(model, qpos):
    """Per body: 6 x nv Jacobian [jacp; jacr] at the body's inertial-frame origin (world axes)."""
    xpos, xquat, xanchor, xaxis = kinematics(model, qpos)
    nbody, nv = model["nbody"], model["nv"]
    xipos = np.array([xpos[b] + rotate(model["body_ipos"][b], xquat[b]) for b in range(nbody)])
    jac = np.zeros((nbody, 6, nv))
    for b in range(1, nbody):
        a = b
        while a > 0:
            for k in range(model["body_jntnum"][a]):
                j = model["body_jntadr"][a] + k
                d = model["jnt_dofadr"][j]
                if model["jnt_type"][j] == JNT_FREE:
                    jac[b, 0:3, d : d + 3] = np.eye(3)
                    rot = quat_to_mat(xquat[a])
                    for c in range(3):
                        ax = rot[:, c]
                        jac[b, 3:6, d + 3 + c] = ax
                        jac[b, 0:3, d + 3 + c] = np.cross(ax, xipos[b] - xanchor[j])
                else:
                    jac[b, 3:6, d] = xaxis[j]
                    jac[b, 0:3, d] = np.cross(xaxis[j], xipos[b] - xanchor[j])
            a = model["body_parentid"][a]
    return jac, xipos, xquat


def mass_matrix(model, qpos):
    """Joint-space inertia M(q) = sum_b J_b^T diag(m, I_b) J_b + diag(armature) (fp64, dense)."""
    jac, _, xquat = body_jacobians(model, qpos)
    nv = model["nv"]
    m = np.diag(model["dof_armature"]).astype(np.float64)
    for b in range(1, model["nbody"]):
        mass = model["body_mass"][b]
        if mass <= 0:
            continue
        r = quat_to_mat(quat_mul(xquat[b], model["body_iquat"][b]))
        inert = r @ np.diag(model["body_inertia"][b]) @ r.T
        jp_, jr = jac[b, 0:3], jac[b, 3:6]
        m += mass * jp_.T @ jp_ + jr.T @ inert @ jr
    return m


def _set_const(model: dict[str, Any]) -> None:
    """`dof_invweight0`, `body_invweight0`, `stat.meaninertia` at qpos0 (MuJoCo mj_setConst)."""
    nv, nbody = model["nv"], model["nbody"]
    m = mass_matrix(model, model["qpos0"])
    minv = np.linalg.inv(m)
    dof_invweight0 = np.diag(minv).copy()
    for j in range(model["njnt"]):
        if model["jnt_type"][j] == JNT_FREE:  # averaged over translational / rotational triplets
            d = model["jnt_dofadr"][j]
            dof_invweight0[d : d + 3] = dof_invweight0[d : d + 3].mean()
            dof_invweight0[d + 3 : d + 6] = dof_invweight0[d + 3 : d + 6].mean()
    jac, _, _ = body_jacobians(model, model["qpos0"])
    body_invweight0 = np.zeros((nbody, 2))
    for b in range(1, nbody):
        if not jac[b].any():  # static body (welded to the world)
            continue
        a = jac[b] @ minv @ jac[b].T
        body_invweight0[b, 0] = (a[0, 0] + a[1, 1] + a[2, 2]) / 3
        body_invweight0[b, 1] = (a[3, 3] + a[4, 4] + a[5, 5]) / 3
    model["dof_invweight0"] = dof_invweight0
    model["body_invweight0"] = body_invweight0
    model["stat_meaninertia"] = float(np.trace(m) / max(nv, 1))
    model["qM0"] = m
