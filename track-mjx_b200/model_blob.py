"""Flat model-constant table ("blob") shared by the CUDA library and the CPU oracle.

Stands in for `mjx.put_model(self.sys.mj_model)` (reference
`track_mjx/environment/task/single_clip_tracking.py:91`): everything the step needs from the
compiled model, as one little-endian buffer of named fp32 / int32 sections.

Layout:
    header   : magic 'TMJX' (u32) | version (u32) | n_sections (i32) | reserved (i32)
    directory: n_sections x { name[24] | dtype (i32: 0=f32, 1=i32) | count (i32) | offset (i64) }
    data     : 16-byte aligned sections
"""

from __future__ import annotations

import struct
from typing import Any

import numpy as np

from . import mjcf

MAGIC = 0x584A4D54  # 'TMJX'
VERSION = 1


def model_sections(model: dict[str, Any]) -> dict[str, np.ndarray]:
    """Select + flatten the model fields consumed by the step into named arrays."""
    geoms = model["geoms"]
    pairs = model["contact_pairs"]
    planes = sorted({p["plane"] for p in pairs})
    if len(planes) > 1:
        raise ValueError("a single collision plane is supported")
    f32, i32 = np.float32, np.int32
    sec: dict[str, np.ndarray] = {}
    cg = sorted({p["geom"] for p in pairs}, key=lambda g: [q["geom"] for q in pairs].index(g))
    nefc = int(np.sum(model["jnt_limited"])) + 4 * model["ncon"]
    sec["dims"] = np.array(
        [model["nq"], model["nv"], model["nu"], model["na"], model["nbody"], model["njnt"], len(cg),
         model["ncon"], nefc, model["ntendon"]], i32)
    o = model["opt"]
    sec["opt"] = np.array(
        [o["timestep"], o["gravity"][0], o["gravity"][1], o["gravity"][2], o["tolerance"], o["ls_tolerance"],
         o["impratio"], model["stat_meaninertia"]], f32)
    for k in ("body_parentid", "body_rootid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum",
              "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "dof_bodyid", "dof_jntid", "dof_parentid"):
        sec[k] = np.asarray(model[k], i32)
    sec["jnt_limited"] = np.asarray(model["jnt_limited"], i32)
    for k in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_invweight0",
              "jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_margin", "jnt_solref", "jnt_solimp",
              "qpos0", "qpos_spring", "dof_armature", "dof_damping", "dof_invweight0",
              "actuator_moment", "actuator_gain", "actuator_biasprm", "actuator_dynprm", "actuator_ctrlrange",
              "actuator_forcerange"):
        sec[k] = np.asarray(model[k], np.float64).astype(f32).ravel()
    sec["actuator_ctrllimited"] = np.asarray(model["actuator_ctrllimited"], i32)
    sec["actuator_forcelimited"] = np.asarray(model["actuator_forcelimited"], i32)
    sec["actuator_bias_affine"] = np.array([t == "affine" for t in model["actuator_biastype"]], i32)
    sec["actuator_dyn_filter"] = np.array([t == "filter" for t in model["actuator_dyntype"]], i32)
    # colliding geoms (non-plane side of each contact pair)
    sec["cgeom_type"] = np.array([geoms[g]["type"] for g in cg], i32)
    sec["cgeom_bodyid"] = np.array([geoms[g]["body"] for g in cg], i32)
    sec["cgeom_pos"] = np.array([geoms[g]["pos"] for g in cg], f32).ravel()
    sec["cgeom_quat"] = np.array([geoms[g]["quat"] for g in cg], f32).ravel()
    sec["cgeom_size"] = np.array([geoms[g]["size"] for g in cg], f32).ravel()
    # the plane: its body must be static, so its world frame is a constant
    if planes:
        pg = geoms[planes[0]]
        b = pg["body"]
        chain_static = True
        pos, quat = np.zeros(3), np.array([1.0, 0, 0, 0])
        stack = []
        while b > 0:
            stack.append(b)
            chain_static &= model["body_dofnum"][b] == 0
            b = model["body_parentid"][b]
        if not chain_static:
            raise ValueError("the collision plane must be attached to a static body")
        for b in reversed(stack):
            pos = pos + mjcf.rotate(model["body_pos"][b], quat)
            quat = mjcf.quat_mul(quat, model["body_quat"][b])
        ppos = pos + mjcf.rotate(pg["pos"], quat)
        pmat = mjcf.quat_to_mat(mjcf.quat_mul(quat, pg["quat"]))
        sec["plane"] = np.concatenate([ppos, pmat[:, 2]]).astype(f32)
        sec["plane_bodyid"] = np.array([pg["body"]], i32)
    # per pair contact parameters
    sec["pair_cgeom"] = np.array([cg.index(p["geom"]) for p in pairs], i32)
    sec["pair_friction"] = np.array(
        [[p["friction"][0], p["friction"][0], p["friction"][1], p["friction"][2], p["friction"][2]] for p in pairs],
        f32).ravel()
    sec["pair_solref"] = np.array([p["solref"] for p in pairs], f32).ravel()
    sec["pair_solimp"] = np.array([p["solimp"] for p in pairs], f32).ravel()
    sec["pair_includemargin"] = np.array([p["includemargin"] for p in pairs], f32)
    return sec


def pack(model: dict[str, Any]) -> bytes:
    sec = model_sections(model)
    names = list(sec)
    dir_size = 16 + 40 * len(names)
    off = (dir_size + 15) // 16 * 16
    entries, chunks = [], []
    for n in names:
        a = np.ascontiguousarray(sec[n])
        dt = 0 if a.dtype == np.float32 else 1
        raw = a.tobytes()
        entries.append(struct.pack("<24siiq", n.encode()[:23], dt, a.size, off))
        pad = (-len(raw)) % 16
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = struct.pack("<IIii", MAGIC, VERSION, len(names), 0) + b"".join(entries)
    head += b"\0" * ((-len(head)) % 16)
    return head + b"".join(chunks)


def unpack(blob: bytes) -> dict[str, np.ndarray]:
    magic, version, n, _ = struct.unpack_from("<IIii", blob, 0)
    if magic != MAGIC or version != VERSION:
        raise ValueError("not a TMJX v1 model blob")
    out = {}
    for i in range(n):
        name, dt, count, off = struct.unpack_from("<24siiq", blob, 16 + 40 * i)
        name = name.split(b"\0")[0].decode()
        out[name] = np.frombuffer(blob, np.float32 if dt == 0 else np.int32, count, off).copy()
    return out


def pack_sections(sec: dict[str, np.ndarray]) -> bytes:
    """Re-pack an (edited) dict of sections as returned by `unpack` (tests use this to build model variants)."""
    names = list(sec)
    dir_size = 16 + 40 * len(names)
    off = (dir_size + 15) // 16 * 16
    entries, chunks = [], []
    for n in names:
        a = np.ascontiguousarray(sec[n])
        if a.dtype not in (np.float32, np.int32):
            raise ValueError(f"section {n}: dtype {a.dtype}")
        raw = a.tobytes()
        entries.append(struct.pack("<24siiq", n.encode()[:23], 0 if a.dtype == np.float32 else 1, a.size, off))
        pad = (-len(raw)) % 16
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = struct.pack("<IIii", MAGIC, VERSION, len(names), 0) + b"".join(entries)
    head += b"\0" * ((-len(head)) % 16)
    return head + b"".join(chunks)


def as_model(sec: dict[str, np.ndarray]) -> dict[str, Any]:
    """View unpacked sections as the dict `mjcf.kinematics / body_jacobians / mass_matrix` expect (fp64)."""
    d = sec["dims"]
    nq, nv, nu, na, nbody, njnt = (int(x) for x in d[:6])
    f = lambda k, *shape: sec[k].astype(np.float64).reshape(*shape) if shape else sec[k].astype(np.float64)  # noqa: E731
    i = lambda k: sec[k].astype(int)  # noqa: E731
    return dict(
        nq=nq, nv=nv, nu=nu, na=na, nbody=nbody, njnt=njnt,
        body_parentid=i("body_parentid"), body_rootid=i("body_rootid"), body_jntadr=i("body_jntadr"),
        body_jntnum=i("body_jntnum"), body_dofadr=i("body_dofadr"), body_dofnum=i("body_dofnum"),
        jnt_type=i("jnt_type"), jnt_qposadr=i("jnt_qposadr"), jnt_dofadr=i("jnt_dofadr"), jnt_bodyid=i("jnt_bodyid"),
        dof_bodyid=i("dof_bodyid"), dof_parentid=i("dof_parentid"),
        body_pos=f("body_pos", nbody, 3), body_quat=f("body_quat", nbody, 4), body_ipos=f("body_ipos", nbody, 3),
        body_iquat=f("body_iquat", nbody, 4), body_mass=f("body_mass"), body_inertia=f("body_inertia", nbody, 3),
        jnt_pos=f("jnt_pos", njnt, 3), jnt_axis=f("jnt_axis", njnt, 3), jnt_range=f("jnt_range", njnt, 2),
        qpos0=f("qpos0"), dof_armature=f("dof_armature"), dof_damping=f("dof_damping"),
    )
