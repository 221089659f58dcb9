"""Reference-clip container and the synthetic "stac-mjx-shaped" clip generator.

`ReferenceClip` mirrors reference `track_mjx/io/load.py:16-38` (same field names, each
`(n_clips, n_frames, d)` as `make_multiclip_data` builds them, load.py:105-137).  The HDF5 reader is
out of scope (no h5py here, no data files); `make_synthetic_clips` produces arrays of the same shape
and semantics from smooth random joint trajectories and the model's own forward kinematics.
"""

from __future__ import annotations

import dataclasses

import numpy as np

JNT_FREE = 0


@dataclasses.dataclass(frozen=True)
class ReferenceClip:
    """Trajectory features used by the tracking task (reference io/load.py:16-38)."""

    position: np.ndarray            # (C, F, 3)   qpos[:3]
    quaternion: np.ndarray          # (C, F, 4)   qpos[3:7]
    joints: np.ndarray              # (C, F, nq-7)
    body_positions: np.ndarray      # (C, F, n_ref_bodies, 3)   stac xpos (no `floor` body)
    velocity: np.ndarray            # (C, F, 3)
    angular_velocity: np.ndarray    # (C, F, 3)
    joints_velocity: np.ndarray     # (C, F, nv-6)
    body_quaternions: np.ndarray    # (C, F, n_ref_bodies, 4)
    original_clip_idx: np.ndarray | None = None

    @property
    def n_clips(self) -> int:
        return self.position.shape[0]

    @property
    def clip_length(self) -> int:
        return self.position.shape[1]


# --------------------------------------------------------------------------- batched quaternion helpers
def _qmul(a, b):
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def _qrot(v, q):
    s, u = q[..., :1], q[..., 1:]
    return 2 * np.sum(u * v, -1, keepdims=True) * u + (s * s - np.sum(u * u, -1, keepdims=True)) * v \
        + 2 * s * np.cross(u, v)


def batched_kinematics(sec: dict[str, np.ndarray], qpos: np.ndarray):
    """Body frames for a batch of qpos `(N, nq)` from the unpacked model table (`model_blob.unpack`).

    fp64 numpy restatement of MuJoCo's kinematics pass, vectorised over frames; returns
    `xpos (N, nbody, 3)`, `xquat (N, nbody, 4)`.
    """
    nbody = int(sec["dims"][4])
    n = qpos.shape[0]
    qpos = qpos.astype(np.float64)
    body_pos = sec["body_pos"].reshape(nbody, 3).astype(np.float64)
    body_quat = sec["body_quat"].reshape(nbody, 4).astype(np.float64)
    jnt_pos = sec["jnt_pos"].reshape(-1, 3).astype(np.float64)
    jnt_axis = sec["jnt_axis"].reshape(-1, 3).astype(np.float64)
    qpos0 = sec["qpos0"].astype(np.float64)
    xpos = np.zeros((n, nbody, 3))
    xquat = np.zeros((n, nbody, 4))
    xquat[:, 0, 0] = 1
    for b in range(1, nbody):
        p = int(sec["body_parentid"][b])
        pos = xpos[:, p] + _qrot(body_pos[b][None], xquat[:, p])
        quat = _qmul(xquat[:, p], np.broadcast_to(body_quat[b], (n, 4)))
        for k in range(int(sec["body_jntnum"][b])):
            j = int(sec["body_jntadr"][b]) + k
            qa = int(sec["jnt_qposadr"][j])
            if int(sec["jnt_type"][j]) == JNT_FREE:
                pos = qpos[:, qa:qa + 3]
                quat = qpos[:, qa + 3:qa + 7] / np.linalg.norm(qpos[:, qa + 3:qa + 7], axis=-1, keepdims=True)
            else:
                anchor = _qrot(jnt_pos[j][None], quat) + pos
                ang = qpos[:, qa] - qpos0[qa]
                qloc = np.concatenate([np.cos(ang / 2)[:, None], jnt_axis[j][None] * np.sin(ang / 2)[:, None]], -1)
                quat = _qmul(quat, qloc)
                pos = anchor - _qrot(jnt_pos[j][None], quat)
        xpos[:, b], xquat[:, b] = pos, quat
    return xpos, xquat


def make_synthetic_clips(sec: dict[str, np.ndarray], n_clips: int, clip_length: int = 250, mocap_hz: float = 50.0,
                         seed: int = 1234, root_height: float = 0.0495, joint_amplitude: float = 0.25,
                         drop_bodies: tuple[int, ...] = (1,)) -> ReferenceClip:
    """Synthetic stac-mjx-shaped clips (the SURVEY.md §8d generator).

    Per clip (numpy `default_rng(seed + clip)`): root xy = low-passed random walk (sigma 2 mm/frame), constant
    height, yaw-only root rotation `0.3 sin(2 pi 0.5 t + phi)`; every hinge follows
    `mid + amplitude * halfrange * sin(2 pi f t + phi)`, `f ~ U(0.5, 2)` Hz; velocities are central finite
    differences (angular part from the quaternion log); `body_positions/quaternions` are the model's forward
    kinematics of that qpos with `drop_bodies` (the static `floor` body, id 1) removed, which gives the 67-row
    arrays stac-mjx stores (reference walker/base.py:254 needs `body_positions.shape[-2] == 67`).
    """
    nq, nv = int(sec["dims"][0]), int(sec["dims"][1])
    nbody = int(sec["dims"][4])
    rng_lo = sec["jnt_range"].reshape(-1, 2)[1:, 0].astype(np.float64)
    rng_hi = sec["jnt_range"].reshape(-1, 2)[1:, 1].astype(np.float64)
    mid, half = 0.5 * (rng_lo + rng_hi), 0.5 * (rng_hi - rng_lo)
    t = np.arange(clip_length) / mocap_hz
    qpos = np.zeros((n_clips, clip_length, nq))
    for c in range(n_clips):
        rng = np.random.default_rng(seed + c)
        steps = rng.normal(0.0, 0.002, (clip_length, 2))
        kernel = np.ones(9) / 9.0
        steps = np.stack([np.convolve(steps[:, i], kernel, mode="same") for i in range(2)], -1)
        qpos[c, :, 0:2] = np.cumsum(steps, 0)
        qpos[c, :, 2] = root_height
        yaw = 0.3 * np.sin(2 * np.pi * 0.5 * t + rng.uniform(0, 2 * np.pi))
        qpos[c, :, 3], qpos[c, :, 6] = np.cos(yaw / 2), np.sin(yaw / 2)
        f = rng.uniform(0.5, 2.0, nq - 7)
        ph = rng.uniform(0, 2 * np.pi, nq - 7)
        q = mid[None] + joint_amplitude * half[None] * np.sin(2 * np.pi * f[None] * t[:, None] + ph[None])
        qpos[c, :, 7:] = np.clip(q, rng_lo[None], rng_hi[None])
    # finite-difference velocities
    qvel = np.zeros((n_clips, clip_length, nv))
    d = np.gradient(qpos, 1.0 / mocap_hz, axis=1)
    qvel[..., 0:3] = d[..., 0:3]
    qvel[..., 6:] = d[..., 7:]
    quat = qpos[..., 3:7]
    qn = np.roll(quat, -1, axis=1)
    qn[:, -1] = quat[:, -1]
    qp = np.roll(quat, 1, axis=1)
    qp[:, 0] = quat[:, 0]
    conj = qp * np.array([1.0, -1, -1, -1])
    dq = _qmul(conj, qn)  # body-frame rotation from previous to next frame
    ang = 2 * np.arctan2(np.linalg.norm(dq[..., 1:], axis=-1), dq[..., 0])
    axis = dq[..., 1:] / np.maximum(np.linalg.norm(dq[..., 1:], axis=-1, keepdims=True), 1e-12)
    span = np.full(clip_length, 2.0 / mocap_hz)
    span[0] = span[-1] = 1.0 / mocap_hz
    qvel[..., 3:6] = axis * (ang / span[None])[..., None]
    xpos, xquat = batched_kinematics(sec, qpos.reshape(-1, nq))
    keep = [b for b in range(nbody) if b not in drop_bodies]
    xpos = xpos[:, keep].reshape(n_clips, clip_length, len(keep), 3)
    xquat = xquat[:, keep].reshape(n_clips, clip_length, len(keep), 4)
    f32 = np.float32
    return ReferenceClip(
        position=qpos[..., :3].astype(f32), quaternion=qpos[..., 3:7].astype(f32), joints=qpos[..., 7:].astype(f32),
        body_positions=xpos.astype(f32), velocity=qvel[..., :3].astype(f32),
        angular_velocity=qvel[..., 3:6].astype(f32), joints_velocity=qvel[..., 6:].astype(f32),
        body_quaternions=xquat.astype(f32), original_clip_idx=np.arange(n_clips, dtype=np.int32))


# --------------------------------------------------------------------------- stac-mjx data -> clip table
_FIELDS = ("position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity", "body_quaternions")


def make_multiclip_data(source, n_frames_per_clip: int | None = None) -> ReferenceClip:
    """reference `io/load.py:105-137`: stac-mjx arrays `qpos (N, nq)`, `qvel (N, nv)`, `xpos (N, B, 3)`, `xquat (N, B, 4)` with
    `N = n_clips * n_frames_per_clip` reshaped to `(clips, frames, dims)` and split into the eight ReferenceClip fields.

    `source` is a mapping with those four arrays, or a path to an `.npz` holding them (plus optionally a scalar
    `n_frames_per_clip`), or a stac-mjx `.h5` file when h5py is importable (it is not in this image; the reference reads the
    clip length from the file's YAML config in that case)."""
    if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__"):
        path = str(source)
        if path.endswith((".h5", ".hdf5")):
            try:
                import h5py  # noqa: PLC0415
            except ImportError as e:  # pragma: no cover - h5py is absent from this image
                raise ImportError("reading stac-mjx HDF5 needs h5py; convert to .npz (qpos, qvel, xpos, xquat) instead") from e
            with h5py.File(path, "r") as f:
                data = {k: f[k][()] for k in ("qpos", "qvel", "xpos", "xquat")}
                if n_frames_per_clip is None:
                    import yaml  # noqa: PLC0415

                    n_frames_per_clip = yaml.safe_load(f["config"][()].decode("utf-8"))["stac"]["n_frames_per_clip"]
        else:
            with np.load(path) as f:
                data = {k: f[k] for k in ("qpos", "qvel", "xpos", "xquat")}
                if n_frames_per_clip is None and "n_frames_per_clip" in f:
                    n_frames_per_clip = int(f["n_frames_per_clip"])
    else:
        data = {k: np.asarray(source[k]) for k in ("qpos", "qvel", "xpos", "xquat")}
    if n_frames_per_clip is None:
        raise ValueError("n_frames_per_clip is required (the reference reads it from the HDF5 config)")
    L_ = int(n_frames_per_clip)
    n = data["qpos"].shape[0]
    if n % L_ or any(data[k].shape[0] != n for k in data):
        raise ValueError("frame count must be a multiple of n_frames_per_clip and equal across qpos / qvel / xpos / xquat")

    def rs(a):
        return np.ascontiguousarray(a.reshape(n // L_, L_, *a.shape[1:]), np.float32)

    qpos, qvel, xpos, xquat = rs(data["qpos"]), rs(data["qvel"]), rs(data["xpos"]), rs(data["xquat"])
    return ReferenceClip(position=qpos[:, :, :3], quaternion=qpos[:, :, 3:7], joints=qpos[:, :, 7:], body_positions=xpos,
                         velocity=qvel[:, :, :3], angular_velocity=qvel[:, :, 3:6], joints_velocity=qvel[:, :, 6:], body_quaternions=xquat)


def to_stac_arrays(clips: ReferenceClip) -> dict[str, np.ndarray]:
    """Inverse of make_multiclip_data: the flat per-frame arrays stac-mjx stores."""
    c, f = clips.position.shape[:2]
    qpos = np.concatenate([clips.position, clips.quaternion, clips.joints], -1).reshape(c * f, -1)
    qvel = np.concatenate([clips.velocity, clips.angular_velocity, clips.joints_velocity], -1).reshape(c * f, -1)
    return {"qpos": qpos, "qvel": qvel, "xpos": clips.body_positions.reshape(c * f, *clips.body_positions.shape[2:]),
            "xquat": clips.body_quaternions.reshape(c * f, *clips.body_quaternions.shape[2:])}


def select_clips(clips: ReferenceClip, indices) -> ReferenceClip:
    """reference `io/load.py:258-278`."""
    idx = np.asarray(indices)
    return ReferenceClip(**{k: getattr(clips, k)[idx] for k in _FIELDS}, original_clip_idx=idx[:, None].astype(np.int32))


def generate_train_test_split(data: ReferenceClip, test_ratio: float = 0.1, rng: np.random.Generator | None = None):
    """reference `io/load.py:187-214` (the reference draws from numpy's global RNG; pass `rng` for a reproducible split)."""
    n = data.position.shape[0]
    indices = np.arange(n)
    chooser = rng.choice if rng is not None else np.random.choice
    test_idx = np.sort(chooser(indices, size=int(n * test_ratio), replace=False))
    train_idx = indices[~np.isin(indices, test_idx)]
    return select_clips(data, train_idx), select_clips(data, test_idx)
