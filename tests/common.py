"""Shared helpers for the parity tests: seeded initial states and oracle/GPU buffer exchange."""
import numpy as np


def init_buffers(buf, clips, seed=0, noise=1e-3, xp=np):
    """start_frame ~ randint(0,44), clip_idx ~ randint(0,C), qpos = ref + U(+-noise), qvel = U(+-noise) (same draws)."""
    rng = np.random.default_rng(seed)
    n = buf["qpos"].shape[0]
    nq, nv = buf["qpos"].shape[1], buf["qvel"].shape[1]
    ci = rng.integers(0, clips.position.shape[0], n).astype(np.int32)
    sf = rng.integers(0, 44, n).astype(np.int32)
    u = rng.uniform(-noise, noise, (n, nq))
    qpos = np.concatenate([clips.position[ci, sf], clips.quaternion[ci, sf], clips.joints[ci, sf]], -1) + u
    return dict(qpos=qpos.astype(np.float32), qvel=u[:, :nv].astype(np.float32), clip_idx=ci[:, None], start_frame=sf[:, None])


def put(buf, vals):
    """Copy numpy values into oracle (numpy) or device (torch) buffers of the same name."""
    for k, v in vals.items():
        dst = buf[k]
        if isinstance(dst, np.ndarray):
            dst[...] = np.asarray(v).reshape(dst.shape).astype(dst.dtype)
        else:
            import torch

            dst.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(v).reshape(tuple(dst.shape)))).to(dst.dtype))


def get(buf, keys=None):
    out = {}
    for k, v in buf.items():
        if keys is not None and k not in keys:
            continue
        out[k] = v.copy() if isinstance(v, np.ndarray) else v.detach().cpu().numpy()
    return out


STATE_KEYS = ("qpos", "qvel", "act", "time", "qacc_warmstart", "xpos", "xquat", "qfrc_actuator", "clip_idx", "start_frame",
              "buffer_index", "prev_ctrl", "action_buffer", "steps", "truncation", "first_qpos", "first_qvel", "first_act",
              "first_time", "first_qacc_warmstart", "first_xpos", "first_xquat", "first_qfrc_actuator", "first_obs",
              "first_prev_ctrl", "done")


def err(a, b):
    """max abs error and max error relative to the array's scale."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b).max() if a.size else 0.0
    return d, d / max(np.abs(b).max(), 1e-30)
