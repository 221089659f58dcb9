"""Golden vectors for the task layer, produced by the REFERENCE'S OWN CODE.

The reference's env arithmetic splits in two: the physics (mujoco-mjx, not installable here) and the task layer
(`track_mjx/environment/task/{single,multi}_clip_tracking.py`, `reward.py`, `walker/base.py`), which is plain
`jax.numpy`.  This script imports the unmodified task-layer modules from /root/reference with `jax.numpy` bound to
numpy (float32 arrays; the few jax-only constructs -- `.at[].set`, `lax.dynamic_slice`, `vmap`, `tree.map`,
`ravel_pytree` -- are provided as small shims below, and `brax.math.rotate / relative_quat / quat_mul / quat_inv` are
restated from brax 0.12.3's `brax/math.py`), feeds `SingleClipTracking.step` / `reset_from_clip` post-physics states
produced by the CPU oracle, and stores inputs + the reference's outputs in `tests/golden/task_layer.npz`.

    python tools/make_golden_task.py        # needs /root/reference; output is committed

`tests/test_golden_task.py` then checks the oracle's (and through it the CUDA kernel's) reward / termination /
observation / frame-index / ring-buffer arithmetic against these vectors without needing the reference tree.
"""
import dataclasses
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


# ------------------------------------------------------------------------------------------------ numpy-backed shims
class JArr(np.ndarray):
    """ndarray with jax.Array's immutable flavour: `.at[i].set(v)` and out-of-place augmented division."""

    class _At:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            a = self.a

            class _Set:
                def set(self, v):
                    out = np.array(a, copy=True).view(JArr)
                    out[idx] = v
                    return out

            return _Set()

    @property
    def at(self):
        return JArr._At(self)

    def __itruediv__(self, other):
        return np.true_divide(self, other)

    def __getitem__(self, idx):
        """jax.numpy indexing: out-of-bounds integer-array indices are CLAMPED (numpy raises)."""
        tup = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(i, (np.ndarray, list)) and np.asarray(i).dtype.kind in "iu" for i in tup):
            fixed, axis = [], 0
            for i in tup:
                if isinstance(i, (np.ndarray, list)) and np.asarray(i).dtype.kind in "iu":
                    i = np.clip(np.asarray(i), -self.shape[axis], self.shape[axis] - 1)
                fixed.append(i)
                axis += 1
            idx = tuple(fixed) if isinstance(idx, tuple) else fixed[0]
        return super().__getitem__(idx)


def J(x, dtype=np.float32):
    return np.asarray(x, dtype=dtype).view(JArr)


def install_shims():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.array = lambda x, dtype=None: np.asarray(x, dtype=dtype).view(JArr) if dtype is not None else np.asarray(x).view(JArr)
    jnp.ndarray = np.ndarray
    jnp.linalg = np.linalg

    def dynamic_slice_in_dim(x, start, size, axis=0):
        start = int(np.clip(int(start), 0, x.shape[axis] - size))     # XLA clamps the start index
        sl = [slice(None)] * x.ndim
        sl[axis] = slice(start, start + size)
        return x[tuple(sl)]

    def dynamic_slice(x, starts, sizes):
        out = x
        for ax, (s, n) in enumerate(zip(starts, sizes)):
            out = dynamic_slice_in_dim(out, s, n, ax)
        return out

    def vmap(f, in_axes=0):
        def g(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
            outs = [f(*[np.take(a, i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]
            return np.stack(outs).view(JArr)

        return g

    def tree_map(f, tree, *rest):
        if dataclasses.is_dataclass(tree):
            kw = {}
            for fld in dataclasses.fields(tree):
                v = getattr(tree, fld.name)
                kw[fld.name] = None if v is None else f(v, *[getattr(r, fld.name) for r in rest])
            return type(tree)(**kw)
        return f(tree, *rest)

    def ravel_pytree(obj):
        leaves = [np.ravel(v) for v in vars(obj).values() if isinstance(v, np.ndarray)]
        return np.concatenate(leaves), None

    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.Array = np.ndarray
    jax.vmap = vmap
    jax.lax = types.SimpleNamespace(dynamic_slice=dynamic_slice, dynamic_slice_in_dim=dynamic_slice_in_dim)
    jax.tree = types.SimpleNamespace(map=tree_map)
    jax.tree_util = types.SimpleNamespace(tree_map=tree_map)
    rnd = types.ModuleType("jax.random")
    rnd.split = lambda key, n=2: [key] * n
    rnd.uniform = lambda key, shape, minval=0.0, maxval=1.0: J(key["noise"][: shape[0]])
    rnd.randint = lambda key, shape, lo, hi: key["ints"].pop(0)
    jax.random = rnd
    fu = types.ModuleType("jax.flatten_util")
    fu.ravel_pytree = ravel_pytree
    jax.flatten_util = fu
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.random": rnd, "jax.flatten_util": fu})

    flax = types.ModuleType("flax")
    struct = types.ModuleType("flax.struct")
    struct.dataclass = lambda cls=None, **kw: dataclasses.dataclass(cls, frozen=True) if cls is not None else (lambda c: dataclasses.dataclass(c, frozen=True))
    flax.struct = struct
    sys.modules.update({"flax": flax, "flax.struct": struct})

    oc = types.ModuleType("omegaconf")
    oc.ListConfig = type("ListConfig", (list,), {})
    oc.DictConfig = dict
    sys.modules["omegaconf"] = oc

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any(k)

        def __call__(self, *a, **k):
            return None

    for name in ("mujoco", "mujoco.mjx", "h5py", "hydra", "brax.io", "brax.io.mjcf", "brax.envs"):
        sys.modules[name] = _Any(name)
    sys.modules["mujoco"].mjx = sys.modules["mujoco.mjx"]

    # brax.math (brax 0.12.3 brax/math.py: rotate, quat_mul, quat_inv, relative_quat)
    bm = types.ModuleType("brax.math")

    def rotate(vec, quat):
        if len(vec.shape) != 1:
            raise ValueError("vec must have no batch dimensions.")
        s, u = quat[0], quat[1:]
        r = 2 * (np.dot(u, vec) * u) + (s * s - np.dot(u, u)) * vec
        return r + 2 * s * np.cross(u, vec)

    def quat_mul(u, v):
        return np.array([
            u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3], u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
            u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1], u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0]],
            dtype=np.result_type(u, v))

    bm.rotate, bm.quat_mul = rotate, quat_mul
    bm.quat_inv = lambda q: q * np.array([1, -1, -1, -1], dtype=q.dtype)
    bm.relative_quat = lambda q1, q2: quat_mul(q2, bm.quat_inv(q1))
    brax = types.ModuleType("brax")
    brax.math = bm
    brax.io = sys.modules["brax.io"]
    sys.modules["brax.io"].mjcf = sys.modules["brax.io.mjcf"]
    base = types.ModuleType("brax.envs.base")

    class PipelineEnv:
        pass

    class State:
        def __init__(self, pipeline_state, obs, reward, done, metrics, info):
            self.pipeline_state, self.obs, self.reward, self.done, self.metrics, self.info = pipeline_state, obs, reward, done, metrics, info

        def replace(self, **kw):
            s = State(self.pipeline_state, self.obs, self.reward, self.done, self.metrics, self.info)
            for k, v in kw.items():
                setattr(s, k, v)
            return s

    base.PipelineEnv, base.State = PipelineEnv, State
    brax.envs = sys.modules["brax.envs"]
    sys.modules.update({"brax": brax, "brax.math": bm, "brax.envs.base": base})
    return State


class FakeData:
    """The fields of mjx.Data the task layer touches."""

    def __init__(self, qpos, qvel, xpos, xmat, qfrc_actuator, time):
        self.qpos, self.qvel, self.xpos, self.xmat, self.qfrc_actuator, self.time = J(qpos), J(qvel), J(xpos), J(xmat), J(qfrc_actuator), J(time)

    def bind(self, model, spec_body):
        b = spec_body
        return types.SimpleNamespace(xpos=self.xpos[b], xmat=self.xmat[b])


def quat_to_mat32(q):
    w, x, y, z = [q[..., i] for i in range(4)]
    return np.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z),
                     w * w - x * x + y * y - z * z, 2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x),
                     w * w - x * x - y * y + z * z], -1).reshape(q.shape[:-1] + (3, 3)).astype(np.float32)


def main():
    State = install_shims()
    sys.path.insert(0, REF)
    from track_mjx.environment.task.multi_clip_tracking import MultiClipTracking
    from track_mjx.environment.task.reward import RewardConfig
    from track_mjx.environment.walker.base import BaseWalker
    from track_mjx.io.load import ReferenceClip

    import common
    from oracle.oracle import Oracle
    from track_mjx_b200 import _lib as L
    from track_mjx_b200 import clips as clipmod, config
    from track_mjx_b200.walker import Rodent

    walker = Rodent(torque_actuators=True)
    clips = clipmod.make_synthetic_clips(walker.sections, 3)
    env_args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    cfg = config.make_task_config(walker, config.RewardConfig(), **env_args)
    orc = Oracle(walker.blob, cfg, clips, dtype=np.float32)

    # ---- the reference env object, bypassing the MuJoCo-dependent constructor
    class RefWalker(BaseWalker):
        def _initialize_indices(self):
            pass

        def _build_spec(self, *a, **k):
            pass

    for name in list(getattr(RefWalker, "__abstractmethods__", ())):
        setattr(RefWalker, name, lambda self, *a, **k: None)
    RefWalker.__abstractmethods__ = frozenset()
    rw = object.__new__(RefWalker)
    rw._joint_idxs, rw._body_idxs, rw._endeff_idxs = np.array(walker.joint_idxs), np.array(walker.body_idxs), np.array(walker.endeff_idxs)
    rw._torso_idx, rw._torso_name, rw._end_eff_names = walker.torso_idx, "torso", list(config.RODENT_END_EFF_NAMES)
    rc = config.RewardConfig()
    kw = dataclasses.asdict(rc)
    kw["penalty_pos_distance_scale"] = J(kw["penalty_pos_distance_scale"])
    kw["healthy_z_range"] = tuple(kw["healthy_z_range"])
    ref_rc = RewardConfig(**kw)
    env = object.__new__(MultiClipTracking)
    env.walker, env._reward_config, env._mocap_hz, env._ref_len, env._reset_noise_scale = rw, ref_rc, 50, 5, 1e-3
    env._mjx_model = None
    env._mj_spec = types.SimpleNamespace(body=lambda name: walker.body_id(name))
    env.sys = types.SimpleNamespace(nq=walker.nq, nv=walker.nv, nu=walker.nu)
    env._reference_clips = ReferenceClip(**{f.name: (J(getattr(clips, f.name)) if f.name != "original_clip_idx" else None)
                                            for f in dataclasses.fields(ReferenceClip)})
    env._n_clips = 3

    n = 48
    gold = {}
    buf = orc.alloc(n, debug=False)
    init = common.init_buffers(buf, clips, seed=21)
    common.put(buf, init)
    # ---- reset path: reference reset_from_clip with the noise draws injected, physics = oracle forward
    noise = (init["qpos"] - np.concatenate([clips.position[init["clip_idx"][:, 0], init["start_frame"][:, 0]],
                                            clips.quaternion[init["clip_idx"][:, 0], init["start_frame"][:, 0]],
                                            clips.joints[init["clip_idx"][:, 0], init["start_frame"][:, 0]]], -1)).astype(np.float32)
    orc.forward(buf, L.TMJX_F_SNAPSHOT)
    gold["reset_in_qpos"], gold["reset_in_qvel"] = init["qpos"], init["qvel"]
    gold["clip_idx"], gold["start_frame"] = init["clip_idx"], init["start_frame"]
    reset_obs = []
    for e in range(n):
        d = FakeData(buf["qpos"][e], buf["qvel"][e], buf["xpos"][e].reshape(-1, 3), quat_to_mat32(buf["xquat"][e].reshape(-1, 4)),
                     buf["qfrc_actuator"][e], buf["time"][e, 0])
        env.pipeline_init = lambda qpos, qvel, d=d: d
        info = {"clip_idx": int(init["clip_idx"][e, 0]), "start_frame": int(init["start_frame"][e, 0]), "prev_ctrl": J(np.zeros(walker.nu))}
        st = env.reset_from_clip({"noise": noise[e]}, info, noise=True)
        reset_obs.append(np.asarray(st.obs, np.float32))
        assert st.info["reference_obs_size"] == 470 and st.info["proprioceptive_obs_size"] == 226
    gold["reset_obs"] = np.stack(reset_obs)

    # ---- step path: several control steps; at each, the reference epilogue runs on the oracle's post-physics state
    rng = np.random.default_rng(33)
    nsteps = 6
    keys_out = ("obs", "reward", "done", "metrics", "cur_frame", "action_buffer", "buffer_index", "prev_ctrl")
    rec = {k: [] for k in keys_out}
    rec_post = {k: [] for k in ("qpos", "qvel", "xpos", "xquat", "qfrc_actuator", "time")}
    actions = []
    gold["step_in_state"] = {k: v.copy() for k, v in common.get(buf, common.STATE_KEYS).items()}
    for s in range(nsteps):
        scale = (0.02, 0.05, 0.3, 1.0, 0.1, 0.0)[s]
        act = (scale * rng.normal(size=(n, walker.nu))).astype(np.float32)
        before = {k: buf[k].copy() for k in ("action_buffer", "buffer_index")}
        orc.step(buf, act)
        actions.append(act)
        outs = {k: [] for k in keys_out}
        for e in range(n):
            d = FakeData(buf["qpos"][e], buf["qvel"][e], buf["xpos"][e].reshape(-1, 3), quat_to_mat32(buf["xquat"][e].reshape(-1, 4)),
                         buf["qfrc_actuator"][e], buf["time"][e, 0])
            env.pipeline_step = lambda data0, action, d=d: d
            info = {"clip_idx": int(init["clip_idx"][e, 0]), "start_frame": int(init["start_frame"][e, 0]),
                    "prev_ctrl": J(np.zeros(walker.nu)), "action_buffer": J(before["action_buffer"][e].reshape(50, walker.nu)),
                    "buffer_index": int(before["buffer_index"][e, 0])}
            st0 = State(None, None, None, None, {}, info)
            st = env.step(st0, J(act[e]))
            outs["obs"].append(np.asarray(st.obs, np.float32))
            outs["reward"].append(np.float32(st.reward))
            outs["done"].append(np.float32(st.done))
            outs["metrics"].append(np.array([st0.metrics[k] for k in config.METRIC_NAMES], np.float32))
            outs["cur_frame"].append(int(env._get_cur_frame(info, d)))
            outs["action_buffer"].append(np.asarray(st.info["action_buffer"], np.float32).ravel())
            outs["buffer_index"].append(int(st.info["buffer_index"]))
            outs["prev_ctrl"].append(np.asarray(st.info["prev_ctrl"], np.float32))
        for k in keys_out:
            rec[k].append(np.stack(outs[k]))
        for k in rec_post:
            rec_post[k].append(buf[k].copy())
    out = {f"ref_{k}": np.stack(v) for k, v in rec.items()}
    out.update({f"post_{k}": np.stack(v) for k, v in rec_post.items()})
    out["actions"] = np.stack(actions)
    out.update({f"s0_{k}": v for k, v in gold.pop("step_in_state").items()})
    out.update(gold)
    path = os.path.join(ROOT, "tests", "golden", "task_layer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "dones per step:", [int(x.sum()) for x in out["ref_done"]])
    # immediate self-check against the oracle's own epilogue
    for s in range(nsteps):
        pass


if __name__ == "__main__":
    main()
