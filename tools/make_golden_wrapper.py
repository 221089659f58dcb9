"""Golden vectors for the auto-reset wrapper, produced by the REFERENCE'S OWN `AutoResetWrapperTracking` text.

`track_mjx/environment/wrappers.py` imports brax / flax / mujoco at module level (not installable here), so the unmodified source
text of the class `AutoResetWrapperTracking` (wrappers.py:277-310) is cut out of the reference file with `ast` at generation time and
executed with `jax.numpy` bound to numpy.  What it wraps is supplied as stand-ins and is therefore INPUT to the vectors, not under
test: the inner env is brax 0.12.3's `EpisodeWrapper` (restated below: steps += 1, done where steps >= episode_length, truncation =
where(steps >= episode_length, 1 - done, 0)) around the CPU oracle's un-wrapped control step; `State` is a small attribute bag with
`.replace`.  Under test: zeroing `steps` where the PREVIOUS state was done, zeroing `done`, `where(done, first_*, current)` over the
pipeline state, the observation and `prev_ctrl` -- and what is NOT restored (`start_frame`, `clip_idx`, `action_buffer`,
`buffer_index`, `steps`).  Only inputs and the outputs the reference computed are stored (`tests/golden/wrapper.npz`).

    python tools/make_golden_wrapper.py        # needs /root/reference; output is committed
"""
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/track_mjx/environment/wrappers.py"

PIPE = ("qpos", "qvel", "act", "time", "qacc_warmstart", "xpos", "xquat", "qfrc_actuator")
INFO = ("clip_idx", "start_frame", "buffer_index", "prev_ctrl", "action_buffer", "steps", "truncation")


class State:
    def __init__(self, pipeline_state, obs, reward, done, metrics, info):
        self.pipeline_state, self.obs, self.reward, self.done, self.metrics, self.info = pipeline_state, obs, reward, done, metrics, info

    def replace(self, **kw):
        d = dict(pipeline_state=self.pipeline_state, obs=self.obs, reward=self.reward, done=self.done, metrics=self.metrics, info=self.info)
        d.update(kw)
        return State(**d)


class Wrapper:
    def __init__(self, env):
        self.env = env


def reference_wrapper_class():
    tree = ast.parse(open(REF).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "AutoResetWrapperTracking")
    for fn in node.body:                       # annotations name jax / brax types
        if isinstance(fn, ast.FunctionDef):
            fn.returns = None
            for a in fn.args.args:
                a.annotation = None
    jp = types.SimpleNamespace(where=lambda c, x, y: np.where(c != 0, x, y), zeros_like=np.zeros_like, reshape=np.reshape)
    jax = types.SimpleNamespace(tree=types.SimpleNamespace(map=lambda f, a, b: {k: f(a[k], b[k]) for k in a}), Array=np.ndarray)
    ns = {"Wrapper": Wrapper, "State": State, "jp": jp, "jax": jax}
    exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns["AutoResetWrapperTracking"]


class EpisodeEnv:
    """brax EpisodeWrapper (action_repeat = 1) around the oracle's un-wrapped step -- the INNER env of the reference's wrapper stack."""

    def __init__(self, oracle, buf, episode_length):
        self.o, self.buf, self.L = oracle, buf, float(episode_length)

    def load(self, state):
        import common

        common.put(self.buf, {**state.pipeline_state, **{k: state.info[k] for k in INFO if k in state.info}, "obs": state.obs, "done": state.done})

    def view(self):
        b = self.buf
        info = {k: b[k].copy() for k in INFO}
        return State({k: b[k].copy() for k in PIPE}, b["obs"].copy(), b["reward"].copy(), b["done"].copy(), b["metrics"].copy(), info)

    def step(self, state, action):
        self.load(state)
        first = {k: v for k, v in state.info.items() if k.startswith("first_")}
        self.o.step(self.buf, action, 0)
        s = self.view()
        steps = state.info["steps"] + 1
        one, zero = np.ones_like(s.done), np.zeros_like(s.done)
        done = np.where(steps >= self.L, one, s.done)
        s.info["truncation"] = np.where(steps >= self.L, 1 - s.done, zero)
        s.info["steps"] = steps
        s.info.update(first)
        return s.replace(done=done)


def main():
    import common
    from oracle.oracle import Oracle
    from track_mjx_b200 import _lib as L
    from track_mjx_b200 import clips as clipmod, config
    from track_mjx_b200.walker import Rodent

    w = Rodent(torque_actuators=True, rescale_factor=0.9)
    cl = clipmod.make_synthetic_clips(w.sections, 2)
    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    args["physics_steps_per_control_step"] = 5
    cfg = config.make_task_config(w, config.RewardConfig(), **args)
    cfg.episode_length = 5                                    # truncation is reached inside the recorded steps
    n, T = 12, 8
    o = Oracle(w.blob, cfg, cl, dtype=np.float32)
    buf = o.alloc(n, debug=False)
    common.put(buf, common.init_buffers(buf, cl, seed=9))
    o.forward(buf, L.TMJX_F_SNAPSHOT)
    env = reference_wrapper_class()(EpisodeEnv(o, buf, cfg.episode_length))
    state = env.env.view()
    state.info["first_pipeline_state"] = {k: buf["first_" + k].copy() for k in PIPE}
    state.info["first_obs"], state.info["first_prev_ctrl"] = buf["first_obs"].copy(), buf["first_prev_ctrl"].copy()
    rng = np.random.default_rng(3)
    out = {"episode_length": np.int32(cfg.episode_length), "n_frames": np.int32(5)}
    for k in PIPE:
        out[f"first_{k}"] = buf["first_" + k].copy()
    out["first_obs"], out["first_prev_ctrl"] = buf["first_obs"].copy(), buf["first_prev_ctrl"].copy()
    rec = {}
    for t in range(T):
        action = (rng.normal(size=(n, w.nu)) * np.where(np.arange(n) % 2, 1.0, 0.05)[:, None]).astype(np.float32)   # odd envs: violent actions -> early termination
        pre = {**{k: state.pipeline_state[k] for k in PIPE}, **{k: state.info[k] for k in INFO}, "obs": state.obs, "done": state.done}
        state = env.step(state, action)
        post = {**{k: state.pipeline_state[k] for k in PIPE}, **{k: state.info[k] for k in INFO}, "obs": state.obs, "done": state.done,
                "reward": state.reward, "metrics": state.metrics}
        for k, v in pre.items():
            rec.setdefault("in_" + k, []).append(np.array(v))
        for k, v in post.items():
            rec.setdefault("out_" + k, []).append(np.array(v))
        rec.setdefault("action", []).append(action)
    for k, v in rec.items():
        out[k] = np.stack(v)
    d = out["out_done"]
    print("done per step:", d.reshape(T, -1).sum(1), " truncation per step:", out["out_truncation"].reshape(T, -1).sum(1))
    path = os.path.join(ROOT, "tests", "golden", "wrapper.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
