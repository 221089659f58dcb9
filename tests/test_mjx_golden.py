"""Consumer of `tests/golden/mjx_step.npz` -- golden vectors recorded from the REAL reference (jax + mujoco-mjx 3.3.2 +
brax 0.12.3 running track-mjx's own `MultiClipTracking.reset_from_clip / step`) by `tools/dump_mjx_golden.py`.

That file cannot be produced in this image (no jax / mujoco) so, until a maintainer commits it, the three real
tests below SKIP with the reason spelled out and the physics oracle stays "parity unpinned".  The plumbing of the consumer is
still exercised here on every run: `test_consumer_plumbing_with_oracle_standin` writes a file of the same layout from the CPU
oracle into a temp directory and runs the same checks on it (it proves nothing about MJX, only that the consumer works).

Tolerances (BASELINE.json north_star): one physics substep rel 1e-5 per quantity (relative to the quantity's largest
entry; 10 x the oracle's own fp32-vs-fp64 noise where that is larger -- qvel on stiff envs); one 10-substep control step within 10 x the oracle's own fp32-vs-fp64 noise (the 5-iteration CG is unconverged and
amplifies rounding; floor 1e-4); `done`, frame indices, ring-buffer indices bit-exact on every env whose state is finite.
"""
import os

import numpy as np
import pytest

import common
from oracle.oracle import Oracle
from track_mjx_b200 import clips as clipmod
from track_mjx_b200 import config, model_blob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mjx_step.npz")
SKIP = ("tests/golden/mjx_step.npz is absent: it has to be recorded from the real reference with "
        "`JAX_PLATFORMS=cpu python tools/dump_mjx_golden.py` (needs jax + mujoco-mjx 3.3.2 + brax 0.12.3 + track-mjx, none of which "
        "exist in this image).  PHYSICS PARITY AGAINST MJX IS UNPINNED until that file is committed.")

CONFIGS = {
    "cg5": dict(solver="cg", iterations=5, ls_iterations=5, physics_steps_per_control_step=10),
    "cg5_sub1": dict(solver="cg", iterations=5, ls_iterations=5, physics_steps_per_control_step=1),
    "newton10_x20": dict(solver="newton", iterations=10, ls_iterations=10, physics_steps_per_control_step=20),
}
FIELDS = ("position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity", "body_quaternions")
PHYS = ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "xquat", "qfrc_actuator")


def _cfg(walker, name):
    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    args.update(CONFIGS[name])
    return config.make_task_config(walker, config.RewardConfig(), **args)


def _clips(g):
    return clipmod.ReferenceClip(**{k: np.asarray(g[f"clip/{k}"], np.float32) for k in FIELDS})


def _pre_state(g, name, t):
    """The State the reference had BEFORE control step t (reset state for t = 0), as buffer values."""
    src = (lambda k: g[f"{name}/reset/{k}"]) if t == 0 else (lambda k: g[f"{name}/step/{k}"][t - 1])
    n = g["reset/clip_idx"].shape[0]
    st = {k: np.asarray(src(k), np.float32).reshape(n, -1) for k in PHYS + ("time", "prev_ctrl", "action_buffer")}
    st["buffer_index"] = np.asarray(src("buffer_index")).reshape(n, 1).astype(np.int32)
    st["clip_idx"] = np.asarray(g["reset/clip_idx"]).reshape(n, 1).astype(np.int32)
    st["start_frame"] = np.asarray(g["reset/start_frame"]).reshape(n, 1).astype(np.int32)
    return st


def check_model_constants(g, walker):
    """our MJCF-subset compiler (torque rewrite + 0.9 rescale) against MuJoCo's `MjSpec.compile()`."""
    u = model_blob.unpack(walker.blob)
    for k in ("body_parentid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "jnt_type", "jnt_qposadr", "jnt_dofadr",
              "jnt_bodyid", "dof_bodyid", "dof_jntid", "dof_parentid"):
        assert np.array_equal(np.asarray(u[k]).ravel(), np.asarray(g[f"model/{k}"]).ravel()), k
    for k, tol in (("body_pos", 1e-6), ("body_quat", 1e-6), ("body_ipos", 1e-5), ("body_iquat", 1e-5), ("body_mass", 1e-5),
                   ("body_inertia", 1e-5), ("jnt_pos", 1e-6), ("jnt_axis", 1e-6), ("jnt_range", 1e-6), ("jnt_stiffness", 1e-6),
                   ("dof_armature", 1e-6), ("dof_damping", 1e-6), ("qpos0", 1e-6), ("qpos_spring", 1e-6), ("body_invweight0", 1e-4),
                   ("dof_invweight0", 1e-4)):
        a, b = np.asarray(u[k], np.float64).ravel(), np.asarray(g[f"model/{k}"], np.float64).ravel()
        if k == "body_iquat":       # q and -q are the same frame; principal axes may also be permuted only when inertias tie
            a = a.reshape(-1, 4) * np.sign((a.reshape(-1, 4) * b.reshape(-1, 4)).sum(1, keepdims=True) + 1e-30)
            a = a.ravel()
        assert np.allclose(a, b, rtol=tol, atol=tol * max(1e-30, np.abs(b).max())), k
    assert np.allclose(np.asarray(u["actuator_gain"]), np.asarray(g["model/actuator_gainprm"])[:, 0], rtol=1e-6)
    assert np.isclose(float(u["opt"][7]), float(g["model/stat_meaninertia"]), rtol=1e-4)
    assert list(np.asarray(g["model/joint_idxs"])) == list(walker.joint_idxs)
    assert list(np.asarray(g["model/body_idxs"])) == list(walker.body_idxs)
    assert list(np.asarray(g["model/endeff_idxs"])) == list(walker.endeff_idxs)
    assert int(g["model/torso_idx"]) == walker.torso_idx


def replay(make_runner, g, walker, name):
    """Feed every recorded pre-step State + action to the implementation under test and compare with the recorded post-step State.
    `make_runner(cfg, clips, n, dtype)` returns (buf, forward_fn, step_fn, get_fn)."""
    cfg, clips = _cfg(walker, name), _clips(g)
    n = g["reset/clip_idx"].shape[0]
    sub1 = CONFIGS[name]["physics_steps_per_control_step"] == 1
    buf, _fwd, step, get = make_runner(cfg, clips, n, np.float32)
    o32, o64 = Oracle(walker.blob, cfg, clips, dtype=np.float32), Oracle(walker.blob, cfg, clips, dtype=np.float64)
    a, b = o32.alloc(n, debug=False), o64.alloc(n, debug=False)
    T = g["step/actions"].shape[0]
    worst = {}
    for t in range(T):
        st = _pre_state(g, name, t)
        act = np.asarray(g["step/actions"][t], np.float32)
        finite_in = np.isfinite(st["qpos"]).all(1) & np.isfinite(st["qvel"]).all(1) & (np.abs(st["qvel"]).max(1) < 1e4)
        common.put(buf, st); common.put(a, st); common.put(b, st)
        step(act); o32.step(a, act); o64.step(b, act)
        out = get()
        gold = {k: np.asarray(g[f"{name}/step/{k}"][t]).reshape(n, -1) for k in PHYS + ("obs", "reward", "done", "cur_frame", "buffer_index")}
        ok = finite_in & np.isfinite(gold["qpos"]).all(1) & (np.abs(gold["qvel"]).max(1) < 1e4)
        assert ok.sum() >= n // 4, f"step {t}: too few finite envs in the golden file"
        for k in ("qpos", "qvel", "act", "xpos", "xquat", "qfrc_actuator", "obs", "reward"):
            e = common.err(out[k][ok], gold[k][ok])[1]
            noise = common.err(a[k][ok], b[k][ok])[1]
            tol = max(1e-5, 10.0 * noise) if sub1 else max(10.0 * noise, 1e-4)
            worst[k] = max(worst.get(k, 0.0), e)
            assert e <= tol, f"{name} step {t} {k}: rel err vs MJX {e:.3e} > {tol:.3e} (oracle fp32-vs-fp64 noise {noise:.3e})"
        assert (out["cur_frame"][ok] == gold["cur_frame"][ok]).all(), f"{name} step {t}: cur_frame"
        assert (out["buffer_index"][ok] == gold["buffer_index"][ok]).all(), f"{name} step {t}: buffer_index"
        agree = ok & (a["done"][:, 0] == b["done"][:, 0])          # envs on which the flag is well defined at fp32 resolution
        assert (out["done"][agree] == gold["done"][agree]).all(), f"{name} step {t}: done"
    return worst


def oracle_runner(walker):
    def make(cfg, clips, n, dtype):
        o = Oracle(walker.blob, cfg, clips, dtype=dtype)
        buf = o.alloc(n, debug=False)
        return buf, (lambda: o.forward(buf)), (lambda act: o.step(buf, act)), (lambda: common.get(buf))
    return make


def gpu_runner(walker):
    import torch

    from track_mjx_b200.env import Stepper

    def make(cfg, clips, n, dtype):
        s = Stepper(walker.blob, cfg, clips, n, 0)
        return s.buf, s.forward, (lambda act: s.step(torch.from_numpy(np.ascontiguousarray(act)).cuda())), (lambda: common.get(s.buf))
    return make


def _load():
    if not os.path.exists(GOLD):
        pytest.skip(SKIP)
    return np.load(GOLD)


def test_model_constants_match_mujoco_compile(walker):
    check_model_constants(_load(), walker)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_matches_mjx_step(walker, name):
    replay(oracle_runner(walker), _load(), walker, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CONFIGS))
def test_cuda_matches_mjx_step(walker, name):
    replay(gpu_runner(walker), _load(), walker, name)


# ----------------------------------------------------------------------------------------------- plumbing self-test
def _write_standin(path, walker, n=12, T=3):
    """Same keys / shapes as tools/dump_mjx_golden.py writes, but produced by the fp32 CPU oracle (NOT a golden file)."""
    rng = np.random.default_rng(0)
    table = clipmod.make_synthetic_clips(walker.sections, 2)
    out = {f"clip/{k}": np.asarray(getattr(table, k), np.float32) for k in FIELDS}
    u = model_blob.unpack(walker.blob)
    for k in ("body_parentid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid",
              "dof_bodyid", "dof_jntid", "dof_parentid", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "jnt_pos",
              "jnt_axis", "jnt_range", "jnt_stiffness", "dof_armature", "dof_damping", "qpos0", "qpos_spring", "body_invweight0", "dof_invweight0"):
        out[f"model/{k}"] = np.asarray(u[k])
    out["model/actuator_gainprm"] = np.concatenate([np.asarray(u["actuator_gain"])[:, None], np.zeros((walker.nu, 9), np.float32)], 1)
    out["model/stat_meaninertia"] = np.asarray(u["opt"][7])
    out["model/joint_idxs"], out["model/body_idxs"], out["model/endeff_idxs"] = map(np.asarray, (walker.joint_idxs, walker.body_idxs, walker.endeff_idxs))
    out["model/torso_idx"] = np.asarray(walker.torso_idx)
    out["reset/clip_idx"] = rng.integers(0, 2, n).astype(np.int32)
    out["reset/start_frame"] = rng.integers(0, 44, n).astype(np.int32)
    out["step/actions"] = (0.05 * rng.normal(size=(T, n, walker.nu))).astype(np.float32)
    for name in CONFIGS:
        o = Oracle(walker.blob, _cfg(walker, name), table, dtype=np.float32)
        buf = o.alloc(n, debug=False)
        init = common.init_buffers(buf, table, seed=1)
        init["clip_idx"], init["start_frame"] = out["reset/clip_idx"][:, None], out["reset/start_frame"][:, None]
        ci, sf = out["reset/clip_idx"], out["reset/start_frame"]
        init["qpos"] = np.concatenate([table.position[ci, sf], table.quaternion[ci, sf], table.joints[ci, sf]], -1).astype(np.float32)
        common.put(buf, init)
        o.forward(buf)
        keys = PHYS + ("time", "obs", "reward", "done", "prev_ctrl", "action_buffer", "buffer_index", "cur_frame")
        for k in keys:
            out[f"{name}/reset/{k}"] = buf[k].copy()
        rec = {k: [] for k in keys}
        for t in range(T):
            o.step(buf, out["step/actions"][t])
            for k in keys:
                rec[k].append(buf[k].copy())
        for k in keys:
            out[f"{name}/step/{k}"] = np.stack(rec[k])
    np.savez(path, **out)


def test_consumer_plumbing_with_oracle_standin(walker, tmp_path):
    path = str(tmp_path / "standin.npz")
    _write_standin(path, walker)
    g = np.load(path)
    check_model_constants(g, walker)
    for name in ("cg5_sub1", "cg5"):
        worst = replay(oracle_runner(walker), g, walker, name)
        assert max(worst.values()) == 0.0          # the oracle reproduces its own recording bit for bit


@pytest.mark.gpu
def test_cuda_consumer_plumbing_with_oracle_standin(walker, tmp_path):
    path = str(tmp_path / "standin.npz")
    _write_standin(path, walker)
    replay(gpu_runner(walker), np.load(path), walker, "cg5_sub1")
