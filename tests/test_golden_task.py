"""Task-layer parity against golden vectors produced by the REFERENCE'S OWN CODE (tools/make_golden_task.py runs
the unmodified track_mjx task modules -- single/multi_clip_tracking.py, reward.py, walker/base.py -- on numpy).

The physics inside these steps is the oracle's own (the reference's MJX cannot run here), so what is pinned is every
line of reward / termination / observation / frame-index / ring-buffer arithmetic."""
import os

import numpy as np
import pytest

import common
from oracle.oracle import Oracle
from track_mjx_b200 import _lib as L
from track_mjx_b200 import clips as clipmod, config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "task_layer.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def clips3(walker):
    return clipmod.make_synthetic_clips(walker.sections, 3)


def _oracle(walker, clips3, task_cfg):
    return Oracle(walker.blob, task_cfg, clips3, dtype=np.float32)


def check_outputs(out, gold, s, exact_physics=True):
    tol = dict(rtol=2e-5, atol=2e-6)
    assert np.allclose(out["obs"], gold["ref_obs"][s], **tol)
    assert np.allclose(out["reward"][:, 0], gold["ref_reward"][s], **tol)
    mt = out["metrics"]
    names = config.METRIC_NAMES
    flags = [names.index(k) for k in ("done", "too_far", "bad_pose", "bad_quat", "fall", "nan")]
    assert (mt[:, flags] == gold["ref_metrics"][s][:, flags]).all()          # termination flags: bit-exact
    assert np.allclose(mt, gold["ref_metrics"][s], rtol=1e-4, atol=2e-6, equal_nan=True)      # NaN envs: NaN in the same places
    assert (out["done"][:, 0] == gold["ref_done"][s]).all()
    assert (out["cur_frame"][:, 0] == gold["ref_cur_frame"][s]).all()        # frame indices: bit-exact
    assert (out["buffer_index"][:, 0] == gold["ref_buffer_index"][s]).all()
    assert (out["action_buffer"] == gold["ref_action_buffer"][s]).all()
    assert (out["prev_ctrl"] == gold["ref_prev_ctrl"][s]).all()


def test_reset_obs_matches_reference_code(walker, clips3, task_cfg, gold):
    o = _oracle(walker, clips3, task_cfg)
    n = gold["reset_in_qpos"].shape[0]
    buf = o.alloc(n, debug=False)
    common.put(buf, dict(qpos=gold["reset_in_qpos"], qvel=gold["reset_in_qvel"], clip_idx=gold["clip_idx"], start_frame=gold["start_frame"]))
    o.forward(buf, L.TMJX_F_SNAPSHOT)
    assert buf["obs"].shape[1] == 696
    assert np.allclose(buf["obs"], gold["reset_obs"], rtol=2e-5, atol=2e-6)


def test_step_epilogue_matches_reference_code(walker, clips3, task_cfg, gold):
    o = _oracle(walker, clips3, task_cfg)
    n = gold["actions"].shape[1]
    buf = o.alloc(n, debug=False)
    common.put(buf, {k[3:]: gold[k] for k in gold.files if k.startswith("s0_")})
    n_done = 0
    for s in range(gold["actions"].shape[0]):
        o.step(buf, gold["actions"][s])
        # the physics is the oracle's own and deterministic: the recorded post-physics state is reproduced
        for k in ("qpos", "qvel", "xpos", "xquat", "qfrc_actuator", "time"):
            assert np.array_equal(buf[k], gold[f"post_{k}"][s], equal_nan=True), k
        check_outputs(buf, gold, s)
        n_done += int(buf["done"].sum())
    assert n_done > 10          # terminating cases are covered


@pytest.mark.gpu
def test_cuda_epilogue_matches_reference_code(walker, clips3, task_cfg, gold):
    """Same golden vectors through the C ABI: feed the recorded pre-step state, run ONE control step on the GPU and
    compare the task-layer outputs for the steps whose physics is well conditioned (first two: free fall)."""
    import torch

    from track_mjx_b200.env import Stepper

    n = gold["actions"].shape[1]
    g = Stepper(walker.blob, task_cfg, clips3, n, 0)
    common.put(g.buf, {k[3:]: gold[k] for k in gold.files if k.startswith("s0_") and k[3:] in g.buf})
    for s in range(2):
        g.step(torch.from_numpy(gold["actions"][s]).cuda())
        out = common.get(g.buf)
        assert np.allclose(out["obs"], gold["ref_obs"][s], rtol=1e-3, atol=2e-4)
        assert np.allclose(out["reward"][:, 0], gold["ref_reward"][s], rtol=1e-3, atol=1e-4)
        assert (out["done"][:, 0] == gold["ref_done"][s]).all()
        assert (out["cur_frame"][:, 0] == gold["ref_cur_frame"][s]).all()
        assert (out["action_buffer"] == gold["ref_action_buffer"][s]).all()
    g.close()
