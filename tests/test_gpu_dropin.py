"""Drop-in surface of the env mirror as the reference's callers use it: `wrappers.wrap`, the rollout loop of
`track_mjx/agent/wandb_logging.py:67-136`, `RenderRolloutWrapperTracking` (`environment/wrappers.py:353-377`)."""
import numpy as np
import pytest
import torch

from track_mjx_b200 import config
from track_mjx_b200.env import MultiClipTracking, wrap

pytestmark = pytest.mark.gpu


def make_env(walker, clips, n, **kw):
    return MultiClipTracking(clips, walker, config.RewardConfig(), num_envs=n, **dict(config.DEFAULT_ENV_ARGS, **kw))


def test_wandb_logging_rollout_call_pattern(walker, clips2):
    """reset -> loop(obs -> ctrl -> step) -> per-state metrics / qpos -> env._get_reference_clip(rollout[0].info) -> hstack."""
    env = wrap(make_env(walker, clips2, 4))
    state = env.reset(3)
    rollout = [state.clone()]                       # the mirror's State aliases live buffers: keep copies like a functional caller would
    episode_length = int(250 * env._steps_for_cur_frame)
    assert episode_length == 250
    for _ in range(6):
        ctrl = 0.05 * torch.randn(4, env.action_size, device=env.device)
        state = env.step(state, ctrl)
        rollout.append(state.clone())
    for name in ("pos_reward", "quat_reward", "joint_reward", "endeff_reward", "too_far", "fall", "summed_pos_distance"):
        series = [float(s.metrics[name][0]) for s in rollout]
        assert len(series) == 7 and np.isfinite(series).all()
    qposes_rollout = np.array([s.pipeline_state.qpos[0].cpu().numpy() for s in rollout])
    assert qposes_rollout.shape == (7, 74)
    info0 = {k: (v[0] if isinstance(v, torch.Tensor) else v) for k, v in rollout[0].info.items() if k != "reference_frame"}
    ref_traj = env._get_reference_clip(info0)                                   # scalar clip_idx -> ONE clip, as at wandb_logging.py:136
    ci = int(info0["clip_idx"])
    assert ref_traj.position.shape == (250, 3) and np.array_equal(ref_traj.position, clips2.position[ci])
    qposes_ref = np.repeat(np.hstack([ref_traj.position, ref_traj.quaternion, ref_traj.joints]), int(env._steps_for_cur_frame), axis=0)
    assert qposes_ref.shape == (250, 74)
    batch = env._get_reference_clip(rollout[0].info)                            # vector clip_idx -> one clip per env (the vmapped form)
    assert batch.joints.shape == (4, 250, 67)
    assert np.array_equal(batch.joints, clips2.joints[rollout[0].info["clip_idx"].cpu().numpy()])
    assert env._get_reference_clip({"clip_idx": 99}).position.shape == (250, 3)    # out of range clamps like jnp indexing


def test_reference_frame_cur_frame_and_obs_accessors(walker, clips2):
    env = wrap(make_env(walker, clips2, 8))
    state = env.reset(1)
    for _ in range(3):
        state = env.step(state, torch.zeros(8, env.action_size, device=env.device))
    ci, fr = state.info["clip_idx"].cpu().numpy(), state.info["cur_frame"].cpu().numpy()
    rf = state.info["reference_frame"]
    assert np.array_equal(rf.position.cpu().numpy(), clips2.position[ci, fr])
    assert np.array_equal(rf.joints_velocity.cpu().numpy(), clips2.joints_velocity[ci, fr])
    assert rf.body_positions.shape == (8, 67, 3)
    alive = (state.done == 0).cpu().numpy()
    assert alive.any()
    got = env._get_cur_frame(state.info, state.pipeline_state).cpu().numpy()
    assert np.array_equal(got[alive], fr[alive])                # a restored env's time is back to 0 while cur_frame reports the pre-reset frame
    ref_obs, prop_obs = env._get_obs(state.pipeline_state, state.info)
    assert ref_obs.shape == (8, state.info["reference_obs_size"]) and prop_obs.shape == (8, state.info["proprioceptive_obs_size"])
    assert torch.equal(torch.cat([ref_obs, prop_obs], -1), state.obs)
    with pytest.raises(ValueError):
        env._get_obs(state.clone().pipeline_state, state.info)
    assert env._mjx_model is walker and env._n_clips == 2 and env.sys.nu == 38 and abs(env.dt - 0.02) < 1e-8


def test_wrap_episode_length_is_pushed_to_the_device_and_side_effect_free(walker, clips2):
    env = make_env(walker, clips2, 16)
    before = env.cfg.episode_length
    assert before == 195
    with pytest.raises(ValueError):
        wrap(env, episode_length=0)
    assert env.cfg.episode_length == before and not env._autoreset        # a failed wrap changes nothing
    env = wrap(env, episode_length=4.5)                                    # brax: truncation when steps >= 4.5, i.e. at step 5
    assert env.cfg.episode_length == 5
    state = env.reset(0)
    for i in range(1, 7):
        state = env.step(state, torch.zeros(16, env.action_size, device=env.device))
        if i < 5:
            assert (state.info["truncation"] == 0).all()
    # at step 5 every env that had not terminated on its own was truncated and restored (steps restart from 0 afterwards)
    assert (state.info["steps"] <= 1).all()
    wrap(env, episode_length=5)                                            # same value again: accepted, no change
    assert env.cfg.episode_length == 5


def test_step_before_reset_after_wrap_raises_and_bad_indices_clamp(walker, clips2):
    env = make_env(walker, clips2, 4)
    state = env.reset(0)
    wrap(env)
    with pytest.raises(RuntimeError):
        env.step(state, torch.zeros(4, env.action_size, device=env.device))
    g = torch.Generator(device=env.device).manual_seed(0)
    info = {"clip_idx": torch.tensor([0, 1, 7, -3], device=env.device, dtype=torch.int32),
            "start_frame": torch.tensor([0, 10, 400, -1], device=env.device, dtype=torch.int32)}
    state = env.reset_from_clip(g, info, noise=False)
    assert state.info["clip_idx"].tolist() == [0, 1, 1, 0] and state.info["start_frame"].tolist() == [0, 10, 249, 0]
    assert torch.isfinite(state.obs).all()
    with pytest.raises(ValueError):
        env.reset_from_clip(g, {"clip_idx": torch.zeros(4, device=env.device), "start_frame": info["start_frame"]})
    env2 = MultiClipTracking(clips2, walker, config.RewardConfig(), num_envs=2, device=torch.device("cuda"), **config.DEFAULT_ENV_ARGS)
    assert env2.device.index is not None and torch.isfinite(env2.reset(0).obs).all()


def test_step_host_equals_step_bit_for_bit(walker, clips2):
    """`step_host` (host buffers in / out; the batch cut at lock-step round boundaries so that the device->host copy of the first part
    overlaps the step kernel of the second) against the plain `step` on a twin env: two launches of 2072 + 128 envs give exactly the
    bits of one launch of 2200 (the step kernel's results are independent of the grid), and the host buffers hold what the device
    state holds.  Runs under the fused auto-reset wrapper, which is what the bench's end-to-end arm steps."""
    n = 2200
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    a, b = wrap(make_env(walker, clips2, n)), wrap(make_env(walker, clips2, n))
    assert n > sm * a.stepper.dims["envs_per_block"]                  # more than one round: the batch IS cut
    sa, sb = a.reset(5), b.reset(5)
    assert torch.equal(sa.obs, sb.obs)
    h_obs, h_rew, h_done = (torch.empty(n, a.observation_size).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n).pin_memory())
    g = torch.Generator().manual_seed(0)
    for t in range(4):
        h_act = (0.3 * torch.randn(n, a.action_size, generator=g)).pin_memory()
        sa = a.step(sa, h_act.cuda())
        sb = b.step_host(sb, h_act, h_obs, h_rew, h_done)
        torch.cuda.synchronize()
        for k in ("qpos", "qvel", "obs", "reward", "done", "metrics", "cur_frame", "steps", "action_buffer"):
            assert torch.equal(a.stepper.buf[k], b.stepper.buf[k]), (t, k)
        assert torch.equal(h_obs, sb.obs.cpu()) and torch.equal(h_rew, sb.reward.cpu()) and torch.equal(h_done, sb.done.cpu())
    with pytest.raises(ValueError):
        b.step_host(sb, h_act, h_obs[:10], h_rew, h_done)
