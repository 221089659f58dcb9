"""Task configuration: the kwargs of the reference env, reward config and walker, and their C mirror.

Mirrors (names, meaning, defaults):
  * `RewardConfig`            reference track_mjx/environment/task/reward.py:15-54
  * env_args / reference_config of reference track_mjx/config/rodent-full-clips.yaml:11-50
  * walker_config name lists   reference track_mjx/config/rodent-full-clips.yaml:118-176
  * `TmjxTaskConfig`          include/tmjx.h
"""

from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Sequence

TMJX_ABI_VERSION = 1
TMJX_MAX_IDX = 80
TMJX_N_METRICS = 20
SOLVERS = {"cg": 0, "newton": 1}

METRIC_NAMES = (
    "pos_reward", "quat_reward", "joint_reward", "angvel_reward", "bodypos_reward", "endeff_reward", "ctrl_cost",
    "ctrl_diff_cost", "energy_cost", "done", "too_far", "bad_pose", "bad_quat", "fall", "nan", "joint_distance",
    "summed_pos_distance", "quat_distance", "var_cost", "jerk_cost",
)  # insertion order of the metrics dict, single_clip_tracking.py:176-197


@dataclasses.dataclass(frozen=True)
class RewardConfig:
    """Weights and scales for the imitation reward terms (reward.py:15-54; yaml values :20-43)."""

    too_far_dist: float = 0.01
    bad_pose_dist: float = 20.0
    bad_quat_dist: float = 1.0
    ctrl_cost_weight: float = 0.02
    ctrl_diff_cost_weight: float = 0.02
    energy_cost_weight: float = 0.01
    pos_reward_weight: float = 1.0
    quat_reward_weight: float = 1.0
    joint_reward_weight: float = 1.0
    angvel_reward_weight: float = 0.0
    bodypos_reward_weight: float = 0.0
    endeff_reward_weight: float = 1.0
    healthy_z_range: tuple[float, float] = (0.0325, 0.5)
    pos_reward_exp_scale: float = 400.0
    quat_reward_exp_scale: float = 4.0
    joint_reward_exp_scale: float = 0.25
    angvel_reward_exp_scale: float = 0.5
    bodypos_reward_exp_scale: float = 8.0
    endeff_reward_exp_scale: float = 500.0
    penalty_pos_distance_scale: tuple[float, float, float] = (1.0, 1.0, 0.5)
    var_window_size: int = 50
    var_coeff: float = 5e-3
    jerk_coeff: float = 5e-4


# walker_config of rodent-full-clips.yaml:118-176
RODENT_JOINT_NAMES = (
    "vertebra_1_extend", "hip_L_supinate", "hip_L_abduct", "hip_L_extend", "knee_L", "ankle_L", "toe_L",
    "hip_R_supinate", "hip_R_abduct", "hip_R_extend", "knee_R", "ankle_R", "toe_R", "vertebra_C11_extend",
    "vertebra_cervical_1_bend", "vertebra_axis_twist", "atlas", "mandible", "scapula_L_supinate", "scapula_L_abduct",
    "scapula_L_extend", "shoulder_L", "shoulder_sup_L", "elbow_L", "wrist_L", "scapula_R_supinate",
    "scapula_R_abduct", "scapula_R_extend", "shoulder_R", "shoulder_sup_R", "elbow_R", "wrist_R", "finger_R",
)
RODENT_BODY_NAMES = (
    "torso", "pelvis", "upper_leg_L", "lower_leg_L", "foot_L", "upper_leg_R", "lower_leg_R", "foot_R", "skull", "jaw",
    "scapula_L", "upper_arm_L", "lower_arm_L", "finger_L", "scapula_R", "upper_arm_R", "lower_arm_R", "finger_R",
)
RODENT_END_EFF_NAMES = ("foot_L", "foot_R", "hand_L", "hand_R", "skull")

# env_args + reference_config of rodent-full-clips.yaml:11-50
DEFAULT_ENV_ARGS = dict(
    solver="cg", iterations=5, ls_iterations=5, physics_steps_per_control_step=10, reset_noise_scale=1e-3,
    mj_model_timestep=0.002, mocap_hz=50, clip_length=250, random_init_range=50, traj_length=5,
)


class TaskConfigC(C.Structure):
    """ctypes mirror of `TmjxTaskConfig` (include/tmjx.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("physics_steps_per_control_step", C.c_int32),
        ("solver", C.c_int32),
        ("iterations", C.c_int32),
        ("ls_iterations", C.c_int32),
        ("mj_model_timestep", C.c_float),
        ("mocap_hz", C.c_float),
        ("clip_length", C.c_int32),
        ("traj_length", C.c_int32),
        ("episode_length", C.c_int32),
        ("too_far_dist", C.c_float), ("bad_pose_dist", C.c_float), ("bad_quat_dist", C.c_float),
        ("ctrl_cost_weight", C.c_float), ("ctrl_diff_cost_weight", C.c_float), ("energy_cost_weight", C.c_float),
        ("pos_reward_weight", C.c_float), ("quat_reward_weight", C.c_float), ("joint_reward_weight", C.c_float),
        ("angvel_reward_weight", C.c_float), ("bodypos_reward_weight", C.c_float), ("endeff_reward_weight", C.c_float),
        ("healthy_z_min", C.c_float), ("healthy_z_max", C.c_float),
        ("pos_reward_exp_scale", C.c_float), ("quat_reward_exp_scale", C.c_float), ("joint_reward_exp_scale", C.c_float),
        ("angvel_reward_exp_scale", C.c_float), ("bodypos_reward_exp_scale", C.c_float),
        ("endeff_reward_exp_scale", C.c_float),
        ("penalty_pos_distance_scale", C.c_float * 3),
        ("var_window_size", C.c_int32),
        ("var_coeff", C.c_float), ("jerk_coeff", C.c_float),
        ("n_joint_idxs", C.c_int32), ("n_body_idxs", C.c_int32), ("n_endeff_idxs", C.c_int32),
        ("joint_idxs", C.c_int32 * TMJX_MAX_IDX),
        ("body_idxs", C.c_int32 * TMJX_MAX_IDX),
        ("endeff_idxs", C.c_int32 * TMJX_MAX_IDX),
        ("torso_idx", C.c_int32),
        ("torso_body_id", C.c_int32),
        ("n_appendages", C.c_int32),
        ("appendage_body_ids", C.c_int32 * 8),
    ]


def episode_length(clip_length: int, random_init_range: int, traj_length: int, steps_for_cur_frame: float) -> int:
    """reference track_mjx/train.py:221-225 keeps the FLOAT product and brax's EpisodeWrapper truncates when `steps >= episode_length`
    with an integer-valued step counter: the first step that satisfies it is ceil(limit), which is what the kernel compares against."""
    import math

    return int(math.ceil((clip_length - random_init_range - traj_length) * steps_for_cur_frame - 1e-9))


def make_task_config(
    walker,
    reward_config: RewardConfig,
    *,
    physics_steps_per_control_step: int,
    solver: str,
    iterations: int,
    ls_iterations: int,
    mj_model_timestep: float,
    mocap_hz: int,
    clip_length: int,
    random_init_range: int,
    traj_length: int,
) -> TaskConfigC:
    """Flatten env kwargs + reward config + walker index tables into the C struct."""
    c = TaskConfigC()
    c.abi_version = TMJX_ABI_VERSION
    c.physics_steps_per_control_step = physics_steps_per_control_step
    c.solver = SOLVERS[solver.lower()]
    c.iterations, c.ls_iterations = iterations, ls_iterations
    c.mj_model_timestep, c.mocap_hz = mj_model_timestep, float(mocap_hz)
    c.clip_length, c.traj_length = clip_length, traj_length
    steps_for_cur_frame = (1.0 / (mocap_hz * mj_model_timestep)) / physics_steps_per_control_step
    c.episode_length = episode_length(clip_length, random_init_range, traj_length, steps_for_cur_frame)
    r = reward_config
    for k in ("too_far_dist", "bad_pose_dist", "bad_quat_dist", "ctrl_cost_weight", "ctrl_diff_cost_weight",
              "energy_cost_weight", "pos_reward_weight", "quat_reward_weight", "joint_reward_weight",
              "angvel_reward_weight", "bodypos_reward_weight", "endeff_reward_weight", "pos_reward_exp_scale",
              "quat_reward_exp_scale", "joint_reward_exp_scale", "angvel_reward_exp_scale",
              "bodypos_reward_exp_scale", "endeff_reward_exp_scale", "var_coeff", "jerk_coeff"):
        setattr(c, k, float(getattr(r, k)))
    c.healthy_z_min, c.healthy_z_max = r.healthy_z_range
    for i in range(3):
        c.penalty_pos_distance_scale[i] = float(r.penalty_pos_distance_scale[i])
    c.var_window_size = r.var_window_size

    def put(dst, src: Sequence[int]) -> int:
        if len(src) > TMJX_MAX_IDX:
            raise ValueError("index table too long")
        for i, v in enumerate(src):
            dst[i] = int(v)
        return len(src)

    c.n_joint_idxs = put(c.joint_idxs, walker.joint_idxs)
    c.n_body_idxs = put(c.body_idxs, walker.body_idxs)
    c.n_endeff_idxs = put(c.endeff_idxs, walker.endeff_idxs)
    c.torso_idx = int(walker.torso_idx)
    c.torso_body_id = int(walker.body_id(walker._torso_name))
    app = [walker.body_id(n) for n in walker._end_eff_names]
    if len(app) > 8:
        raise ValueError("at most 8 appendages")
    c.n_appendages = len(app)
    for i, v in enumerate(app):
        c.appendage_body_ids[i] = v
    return c
