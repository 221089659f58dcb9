"""ctypes binding of the C-ABI library `libtmjx.so` (include/tmjx.h).

This is the reference-side stub a maintainer would add: plain pointers and sizes,
no torch types in the signatures.  The library is mandatory: importing the compute path without it
raises, there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os

from .config import TaskConfigC, TMJX_N_METRICS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TMJX_LIB_PATH") or os.path.join(_HERE, "csrc", "libtmjx.so")   # TMJX_LIB_PATH: development builds (tools/gpu_phase_timing.py)

TMJX_F_AUTORESET = 1
TMJX_F_SNAPSHOT = 2
TMJX_F_EPILOGUE_ONLY = 4

STATE_FIELDS = (
    # name, per-env shape key, dtype ('f' = float, 'i' = int32)
    ("qpos", "nq", "f"), ("qvel", "nv", "f"), ("act", "na", "f"), ("time", 1, "f"), ("qacc_warmstart", "nv", "f"),
    ("xpos", "nbody*3", "f"), ("xquat", "nbody*4", "f"), ("qfrc_actuator", "nv", "f"),
    ("clip_idx", 1, "i"), ("start_frame", 1, "i"), ("buffer_index", 1, "i"),
    ("prev_ctrl", "nu", "f"), ("action_buffer", "var_window_size*nu", "f"),
    ("steps", 1, "f"), ("truncation", 1, "f"),
    ("first_qpos", "nq", "f"), ("first_qvel", "nv", "f"), ("first_act", "na", "f"), ("first_time", 1, "f"),
    ("first_qacc_warmstart", "nv", "f"), ("first_xpos", "nbody*3", "f"), ("first_xquat", "nbody*4", "f"),
    ("first_qfrc_actuator", "nv", "f"), ("first_obs", "obs_size", "f"), ("first_prev_ctrl", "nu", "f"),
)
OUT_FIELDS = (
    ("obs", "obs_size", "f"), ("reward", 1, "f"), ("done", 1, "f"), ("metrics", "n_metrics", "f"),
    ("cur_frame", 1, "i"),
)
DEBUG_FIELDS = (
    ("dbg_qacc", "nv", "f"), ("dbg_qacc_smooth", "nv", "f"), ("dbg_qfrc_bias", "nv", "f"),
    ("dbg_qfrc_constraint", "nv", "f"), ("dbg_contact_dist", "ncon", "f"), ("dbg_efc_force", "nefc", "f"),
    ("dbg_qM", "nv*nv", "f"), ("dbg_subtree_com", 3, "f"),
)


class StateC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _, _ in STATE_FIELDS]


class OutC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n, _, _ in OUT_FIELDS + DEBUG_FIELDS]


class PolicyDescC(C.Structure):
    _fields_ = [("obs_size", C.c_int32), ("reference_obs_size", C.c_int32), ("latent_size", C.c_int32), ("action_size", C.c_int32),
                ("n_encoder_layers", C.c_int32), ("encoder_layers", C.c_int32 * 8),
                ("n_decoder_layers", C.c_int32), ("decoder_layers", C.c_int32 * 8)]


class ValueDescC(C.Structure):
    _fields_ = [("obs_size", C.c_int32), ("n_hidden_layers", C.c_int32), ("hidden_layers", C.c_int32 * 8)]


class DimsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nq", "nv", "nu", "na", "nbody", "njnt", "ncon", "nefc", "obs_size", "reference_obs_size",
        "proprioceptive_obs_size", "var_window_size", "n_metrics", "smem_bytes_per_env", "envs_per_block",
        "threads_per_env")]


def field_size(spec, dims: dict) -> int:
    """Evaluate a per-env element count such as 'nbody*3' against a dims dict."""
    if isinstance(spec, int):
        return spec
    n = 1
    for tok in spec.split("*"):
        n *= int(tok) if tok.isdigit() else int(dims[tok])
    return n


def dims_dict(d: DimsC) -> dict:
    return {n: int(getattr(d, n)) for n, _ in DimsC._fields_}


def fill_struct(struct, fields, arrays: dict, ptr_of):
    for name, _, _ in fields:
        a = arrays.get(name)
        setattr(struct, name, None if a is None else ptr_of(a))
    return struct


_lib = None


def load() -> C.CDLL:
    """Load libtmjx.so; raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback for the environment step.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint, C.c_size_t
    fp = C.POINTER(C.c_float)
    lib.tmjx_abi_version.restype = i32
    lib.tmjx_last_error.restype = C.c_char_p
    lib.tmjx_model_create.argtypes = [vp, sz, C.POINTER(TaskConfigC), i32, C.POINTER(vp)]
    lib.tmjx_model_destroy.argtypes = [vp]
    lib.tmjx_model_destroy.restype = None
    lib.tmjx_model_dims.argtypes = [vp, C.POINTER(DimsC)]
    lib.tmjx_model_set_episode_length.argtypes = [vp, i32]
    lib.tmjx_clips_create.argtypes = [vp] + [fp] * 8 + [i32, i32, i32, C.POINTER(vp)]
    lib.tmjx_clips_destroy.argtypes = [vp]
    lib.tmjx_clips_destroy.restype = None
    lib.tmjx_clips_device_bytes.argtypes = [vp]
    lib.tmjx_clips_device_bytes.restype = sz
    lib.tmjx_forward.argtypes = [vp, vp, C.POINTER(StateC), C.POINTER(OutC), i32, u32, vp]
    lib.tmjx_step.argtypes = [vp, vp, vp, C.POINTER(StateC), C.POINTER(OutC), i32, u32, vp]
    lib.tmjx_fp32_peak_tflops.argtypes = [i32, vp]
    lib.tmjx_fp32_peak_tflops.restype = C.c_double
    lib.tmjx_policy_param_count.argtypes = [C.POINTER(PolicyDescC)]
    lib.tmjx_policy_param_count.restype = sz
    lib.tmjx_policy_create.argtypes = [C.POINTER(PolicyDescC), fp, sz, i32, i32, C.POINTER(vp)]
    lib.tmjx_policy_destroy.argtypes = [vp]
    lib.tmjx_value_param_count.argtypes = [C.POINTER(ValueDescC)]
    lib.tmjx_value_param_count.restype = sz
    lib.tmjx_value_create.argtypes = [C.POINTER(ValueDescC), fp, sz, i32, i32, C.POINTER(vp)]
    lib.tmjx_value_apply.argtypes = [vp, vp, vp, i32, vp]
    lib.tmjx_policy_destroy.restype = None
    lib.tmjx_policy_last_error.restype = C.c_char_p
    lib.tmjx_policy_act.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.tmjx_policy_linear.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp]
    lib.tmjx_policy_launches_per_act.argtypes = [vp]
    lib.tmjx_running_stats_scratch_floats.argtypes = [i32]
    lib.tmjx_running_stats_scratch_floats.restype = sz
    lib.tmjx_running_stats_sums.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.tmjx_running_stats_mean.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    lib.tmjx_running_stats_apply.argtypes = [vp, i32, C.c_float, C.c_float, vp, vp, vp, vp, vp]
    lib.tmjx_ppo_loss_scratch_floats.argtypes = [i32, i32]
    lib.tmjx_ppo_loss_scratch_floats.restype = sz
    lib.tmjx_ppo_loss_head.argtypes = [vp] * 11 + [i32] * 4 + [vp] * 10
    lib.tmjx_adam_scratch_floats.argtypes = []
    lib.tmjx_adam_scratch_floats.restype = sz
    lib.tmjx_adam_step.argtypes = [vp, vp, vp, vp, sz] + [C.c_float] * 6 + [i32, vp, vp, vp]
    lib.tmjx_gae.argtypes = [vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, i32, i32, vp]
    lib.tmjx_policy_set_params.argtypes = [vp, vp, vp]
    lib.tmjx_xla_ffi_available.restype = i32
    lib.tmjx_ffi_selftest.argtypes = [i32, vp, vp, vp, i32, C.POINTER(StateC), C.POINTER(OutC), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, u32, vp, i32]
    lib.tmjx_ffi_selftest_error.restype = C.c_char_p
    lib.tmjx_trainer_create.argtypes = [C.POINTER(PolicyDescC), C.POINTER(ValueDescC), fp, fp, i32, i32, C.POINTER(vp)]
    lib.tmjx_trainer_destroy.argtypes = [vp]
    lib.tmjx_trainer_destroy.restype = None
    lib.tmjx_trainer_param_count.argtypes = [vp]
    lib.tmjx_trainer_param_count.restype = sz
    lib.tmjx_trainer_policy_param_count.argtypes = [vp]
    lib.tmjx_trainer_policy_param_count.restype = sz
    lib.tmjx_trainer_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.tmjx_trainer_sync.argtypes = [vp, vp]
    lib.tmjx_trainer_policy_forward.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    lib.tmjx_trainer_policy_backward.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.tmjx_trainer_value_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.tmjx_trainer_value_backward.argtypes = [vp, vp, i32, vp]
    if lib.tmjx_abi_version() != 1:
        raise ImportError("libtmjx.so ABI version mismatch")
    _lib = lib
    return lib


def check(lib, rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib.tmjx_last_error().decode()}")
