"""Brax-style `Env` mirror of the reference tracking env, backed by the B200 step library.

Drop-in surface (names, argument meaning and conventions follow the reference):
  * `MultiClipTracking(reference_clip, walker, reward_config, physics_steps_per_control_step, ...)`
        reference track_mjx/environment/task/multi_clip_tracking.py:13-109
  * `.reset(rng[, clip_idx]) -> State`, `.reset_from_clip(rng, info, noise=True) -> State`,
    `.step(state, action) -> State`      reference .../task/single_clip_tracking.py:121-320
  * `State(pipeline_state, obs, reward, done, metrics, info)`   (brax.envs.base.State)
  * `wrap(env, episode_length)`: EpisodeWrapper + auto-reset fused into the same launch
        reference track_mjx/environment/wrappers.py:18-56, 104-144, 288-310

Differences that are inherent to leaving JAX (documented in DESIGN.md): the env is *batched* (the leading
axis of every array is the env axis that `jax.vmap` would add), arrays are `torch` CUDA tensors, `rng` is a
`torch.Generator` or an int seed (threefry parity is not claimed), and `State` objects returned by `step`
share the env's device buffers (the functional-purity of the JAX original is traded for zero-copy stepping;
`State.clone()` gives an independent copy).
"""

from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any

import numpy as np
import torch

from . import _lib as L
from . import config as _config
from .clips import ReferenceClip


@dataclasses.dataclass
class PipelineState:
    """The slice of `mjx.Data` the task reads and carries (lazily the rest is not materialised)."""

    qpos: torch.Tensor
    qvel: torch.Tensor
    act: torch.Tensor
    time: torch.Tensor
    qacc_warmstart: torch.Tensor
    xpos: torch.Tensor
    xquat: torch.Tensor
    qfrc_actuator: torch.Tensor

    # brax aliases
    @property
    def q(self):
        return self.qpos

    @property
    def qd(self):
        return self.qvel


@dataclasses.dataclass
class State:
    pipeline_state: PipelineState
    obs: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    metrics: dict[str, torch.Tensor]
    info: dict[str, Any]

    def clone(self) -> "State":
        cl = lambda x: x.clone() if isinstance(x, torch.Tensor) else x  # noqa: E731
        ps = PipelineState(**{f.name: cl(getattr(self.pipeline_state, f.name)) for f in dataclasses.fields(PipelineState)})
        return State(ps, cl(self.obs), cl(self.reward), cl(self.done), {k: cl(v) for k, v in self.metrics.items()},
                     {k: cl(v) for k, v in self.info.items()})


class Stepper:
    """Owns the device model, clip table and per-env buffers; thin wrapper over the C ABI."""

    def __init__(self, blob: bytes, cfg: _config.TaskConfigC, clips: ReferenceClip, n_env: int, device: int | torch.device = 0,
                 debug: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("track_mjx_b200 needs a CUDA device: the env step has no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:                      # torch.device("cuda") -> the current device, never None into ctypes
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_env = n_env
        self.cfg = cfg
        self._model = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib, self.lib.tmjx_model_create(blob, len(blob), C.byref(cfg), self.device.index, C.byref(self._model)),
                    "tmjx_model_create")
            d = L.DimsC()
            L.check(self.lib, self.lib.tmjx_model_dims(self._model, C.byref(d)), "tmjx_model_dims")
            self.dims = L.dims_dict(d)
            fp = C.POINTER(C.c_float)
            arrs = [np.ascontiguousarray(getattr(clips, k), np.float32) for k in (
                "position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity",
                "body_quaternions")]
            self._clips = C.c_void_p()
            L.check(self.lib, self.lib.tmjx_clips_create(
                self._model, *[a.ctypes.data_as(fp) for a in arrs], clips.position.shape[0], clips.position.shape[1],
                clips.body_positions.shape[2], C.byref(self._clips)), "tmjx_clips_create")
        self.n_clips, self.clip_length = clips.position.shape[:2]
        self.buf: dict[str, torch.Tensor] = {}
        fields = L.STATE_FIELDS + L.OUT_FIELDS + (tuple(f for f in L.DEBUG_FIELDS if f[0] != "dbg_qM") if debug else ())
        for name, spec, kind in fields:
            self.buf[name] = torch.zeros((n_env, L.field_size(spec, self.dims)), device=self.device,
                                         dtype=torch.float32 if kind == "f" else torch.int32)
        self._own_obs = self.buf["obs"]
        self._state_c = L.fill_struct(L.StateC(), L.STATE_FIELDS, self.buf, lambda t: t.data_ptr())
        self._out_c = L.fill_struct(L.OutC(), L.OUT_FIELDS + L.DEBUG_FIELDS, self.buf, lambda t: t.data_ptr())

    def close(self):
        if getattr(self, "_clips", None):
            self.lib.tmjx_clips_destroy(self._clips)
            self._clips = None
        if getattr(self, "_model", None):
            self.lib.tmjx_model_destroy(self._model)
            self._model = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, flags: int = 0):
        L.check(self.lib, self.lib.tmjx_forward(self._model, self._clips, C.byref(self._state_c), C.byref(self._out_c),
                                                self.n_env, flags, self._stream()), "tmjx_forward")

    def step(self, action: torch.Tensor, flags: int = 0):
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(self.device, torch.float32).contiguous()
        if action.shape != (self.n_env, self.dims["nu"]):
            raise ValueError(f"action must have shape ({self.n_env}, {self.dims['nu']})")
        L.check(self.lib, self.lib.tmjx_step(self._model, self._clips, C.c_void_p(action.data_ptr()), C.byref(self._state_c),
                                             C.byref(self._out_c), self.n_env, flags, self._stream()), "tmjx_step")

    def _range_structs(self, lo: int, hi: int):
        """TmjxState / TmjxOut whose pointers start at env `lo` (every per-env buffer is `[n_env, width]` row-major)."""
        key = (lo, hi, self.buf["obs"].data_ptr())
        cache = self.__dict__.setdefault("_range_cache", {})
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            off = lambda t: t[lo:].data_ptr() if lo < t.shape[0] else t.data_ptr()
            cache[key] = (L.fill_struct(L.StateC(), L.STATE_FIELDS, self.buf, off), L.fill_struct(L.OutC(), L.OUT_FIELDS + L.DEBUG_FIELDS, self.buf, off))
        return cache[key]

    def step_range(self, action: torch.Tensor, lo: int, hi: int, flags: int = 0):
        """Step the environments `lo <= e < hi` only (`action` is the full `[n_env, nu]` device tensor).  Results are independent of how
        the batch is cut into launches (no cross-env communication, fixed-order arithmetic): `MultiClipTracking.step_host` uses this to
        copy the first part of a batch to the host while the second part is still in the step kernel."""
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device or action.shape != (self.n_env, self.dims["nu"]):
            raise ValueError(f"action must be a contiguous fp32 ({self.n_env}, {self.dims['nu']}) tensor on {self.device}")
        if not 0 <= lo < hi <= self.n_env:
            raise ValueError("bad env range")
        st, out = self._range_structs(lo, hi)
        L.check(self.lib, self.lib.tmjx_step(self._model, self._clips, C.c_void_p(action[lo:].data_ptr()), C.byref(st), C.byref(out), hi - lo, flags,
                                             self._stream()), "tmjx_step")

    def redirect_obs(self, obs: torch.Tensor | None = None):
        """Point TmjxOut.obs at a caller-owned `[n_env, obs_size]` tensor (e.g. slot t+1 of a rollout buffer) so that the step
        kernel writes the observation where its consumer wants it; `None` restores the stepper's own buffer."""
        if obs is None:
            obs = self._own_obs
        if obs.shape != self._own_obs.shape or obs.dtype != torch.float32 or not obs.is_contiguous() or obs.device != self.device:
            raise ValueError("obs target must be a contiguous fp32 [n_env, obs_size] tensor on the env's device")
        self.buf["obs"] = obs
        self._out_c.obs = obs.data_ptr()

    def set_episode_length(self, episode_length: int):
        L.check(self.lib, self.lib.tmjx_model_set_episode_length(self._model, int(episode_length)), "tmjx_model_set_episode_length")
        self.cfg.episode_length = int(episode_length)

    def clips_device_bytes(self) -> int:
        return int(self.lib.tmjx_clips_device_bytes(self._clips))

    def fp32_peak_tflops(self) -> float:
        return float(self.lib.tmjx_fp32_peak_tflops(self.device.index, self._stream()))


class MultiClipTracking:
    """Batched multi-clip tracking env (reference multi_clip_tracking.py:13; single-clip = a 1-clip table)."""

    def __init__(
        self,
        reference_clip: ReferenceClip,
        walker,
        reward_config: _config.RewardConfig | None,
        physics_steps_per_control_step: int,
        reset_noise_scale: float,
        solver: str = "cg",
        iterations: int = 4,
        ls_iterations: int = 4,
        mj_model_timestep: float = 0.002,
        mocap_hz: int = 50,
        clip_length: int = 250,
        random_init_range: int = 50,
        traj_length: int = 5,
        *,
        num_envs: int,
        device: int = 0,
        debug: bool = False,
    ):
        self.walker = walker
        self._reward_config = reward_config or _config.RewardConfig()
        self._reference_clips = reference_clip
        self._n_clips = reference_clip.position.shape[0]
        self._reset_noise_scale = reset_noise_scale
        self._mocap_hz = mocap_hz
        self._ref_len = traj_length
        self._steps_for_cur_frame = (1.0 / (mocap_hz * mj_model_timestep)) / physics_steps_per_control_step
        self._n_frames = physics_steps_per_control_step
        self.cfg = _config.make_task_config(
            walker, self._reward_config, physics_steps_per_control_step=physics_steps_per_control_step, solver=solver,
            iterations=iterations, ls_iterations=ls_iterations, mj_model_timestep=mj_model_timestep, mocap_hz=mocap_hz,
            clip_length=clip_length, random_init_range=random_init_range, traj_length=traj_length)
        self.num_envs = num_envs
        self.stepper = Stepper(walker.blob, self.cfg, reference_clip, num_envs, device, debug=debug)
        self.device = self.stepper.device
        self._clip_pos = torch.from_numpy(np.ascontiguousarray(reference_clip.position)).to(self.device)
        self._clip_quat = torch.from_numpy(np.ascontiguousarray(reference_clip.quaternion)).to(self.device)
        self._clip_joints = torch.from_numpy(np.ascontiguousarray(reference_clip.joints)).to(self.device)
        self._clip_dev = {"position": self._clip_pos, "quaternion": self._clip_quat, "joints": self._clip_joints}   # other fields on demand
        self._autoreset = False
        self._snapshot_taken = False
        self._mjx_model = walker          # the compiled model constants (the reference keeps `mjx.put_model(...)` here, single_clip_tracking.py:91)

    # ---- attributes callers use
    @property
    def dt(self) -> float:
        return self.cfg.mj_model_timestep * self._n_frames

    @property
    def action_size(self) -> int:
        return self.stepper.dims["nu"]

    @property
    def observation_size(self) -> int:
        return self.stepper.dims["obs_size"]

    @property
    def sys(self):
        return self.walker  # exposes nq / nv / nu like brax `System`

    def _gen(self, rng) -> torch.Generator:
        if isinstance(rng, torch.Generator):
            return rng
        g = torch.Generator(device=self.device)
        g.manual_seed(int(rng))
        return g

    def reset(self, rng, clip_idx: torch.Tensor | int | None = None) -> State:
        """reference multi_clip_tracking.py:74-96 (start_frame ~ randint(0, 44), clip_idx ~ randint(0, n_clips))."""
        g = self._gen(rng)
        n = self.num_envs
        start_frame = torch.randint(0, 44, (n,), generator=g, device=self.device, dtype=torch.int32)
        if clip_idx is None:
            clip_idx = torch.randint(0, self._n_clips, (n,), generator=g, device=self.device, dtype=torch.int32)
        elif not isinstance(clip_idx, torch.Tensor):
            clip_idx = torch.full((n,), int(clip_idx), device=self.device, dtype=torch.int32)
        info = {"clip_idx": clip_idx.to(device=self.device, dtype=torch.int32), "start_frame": start_frame}
        return self.reset_from_clip(g, info, noise=True)

    def reset_from_clip(self, rng, info: dict[str, Any], noise: bool = True) -> State:
        """reference single_clip_tracking.py:121-205."""
        g = self._gen(rng)
        b = self.stepper.buf
        n, nq, nv = self.num_envs, self.stepper.dims["nq"], self.stepper.dims["nv"]
        # jnp indexing clamps out-of-range indices (the reference never raises at run time); an index that is not even integral is a bug
        ci_in, sf_in = torch.as_tensor(info["clip_idx"], device=self.device), torch.as_tensor(info["start_frame"], device=self.device)
        if ci_in.is_floating_point() or sf_in.is_floating_point() or ci_in.numel() not in (1, n) or sf_in.numel() not in (1, n):
            raise ValueError("info['clip_idx'] / info['start_frame'] must be integer tensors with one entry per env (or one for all)")
        ci = ci_in.reshape(-1).expand(n).long().clamp(0, self._n_clips - 1)
        sf = sf_in.reshape(-1).expand(n).long().clamp(0, self._clip_pos.shape[1] - 1)
        info = dict(info, clip_idx=ci.to(torch.int32), start_frame=sf.to(torch.int32))
        new_qpos = torch.cat([self._clip_pos[ci, sf], self._clip_quat[ci, sf], self._clip_joints[ci, sf]], dim=-1)
        s = self._reset_noise_scale
        # the reference draws qpos and qvel noise from the SAME key (:153-161): the first nv qvel draws equal the qpos draws
        u = (torch.rand((n, nq), generator=g, device=self.device) * 2 - 1) * s
        b["qpos"].copy_(new_qpos + u)
        b["qvel"].copy_(u[:, :nv] if noise else torch.zeros((n, nv), device=self.device))
        b["clip_idx"].copy_(info["clip_idx"].view(n, 1))
        b["start_frame"].copy_(info["start_frame"].view(n, 1))
        self.stepper.forward(L.TMJX_F_SNAPSHOT if self._autoreset else 0)
        self._snapshot_taken = self._autoreset
        return self._state()

    def step(self, state: State, action: torch.Tensor) -> State:
        """reference single_clip_tracking.py:207-320 (+ fused wrappers after `wrap`)."""
        if self._autoreset and not self._snapshot_taken:
            raise RuntimeError("wrap(env) was applied after the last reset: the auto-reset wrapper has no first_* snapshot to restore "
                               "(the reference takes it in the wrapper's reset, wrappers.py:281-286); call env.reset(...) again")
        self.stepper.step(action, L.TMJX_F_AUTORESET if self._autoreset else 0)
        return self._state()

    def step_host(self, state: State, h_action: torch.Tensor, h_obs: torch.Tensor, h_reward: torch.Tensor, h_done: torch.Tensor,
                  parts: int = 2) -> State:
        """`step` for a caller whose actions and results live in (pinned) HOST memory -- the situation of a host-side policy or of the
        reference's own API boundary.  Copies the actions in, steps, and copies obs / reward / done out, with the batch cut at lock-step
        round boundaries of the step kernel (one round = SM count x environments per block) so that the device->host copy of the first
        part overlaps the kernel of the next one.  Returns when the host buffers are valid.  Same results as `step` (the step kernel's
        results do not depend on how the batch is cut into launches)."""
        if self._autoreset and not self._snapshot_taken:
            raise RuntimeError("wrap(env) was applied after the last reset; call env.reset(...) again")
        n, sp = self.num_envs, self.stepper
        for name, t, shape in (("h_action", h_action, (n, self.action_size)), ("h_obs", h_obs, (n, self.observation_size)), ("h_reward", h_reward, (n,)),
                               ("h_done", h_done, (n,))):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"{name} must be a contiguous fp32 host tensor of shape {shape}")
        if not hasattr(self, "_host"):
            props = torch.cuda.get_device_properties(self.device)
            self._host = {"act": torch.empty(n, self.action_size, device=self.device), "copy": torch.cuda.Stream(device=self.device),
                          "round": props.multi_processor_count * int(sp.dims.get("envs_per_block", 14))}
        h = self._host
        main = torch.cuda.current_stream(self.device)
        h["act"].copy_(h_action, non_blocking=True)
        rounds = -(-n // h["round"])
        parts = max(1, min(parts, rounds))
        cuts = [min(n, -(-rounds * p // parts) * h["round"]) for p in range(parts + 1)]
        flags = L.TMJX_F_AUTORESET if self._autoreset else 0
        b = sp.buf
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            if hi <= lo:
                continue
            sp.step_range(h["act"], lo, hi, flags)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(h["copy"]):
                h["copy"].wait_event(ev)
                h_obs[lo:hi].copy_(b["obs"][lo:hi], non_blocking=True)
                h_reward[lo:hi].copy_(b["reward"][lo:hi, 0], non_blocking=True)
                h_done[lo:hi].copy_(b["done"][lo:hi, 0], non_blocking=True)
        h["copy"].synchronize()
        main.wait_stream(h["copy"])
        return self._state()

    def _state(self) -> State:
        b = self.stepper.buf
        ps = PipelineState(b["qpos"], b["qvel"], b["act"], b["time"][:, 0], b["qacc_warmstart"],
                           b["xpos"].view(self.num_envs, -1, 3), b["xquat"].view(self.num_envs, -1, 4), b["qfrc_actuator"])
        metrics = {k: b["metrics"][:, i] for i, k in enumerate(_config.METRIC_NAMES)}
        info = {
            "clip_idx": b["clip_idx"][:, 0], "start_frame": b["start_frame"][:, 0], "prev_ctrl": b["prev_ctrl"],
            "action_buffer": b["action_buffer"].view(self.num_envs, self.cfg.var_window_size, -1),
            "buffer_index": b["buffer_index"][:, 0], "cur_frame": b["cur_frame"][:, 0],
            "reference_obs_size": self.stepper.dims["reference_obs_size"],
            "proprioceptive_obs_size": self.stepper.dims["proprioceptive_obs_size"],
        }
        info["reference_frame"] = ReferenceFrameView(self, info["clip_idx"], info["cur_frame"])     # lazy: gathered only when read
        if self._autoreset:
            info.update(steps=b["steps"][:, 0], truncation=b["truncation"][:, 0])
        return State(ps, b["obs"], b["reward"][:, 0], b["done"][:, 0], metrics, info)

    def _clip_field(self, name: str) -> torch.Tensor:
        """Device copy of one ReferenceClip field `[n_clips, clip_len, ...]`, uploaded on first use (the step kernel has its own packed table)."""
        if name not in self._clip_dev:
            self._clip_dev[name] = torch.from_numpy(np.ascontiguousarray(getattr(self._reference_clips, name), np.float32)).to(self.device)
        return self._clip_dev[name]

    def _get_reference_clip(self, info) -> ReferenceClip:
        """reference multi_clip_tracking.py:98-109: `tree.map(lambda x: x[info["clip_idx"]], self._reference_clips)` -- the clip(s) the
        given env(s) track.  A scalar `clip_idx` gives one clip `(clip_len, d)` (the call at wandb_logging.py:136), a vector gives
        `(n, clip_len, d)` per field, like the reference under vmap.  Out-of-range indices clamp, as jnp indexing does.  Host arrays."""
        idx = info["clip_idx"]
        idx = idx.detach().cpu().numpy() if isinstance(idx, torch.Tensor) else np.asarray(idx)
        idx = np.clip(idx.astype(np.int64), 0, self._n_clips - 1)
        fields = {f.name: getattr(self._reference_clips, f.name) for f in dataclasses.fields(self._reference_clips)}
        return type(self._reference_clips)(**{k: (None if v is None else np.asarray(v)[idx]) for k, v in fields.items()})

    def _get_cur_frame(self, info, data) -> torch.Tensor:
        """reference single_clip_tracking.py:452-454 (fp32 multiply, add, floor): same arithmetic as the kernel's frame index."""
        t = torch.as_tensor(data.time, dtype=torch.float32, device=self.device)
        return torch.floor(t * torch.tensor(float(self._mocap_hz), dtype=torch.float32, device=self.device)
                           + info["start_frame"].to(torch.float32)).to(torch.int32)

    def _get_obs(self, data, info):
        """reference single_clip_tracking.py:394-450 -> (reference_obs, proprioceptive_obs).  The observation is produced inside the step /
        forward launch; this accessor serves it for the env's LIVE state (the only state for which it exists on the device) and refuses
        a foreign `data` instead of returning a stale observation."""
        live = self.stepper.buf["qpos"]
        q = getattr(data, "qpos", None)
        if not (isinstance(q, torch.Tensor) and q.data_ptr() == live.data_ptr()):
            raise ValueError("_get_obs serves the env's live pipeline_state only: step / reset_from_clip compute the observation in the same "
                             "launch as the physics; put the state into the env (reset_from_clip) to observe it")
        r = self.stepper.dims["reference_obs_size"]
        obs = self.stepper.buf["obs"]
        return obs[:, :r], obs[:, r:]


class ReferenceFrameView:
    """`info["reference_frame"]` (single_clip_tracking.py:223-226: `tree.map(lambda x: x[cur_frame], reference_clip)`) without
    materialising 616 floats per env per step: each ReferenceClip field is gathered from the device clip table when it is read
    (`state.info["reference_frame"].position`, as wrappers.py:353-363 does).  Index clamping follows jnp."""

    FIELDS = ("position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity", "body_quaternions")

    def __init__(self, env: "MultiClipTracking", clip_idx: torch.Tensor, cur_frame: torch.Tensor):
        self._env, self._clip, self._frame = env, clip_idx, cur_frame

    def __getattr__(self, name):
        if name.startswith("_") or name not in self.FIELDS:
            raise AttributeError(name)
        tab = self._env._clip_field(name)
        ci = self._clip.long().clamp(0, tab.shape[0] - 1)
        fr = self._frame.long().clamp(0, tab.shape[1] - 1)
        return tab[ci, fr]


def wrap(env: MultiClipTracking, episode_length: int | None = None) -> MultiClipTracking:
    """reference wrappers.wrap (wrappers.py:18-56): EpisodeWrapper + VmapWrapper + auto-reset.

    The env is already batched, and the episode counter / truncation / `where(done, first_*, cur)` restore are
    executed inside the step launch (TMJX_F_AUTORESET), so this only switches the fused path on.
    """
    if episode_length is not None:
        import math

        limit = int(math.ceil(float(episode_length) - 1e-9))     # brax: `steps >= episode_length` on an integer-valued counter
        if limit <= 0:
            raise ValueError("episode_length must be positive")
        if limit != env.cfg.episode_length:
            env.stepper.set_episode_length(limit)            # pushed to the device model; nothing is changed when this raises
            env.cfg.episode_length = limit
    env._autoreset = True
    return env
