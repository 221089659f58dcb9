"""Host-side mirror of the reference's intention-network acting policy, driving the tcgen05 kernels in csrc/tmjx_policy.cu.

Reference: `track_mjx/agent/mlp_ppo/intention_network.py:14-142` (Encoder / Decoder / IntentionNetwork),
`ppo_networks.py:34-100` (make_inference_fn: sample, log_prob, postprocess), `masked_running_statistics.py:217-236`
(normalize).  Parameters are held as a dict that mirrors the flax tree (`encoder/hidden_i/{kernel,bias}`,
`encoder/LayerNorm_i/{scale,bias}`, `encoder/fc2_mean`, `encoder/fc2_logvar`, `decoder/...`); `flatten_params` lays them out in
the order include/tmjx.h documents.  Inference only: there is no CPU fallback and no autograd.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Sequence

import numpy as np

from . import _lib as L


@dataclass
class IntentionNetworkConfig:
    """network_config of config/rodent-full-clips.yaml:50-57 plus the env's observation sizes."""
    obs_size: int = 696
    reference_obs_size: int = 470
    action_size: int = 38
    latent_size: int = 60
    encoder_layers: Sequence[int] = (1024, 512, 512, 512, 512)
    decoder_layers: Sequence[int] = (512, 512, 512, 256, 256)


def init_params(cfg: IntentionNetworkConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """LeCun-uniform kernels (flax Dense default in the reference: `jax.nn.initializers.lecun_uniform()`), zero biases,
    LayerNorm scale 1 / bias 0, identity normaliser.  numpy RNG: stream parity with jax.random is not claimed."""
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {"norm/mean": np.zeros(cfg.obs_size, np.float32), "norm/std": np.ones(cfg.obs_size, np.float32)}

    def dense(name, k, n):
        lim = np.sqrt(3.0 / k)
        p[f"{name}/kernel"] = rng.uniform(-lim, lim, size=(k, n)).astype(np.float32)
        p[f"{name}/bias"] = np.zeros(n, np.float32)

    k = cfg.reference_obs_size
    for i, n in enumerate(cfg.encoder_layers):
        dense(f"encoder/hidden_{i}", k, n)
        p[f"encoder/LayerNorm_{i}/scale"] = np.ones(n, np.float32)
        p[f"encoder/LayerNorm_{i}/bias"] = np.zeros(n, np.float32)
        k = n
    dense("encoder/fc2_mean", k, cfg.latent_size)
    dense("encoder/fc2_logvar", k, cfg.latent_size)
    k = cfg.latent_size + cfg.obs_size - cfg.reference_obs_size
    for i, n in enumerate(cfg.decoder_layers):
        dense(f"decoder/hidden_{i}", k, n)
        p[f"decoder/LayerNorm_{i}/scale"] = np.ones(n, np.float32)
        p[f"decoder/LayerNorm_{i}/bias"] = np.zeros(n, np.float32)
        k = n
    dense(f"decoder/hidden_{len(cfg.decoder_layers)}", k, 2 * cfg.action_size)
    return p


def flatten_params(cfg: IntentionNetworkConfig, p: Dict[str, np.ndarray]) -> np.ndarray:
    parts = [p["norm/mean"], p["norm/std"]]
    for i in range(len(cfg.encoder_layers)):
        parts += [p[f"encoder/hidden_{i}/kernel"], p[f"encoder/hidden_{i}/bias"], p[f"encoder/LayerNorm_{i}/scale"], p[f"encoder/LayerNorm_{i}/bias"]]
    parts += [p["encoder/fc2_mean/kernel"], p["encoder/fc2_mean/bias"], p["encoder/fc2_logvar/kernel"], p["encoder/fc2_logvar/bias"]]
    for i in range(len(cfg.decoder_layers)):
        parts += [p[f"decoder/hidden_{i}/kernel"], p[f"decoder/hidden_{i}/bias"], p[f"decoder/LayerNorm_{i}/scale"], p[f"decoder/LayerNorm_{i}/bias"]]
    n = len(cfg.decoder_layers)
    parts += [p[f"decoder/hidden_{n}/kernel"], p[f"decoder/hidden_{n}/bias"]]
    return np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).ravel() for a in parts]))


def make_desc(cfg: IntentionNetworkConfig) -> L.PolicyDescC:
    d = L.PolicyDescC()
    d.obs_size, d.reference_obs_size, d.latent_size, d.action_size = cfg.obs_size, cfg.reference_obs_size, cfg.latent_size, cfg.action_size
    d.n_encoder_layers = len(cfg.encoder_layers)
    d.n_decoder_layers = len(cfg.decoder_layers)
    for i, n in enumerate(cfg.encoder_layers):
        d.encoder_layers[i] = int(n)
    for i, n in enumerate(cfg.decoder_layers):
        d.decoder_layers[i] = int(n)
    return d


class IntentionPolicy:
    """`policy(observations, key) -> (action, extras)` of make_inference_fn on the GPU.  Torch is plumbing only (device
    buffers and the Gaussian noise the reference draws with jax.random)."""

    def __init__(self, cfg: IntentionNetworkConfig, params: Dict[str, np.ndarray], max_env: int, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("IntentionPolicy needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.cfg, self.max_env = cfg, int(max_env)
        self.device = torch.device("cuda", device)
        self.lib = L.load()
        flat = flatten_params(cfg, params)
        desc = make_desc(cfg)
        if flat.size != self.lib.tmjx_policy_param_count(C.byref(desc)):
            raise ValueError("parameter tree does not match the network config")
        self._p = C.c_void_p()
        rc = self.lib.tmjx_policy_create(C.byref(desc), flat.ctypes.data_as(C.POINTER(C.c_float)), flat.size, device, self.max_env, C.byref(self._p))
        if rc != 0:
            raise RuntimeError(f"tmjx_policy_create failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")
        f = dict(dtype=torch.float32, device=self.device)
        n, a, z = self.max_env, cfg.action_size, cfg.latent_size
        self.out = {"action": torch.empty(n, a, **f), "raw_action": torch.empty(n, a, **f), "log_prob": torch.empty(n, **f),
                    "logits": torch.empty(n, 2 * a, **f), "latent_mean": torch.empty(n, z, **f), "latent_logvar": torch.empty(n, z, **f)}
        self.launches_per_act = int(self.lib.tmjx_policy_launches_per_act(self._p))

    def act(self, obs, eps_latent=None, eps_action=None, deterministic: bool = False, out=None):
        """obs: [n, obs_size] CUDA tensor.  Returns (action, extras) views into buffers owned by the policy, or into the
        caller's `out` tensors (keys of self.out; e.g. slot t of a rollout buffer) when given."""
        t = self.torch
        n = int(obs.shape[0])
        if not deterministic:
            if eps_latent is None:
                eps_latent = t.randn(n, self.cfg.latent_size, device=self.device)
            if eps_action is None:
                eps_action = t.randn(n, self.cfg.action_size, device=self.device)
        for name, x, width in (("obs", obs, self.cfg.obs_size), ("eps_latent", eps_latent, self.cfg.latent_size), ("eps_action", eps_action, self.cfg.action_size)):
            if x is not None and (x.dtype != t.float32 or not x.is_contiguous() or x.device != self.device or tuple(x.shape) != (n, width)):
                raise ValueError(f"{name} must be a contiguous float32 [{n}, {width}] tensor on {self.device}")
        if n > self.max_env:
            raise ValueError(f"{n} rows exceed the policy's max_env = {self.max_env}")
        ptr = lambda x: None if x is None else C.c_void_p(x.data_ptr())
        o = self.out if out is None else {**self.out, **out}
        for k, v in o.items():
            if v.dtype != t.float32 or not v.is_contiguous() or v.shape[0] < n:
                raise ValueError(f"policy output '{k}' must be a contiguous fp32 tensor with >= {n} rows")
        rc = self.lib.tmjx_policy_act(self._p, ptr(obs), ptr(eps_latent), ptr(eps_action), int(deterministic), ptr(o["action"]),
                                      ptr(o["raw_action"]), ptr(o["log_prob"]), ptr(o["logits"]), ptr(o["latent_mean"]),
                                      ptr(o["latent_logvar"]), n, C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"tmjx_policy_act failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")
        return o["action"][:n], {k: v[:n] for k, v in o.items() if k != "action"}

    def set_params(self, flat_params):
        """Refresh the weights and the normaliser from a flat fp32 CUDA vector in `flatten_params` order (e.g. the policy slice of
        `learner.Trainer.params` after an optimiser step): `tmjx_policy_set_params`, an in-place repack, no reallocation."""
        t = self.torch
        if flat_params.dtype != t.float32 or not flat_params.is_contiguous() or flat_params.device != self.device:
            raise ValueError("flat_params must be a contiguous float32 tensor on the policy's device")
        desc = make_desc(self.cfg)
        if flat_params.numel() != self.lib.tmjx_policy_param_count(C.byref(desc)):
            raise ValueError("flat parameter vector has the wrong length")
        rc = self.lib.tmjx_policy_set_params(self._p, C.c_void_p(flat_params.data_ptr()), C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"tmjx_policy_set_params failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")

    def linear(self, which: int, x, y):
        rc = self.lib.tmjx_policy_linear(self._p, which, C.c_void_p(x.data_ptr()), int(x.stride(0)), C.c_void_p(y.data_ptr()), int(y.stride(0)),
                                         int(x.shape[0]), C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"tmjx_policy_linear failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")

    def close(self):
        if self._p:
            self.lib.tmjx_policy_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def init_value_params(obs_size: int, hidden_layers: Sequence[int] = (1024, 1024), seed: int = 0) -> Dict[str, np.ndarray]:
    """Random-init parameter tree of the value MLP in flax naming (`hidden_i/kernel` [in, out], `hidden_i/bias`; the last Dense has
    one output), lecun-uniform like brax's `make_value_network`; identity normaliser."""
    rng = np.random.default_rng(seed)
    p = {"norm/mean": np.zeros(obs_size, np.float32), "norm/std": np.ones(obs_size, np.float32)}
    k = obs_size
    for i, n in enumerate(list(hidden_layers) + [1]):
        lim = np.sqrt(3.0 / k)
        p[f"hidden_{i}/kernel"] = rng.uniform(-lim, lim, (k, n)).astype(np.float32)
        p[f"hidden_{i}/bias"] = np.zeros(n, np.float32)
        k = n
    return p


class ValueNetwork:
    """`value_network.apply(normalizer_params, params.value, obs)` of `ppo_networks.py:180-185` (brax `make_value_network`: normalise,
    Dense + swish per hidden layer, Dense to 1, squeeze) on the GPU through `tmjx_value_apply`; default hidden sizes are the
    reference's `value_hidden_layer_sizes=(1024,) * 2` (`ppo_networks.py:165`).  No CPU fallback."""

    def __init__(self, obs_size: int, params: Dict[str, np.ndarray], max_env: int, hidden_layers: Sequence[int] = (1024, 1024), device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("ValueNetwork needs a CUDA device: there is no CPU fallback")
        self.torch, self.lib, self.max_env = torch, L.load(), int(max_env)
        self.device = torch.device("cuda", device)
        self.obs_size = int(obs_size)
        desc = L.ValueDescC()
        desc.obs_size, desc.n_hidden_layers = int(obs_size), len(hidden_layers)
        for i, n in enumerate(hidden_layers):
            desc.hidden_layers[i] = int(n)
        parts = [params["norm/mean"], params["norm/std"]]
        for i in range(len(hidden_layers) + 1):
            parts += [params[f"hidden_{i}/kernel"], params[f"hidden_{i}/bias"]]
        flat = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).ravel() for a in parts]))
        if flat.size != self.lib.tmjx_value_param_count(C.byref(desc)):
            raise ValueError("parameter tree does not match the value-network shape")
        self._p = C.c_void_p()
        rc = self.lib.tmjx_value_create(C.byref(desc), flat.ctypes.data_as(C.POINTER(C.c_float)), flat.size, device, self.max_env, C.byref(self._p))
        if rc != 0:
            raise RuntimeError(f"tmjx_value_create failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")

    def apply(self, obs, out=None):
        """obs: [..., obs_size] fp32 CUDA tensor (any leading dimensions, e.g. [T, B]) -> value [...]."""
        t = self.torch
        lead = tuple(obs.shape[:-1])
        x = obs.reshape(-1, self.obs_size).to(t.float32).contiguous()
        n = int(x.shape[0])
        if out is not None and (not out.is_contiguous() or out.dtype != t.float32 or out.numel() != n or out.device != self.device):
            raise ValueError("out must be a contiguous float32 tensor with one element per observation row, on the network's device")
        v = t.empty(n, dtype=t.float32, device=self.device) if out is None else out.view(-1)
        st = C.c_void_p(t.cuda.current_stream(self.device).cuda_stream)
        for i in range(0, n, self.max_env):                          # row chunks of at most max_env (the activation buffers' size)
            m = min(self.max_env, n - i)
            rc = self.lib.tmjx_value_apply(self._p, C.c_void_p(x.data_ptr() + 4 * i * self.obs_size), C.c_void_p(v.data_ptr() + 4 * i), m, st)
            if rc != 0:
                raise RuntimeError(f"tmjx_value_apply failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")
        return v.reshape(lead)

    def close(self):
        if self._p:
            self.lib.tmjx_policy_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------- golden cases
# Seeded inputs of tests/golden/policy.npz (outputs computed by the reference's own module text, tools/make_golden_policy.py).
GOLDEN_CASES = (
    ("shipped", IntentionNetworkConfig(), 11, 48),                                                    # rodent-full-clips.yaml:50-57
    ("small", IntentionNetworkConfig(obs_size=100, reference_obs_size=60, action_size=6, latent_size=12, encoder_layers=(96, 64),
                                     decoder_layers=(64, 32)), 12, 40),
)


def golden_case(cfg: IntentionNetworkConfig, seed: int, rows: int):
    """(params, obs [rows, obs_size], eps_latent [rows, latent]) regenerated from a seed (numpy PCG64: stable across versions).  The
    parameters are random-init kernels with NON-trivial biases, LayerNorm scale / bias and normaliser, so that every term is exercised."""
    rng = np.random.default_rng(seed)
    p = init_params(cfg, seed)
    for k in sorted(p):
        if k.endswith("/bias"):
            p[k] = (0.1 * rng.normal(size=p[k].shape)).astype(np.float32)
        if k.endswith("/scale"):
            p[k] = (1.0 + 0.2 * rng.normal(size=p[k].shape)).astype(np.float32)
    p["norm/mean"] = (0.3 * rng.normal(size=cfg.obs_size)).astype(np.float32)
    p["norm/std"] = (0.5 + rng.uniform(size=cfg.obs_size)).astype(np.float32)
    obs = rng.normal(size=(rows, cfg.obs_size)).astype(np.float32)
    eps = rng.normal(size=(rows, cfg.latent_size)).astype(np.float32)
    return p, obs, eps
