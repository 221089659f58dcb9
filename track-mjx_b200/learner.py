"""Learner-side pieces that sit next to the acting loop: GAE, the observation-normaliser update and the
PPO loss head (loss terms + the gradients the network backward pass starts from), the KL-weight schedule and the optimiser step
(global-norm clip + Adam, gradient all-reduce), and `Trainer`: the forward + backward pass through the networks (csrc/tmjx_train.cuh).

`compute_gae` mirrors `track_mjx/agent/mlp_ppo/losses.py:39-101`: same argument names and meaning, time-major `[T, B]` fp32 CUDA
tensors in, `(vs, advantages)` out, computed by the `tmjx_gae` kernel (csrc/tmjx_policy.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float = 1.0, discount: float = 0.99):
    import torch

    if not rewards.is_cuda:
        raise RuntimeError("compute_gae needs CUDA tensors: there is no CPU fallback")
    T, B = rewards.shape
    args = [a.to(torch.float32).contiguous() for a in (truncation, termination, rewards, values, bootstrap_value)]
    for a in args[:4]:
        if a.shape != (T, B):
            raise ValueError("truncation, termination, rewards and values must all be [T, B]")
    if args[4].shape != (B,):
        raise ValueError("bootstrap_value must be [B]")
    vs, adv = torch.empty_like(args[2]), torch.empty_like(args[2])
    lib = L.load()
    ptr = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.tmjx_gae(*[ptr(a) for a in args], float(lambda_), float(discount), ptr(vs), ptr(adv), int(T), int(B),
                      C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"tmjx_gae failed ({rc}): {lib.tmjx_policy_last_error().decode()}")
    return vs, adv


def shuffle_minibatches(data, permutation, num_minibatches: int):
    """`convert_data` of the reference's `sgd_step` (`track_mjx/agent/mlp_ppo/ppo.py:305-310`) on the time-major layout of the
    acting loop: every tensor of `data` is `[T, B, ...]` (the reference holds `[B, T, ...]` and permutes / splits axis 0 = B);
    `permutation` is the `[B]` index vector the reference draws with `jax.random.permutation(key_perm, ...)` (the same one for every
    leaf).  Returns tensors `[num_minibatches, T, B / num_minibatches, ...]`: minibatch `i` holds environments
    `permutation[i * B / num_minibatches : (i + 1) * B / num_minibatches]`, ready for `ppo_loss_head`.  Works on any torch device
    (one gather per leaf; plumbing, not a kernel of this library)."""
    out = {}
    for k, x in data.items():
        if isinstance(x, dict):
            out[k] = shuffle_minibatches(x, permutation, num_minibatches)
            continue
        T, B = x.shape[0], x.shape[1]
        if B % num_minibatches != 0 or permutation.shape != (B,):
            raise ValueError(f"{k}: batch of {B} environments cannot be split into {num_minibatches} minibatches with this permutation")
        y = x.index_select(1, permutation)                                      # jax.random.permutation(key_perm, x) on the env axis
        y = y.reshape((T, num_minibatches, B // num_minibatches) + tuple(x.shape[2:]))
        out[k] = y.movedim(1, 0).contiguous()                                   # jnp.reshape(x, (num_minibatches, -1) + x.shape[1:])
    return out


class _PpoHyper(C.Structure):
    _fields_ = [("entropy_cost", C.c_float), ("kl_weight", C.c_float), ("discounting", C.c_float), ("reward_scaling", C.c_float),
                ("gae_lambda", C.c_float), ("clipping_epsilon", C.c_float), ("normalize_advantage", C.c_int32)]


def ppo_loss_head(policy_logits, latent_mean, latent_logvar, baseline, bootstrap_value, reward, discount, truncation, raw_action,
                  behaviour_log_prob, eps_entropy, entropy_cost: float = 1e-4, kl_weight: float = 1e-3, discounting: float = 0.9,
                  reward_scaling: float = 1.0, gae_lambda: float = 0.95, clipping_epsilon: float = 0.3, normalize_advantage: bool = True):
    """`compute_ppo_loss` of `track_mjx/agent/mlp_ppo/losses.py:104-245` from the point where the networks have been applied (same
    hyper-parameter names and defaults), on time-major `[T, B, ...]` fp32 CUDA tensors (the reference swaps `[B, T]` to
    time-major itself at :147).  `eps_entropy` is the standard-normal draw the reference takes from `entropy_key`.

    Returns the reference's metrics (`total_loss`, `policy_loss`, `v_loss`, `kl_latent_loss`, `entropy_loss`: 0-d CUDA tensors,
    views of `losses`), `vs`, `advantages` (normalised when asked) and `d_logits`, `d_latent_mean`, `d_latent_logvar`,
    `d_baseline` = d total_loss / d input.  One call = five kernel launches (`tmjx_ppo_loss_head`); no CPU fallback."""
    import torch

    if not policy_logits.is_cuda:
        raise RuntimeError("ppo_loss_head needs CUDA tensors: there is no CPU fallback")
    T, B, A2 = policy_logits.shape
    A, Lz = A2 // 2, latent_mean.shape[-1]
    cont = lambda a: a.to(torch.float32).contiguous()
    ins = [cont(a) for a in (policy_logits, latent_mean, latent_logvar, baseline, bootstrap_value, reward, discount, truncation,
                             raw_action, behaviour_log_prob, eps_entropy)]
    want = [(T, B, 2 * A), (T, B, Lz), (T, B, Lz), (T, B), (B,), (T, B), (T, B), (T, B), (T, B, A), (T, B), (T, B, A)]
    for a, w in zip(ins, want):
        if tuple(a.shape) != w:
            raise ValueError(f"ppo_loss_head: got shape {tuple(a.shape)}, expected {w}")
    f = dict(dtype=torch.float32, device=policy_logits.device)
    lib = L.load()
    losses, vs, adv = torch.empty(8, **f), torch.empty(T, B, **f), torch.empty(T, B, **f)
    d_logits, d_mean, d_logvar, d_base = torch.empty_like(ins[0]), torch.empty_like(ins[1]), torch.empty_like(ins[2]), torch.empty(T, B, **f)
    scratch = torch.empty(int(lib.tmjx_ppo_loss_scratch_floats(T, B)), **f)
    hp = _PpoHyper(entropy_cost, kl_weight, discounting, reward_scaling, gae_lambda, clipping_epsilon, int(bool(normalize_advantage)))
    ptr = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.tmjx_ppo_loss_head(*[ptr(a) for a in ins], int(T), int(B), int(A), int(Lz), C.cast(C.byref(hp), C.c_void_p), ptr(losses),
                                ptr(vs), ptr(adv), ptr(d_logits), ptr(d_mean), ptr(d_logvar), ptr(d_base), ptr(scratch),
                                C.c_void_p(torch.cuda.current_stream(policy_logits.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"tmjx_ppo_loss_head failed ({rc}): {lib.tmjx_policy_last_error().decode()}")
    return {"total_loss": losses[0], "policy_loss": losses[1], "v_loss": losses[2], "kl_latent_loss": losses[3], "entropy_loss": losses[4],
            "kl_weight": kl_weight, "losses": losses, "vs": vs, "advantages": adv, "d_logits": d_logits, "d_latent_mean": d_mean,
            "d_latent_logvar": d_logvar, "d_baseline": d_base}


def create_ramp_schedule(max_value: float = 0.1, min_value: float = 0.0001, ramp_steps: int = 1000, warmup_steps: int = 0,
                         schedule: str = "linear", period: int = 45):
    """KL-weight schedule of `track_mjx/agent/mlp_ppo/losses.py:248-290` (host arithmetic, float32 like the reference; its quirks
    kept: the linear progress is clipped below at `min_value`, not 0, and the cyclic forms add `min_value` to the midpoint)."""
    import numpy as np

    f = np.float32
    if schedule not in ("linear", "cosine", "sine"):
        raise ValueError(f"schedule must be either 'linear' 'cosine', or 'sine', not {schedule}")

    def schedule_fn(step):
        step = f(step)
        if schedule == "linear":
            progress = np.clip((step - f(warmup_steps)) / f(ramp_steps), f(min_value), f(1))
            return f(min_value) if step < warmup_steps else f(progress * f(max_value))
        amplitude, midpoint = f((max_value - min_value) / 2), f((max_value + min_value) / 2)
        angle = (f(2 * np.pi) * step) / f(period)
        if schedule == "cosine":
            return f(midpoint + f(min_value) + amplitude * np.cos(angle, dtype=f))
        return f(midpoint + f(min_value) + amplitude * np.sin(angle - f(np.pi / 2), dtype=f))

    return schedule_fn


class Adam:
    """`optax.chain(optax.clip_by_global_norm(max_grad_norm), optax.adam(learning_rate))` of `ppo.py:517-520` on one flat fp32
    CUDA parameter buffer (`tmjx_adam_step`).  `step(grads)` updates `params` in place; with an initialised `torch.distributed`
    group the gradients are SUM all-reduced (NCCL) first and scaled by 1 / world_size inside the kernel, which is the reference's
    `jax.lax.pmean(grads)` (brax `gradients.py`, called at ppo.py:371-376).  No CPU fallback."""

    def __init__(self, params, learning_rate: float = 1e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
                 max_grad_norm: float = 10.0):
        import torch

        if not params.is_cuda or params.dtype != torch.float32 or not params.is_contiguous():
            raise RuntimeError("Adam needs a contiguous float32 CUDA parameter buffer: there is no CPU fallback")
        self.torch, self.lib, self.params = torch, L.load(), params
        self.mu, self.nu = torch.zeros_like(params), torch.zeros_like(params)
        self.count = 0
        self.hp = (float(learning_rate), float(b1), float(b2), float(eps), float(max_grad_norm))
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=params.device)
        self._scratch = torch.zeros(int(self.lib.tmjx_adam_scratch_floats()), dtype=torch.float32, device=params.device)

    def step(self, grads, all_reduce: bool | None = None, grad_scale: float | None = None):
        """`grads` is all-reduced IN PLACE (it holds the SUM over ranks afterwards; the 1 / world_size of `pmean` is applied inside the
        kernel, not to the buffer).  A caller that has already all-reduced (e.g. bucket by bucket, overlapped with the backward pass)
        passes `all_reduce=False, grad_scale=1 / world_size`."""
        t = self.torch
        if grads.shape != self.params.shape or grads.dtype != t.float32 or not grads.is_cuda or grads.device != self.params.device:
            raise ValueError("grads must be a float32 CUDA tensor shaped like params, on the same device")
        if not grads.is_contiguous():
            raise ValueError("grads must be contiguous (the all-reduce and the kernel work in place)")
        dist = t.distributed
        if all_reduce is None:
            all_reduce = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        scale = 1.0 if grad_scale is None else float(grad_scale)
        if all_reduce:
            dist.all_reduce(grads)
            scale = 1.0 / dist.get_world_size()
        self.count += 1
        ptr = lambda a: C.c_void_p(a.data_ptr())
        rc = self.lib.tmjx_adam_step(ptr(self.params), ptr(grads), ptr(self.mu), ptr(self.nu), self.params.numel(), *self.hp, scale,
                                     self.count, ptr(self.grad_norm), ptr(self._scratch),
                                     C.c_void_p(t.cuda.current_stream(self.params.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"tmjx_adam_step failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")
        return self.params


class RunningStatistics:
    """`RunningStatisticsState` + `update` of `track_mjx/agent/masked_running_statistics.py:34-214` on the GPU (no mask / weights,
    the way `ppo.py:357-361` calls it).  `update(batch)` accepts any leading batch dimensions; with an initialised
    `torch.distributed` process group the column sums and the row count are all-reduced (NCCL) between the two kernels, which is
    what `pmap_axis_name` does in the reference.  `mean` / `std` are what `IntentionPolicy` takes as `norm/mean`, `norm/std`."""

    def __init__(self, size: int, device: int = 0, std_min_value: float = 1e-6, std_max_value: float = 1e6):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("RunningStatistics needs a CUDA device: there is no CPU fallback")
        self.torch, self.lib, self.D = torch, L.load(), int(size)
        dev = torch.device("cuda", device)
        f = dict(dtype=torch.float32, device=dev)
        self.count = torch.zeros(1, **f)
        self.mean = torch.zeros(size, **f)
        self.summed_variance = torch.zeros(size, **f)
        self.std = torch.ones(size, **f)
        self._buf = torch.zeros(3 * size + 1, **f)           # row count | sum(x - mean) | batch mean | batch M2
        self._scratch = torch.zeros(int(self.lib.tmjx_running_stats_scratch_floats(size)), **f)
        self.std_min, self.std_max = float(std_min_value), float(std_max_value)
        self._frozen = None

    def freeze_tail(self, mean, std, summed_variance):
        """Keep the last `len(mean)` features (the proprioceptive block) at frozen statistics: after every `update` they are written
        back over the tail of mean / std / summed_variance, as `ppo.py:364-382` does with `frozen_proprioceptive_normalizer_params`."""
        t = self.torch
        self._frozen = tuple(t.as_tensor(a, dtype=t.float32, device=self.mean.device).reshape(-1).clone() for a in (mean, std, summed_variance))
        if not (0 < self._frozen[0].numel() <= self.D) or any(a.numel() != self._frozen[0].numel() for a in self._frozen):
            raise ValueError("frozen statistics must be three vectors of the same length <= the observation size")
        return self

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")

    def update(self, batch, all_reduce: bool | None = None):
        t, lib, D = self.torch, self.lib, self.D
        if batch.device != self.mean.device:
            raise ValueError(f"batch is on {batch.device}, the statistics on {self.mean.device}")
        x = batch.reshape(-1, D).to(t.float32).contiguous()
        n = int(x.shape[0])
        st = C.c_void_p(t.cuda.current_stream(x.device).cuda_stream)
        ptr = lambda a, off=0: C.c_void_p(a.data_ptr() + 4 * off)
        dist = t.distributed
        if all_reduce is None:
            all_reduce = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        sums, inc = ptr(self._buf, 1), ptr(self._buf, 0)
        self._check(lib.tmjx_running_stats_sums(ptr(x), n, D, ptr(self.mean), sums, ptr(self._scratch), st), "tmjx_running_stats_sums")
        self._buf[0] = float(n)
        if all_reduce:
            dist.all_reduce(self._buf[: D + 1])               # SUM over ranks: row count and sum(x - mean)   (psum at :163-166)
        self._check(lib.tmjx_running_stats_mean(sums, inc, n, D, ptr(self.count), ptr(self.mean), ptr(self._scratch), st),
                    "tmjx_running_stats_mean")
        if all_reduce:
            dist.all_reduce(self._buf[1 : D + 1])             # SUM over ranks: variance update                (psum at :176-177)
        self._check(lib.tmjx_running_stats_apply(sums, D, self.std_min, self.std_max, ptr(self.count), ptr(self.summed_variance),
                                                 ptr(self.std), ptr(self._scratch), st), "tmjx_running_stats_apply")
        if self._frozen is not None:
            k = self._frozen[0].numel()
            self.mean[-k:].copy_(self._frozen[0]); self.std[-k:].copy_(self._frozen[1]); self.summed_variance[-k:].copy_(self._frozen[2])
        return self


class _DevArray:
    """A library-owned device buffer exposed through `__cuda_array_interface__` so that torch can alias it without a copy."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class Trainer:
    """Forward + backward of the intention network and the value network for one PPO minibatch on the GPU (`tmjx_trainer_*`,
    csrc/tmjx_train.cuh): the `jax.value_and_grad(loss_fn)` half of the reference's `gradient_update_fn` (`ppo.py:263-272, 621-623`).

    `params` / `grads` are ONE flat fp32 CUDA tensor each (views of buffers the library owns): [policy vector | value vector], every
    vector in the layout of `policy.flatten_params` / `ValueNetwork` (normaliser mean, std first; their gradient is zero), so the NCCL
    all-reduce and `Adam.step` act on a single tensor.  After changing `params` (optimiser step, normaliser update) call `sync()`.
    No CPU fallback."""

    def __init__(self, cfg, policy_params, value_params, value_layers, max_rows: int, device: int = 0):
        import numpy as np
        import torch

        from . import policy as P

        if not torch.cuda.is_available():
            raise RuntimeError("Trainer needs a CUDA device: there is no CPU fallback")
        self.torch, self.lib, self.cfg = torch, L.load(), cfg
        self.device = torch.device("cuda", device)
        self.max_rows, self.value_layers = int(max_rows), tuple(int(x) for x in value_layers)
        pd = P.make_desc(cfg)
        vd = L.ValueDescC()
        vd.obs_size, vd.n_hidden_layers = int(cfg.obs_size), len(self.value_layers)
        for i, n in enumerate(self.value_layers):
            vd.hidden_layers[i] = n
        flat_p = P.flatten_params(cfg, policy_params)
        parts = [value_params["norm/mean"], value_params["norm/std"]]
        for i in range(len(self.value_layers) + 1):
            parts += [value_params[f"hidden_{i}/kernel"], value_params[f"hidden_{i}/bias"]]
        flat_v = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).ravel() for a in parts]))
        self._t = C.c_void_p()
        fp = C.POINTER(C.c_float)
        with torch.cuda.device(self.device):
            rc = self.lib.tmjx_trainer_create(C.byref(pd), C.byref(vd), flat_p.ctypes.data_as(fp), flat_v.ctypes.data_as(fp), device, self.max_rows,
                                              C.byref(self._t))
        self._check(rc, "tmjx_trainer_create")
        self.n_params = int(self.lib.tmjx_trainer_param_count(self._t))
        self.n_policy = int(self.lib.tmjx_trainer_policy_param_count(self._t))
        pp, gp = C.c_void_p(), C.c_void_p()
        self._check(self.lib.tmjx_trainer_buffers(self._t, C.byref(pp), C.byref(gp)), "tmjx_trainer_buffers")
        self.params = torch.as_tensor(_DevArray(pp.value, self.n_params), device=self.device)
        self.grads = torch.as_tensor(_DevArray(gp.value, self.n_params), device=self.device)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.tmjx_policy_last_error().decode()}")

    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _arg(self, x, shape):
        t = self.torch
        if x.dtype != t.float32 or not x.is_contiguous() or x.device != self.device or tuple(x.shape) != tuple(shape):
            raise ValueError(f"expected a contiguous float32 tensor of shape {tuple(shape)} on {self.device}, got {tuple(x.shape)} {x.dtype} {x.device}")
        return C.c_void_p(x.data_ptr())

    def set_normalizer(self, mean, std):
        """`normalizer_params` of both networks <- the running statistics (`ppo.py:357-383` hands the same state to policy and value)."""
        D = self.cfg.obs_size
        for base in (0, self.n_policy):
            self.params[base:base + D].copy_(mean)
            self.params[base + D:base + 2 * D].copy_(std)

    def sync(self):
        self._check(self.lib.tmjx_trainer_sync(self._t, self._st()), "tmjx_trainer_sync")

    def policy_forward(self, obs, eps_latent):
        t, c = self.torch, self.cfg
        rows = int(obs.shape[0])
        f = dict(dtype=t.float32, device=self.device)
        logits, mean, logvar = t.empty(rows, 2 * c.action_size, **f), t.empty(rows, c.latent_size, **f), t.empty(rows, c.latent_size, **f)
        self._check(self.lib.tmjx_trainer_policy_forward(self._t, self._arg(obs, (rows, c.obs_size)), self._arg(eps_latent, (rows, c.latent_size)), rows,
                                                         C.c_void_p(logits.data_ptr()), C.c_void_p(mean.data_ptr()), C.c_void_p(logvar.data_ptr()),
                                                         self._st()), "tmjx_trainer_policy_forward")
        return logits, mean, logvar

    def policy_backward(self, d_logits, d_latent_mean, d_latent_logvar):
        c = self.cfg
        rows = int(d_logits.shape[0])
        self._check(self.lib.tmjx_trainer_policy_backward(self._t, self._arg(d_logits, (rows, 2 * c.action_size)),
                                                          self._arg(d_latent_mean, (rows, c.latent_size)), self._arg(d_latent_logvar, (rows, c.latent_size)),
                                                          rows, self._st()), "tmjx_trainer_policy_backward")

    def value_forward(self, obs):
        t = self.torch
        rows = int(obs.shape[0])
        v = t.empty(rows, dtype=t.float32, device=self.device)
        self._check(self.lib.tmjx_trainer_value_forward(self._t, self._arg(obs, (rows, self.cfg.obs_size)), rows, C.c_void_p(v.data_ptr()), self._st()),
                    "tmjx_trainer_value_forward")
        return v

    def value_backward(self, d_value):
        rows = int(d_value.shape[0])
        self._check(self.lib.tmjx_trainer_value_backward(self._t, self._arg(d_value, (rows,)), rows, self._st()), "tmjx_trainer_value_backward")

    def close(self):
        if getattr(self, "_t", None):
            self.params = self.grads = None
            self.lib.tmjx_trainer_destroy(self._t)
            self._t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
