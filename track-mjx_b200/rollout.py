"""Acting loop with zero-copy rollout buffers: the env <-> policy hand-off of `acting.generate_unroll` as the reference's PPO
calls it (`track_mjx/agent/mlp_ppo/ppo.py:330-354`, upstream brax 0.12.3 `training/acting.py`).

The reference scans `actor_step` and then transposes / reshapes the stacked Transition pytree (`ppo.py:350-353`).  Here the
step kernel writes observation t+1 straight into slot t+1 of the `[T+1, B, obs]` buffer (TmjxOut.obs is a caller-owned
pointer) and the policy kernels write action / raw_action / log_prob / logits into slot t of theirs, so a Transition is a set
of views: `observation = obs[:-1]`, `next_observation = obs[1:]`, no copies of the 2.8 KB/env observation.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch

from .env import MultiClipTracking, State
from .policy import IntentionPolicy


@dataclass
class Transition:
    """brax.training.types.Transition with time-major `[T, B, ...]` leaves (views into the Rollout's buffers)."""
    observation: torch.Tensor
    action: torch.Tensor
    reward: torch.Tensor
    discount: torch.Tensor
    next_observation: torch.Tensor
    extras: Dict[str, Dict[str, torch.Tensor]]


class Rollout:
    def __init__(self, env: MultiClipTracking, policy: IntentionPolicy, unroll_length: int):
        self.env, self.policy, self.T = env, policy, int(unroll_length)
        B, dev = env.num_envs, env.device
        f = dict(dtype=torch.float32, device=dev)
        a, z = policy.cfg.action_size, policy.cfg.latent_size
        T = self.T
        self.obs = torch.zeros(T + 1, B, env.observation_size, **f)
        self.action = torch.zeros(T, B, a, **f)
        self.raw_action = torch.zeros(T, B, a, **f)
        self.logits = torch.zeros(T, B, 2 * a, **f)
        self.log_prob = torch.zeros(T, B, **f)
        self.latent_mean = torch.zeros(T, B, z, **f)
        self.latent_logvar = torch.zeros(T, B, z, **f)
        self.reward = torch.zeros(T, B, **f)
        self.done = torch.zeros(T, B, **f)
        self.truncation = torch.zeros(T, B, **f)

    def generate(self, state: State, generator: torch.Generator | None = None, eps=None, deterministic: bool = False):
        """`unroll_length` actor steps from `state`.  `eps` = optional (eps_latent [T, B, latent], eps_action [T, B, action])
        for reproducible runs; otherwise N(0, 1) draws from `generator`.  Returns (final State, Transition)."""
        env, pol, T = self.env, self.policy, self.T
        B = env.num_envs
        self.obs[0].copy_(state.obs)                       # the only observation copy of the unroll
        for t in range(T):
            if deterministic:
                ez = ea = None
            elif eps is not None:
                ez, ea = eps[0][t], eps[1][t]
            else:
                ez = torch.randn(B, pol.cfg.latent_size, device=env.device, generator=generator)
                ea = torch.randn(B, pol.cfg.action_size, device=env.device, generator=generator)
            out = {"action": self.action[t], "raw_action": self.raw_action[t], "log_prob": self.log_prob[t], "logits": self.logits[t],
                   "latent_mean": self.latent_mean[t], "latent_logvar": self.latent_logvar[t]}
            pol.act(self.obs[t], ez, ea, deterministic=deterministic, out=out)
            env.stepper.redirect_obs(self.obs[t + 1])
            state = env.step(state, self.action[t])
            self.reward[t].copy_(state.reward)
            self.done[t].copy_(state.done)
            if "truncation" in state.info:
                self.truncation[t].copy_(state.info["truncation"])
        env.stepper.redirect_obs(None)
        env.stepper.buf["obs"].copy_(self.obs[T])          # keep the stepper's own view current for plain env.step callers
        state = env._state()
        tr = Transition(
            observation=self.obs[:-1], action=self.action, reward=self.reward, discount=1.0 - self.done, next_observation=self.obs[1:],
            extras={"policy_extras": {"raw_action": self.raw_action, "log_prob": self.log_prob, "logits": self.logits,
                                      "latent_mean": self.latent_mean, "latent_logvar": self.latent_logvar},
                    "state_extras": {"truncation": self.truncation}})
        return state, tr
