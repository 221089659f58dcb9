"""PPO training step on the B200 kernels: host-side mirror of the reference's `training_step` / `sgd_step` / `minibatch_step`
(`track_mjx/agent/mlp_ppo/ppo.py:279-395`), the caller of the env hot path in BASELINE `configs[3]`.

One `training_step()` is, per GPU (one process per GPU, env shard = contiguous block of `num_envs / world_size` environments, as
`ppo.py:453,478-479`):
  1. `batch_size * num_minibatches // num_envs` unrolls of `unroll_length` env steps with the acting policy (`rollout.Rollout`:
     the step kernel and the policy kernels write the `[T, B, ...]` rollout in place)                          ppo.py:330-354
  2. observation-normaliser update over the rollout's observations, two small NCCL all-reduces              ppo.py:357-383
  3. `num_updates_per_batch` epochs x `num_minibatches` minibatches (a fresh env permutation per epoch):     ppo.py:279-318
     network forward (training mode; policy and critic on two streams) -> PPO loss head -> backward (two streams) -> NCCL all-reduce
     of the flat gradient (the value-network bucket goes out while the policy backward is still running) -> global-norm clip + Adam ->
     operand refresh
  4. the acting policy picks the new parameters up in place.
Torch is the plumbing (device buffers, the permutation gather, `torch.distributed` over NCCL); every FLOP of the networks, the
loss, the optimiser and the env runs in this repository's kernels.  No CPU fallback.
"""
from __future__ import annotations

import dataclasses
from typing import Sequence

import torch

from . import learner as LN
from . import policy as P
from .rollout import Rollout


@dataclasses.dataclass
class PPOConfig:
    """train_config + network_config of config/rodent-full-clips.yaml:50-90 (defaults = the shipped values)."""
    unroll_length: int = 20
    num_minibatches: int = 16
    num_updates_per_batch: int = 4
    unrolls_per_step: int = 1                      # batch_size * num_minibatches // num_envs
    learning_rate: float = 1e-4
    entropy_cost: float = 1e-2
    discounting: float = 0.98
    reward_scaling: float = 1.0
    clipping_epsilon: float = 0.2
    gae_lambda: float = 0.95
    normalize_advantage: bool = True
    kl_weight: float = 1e-1
    use_kl_schedule: bool = True
    kl_ramp_steps: int = 1000
    max_grad_norm: float = 10.0
    critic_layer_sizes: Sequence[int] = (512, 512, 512, 512, 512, 256)
    seed: int = 0


class PPO:
    def __init__(self, env, net_cfg: P.IntentionNetworkConfig | None = None, cfg: PPOConfig | None = None):
        self.env, self.cfg = env, cfg or PPOConfig()
        self.net_cfg = net_cfg or P.IntentionNetworkConfig(obs_size=env.observation_size, action_size=env.action_size)
        c, dev = self.cfg, env.device
        B = env.num_envs
        if B % c.num_minibatches:
            raise ValueError("num_envs per GPU must be divisible by num_minibatches")
        self.B, self.Bm, self.T = B, B * c.unrolls_per_step // c.num_minibatches, c.unroll_length
        dist = torch.distributed
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        pol_params = P.init_params(self.net_cfg, c.seed)                       # identical on every rank (same seed), like pmap's replicated init
        val_params = P.init_value_params(self.net_cfg.obs_size, c.critic_layer_sizes, c.seed + 1)
        self.trainer = LN.Trainer(self.net_cfg, pol_params, val_params, c.critic_layer_sizes, max_rows=self.T * self.Bm, device=dev.index)
        self.policy = P.IntentionPolicy(self.net_cfg, pol_params, max_env=B, device=dev.index)
        self.stats = LN.RunningStatistics(self.net_cfg.obs_size, device=dev.index)
        self.adam = LN.Adam(self.trainer.params, learning_rate=c.learning_rate, max_grad_norm=c.max_grad_norm)
        self.rollouts = [Rollout(env, self.policy, self.T) for _ in range(c.unrolls_per_step)]
        self.kl_schedule = LN.create_ramp_schedule(max_value=c.kl_weight, ramp_steps=c.kl_ramp_steps) if c.use_kl_schedule else None
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(c.seed * 1000003 + (dist.get_rank() if self.world > 1 else 0))     # per-rank noise, like pmap's split keys
        self.perm_gen = torch.Generator(device=dev)
        self.perm_gen.manual_seed(c.seed + 17)
        self.it = 0
        self.state = None
        from .sharding import GradientBuckets

        self.buckets = GradientBuckets(self.trainer.grads, [self.trainer.n_policy])     # bucket 0 = policy vector, 1 = value vector
        self.value_stream = torch.cuda.Stream(device=dev)
        self.timing, self.marks = False, []
        self.all_reduce = self.world > 1

    # ------------------------------------------------------------------------------------------------------------------
    def reset(self, seed: int = 0):
        self.state = self.env.reset(seed)
        return self.state

    def _minibatch(self, data, idx):
        """Rows of minibatch `idx` (env indices) as contiguous [T * Bm, ...] tensors (time-major, like the loss head wants)."""
        sel = lambda x: x.index_select(1, idx)
        T, Bm = self.T, idx.numel()
        obs = sel(data["observation"]).reshape(T * Bm, -1)
        return dict(
            obs=obs, next_obs_last=data["next_observation_last"].index_select(0, idx),
            reward=sel(data["reward"]), discount=sel(data["discount"]), truncation=sel(data["truncation"]),
            raw_action=sel(data["raw_action"]), log_prob=sel(data["log_prob"]))

    def _update_minibatch(self, mb):
        c, tr, T = self.cfg, self.trainer, self.T
        Bm = mb["reward"].shape[1]
        rows = T * Bm
        A, Lz = self.net_cfg.action_size, self.net_cfg.latent_size
        eps_z = torch.randn(rows, Lz, device=self.env.device, generator=self.gen)                # policy_key  (losses.py:143, 148-150)
        eps_e = torch.randn(T, Bm, A, device=self.env.device, generator=self.gen)                # entropy_key
        # The two networks are independent until the loss head and after it: they run on two streams (a 10240-row minibatch is 80 CTAs
        # per GEMM on 148 SMs; the policy's and the critic's GEMMs together fill the machine), each with its own backward scratch.
        main = torch.cuda.current_stream(self.env.device)
        sv = self.value_stream
        sv.wait_stream(main)
        with torch.cuda.stream(sv):
            bootstrap = tr.value_forward(mb["next_obs_last"])                                    # losses.py:154-156 (no gradient reaches it)
            baseline = tr.value_forward(mb["obs"])
        logits, mean, logvar = tr.policy_forward(mb["obs"], eps_z)
        main.wait_stream(sv)
        kl_w = float(self.kl_schedule(self.it)) if self.kl_schedule else c.kl_weight
        out = LN.ppo_loss_head(logits.view(T, Bm, 2 * A), mean.view(T, Bm, Lz), logvar.view(T, Bm, Lz), baseline.view(T, Bm), bootstrap,
                               mb["reward"], mb["discount"], mb["truncation"], mb["raw_action"], mb["log_prob"], eps_e,
                               entropy_cost=c.entropy_cost, kl_weight=kl_w, discounting=c.discounting, reward_scaling=c.reward_scaling,
                               gae_lambda=c.gae_lambda, clipping_epsilon=c.clipping_epsilon, normalize_advantage=c.normalize_advantage)
        d_base = out["d_baseline"].reshape(rows)
        sv.wait_stream(main)
        with torch.cuda.stream(sv):
            tr.value_backward(d_base)
            if self.all_reduce:
                self.buckets.reduce(1)    # value-network bucket: reduced on NCCL's stream as soon as the critic's backward is done
        tr.policy_backward(out["d_logits"].view(rows, 2 * A), out["d_latent_mean"].view(rows, Lz), out["d_latent_logvar"].view(rows, Lz))
        main.wait_stream(sv)
        bootstrap.record_stream(main); baseline.record_stream(main)      # allocator: allocated on the side stream, read by the loss head on main
        d_base.record_stream(sv)                                         # allocated on main, read by the critic's backward on the side stream
        scale = 1.0
        if self.all_reduce:
            self.buckets.reduce(0)
            scale = self.buckets.wait()                                            # the compute stream waits, the host does not
        self.adam.step(tr.grads, all_reduce=False, grad_scale=scale)                # pmean = SUM (bucket by bucket) / world_size
        tr.sync()
        return out["losses"]

    def training_step(self):
        """One `training_step` of the reference (ppo.py:320-395).  Returns the losses of the last minibatch (device tensor [8])."""
        c, T = self.cfg, self.T
        if self.state is None:
            self.reset(c.seed)
        self._mark("start")
        # 1. acting: unrolls_per_step x unroll_length env steps on this GPU's env shard
        trs = []
        for r in self.rollouts:
            self.state, tr = r.generate(self.state, generator=self.gen)
            trs.append(tr)
        cat = (lambda f: torch.cat([f(t) for t in trs], dim=1)) if len(trs) > 1 else (lambda f: f(trs[0]))
        data = dict(
            observation=cat(lambda t: t.observation), next_observation_last=cat(lambda t: t.next_observation[-1:])[0],
            reward=cat(lambda t: t.reward), discount=cat(lambda t: t.discount),
            truncation=cat(lambda t: t.extras["state_extras"]["truncation"]),
            raw_action=cat(lambda t: t.extras["policy_extras"]["raw_action"]), log_prob=cat(lambda t: t.extras["policy_extras"]["log_prob"]))
        self._mark("acting")
        # 2. normaliser update (two psum's inside), handed to both networks
        self.stats.update(data["observation"], all_reduce=self.all_reduce)
        self.trainer.set_normalizer(self.stats.mean, self.stats.std)
        self.trainer.sync()
        self._mark("normalizer")
        # 3. SGD epochs
        n_traj = data["reward"].shape[1]
        losses = None
        for _ in range(c.num_updates_per_batch):
            perm = torch.randperm(n_traj, device=self.env.device, generator=self.perm_gen)
            for i in range(c.num_minibatches):
                idx = perm[i * self.Bm:(i + 1) * self.Bm]
                losses = self._update_minibatch(self._minibatch(data, idx))
        # 4. the actor picks up the new parameters (and the new normaliser)
        self.policy.set_params(self.trainer.params[: self.trainer.n_policy])
        self._mark("sgd")
        self.it += 1
        return losses

    def _mark(self, name):
        """CUDA-event phase marks of the current training step (bench.py reads `phase_ms()` after a synchronize)."""
        if self.timing:
            if name == "start":
                self.marks = []
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def phase_ms(self):
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks[:-1], self.marks[1:])}

    def env_steps_per_training_step(self) -> int:
        return self.B * self.T * self.cfg.unrolls_per_step * self.world

    def close(self):
        self.trainer.close()
        self.policy.close()
