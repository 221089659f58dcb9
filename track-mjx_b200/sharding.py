"""Environment sharding across GPUs (one process per GPU) and the only collectives of the env path.

The reference shards envs over devices with `jax.pmap`: `num_envs // device_count` envs per device, per-device
reset keys (reference track_mjx/agent/mlp_ppo/ppo.py:242-257, 453, 477-480).  Environments are independent, so
stepping needs NO data-path collective; what crosses NVLink is the episode statistics (a few scalars, SUM) and,
for timing, the max over ranks.  With backend "nccl" the all-reduce runs over NVLink/NVSwitch; the same code runs
under "gloo" on CPU tensors, which is how tests/test_sharding_gloo.py covers the N > 1 path without GPUs.
"""
from __future__ import annotations

import dataclasses

import torch


@dataclasses.dataclass(frozen=True)
class Shard:
    """Contiguous block of global env ids owned by `rank` (ppo.py:453: envs are split evenly; the remainder, which
    the reference forbids by assertion, goes to the lowest ranks here)."""

    rank: int
    world: int
    global_envs: int

    @property
    def start(self) -> int:
        base, rem = divmod(self.global_envs, self.world)
        return self.rank * base + min(self.rank, rem)

    @property
    def count(self) -> int:
        base, rem = divmod(self.global_envs, self.world)
        return base + (1 if self.rank < rem else 0)

    @property
    def stop(self) -> int:
        return self.start + self.count

    def env_ids(self) -> range:
        return range(self.start, self.stop)

    def seed(self, base_seed: int) -> int:
        """Per-rank reset seed (the reference folds the process index into the key: ppo.py:446)."""
        return int(base_seed) * 1000003 + self.rank


def reduce_episode_stats(reward_sum: torch.Tensor, done_sum: torch.Tensor, n_steps: int, shard: Shard, group=None, nan_sum=None) -> dict:
    """Whole-job mean reward / done fraction (/ NaN fraction): SUM all-reduce of a few scalars (the only env-path collective)."""
    import torch.distributed as dist

    parts = [reward_sum.double().reshape(()), done_sum.double().reshape(())]
    if nan_sum is not None:
        parts.append(nan_sum.double().reshape(()))
    stats = torch.stack(parts)
    if shard.world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    denom = float(shard.global_envs * n_steps)
    out = {"mean_reward": float(stats[0].item()) / denom, "done_frac": float(stats[1].item()) / denom}
    if nan_sum is not None:
        out["nan_frac"] = float(stats[2].item()) / denom
    return out


def max_over_ranks(value: float, device, shard: Shard, group=None) -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-timed duration."""
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    if shard.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class GradientBuckets:
    """`jax.lax.pmean(grads, axis_name)` of the reference's `gradient_update_fn` (ppo.py:621-623, brax gradients.py) as bucketed SUM
    all-reduces of ONE flat gradient tensor: a bucket is reduced as soon as the backward pass has produced it (`reduce(i)` right
    after the kernels that write it were enqueued; the collective runs on the communicator's stream while later buckets are still
    being computed), `wait()` makes the compute stream wait for all of them and returns the 1 / world_size scale the optimiser
    applies (the buffer keeps the SUM).  Works on NCCL (CUDA tensors) and gloo (CPU tensors, tests/test_sharding_gloo.py)."""

    def __init__(self, grads: torch.Tensor, boundaries, group=None):
        if grads.dim() != 1 or not grads.is_contiguous():
            raise ValueError("grads must be a flat contiguous tensor")
        b = [0] + [int(x) for x in boundaries] + [grads.numel()]
        if any(b[i] >= b[i + 1] for i in range(len(b) - 1)):
            raise ValueError("bucket boundaries must be strictly increasing offsets inside the gradient buffer")
        self.views = [grads[b[i]:b[i + 1]] for i in range(len(b) - 1)]
        self.group, self._work = group, []
        import torch.distributed as dist

        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def __len__(self):
        return len(self.views)

    def reduce(self, i: int):
        if self.world > 1:
            import torch.distributed as dist

            self._work.append(dist.all_reduce(self.views[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self) -> float:
        for w in self._work:
            w.wait()
        self._work = []
        return 1.0 / self.world
