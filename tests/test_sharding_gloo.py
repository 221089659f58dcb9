"""N > 1 host logic on CPU: world_size-2 `gloo` process group -- shard arithmetic, per-rank seeds, the episode
statistics all-reduce and the max-over-ranks timing rule (the only collectives of the env path)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, global_envs, out):
    import torch.distributed as dist

    from track_mjx_b200.sharding import Shard, max_over_ranks, reduce_episode_stats

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = Shard(rank, world, global_envs)
    ids = torch.tensor(list(sh.env_ids()), dtype=torch.float64)
    # a per-env "reward" that depends only on the GLOBAL env id: the whole-job statistics must not depend on the sharding
    rew = (ids * 0.5 + 1.0).sum()
    done = (ids % 3 == 0).double().sum()
    stats = reduce_episode_stats(rew, done, n_steps=1, shard=sh)
    tmax = max_over_ranks(10.0 + rank, torch.device("cpu"), sh)
    out.put((rank, sh.start, sh.count, sh.seed(7), stats, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_collectives():
    world, global_envs = 2, 4097   # odd on purpose: remainder goes to rank 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, seed0, st0, t0), (r1, s1, c1, seed1, st1, t1) = res
    assert (s0, c0, s1, c1) == (0, 2049, 2049, 2048) and s1 == s0 + c0 and c0 + c1 == global_envs
    assert seed0 != seed1
    ids = torch.arange(global_envs, dtype=torch.float64)
    want_r = float((ids * 0.5 + 1.0).sum()) / global_envs
    want_d = float((ids % 3 == 0).double().sum()) / global_envs
    for st in (st0, st1):
        assert abs(st["mean_reward"] - want_r) < 1e-9 and abs(st["done_frac"] - want_d) < 1e-12
    assert t0 == t1 == 11.0


def test_shard_partition_properties():
    from track_mjx_b200.sharding import Shard

    for n in (1, 7, 64, 4096, 65536):
        for w in (1, 2, 4, 8):
            shards = [Shard(r, w, n) for r in range(w)]
            assert sum(s.count for s in shards) == n
            assert all(shards[i].stop == shards[i + 1].start for i in range(w - 1))
            assert max(s.count for s in shards) - min(s.count for s in shards) <= 1


def _bucket_worker(rank, world, port, out):
    import torch.distributed as dist

    from track_mjx_b200.sharding import GradientBuckets

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)        # rank r holds (r + 1) x [0..9]
    b = GradientBuckets(g, [4])
    b.reduce(1)                                                   # the later part of the buffer first, as the backward pass produces it
    g[:4] += 100.0                                                # "policy backward" still writing bucket 0 while bucket 1 is in flight
    b.reduce(0)
    scale = b.wait()
    out.put((rank, g.tolist(), scale, len(b)))
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_is_the_pmean():
    """The N > 1 learner path (BASELINE configs[3]): SUM all-reduce per bucket + 1 / world_size scale = jax.lax.pmean of the gradients."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = torch.arange(10, dtype=torch.float32)
    want = (base * 3).tolist()                                    # (1 + 2) x [0..9]
    want[:4] = [v + 200.0 for v in want[:4]]
    for rank, got, scale, nb in res:
        assert got == want and scale == 0.5 and nb == 2
    import pytest

    from track_mjx_b200.sharding import GradientBuckets
    with pytest.raises(ValueError):
        GradientBuckets(torch.zeros(8), [9])
    assert GradientBuckets(torch.zeros(8), [3]).wait() == 1.0     # no process group: identity, scale 1
