"""Multi-GPU check of the learner's two exchange steps (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_learner_2gpu.py

1. `RunningStatistics.update` with each rank holding a different shard of the observations (two all-reduces of [D + 1] and [D]
   floats between the kernels = the reference's `psum`s, masked_running_statistics.py:163-177) must equal the single-GPU update over
   the concatenated batch (2e-6 relative + 2e-6 absolute on mean, 5e-6 on std: the merge order differs), two updates in a row.
2. `Adam.step` with a different gradient on each rank (SUM all-reduce + 1 / world_size = `pmean`, then clip + Adam) must equal a
   local step on the mean gradient (2e-6), and the parameters must stay bitwise identical across ranks.
Rank 0 prints one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_mjx_b200.learner import Adam, RunningStatistics  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.cuda.current_device()
    dist.init_process_group("nccl")
    D, rows = 696, 20 * 1024
    rng = np.random.default_rng(0)                      # the same stream on every rank
    res = {"world": world}
    sharded, whole = RunningStatistics(D, device=dev), RunningStatistics(D, device=dev)
    for it in range(2):
        full = torch.from_numpy((rng.normal(size=(world * rows, D)) * (1.5 + it) + 0.7 - it).astype(np.float32)).cuda()
        sharded.update(full[rank * rows:(rank + 1) * rows])          # NCCL all-reduce inside
        whole.update(full, all_reduce=False)
        torch.cuda.synchronize()
        assert float(sharded.count.item()) == float(whole.count.item()) == (it + 1) * world * rows
        assert torch.allclose(sharded.mean, whole.mean, rtol=2e-6, atol=2e-6), (sharded.mean - whole.mean).abs().max()
        assert torch.allclose(sharded.std, whole.std, rtol=5e-6), ((sharded.std - whole.std) / whole.std).abs().max()
        res[f"stats_update{it}_max_rel_std"] = float(((sharded.std - whole.std) / whole.std).abs().max())
    n = 2_600_003
    p0 = torch.from_numpy(rng.normal(size=n).astype(np.float32)).cuda()
    grads = [torch.from_numpy((rng.normal(size=n) * (3.0 if k == 0 else 0.01)).astype(np.float32)).cuda() for k in range(world * 2)]
    a, b = Adam(p0.clone(), learning_rate=1e-3), Adam(p0.clone(), learning_rate=1e-3)
    for it in range(2):                                  # both steps clip (norms 2420 and 18 > 10); the unclipped path: tests/test_optimizer.py
        mine = grads[it * world + rank] if it == 0 else grads[it * world + rank] * (rank + 1)
        every = [grads[it * world + r] if it == 0 else grads[it * world + r] * (r + 1) for r in range(world)]
        a.step(mine.clone())                             # all-reduce inside
        b.step(torch.stack(every).sum(0) / world, all_reduce=False)
        torch.cuda.synchronize()
        assert torch.allclose(a.params, b.params, rtol=2e-6, atol=2e-7), (a.params - b.params).abs().max()
        assert abs(float(a.grad_norm.item()) - float(b.grad_norm.item())) <= 1e-5 * float(b.grad_norm.item())
        res[f"adam_step{it}_grad_norm"] = float(a.grad_norm.item())
    gathered = [torch.empty_like(a.params) for _ in range(world)]
    dist.all_gather(gathered, a.params)
    assert all(torch.equal(gathered[0], g) for g in gathered)
    res["params_bitwise_identical_across_ranks"] = True
    dist.barrier()
    if rank == 0:
        res["ok"] = True
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
