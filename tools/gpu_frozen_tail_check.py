"""Tiny GPU check of RunningStatistics.freeze_tail (ppo.py:364-382): the tail keeps the frozen statistics across updates, the head
matches an unfrozen instance bitwise."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_mjx_b200.learner import RunningStatistics  # noqa: E402

x = torch.randn(4096, 696, device="cuda") * 2 + 1
a, b = RunningStatistics(696), RunningStatistics(696)
fm, fs, fv = torch.full((226,), 0.5), torch.full((226,), 2.0), torch.full((226,), 7.0)
b.freeze_tail(fm, fs, fv)
for _ in range(2):
    a.update(x); b.update(x)
torch.cuda.synchronize()
assert torch.equal(a.mean[:470], b.mean[:470]) and torch.equal(a.std[:470], b.std[:470])
assert torch.equal(b.mean[470:].cpu(), fm) and torch.equal(b.std[470:].cpu(), fs) and torch.equal(b.summed_variance[470:].cpu(), fv)
assert float(b.count.item()) == 8192.0
print("frozen tail ok")
