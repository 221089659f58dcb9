"""CPU restatement of the reference's observation-normaliser update -- TEST INFRASTRUCTURE (only tests/ may import it).

Follows `track_mjx/agent/masked_running_statistics.py:80-214` (`update`, no mask / weights: the way `ppo.py:357-361` calls it) and
`:217-236` (`normalize`); pinned by `tests/golden/running_stats.npz` (outputs of the reference's own module text,
tools/make_golden_running_stats.py).  float32 throughout, like the reference without jax_enable_x64.
"""
import numpy as np


def update(count, mean, summed_variance, batch, std_min_value=1e-6, std_max_value=1e6):
    f = np.float32
    batch = np.asarray(batch, f).reshape(-1, mean.shape[0])
    count = f(count) + f(batch.shape[0])                                   # :139-146 (step_increment = prod(batch_dims))
    diff_to_old_mean = batch - mean                                        # :162
    mean_update = np.sum(diff_to_old_mean, axis=0) / count                 # :168
    new_mean = (mean + mean_update).astype(f)                              # :171
    diff_to_new_mean = batch - new_mean                                    # :173
    variance_update = np.sum(diff_to_old_mean * diff_to_new_mean, axis=0)  # :174-175
    new_sv = (summed_variance + variance_update).astype(f)                 # :178
    std = np.sqrt(np.maximum(new_sv, 0) / count)                           # :190-193
    std = np.clip(std, f(std_min_value), f(std_max_value)).astype(f)
    return count, new_mean, new_sv, std


def normalize(batch, mean, std):
    return ((np.asarray(batch, np.float32) - mean) / std).astype(np.float32)   # :230
