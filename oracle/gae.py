"""CPU restatement of the reference's Generalised Advantage Estimation -- TEST INFRASTRUCTURE (only tests/ may import it).

Follows `track_mjx/agent/mlp_ppo/losses.py:39-101` (`compute_gae`) operation by operation in float32 (no fused multiply-add),
pinned by `tests/golden/gae.npz`, which holds outputs of the reference's own function text (tools/make_golden_gae.py).
"""
import numpy as np


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_=1.0, discount=0.99):
    f = np.float32
    truncation, termination, rewards, values = (np.asarray(a, f) for a in (truncation, termination, rewards, values))
    bootstrap_value = np.asarray(bootstrap_value, f)
    lambda_, discount = f(lambda_), f(discount)
    T = truncation.shape[0]
    truncation_mask = f(1) - truncation                                                   # losses.py:69
    values_t_plus_1 = np.concatenate([values[1:], bootstrap_value[None]], axis=0)         # :71-73
    deltas = rewards + discount * (f(1) - termination) * values_t_plus_1 - values         # :74
    deltas = deltas * truncation_mask                                                     # :75
    acc = np.zeros_like(bootstrap_value)
    vs_minus_v = np.zeros_like(values)
    for t in range(T - 1, -1, -1):                                                        # reverse scan, :80-92
        acc = deltas[t] + discount * (f(1) - termination[t]) * truncation_mask[t] * lambda_ * acc
        vs_minus_v[t] = acc
    vs = vs_minus_v + values                                                              # :94
    vs_t_plus_1 = np.concatenate([vs[1:], bootstrap_value[None]], axis=0)                 # :96
    advantages = (rewards + discount * (f(1) - termination) * vs_t_plus_1 - values) * truncation_mask   # :97-99
    return vs.astype(f), advantages.astype(f)
