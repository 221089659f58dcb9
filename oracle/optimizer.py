"""CPU restatement of the reference's optimiser update -- TEST INFRASTRUCTURE (only tests/ may import it).

`track_mjx/agent/mlp_ppo/ppo.py:517-520` builds `optax.chain(optax.clip_by_global_norm(10.0), optax.adam(learning_rate))`; optax
(pinned 0.2.5 in the reference's pyproject.toml:28) is not vendored and not installable here, so this restates its published
algorithm: `clip_by_global_norm` (optax/_src/clipping.py: g_norm = sqrt(sum g^2); g unchanged where g_norm < max_norm, else
(g / g_norm) * max_norm) and `scale_by_adam` + `scale(-lr)` (optax/_src/transform.py: bias-corrected moments, eps outside the
square root, eps_root = 0).  PARITY UNPINNED against optax itself; cross-checked against torch.optim.Adam in tests.
"""
import numpy as np


def adam_step(params, grads, mu, nu, count, learning_rate=1e-4, b1=0.9, b2=0.999, eps=1e-8, max_grad_norm=10.0, dtype=np.float32):
    """count: the step number BEFORE this update (optax increments it first).  Returns (params, mu, nu, count, g_norm)."""
    f = dtype
    p, g, mu, nu = (np.asarray(a, f) for a in (params, grads, mu, nu))
    g_norm = f(np.sqrt(np.sum(np.square(g.astype(np.float64)))))
    if max_grad_norm and max_grad_norm > 0 and not g_norm < f(max_grad_norm):
        g = (g / g_norm) * f(max_grad_norm)
    mu = f(b1) * mu + (f(1) - f(b1)) * g
    nu = f(b2) * nu + (f(1) - f(b2)) * (g * g)
    count = count + 1
    c1, c2 = f(1) - f(b1) ** f(count), f(1) - f(b2) ** f(count)
    update = (mu / c1) / (np.sqrt(nu / c2) + f(eps))
    return (p + f(-learning_rate) * update).astype(f), mu.astype(f), nu.astype(f), count, g_norm
