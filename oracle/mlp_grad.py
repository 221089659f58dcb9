"""CPU restatement of the forward AND backward pass of the reference's networks -- TEST INFRASTRUCTURE (only tests/ may import it).

Forward: `IntentionNetwork.__call__` (`track_mjx/agent/mlp_ppo/intention_network.py:14-142`: Dense -> SiLU -> LayerNorm(eps 1e-6)
per hidden layer, fc2_mean / fc2_logvar heads, z = mean + eps exp(logvar / 2), decoder over [z | proprioception]) and brax's
`make_value_network` (Dense -> swish per hidden layer, Dense to 1).  Backward: the hand-derived chain the CUDA backward pass of the
next round has to implement (dgrad, wgrad, SiLU', LayerNorm backward, the reparameterisation's contribution to d mean / d logvar),
written layer by layer so that each kernel has its own checker.  The reference differentiates with `jax.grad` (losses are built at
`ppo.py:263-272`); parity here is pinned against torch autograd in float64 (tests/test_mlp_grad.py), which is the same
mathematical object.  Seeds are what `tmjx_ppo_loss_head` returns (d total_loss / d logits, latent moments, baseline).
Parameter names and layouts are those of `track-mjx_b200/policy.py` (flax: kernel [in, out]).
"""
import numpy as np

LN_EPS = 1e-6


def silu(x):
    return x / (1.0 + np.exp(-x))


def silu_grad(x):
    s = 1.0 / (1.0 + np.exp(-x))
    return s * (1.0 + x * (1.0 - s))


def layernorm_fwd(a, scale, bias):
    mu = a.mean(-1, keepdims=True)
    var = ((a - mu) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + LN_EPS)
    xhat = (a - mu) * rstd
    return xhat * scale + bias, (xhat, rstd)


def layernorm_bwd(dy, scale, cache):
    xhat, rstd = cache
    g = dy * scale
    da = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    return da, (dy * xhat).sum(0), dy.sum(0)                     # d input, d scale, d bias


def hidden_fwd(x, W, b, scale, bias):
    pre = x @ W + b
    y, ln = layernorm_fwd(silu(pre), scale, bias)
    return y, (x, pre, ln)


def hidden_bwd(dy, W, scale, cache):
    x, pre, ln = cache
    da, dscale, dbias = layernorm_bwd(dy, scale, ln)
    dpre = da * silu_grad(pre)
    return dpre @ W.T, x.T @ dpre, dpre.sum(0), dscale, dbias     # dgrad, wgrad, d Dense bias, d LN scale, d LN bias


def intention_fwd(cfg, p, obs, eps_z):
    x = (obs - p["norm/mean"]) / p["norm/std"]
    r = cfg.reference_obs_size
    h, caches = x[:, :r], {}
    for i in range(len(cfg.encoder_layers)):
        h, caches[f"encoder/{i}"] = hidden_fwd(h, p[f"encoder/hidden_{i}/kernel"], p[f"encoder/hidden_{i}/bias"],
                                               p[f"encoder/LayerNorm_{i}/scale"], p[f"encoder/LayerNorm_{i}/bias"])
    mean = h @ p["encoder/fc2_mean/kernel"] + p["encoder/fc2_mean/bias"]
    logvar = h @ p["encoder/fc2_logvar/kernel"] + p["encoder/fc2_logvar/bias"]
    sd = np.exp(0.5 * logvar)
    caches["enc_out"], caches["sd"], caches["eps_z"] = h, sd, eps_z
    h = np.concatenate([mean + eps_z * sd, x[:, r:]], -1)
    nd = len(cfg.decoder_layers)
    for i in range(nd):
        h, caches[f"decoder/{i}"] = hidden_fwd(h, p[f"decoder/hidden_{i}/kernel"], p[f"decoder/hidden_{i}/bias"],
                                               p[f"decoder/LayerNorm_{i}/scale"], p[f"decoder/LayerNorm_{i}/bias"])
    caches["dec_out"] = h
    logits = h @ p[f"decoder/hidden_{nd}/kernel"] + p[f"decoder/hidden_{nd}/bias"]
    return logits, mean, logvar, caches


def intention_bwd(cfg, p, caches, d_logits, d_mean, d_logvar):
    """Seeds -> gradient per parameter (keys of `policy.init_params`, normaliser excluded: it is not trained)."""
    g = {}
    nd, lat = len(cfg.decoder_layers), cfg.latent_size
    h = caches["dec_out"]
    g[f"decoder/hidden_{nd}/kernel"], g[f"decoder/hidden_{nd}/bias"] = h.T @ d_logits, d_logits.sum(0)
    dh = d_logits @ p[f"decoder/hidden_{nd}/kernel"].T
    for i in reversed(range(nd)):
        dh, g[f"decoder/hidden_{i}/kernel"], g[f"decoder/hidden_{i}/bias"], g[f"decoder/LayerNorm_{i}/scale"], g[f"decoder/LayerNorm_{i}/bias"] = \
            hidden_bwd(dh, p[f"decoder/hidden_{i}/kernel"], p[f"decoder/LayerNorm_{i}/scale"], caches[f"decoder/{i}"])
    dz = dh[:, :lat]                                             # the proprioceptive columns end at the (untrained) normaliser
    d_mean = d_mean + dz                                         # z = mean + eps exp(logvar / 2)
    d_logvar = d_logvar + dz * caches["eps_z"] * 0.5 * caches["sd"]
    h = caches["enc_out"]
    g["encoder/fc2_mean/kernel"], g["encoder/fc2_mean/bias"] = h.T @ d_mean, d_mean.sum(0)
    g["encoder/fc2_logvar/kernel"], g["encoder/fc2_logvar/bias"] = h.T @ d_logvar, d_logvar.sum(0)
    dh = d_mean @ p["encoder/fc2_mean/kernel"].T + d_logvar @ p["encoder/fc2_logvar/kernel"].T
    for i in reversed(range(len(cfg.encoder_layers))):
        dh, g[f"encoder/hidden_{i}/kernel"], g[f"encoder/hidden_{i}/bias"], g[f"encoder/LayerNorm_{i}/scale"], g[f"encoder/LayerNorm_{i}/bias"] = \
            hidden_bwd(dh, p[f"encoder/hidden_{i}/kernel"], p[f"encoder/LayerNorm_{i}/scale"], caches[f"encoder/{i}"])
    return g


def value_fwd(p, obs, n_hidden):
    h, caches = (obs - p["norm/mean"]) / p["norm/std"], []
    for i in range(n_hidden):
        pre = h @ p[f"hidden_{i}/kernel"] + p[f"hidden_{i}/bias"]
        caches.append((h, pre))
        h = silu(pre)
    caches.append((h, None))
    return (h @ p[f"hidden_{n_hidden}/kernel"] + p[f"hidden_{n_hidden}/bias"])[:, 0], caches


def value_bwd(p, caches, d_value, n_hidden):
    g = {}
    h, _ = caches[n_hidden]
    dy = d_value[:, None]
    g[f"hidden_{n_hidden}/kernel"], g[f"hidden_{n_hidden}/bias"] = h.T @ dy, dy.sum(0)
    dh = dy @ p[f"hidden_{n_hidden}/kernel"].T
    for i in reversed(range(n_hidden)):
        x, pre = caches[i]
        dpre = dh * silu_grad(pre)
        g[f"hidden_{i}/kernel"], g[f"hidden_{i}/bias"] = x.T @ dpre, dpre.sum(0)
        dh = dpre @ p[f"hidden_{i}/kernel"].T
    return g
