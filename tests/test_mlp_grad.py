"""The hand-derived backward pass of the intention network and the value network (oracle/mlp_grad.py: the checker the CUDA backward
kernels of the next round will be held to) against torch autograd in float64 on the same parameters: 1e-9 relative to the largest
entry of each gradient.  CPU only; the forward restatement is also compared with the torch forward the GPU policy tests use."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import mlp_grad as mg

torch = pytest.importorskip("torch")


def make(seed=0):
    cfg = SimpleNamespace(obs_size=30, reference_obs_size=20, action_size=4, latent_size=5, encoder_layers=(16, 12), decoder_layers=(14, 10))
    rng = np.random.default_rng(seed)
    p = {"norm/mean": rng.normal(0, 0.2, 30), "norm/std": rng.uniform(0.5, 2, 30)}

    def dense(name, k, n):
        p[f"{name}/kernel"], p[f"{name}/bias"] = rng.normal(0, 1 / np.sqrt(k), (k, n)), rng.normal(0, 0.1, n)

    k = 20
    for i, n in enumerate(cfg.encoder_layers):
        dense(f"encoder/hidden_{i}", k, n)
        p[f"encoder/LayerNorm_{i}/scale"], p[f"encoder/LayerNorm_{i}/bias"] = rng.uniform(0.7, 1.3, n), rng.normal(0, 0.1, n)
        k = n
    dense("encoder/fc2_mean", k, 5)
    dense("encoder/fc2_logvar", k, 5)
    k = 5 + 10
    for i, n in enumerate(cfg.decoder_layers):
        dense(f"decoder/hidden_{i}", k, n)
        p[f"decoder/LayerNorm_{i}/scale"], p[f"decoder/LayerNorm_{i}/bias"] = rng.uniform(0.7, 1.3, n), rng.normal(0, 0.1, n)
        k = n
    dense("decoder/hidden_2", k, 8)
    return cfg, p, rng


def test_intention_network_backward_matches_autograd():
    cfg, p, rng = make()
    n = 37
    obs, eps = rng.normal(size=(n, 30)), rng.normal(size=(n, 5))
    seeds = rng.normal(size=(n, 8)), rng.normal(size=(n, 5)), rng.normal(size=(n, 5))
    logits, mean, logvar, caches = mg.intention_fwd(cfg, p, obs, eps)
    grads = mg.intention_bwd(cfg, p, caches, *seeds)

    F = torch.nn.functional
    tp = {k: torch.tensor(v, requires_grad=not k.startswith("norm/")) for k, v in p.items()}
    x = (torch.tensor(obs) - tp["norm/mean"]) / tp["norm/std"]
    h = x[:, :20]
    for i, w in enumerate(cfg.encoder_layers):
        h = F.layer_norm(F.silu(h @ tp[f"encoder/hidden_{i}/kernel"] + tp[f"encoder/hidden_{i}/bias"]), (w,),
                         tp[f"encoder/LayerNorm_{i}/scale"], tp[f"encoder/LayerNorm_{i}/bias"], eps=1e-6)
    tmean = h @ tp["encoder/fc2_mean/kernel"] + tp["encoder/fc2_mean/bias"]
    tlogvar = h @ tp["encoder/fc2_logvar/kernel"] + tp["encoder/fc2_logvar/bias"]
    h = torch.cat([tmean + torch.tensor(eps) * torch.exp(0.5 * tlogvar), x[:, 20:]], -1)
    for i, w in enumerate(cfg.decoder_layers):
        h = F.layer_norm(F.silu(h @ tp[f"decoder/hidden_{i}/kernel"] + tp[f"decoder/hidden_{i}/bias"]), (w,),
                         tp[f"decoder/LayerNorm_{i}/scale"], tp[f"decoder/LayerNorm_{i}/bias"], eps=1e-6)
    tlogits = h @ tp["decoder/hidden_2/kernel"] + tp["decoder/hidden_2/bias"]
    assert np.allclose(logits, tlogits.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert np.allclose(mean, tmean.detach().numpy(), rtol=1e-12, atol=1e-12) and np.allclose(logvar, tlogvar.detach().numpy(), rtol=1e-12, atol=1e-12)
    ((tlogits * torch.tensor(seeds[0])).sum() + (tmean * torch.tensor(seeds[1])).sum() + (tlogvar * torch.tensor(seeds[2])).sum()).backward()
    assert set(grads) == {k for k in p if not k.startswith("norm/")}
    for k, g in grads.items():
        want = tp[k].grad.numpy()
        assert g.shape == want.shape, k
        assert np.abs(g - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-12), (k, np.abs(g - want).max())


def test_value_network_backward_matches_autograd():
    rng = np.random.default_rng(4)
    sizes, k = (12, 9), 30
    p = {"norm/mean": rng.normal(0, 0.2, 30), "norm/std": rng.uniform(0.5, 2, 30)}
    for i, n in enumerate(sizes + (1,)):
        p[f"hidden_{i}/kernel"], p[f"hidden_{i}/bias"] = rng.normal(0, 1 / np.sqrt(k), (k, n)), rng.normal(0, 0.1, n)
        k = n
    obs, seed = rng.normal(size=(23, 30)), rng.normal(size=23)
    value, caches = mg.value_fwd(p, obs, 2)
    grads = mg.value_bwd(p, caches, seed, 2)
    tp = {k: torch.tensor(v, requires_grad=not k.startswith("norm/")) for k, v in p.items()}
    h = (torch.tensor(obs) - tp["norm/mean"]) / tp["norm/std"]
    for i in range(2):
        h = torch.nn.functional.silu(h @ tp[f"hidden_{i}/kernel"] + tp[f"hidden_{i}/bias"])
    tv = (h @ tp["hidden_2/kernel"] + tp["hidden_2/bias"])[:, 0]
    assert np.allclose(value, tv.detach().numpy(), rtol=1e-12, atol=1e-12)
    (tv * torch.tensor(seed)).sum().backward()
    for k, g in grads.items():
        want = tp[k].grad.numpy()
        assert np.abs(g - want).max() <= 1e-9 * np.abs(want).max(), k
