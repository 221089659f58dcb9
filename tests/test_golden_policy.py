"""Intention network against golden vectors computed by the REFERENCE'S OWN module text (tools/make_golden_policy.py executes the
unmodified `track_mjx/agent/mlp_ppo/intention_network.py` with numpy stand-ins for the flax / jax primitives it calls).

Pinned here: which observation slice feeds the encoder, Dense -> SiLU -> LayerNorm order, the un-activated last decoder layer, the
(mean | logvar) heads, z = mean + eps exp(logvar / 2), [z | egocentric obs] into the decoder.  CPU: the float64 restatement
oracle/mlp_grad.py (the checker of the CUDA forward / backward) within fp32 round-off of the golden outputs.  GPU: the acting policy
(`tmjx_policy_act`) and the training forward (`tmjx_trainer_policy_forward`) within the TF32 tolerance of tests/test_gpu_policy.py."""
import os

import numpy as np
import pytest

from oracle import mlp_grad as G
from track_mjx_b200 import policy as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy.npz")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("case", [c[0] for c in P.GOLDEN_CASES])
def test_restatement_matches_the_reference_module(case):
    g = np.load(GOLD)
    name, cfg, seed, rows = next(c for c in P.GOLDEN_CASES if c[0] == case)
    p, obs, eps = P.golden_case(cfg, seed, rows)
    p64 = {k: a.astype(np.float64) for k, a in p.items()}
    logits, mean, logvar, caches = G.intention_fwd(cfg, p64, obs.astype(np.float64), eps.astype(np.float64))
    assert rel(logits, g[f"{name}/logits"]) < 2e-5 and rel(mean, g[f"{name}/latent_mean"]) < 2e-5 and rel(logvar, g[f"{name}/latent_logvar"]) < 2e-5
    det, _, _, _ = G.intention_fwd(cfg, p64, obs.astype(np.float64), np.zeros_like(eps, np.float64))     # deterministic: z = mean
    assert rel(det, g[f"{name}/deterministic_logits"]) < 2e-5
    z = mean + eps * np.exp(0.5 * logvar)
    assert rel(z, g[f"{name}/intention"]) < 2e-5
    assert g[f"{name}/logits"].shape == (rows, 2 * cfg.action_size) and g[f"{name}/latent_mean"].shape == (rows, cfg.latent_size)
    # hidden activations are LayerNorm outputs: mean 0 / variance 1 before the (non-trivial) scale and bias
    h = g[f"{name}/encoder_layer_0"]
    assert h.shape == (rows, cfg.encoder_layers[0])
    hn = (h - p["encoder/LayerNorm_0/bias"]) / p["encoder/LayerNorm_0/scale"]
    assert np.abs(hn.mean(-1)).max() < 1e-3 and np.abs(hn.var(-1) - 1).max() < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c[0] for c in P.GOLDEN_CASES])
def test_cuda_policy_matches_the_reference_module(case):
    import torch

    from track_mjx_b200.learner import Trainer

    g = np.load(GOLD)
    name, cfg, seed, rows = next(c for c in P.GOLDEN_CASES if c[0] == case)
    p, obs, eps = P.golden_case(cfg, seed, rows)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pol = P.IntentionPolicy(cfg, p, max_env=rows)
    ea = torch.zeros(rows, cfg.action_size, device="cuda")
    _, ex = pol.act(cu(obs), cu(eps), ea)
    tol = 2e-2          # TF32 operands through up to 11 Dense layers (tests/test_gpu_policy.py derives it)
    assert rel(ex["logits"].cpu().numpy(), g[f"{name}/logits"]) < tol
    assert rel(ex["latent_mean"].cpu().numpy(), g[f"{name}/latent_mean"]) < tol and rel(ex["latent_logvar"].cpu().numpy(), g[f"{name}/latent_logvar"]) < tol
    _, exd = pol.act(cu(obs), None, None, deterministic=True)
    assert rel(exd["logits"].cpu().numpy(), g[f"{name}/deterministic_logits"]) < tol
    critic = (64, 32)
    tr = Trainer(cfg, p, P.init_value_params(cfg.obs_size, critic, 1), critic, max_rows=rows)
    logits, mean, logvar = tr.policy_forward(cu(obs), cu(eps))
    assert rel(logits.cpu().numpy(), g[f"{name}/logits"]) < tol and rel(mean.cpu().numpy(), g[f"{name}/latent_mean"]) < tol
    assert rel(logvar.cpu().numpy(), g[f"{name}/latent_logvar"]) < tol
    tr.close(); pol.close()
