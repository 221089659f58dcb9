"""PPO loss head: the CPU restatement (oracle/ppo_loss.py) against golden vectors computed by the reference's own
`compute_ppo_loss` text (tools/make_golden_ppo_loss.py), and -- on a GPU -- the CUDA kernels behind `learner.ppo_loss_head`
(loss terms AND the gradients w.r.t. logits / baseline / latent moments that seed the network backward pass) against it.

Tolerances: all five loss terms are float32 means over up to T*B*A elements; restatement vs reference text 2e-6 relative (same
numpy operations), CUDA vs golden 2e-5 relative + 1e-6 absolute (different summation order, fast-math free).  Gradients are
compared with torch autograd of a float64 torch transcription of the restatement (checked against the restatement in the same
test): 1e-4 relative to the largest gradient entry of each tensor, elementwise."""
import os

import numpy as np
import pytest

from oracle import ppo_loss as pl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_loss.npz")
TERMS = ("total_loss", "policy_loss", "v_loss", "kl_latent_loss", "entropy_loss")
INPUTS = ("logits", "latent_mean", "latent_logvar", "baseline", "bootstrap", "reward", "discount", "truncation", "raw_action",
          "behaviour_log_prob", "eps")
HP = ("entropy_cost", "kl_weight", "discounting", "reward_scaling", "gae_lambda", "clipping_epsilon", "normalize_advantage")


def load_case(g, i):
    c = {k: g[f"c{i}_{k}"] for k in INPUTS}
    hp = {k: float(g[f"c{i}_hp_{k}"]) for k in HP}
    hp["normalize_advantage"] = bool(hp["normalize_advantage"])
    return c, hp


def restated(c, hp, dtype=np.float32):
    return pl.ppo_loss(c["logits"], c["latent_mean"], c["latent_logvar"], c["baseline"], c["bootstrap"], c["reward"], c["discount"],
                       c["truncation"], c["raw_action"], c["behaviour_log_prob"], c["eps"], dtype=dtype, **hp)


def test_restatement_matches_reference_outputs():
    g = np.load(GOLD)
    for i in range(3):
        c, hp = load_case(g, i)
        out = restated(c, hp)
        for k in TERMS:
            assert np.isclose(out[k], g[f"c{i}_{k}"], rtol=2e-6, atol=1e-9), (i, k, out[k], g[f"c{i}_{k}"])
    # both sides of the clipping range are exercised by case 0
    c, hp = load_case(g, 0)
    rho = np.exp(pl.tanh_normal_log_prob(c["logits"], c["raw_action"]) - c["behaviour_log_prob"])
    assert (rho < 1 - hp["clipping_epsilon"]).mean() > 0.05 and (rho > 1 + hp["clipping_epsilon"]).mean() > 0.05


def torch_loss(torch, c, hp, vs, adv):
    """float64 torch transcription of oracle/ppo_loss.py with vs / advantages given (they are stop_gradient in the reference)."""
    t = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=k in ("logits", "latent_mean", "latent_logvar", "baseline"))
         for k, v in c.items()}
    A = c["raw_action"].shape[-1]
    sp = torch.nn.functional.softplus
    ldj = lambda x: 2.0 * (np.log(2.0) - x - sp(-2.0 * x))
    loc, scale = t["logits"][..., :A], sp(t["logits"][..., A:]) + pl.MIN_STD
    z = (t["raw_action"] - loc) / scale
    lp = (-0.5 * z * z - torch.log(scale) - 0.5 * np.log(2 * np.pi) - ldj(t["raw_action"])).sum(-1)
    rho = torch.exp(lp - t["behaviour_log_prob"])
    adv = torch.tensor(np.asarray(adv, np.float64))
    eps_c = hp["clipping_epsilon"]
    policy = -torch.minimum(rho * adv, torch.clamp(rho, 1 - eps_c, 1 + eps_c) * adv).mean()
    v = ((torch.tensor(np.asarray(vs, np.float64)) - t["baseline"]) ** 2).mean() * 0.25
    ent = (0.5 + 0.5 * np.log(2 * np.pi) + torch.log(scale) + ldj(loc + scale * t["eps"])).sum(-1).mean()
    mu, lv = t["latent_mean"], t["latent_logvar"]
    pv, T = 1 - 0.95 ** 2, mu.shape[0]
    kl0 = -0.5 * (1 + lv[0] - mu[0] ** 2 - torch.exp(lv[0])).mean()
    if T > 1:
        klt = 0.5 * (torch.exp(lv[1:]) / pv + (0.95 * mu[:-1] - mu[1:]) ** 2 / pv - 1 + (np.log(pv) - lv[1:])).mean()
        kl = hp["kl_weight"] * (kl0 + klt * (T - 1)) / T
    else:
        kl = hp["kl_weight"] * kl0
    total = policy + v - hp["entropy_cost"] * ent + kl
    total.backward()
    return float(total.detach()), {k: t[k].grad.numpy() for k in ("logits", "latent_mean", "latent_logvar", "baseline")}


def test_torch_transcription_matches_restatement():
    torch = pytest.importorskip("torch")
    g = np.load(GOLD)
    for i in range(3):
        c, hp = load_case(g, i)
        out = restated(c, hp)
        total, grads = torch_loss(torch, c, hp, out["vs"], out["advantages"])
        assert np.isclose(total, g[f"c{i}_total_loss"], rtol=5e-6)
        assert all(np.isfinite(v).all() for v in grads.values())


@pytest.mark.gpu
def test_cuda_loss_head_matches_golden_and_autograd():
    torch = pytest.importorskip("torch")
    from track_mjx_b200.learner import ppo_loss_head

    g = np.load(GOLD)
    for i in range(3):
        c, hp = load_case(g, i)
        d = {k: torch.from_numpy(v).cuda() for k, v in c.items()}
        out = ppo_loss_head(d["logits"], d["latent_mean"], d["latent_logvar"], d["baseline"], d["bootstrap"], d["reward"], d["discount"],
                            d["truncation"], d["raw_action"], d["behaviour_log_prob"], d["eps"], **hp)
        torch.cuda.synchronize()
        for k in TERMS:
            assert np.isclose(float(out[k]), g[f"c{i}_{k}"], rtol=2e-5, atol=1e-6), (i, k, float(out[k]), g[f"c{i}_{k}"])
        ref = restated(c, hp)
        assert np.allclose(out["vs"].cpu().numpy(), ref["vs"], rtol=1e-6, atol=1e-6)
        assert np.allclose(out["advantages"].cpu().numpy(), ref["advantages"], rtol=1e-4, atol=1e-5)
        _, grads = torch_loss(torch, c, hp, ref["vs"], ref["advantages"])
        for k, name in (("logits", "d_logits"), ("latent_mean", "d_latent_mean"), ("latent_logvar", "d_latent_logvar"), ("baseline", "d_baseline")):
            got, want = out[name].cpu().numpy().astype(np.float64), grads[k]
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (i, k, np.abs(got - want).max(), np.abs(want).max())
    # full size (one PPO minibatch pass over 20 x 16384 transitions): finite, deterministic, and linear in the hyper-parameters
    rng = np.random.default_rng(5)
    import tools.make_golden_ppo_loss as mk

    c = mk.make_case(rng, 20, 16384, 38, 60, 0.3)
    d = {k: torch.from_numpy(v).cuda() for k, v in c.items()}
    args = (d["logits"], d["latent_mean"], d["latent_logvar"], d["baseline"], d["bootstrap"], d["reward"], d["discount"], d["truncation"],
            d["raw_action"], d["behaviour_log_prob"], d["eps"])
    a = ppo_loss_head(*args, entropy_cost=1e-2, kl_weight=0.1)
    a = {k: (v.clone() if hasattr(v, "clone") else v) for k, v in a.items()}
    b = ppo_loss_head(*args, entropy_cost=1e-2, kl_weight=0.1)
    assert all(torch.equal(a[k], b[k]) for k in ("d_logits", "d_latent_mean", "d_latent_logvar", "d_baseline", "losses"))
    b2 = ppo_loss_head(*args, entropy_cost=2e-2, kl_weight=0.2)
    assert np.isclose(float(b2["entropy_loss"]), 2 * float(a["entropy_loss"]), rtol=1e-6)
    assert np.isclose(float(b2["kl_latent_loss"]), 2 * float(a["kl_latent_loss"]), rtol=1e-6)
    assert torch.allclose(b2["d_latent_logvar"], 2 * a["d_latent_logvar"], rtol=1e-5, atol=1e-12)
    ref = pl.ppo_loss(*[c[k] for k in INPUTS], entropy_cost=1e-2, kl_weight=0.1)
    for k in TERMS:
        assert np.isclose(float(a[k]), ref[k], rtol=5e-5, atol=1e-6), (k, float(a[k]), ref[k])
