"""Size-independent properties of the learner-side restatements (oracle/gae.py, running_stats.py, ppo_loss.py, optimizer.py) and
their edge cases (T = 1, a single row, every step terminal, constant features).  CPU only; the CUDA kernels are held to these
restatements by the `-m gpu` tests."""
import numpy as np

from oracle import gae, optimizer, ppo_loss, running_stats

f = np.float32


def _rollout(rng, T, B):
    term = (rng.random((T, B)) < 0.1).astype(f)
    trunc = ((rng.random((T, B)) < 0.05) & (term == 0)).astype(f)
    return trunc, term, rng.normal(0.5, 1, (T, B)).astype(f), rng.normal(1, 2, (T, B)).astype(f), rng.normal(1, 2, B).astype(f)


def test_gae_edge_cases_and_linearity():
    rng = np.random.default_rng(0)
    trunc, term, r, v, boot = _rollout(rng, 12, 9)
    # lambda = 0: one-step TD targets
    vs, adv = gae.compute_gae(trunc, term, r, v, boot, 0.0, 0.97)
    v_next = np.concatenate([v[1:], boot[None]])
    assert np.allclose(vs, (r + f(0.97) * (1 - term) * v_next - v) * (1 - trunc) + v, rtol=1e-6, atol=1e-6)
    # every step terminal: the target is the reward, the advantage r - v
    ones, zeros = np.ones_like(term), np.zeros_like(term)
    vs, adv = gae.compute_gae(zeros, ones, r, v, boot, 0.95, 0.99)
    assert np.allclose(vs, r, atol=1e-6) and np.allclose(adv, r - v, atol=1e-6)
    # T = 1 and a single environment
    vs1, adv1 = gae.compute_gae(zeros[:1, :1], zeros[:1, :1], r[:1, :1], v[:1, :1], boot[:1], 0.95, 0.9)
    assert np.allclose(vs1, r[:1, :1] + f(0.9) * boot[:1], rtol=1e-6) and np.allclose(adv1, vs1 - v[:1, :1], rtol=1e-5, atol=1e-6)
    # linear in (rewards, values, bootstrap)
    a = gae.compute_gae(trunc, term, r, v, boot, 0.9, 0.95)
    b = gae.compute_gae(trunc, term, 3 * r, 3 * v, 3 * boot, 0.9, 0.95)
    assert np.allclose(b[0], 3 * a[0], rtol=1e-5, atol=1e-5) and np.allclose(b[1], 3 * a[1], rtol=1e-5, atol=1e-5)


def test_running_statistics_chunking_single_row_and_constant_feature():
    rng = np.random.default_rng(1)
    x = (rng.normal(size=(600, 17)) * rng.uniform(0.1, 3, 17) + rng.normal(size=17)).astype(f)
    x[:, 3] = -1.25
    z = np.zeros(17, f)
    whole = running_stats.update(0.0, z, z, x)
    c, m, sv, std = running_stats.update(0.0, z, z, x[:1])            # a single row: std sits on std_min
    assert c == 1 and np.array_equal(m, x[0]) and np.all(std == f(1e-6))
    for cut in (1, 250, 599):
        c, m, sv, _ = running_stats.update(0.0, z, z, x[:cut])
        c, m, sv, std = running_stats.update(c, m, sv, x[cut:])
        assert c == 600 and np.allclose(m, whole[1], rtol=1e-5, atol=1e-6) and np.allclose(std, whole[3], rtol=2e-4, atol=1e-6)
    assert whole[3][3] == f(1e-6) and np.allclose(whole[3], np.clip(x.astype(np.float64).std(0), 1e-6, 1e6), rtol=1e-4)
    assert np.allclose(running_stats.normalize(x, whole[1], whole[3])[:, :3].mean(0), 0, atol=1e-4)


def test_ppo_loss_scale_invariance_and_single_timestep():
    import tools.make_golden_ppo_loss as mk

    rng = np.random.default_rng(2)
    c = mk.make_case(rng, 6, 40, 5, 7, 0.3)
    args = lambda s: (c["logits"], c["latent_mean"], c["latent_logvar"], s * c["baseline"], s * c["bootstrap"], s * c["reward"], c["discount"],
                      c["truncation"], c["raw_action"], c["behaviour_log_prob"], c["eps"])
    a, b = ppo_loss.ppo_loss(*args(f(1))), ppo_loss.ppo_loss(*args(f(4)))
    # normalised advantages: the policy term does not see a common scale of rewards and values, the value term sees its square
    assert np.isclose(a["policy_loss"], b["policy_loss"], rtol=1e-4, atol=1e-6) and np.isclose(b["v_loss"], 16 * a["v_loss"], rtol=1e-4)
    assert np.isclose(a["kl_latent_loss"], b["kl_latent_loss"]) and np.isclose(a["entropy_loss"], b["entropy_loss"])
    assert abs(a["advantages"].mean()) < 1e-5 and np.isclose(a["advantages"].std(), 1, rtol=1e-4)
    # identical behaviour policy: rho = 1 everywhere, nothing is clipped, the policy term is -mean(normalised advantages) = 0
    c2 = dict(c, behaviour_log_prob=ppo_loss.tanh_normal_log_prob(c["logits"], c["raw_action"]).astype(f))
    on = ppo_loss.ppo_loss(c2["logits"], c2["latent_mean"], c2["latent_logvar"], c2["baseline"], c2["bootstrap"], c2["reward"], c2["discount"],
                           c2["truncation"], c2["raw_action"], c2["behaviour_log_prob"], c2["eps"])
    assert abs(on["policy_loss"]) < 1e-5
    # T = 1: the latent prior is the standard normal only (losses.py:233-235)
    one = ppo_loss.ppo_loss(c["logits"][:1], c["latent_mean"][:1], c["latent_logvar"][:1], c["baseline"][:1], c["bootstrap"], c["reward"][:1],
                            c["discount"][:1], c["truncation"][:1], c["raw_action"][:1], c["behaviour_log_prob"][:1], c["eps"][:1], kl_weight=1.0)
    mu, lv = c["latent_mean"][0].astype(np.float64), c["latent_logvar"][0].astype(np.float64)
    assert np.isclose(one["kl_latent_loss"], -0.5 * np.mean(1 + lv - mu ** 2 - np.exp(lv)), rtol=1e-5)


def test_adam_clip_equals_prescaled_gradient_and_bias_correction():
    rng = np.random.default_rng(3)
    p, g = rng.normal(size=100), rng.normal(size=100) * 7                  # norm ~ 70 > 10
    z = np.zeros(100)
    clipped = optimizer.adam_step(p, g, z, z, 0, max_grad_norm=10.0, dtype=np.float64)
    pre = optimizer.adam_step(p, g * (10.0 / np.linalg.norm(g)), z, z, 0, max_grad_norm=0.0, dtype=np.float64)
    assert np.allclose(clipped[0], pre[0], rtol=1e-12) and np.isclose(clipped[4], np.linalg.norm(g))
    # first step with zero moments: the update is -lr sign(g) up to eps
    first = optimizer.adam_step(p, g, z, z, 0, learning_rate=1e-3, max_grad_norm=0.0, dtype=np.float64)
    assert np.allclose(first[0] - p, -1e-3 * np.sign(g), rtol=1e-6) and first[3] == 1
