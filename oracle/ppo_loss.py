"""CPU restatement of the reference's PPO loss head -- TEST INFRASTRUCTURE (only tests/ may import it).

Follows `track_mjx/agent/mlp_ppo/losses.py:104-245` (`compute_ppo_loss`) from the point where the networks have been applied:
inputs are the policy logits, latent mean / log-variance, value baseline and bootstrap value, the Transition fields and the
entropy noise (the reference draws it from `entropy_key`; here it is an input).  `NormalTanhDistribution` is upstream brax 0.12.3
`training/distribution.py` (not vendored in /root/reference): scale = softplus(raw) + 0.001, tanh bijector with
log|d tanh / dx| = 2 (log 2 - x - softplus(-2 x)), entropy = Normal entropy + log-det at a sample, summed over the action axis.
Pinned by `tests/golden/ppo_loss.npz` (outputs of the reference's own `compute_ppo_loss` text, tools/make_golden_ppo_loss.py).
Time-major arrays [T, B, ...]; `dtype` float32 like the reference or float64 for error measurements.
"""
import numpy as np

from oracle import gae as _gae

MIN_STD = 0.001


def softplus(x):
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def log_det_tanh(x):
    return 2.0 * (np.log(2.0) - x - softplus(-2.0 * x))


def tanh_normal_log_prob(logits, raw_action):
    a = raw_action.shape[-1]
    loc, scale = logits[..., :a], softplus(logits[..., a:]) + MIN_STD
    z = (raw_action - loc) / scale
    lp = -0.5 * z * z - np.log(scale) - 0.5 * np.log(2 * np.pi) - log_det_tanh(raw_action)
    return lp.sum(-1)


def tanh_normal_entropy(logits, eps):
    a = eps.shape[-1]
    loc, scale = logits[..., :a], softplus(logits[..., a:]) + MIN_STD
    ent = 0.5 + 0.5 * np.log(2 * np.pi) + np.log(scale) + log_det_tanh(loc + scale * eps)
    return ent.sum(-1)


def ppo_loss(logits, latent_mean, latent_logvar, baseline, bootstrap_value, reward, discount, truncation, raw_action,
             behaviour_log_prob, eps_entropy, entropy_cost=1e-4, kl_weight=1e-3, discounting=0.9, reward_scaling=1.0,
             gae_lambda=0.95, clipping_epsilon=0.3, normalize_advantage=True, dtype=np.float32):
    f = dtype
    c = lambda x: np.asarray(x, f)
    logits, latent_mean, latent_logvar, baseline, bootstrap_value = map(c, (logits, latent_mean, latent_logvar, baseline, bootstrap_value))
    reward, discount, truncation, raw_action, behaviour_log_prob, eps_entropy = map(
        c, (reward, discount, truncation, raw_action, behaviour_log_prob, eps_entropy))
    rewards = reward * f(reward_scaling)                                              # :157
    termination = (1 - discount) * (1 - truncation)                                   # :159
    target_lp = tanh_normal_log_prob(logits, raw_action)                              # :161-163
    vs, adv = _gae.compute_gae(truncation, termination, rewards, baseline, bootstrap_value, f(gae_lambda), f(discounting))  # :166-174
    vs, adv = c(vs), c(adv)
    if normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std() + f(1e-8))                              # :175-176
    rho = np.exp(target_lp - behaviour_log_prob)                                      # :177
    s1 = rho * adv
    s2 = np.clip(rho, f(1 - clipping_epsilon), f(1 + clipping_epsilon)) * adv         # :179-182
    policy_loss = -np.mean(np.minimum(s1, s2))                                        # :184
    v_err = vs - baseline
    v_loss = np.mean(v_err * v_err) * f(0.5) * f(0.5)                                 # :187-188
    entropy_loss = f(entropy_cost) * -np.mean(tanh_normal_entropy(logits, eps_entropy))   # :191-194
    alpha = f(0.95)
    pv = f(1 - 0.95 ** 2)                                                             # :201-202
    kl0 = f(-0.5) * np.mean(1 + latent_logvar[0] - np.square(latent_mean[0]) - np.exp(latent_logvar[0]))   # :206-208
    T = latent_mean.shape[0]
    if T > 1:
        zp, mu, lv = latent_mean[:-1], latent_mean[1:], latent_logvar[1:]            # :214-216
        klt = f(0.5) * np.mean(np.exp(lv) / pv + np.square(alpha * zp - mu) / pv - 1 + (np.log(pv) - lv))   # :221-226
        kl_latent = f(kl_weight) * ((kl0 + klt * (T - 1)) / T)                        # :229-232
    else:
        kl_latent = f(kl_weight) * kl0                                                # :235
    total = policy_loss + v_loss + entropy_loss + kl_latent                           # :237
    return {"total_loss": total, "policy_loss": policy_loss, "v_loss": v_loss, "kl_latent_loss": kl_latent,
            "entropy_loss": entropy_loss, "vs": vs, "advantages": adv}
