"""Golden vectors for the PPO loss head, produced by the REFERENCE'S OWN `compute_ppo_loss` (and `compute_gae`) text.

`track_mjx/agent/mlp_ppo/losses.py` imports brax / flax at module level (not installable here), so the unmodified source text of
the two functions (losses.py:39-101, 104-245; type annotations dropped, they name brax types) is cut out of the reference file
with `ast` at generation time and executed with
`jax.numpy` bound to numpy (float32).  What the reference obtains from other modules is supplied as stand-ins and is therefore
INPUT to the vectors, not under test: the network applications (`policy_apply` / `value_apply` return the stored logits, latent
moments, baseline and bootstrap value), brax 0.12.3's `NormalTanhDistribution` (restated in oracle/ppo_loss.py; its entropy noise
is the stored `eps`), `jax.random.split`.  Nothing of the reference is copied into the repository: only inputs and the outputs it
computed are stored in `tests/golden/ppo_loss.npz`.

    python tools/make_golden_ppo_loss.py        # needs /root/reference; output is committed
"""
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/track_mjx/agent/mlp_ppo/losses.py"

from oracle import ppo_loss as restated  # noqa: E402  (NormalTanhDistribution stand-in only)


class Data:
    """Transition stand-in: attribute access + the nested extras dict; tree_map applies f to every array."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def tree_map(f, d):
    return Data(observation=f(d.observation), next_observation=f(d.next_observation), reward=f(d.reward), discount=f(d.discount),
                extras={"state_extras": {"truncation": f(d.extras["state_extras"]["truncation"])},
                        "policy_extras": {k: f(v) for k, v in d.extras["policy_extras"].items()}})


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    jnp = types.SimpleNamespace(
        ndarray=np.ndarray, concatenate=np.concatenate, expand_dims=np.expand_dims, zeros_like=np.zeros_like, add=np.add,
        swapaxes=np.swapaxes, exp=np.exp, clip=np.clip, mean=np.mean, minimum=np.minimum, square=np.square,
        log=lambda x: np.log(np.float32(x)) if np.isscalar(x) else np.log(x))

    def scan(f, init, xs, length=None, reverse=False):
        n = length if length is not None else len(xs[0])
        order = range(n - 1, -1, -1) if reverse else range(n)
        carry, ys = init, [None] * n
        for t in order:
            carry, y = f(carry, tuple(x[t] for x in xs))
            ys[t] = y
        return carry, np.stack(ys).astype(np.float32)

    jax = types.SimpleNamespace(lax=types.SimpleNamespace(scan=scan, stop_gradient=lambda x: x),
                                random=types.SimpleNamespace(split=lambda rng, n: [rng] * n),
                                tree_util=types.SimpleNamespace(tree_map=tree_map))
    ns = {"jnp": jnp, "jax": jax, "Any": object, "Tuple": tuple, "Callable": object,
          "types": types.SimpleNamespace(Transition=object, Metrics=object), "PPONetworkParams": object,
          "ppo_networks": types.SimpleNamespace(PPONetworks=object)}
    for name in ("compute_gae", "compute_ppo_loss"):
        fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
        fn.returns = None
        for a in fn.args.args + fn.args.kwonlyargs:
            a.annotation = None
        exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["compute_ppo_loss"]


def make_case(rng, T, B, A, L, spread):
    f = np.float32
    c = {"logits": np.concatenate([rng.normal(0, 0.5, (T, B, A)), rng.normal(-0.5, 0.7, (T, B, A))], -1).astype(f),
         "latent_mean": rng.normal(0, 0.8, (T, B, L)).astype(f), "latent_logvar": rng.normal(-1.0, 0.6, (T, B, L)).astype(f),
         "baseline": rng.normal(2.0, 3.0, (T, B)).astype(f), "bootstrap": rng.normal(2.0, 3.0, (B,)).astype(f),
         "reward": rng.normal(0.5, 1.0, (T, B)).astype(f), "eps": rng.normal(0, 1, (T, B, A)).astype(f)}
    done = rng.random((T, B)) < 0.08
    c["truncation"] = ((rng.random((T, B)) < 0.05) & done).astype(f)
    c["discount"] = (1.0 - (done & (c["truncation"] == 0))).astype(f)
    scale = restated.softplus(c["logits"][..., A:]) + restated.MIN_STD
    # behaviour policy = a perturbed copy of the current one, so that rho spreads over both sides of the clipping range
    c["raw_action"] = (c["logits"][..., :A] + scale * rng.normal(0, 1, (T, B, A))).astype(f)
    c["behaviour_log_prob"] = (restated.tanh_normal_log_prob(c["logits"], c["raw_action"]) + rng.normal(0, spread, (T, B))).astype(f)
    return c


def main():
    loss = reference_functions()
    rng = np.random.default_rng(11)
    out = {}
    cases = [(20, 24, 38, 60, 0.3, dict(entropy_cost=1e-2, kl_weight=1e-1, discounting=0.95, reward_scaling=1.0, gae_lambda=0.95,
                                        clipping_epsilon=0.2, normalize_advantage=True)),
             (1, 7, 3, 4, 0.5, dict(entropy_cost=1e-4, kl_weight=1e-3, discounting=0.9, reward_scaling=0.5, gae_lambda=0.9,
                                    clipping_epsilon=0.3, normalize_advantage=False)),
             (5, 33, 38, 60, 0.1, dict(entropy_cost=0.0, kl_weight=1.0, discounting=0.99, reward_scaling=2.0, gae_lambda=1.0,
                                       clipping_epsilon=0.3, normalize_advantage=True))]
    for i, (T, B, A, L, spread, hp) in enumerate(cases):
        c = make_case(rng, T, B, A, L, spread)
        sw = lambda x: np.swapaxes(x, 0, 1)          # the reference takes [B, T, ...] and swaps to time-major itself
        obs = np.zeros((B, T, 1), np.float32)
        data = Data(observation=obs, next_observation=obs, reward=sw(c["reward"]), discount=sw(c["discount"]),
                    extras={"state_extras": {"truncation": sw(c["truncation"])},
                            "policy_extras": {"raw_action": sw(c["raw_action"]), "log_prob": sw(c["behaviour_log_prob"])}})
        dist = types.SimpleNamespace(log_prob=restated.tanh_normal_log_prob,
                                     entropy=lambda logits, key, c=c: restated.tanh_normal_entropy(logits, c["eps"]))

        def value_apply(norm, params, o, c=c):
            return c["baseline"] if o.ndim == 3 else c["bootstrap"]

        net = types.SimpleNamespace(parametric_action_distribution=dist,
                                    policy_network=types.SimpleNamespace(apply=lambda n, p, o, k, c=c: (c["logits"], c["latent_mean"], c["latent_logvar"])),
                                    value_network=types.SimpleNamespace(apply=value_apply))
        total, metrics = loss(types.SimpleNamespace(policy=None, value=None), None, data, 0, 0, net, **hp)
        for k, v in c.items():
            out[f"c{i}_{k}"] = v
        for k, v in hp.items():
            out[f"c{i}_hp_{k}"] = np.float32(v)
        for k in ("total_loss", "policy_loss", "v_loss", "kl_latent_loss", "entropy_loss"):
            out[f"c{i}_{k}"] = np.float32(metrics[k])
        print(i, {k: float(metrics[k]) for k in ("total_loss", "policy_loss", "v_loss", "kl_latent_loss", "entropy_loss")})
    path = os.path.join(ROOT, "tests", "golden", "ppo_loss.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
