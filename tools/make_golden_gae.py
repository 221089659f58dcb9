"""Golden vectors for Generalised Advantage Estimation, produced by the REFERENCE'S OWN `compute_gae`.

`track_mjx/agent/mlp_ppo/losses.py` imports brax / flax at module level (not installable here), so the unmodified source text of
`compute_gae` (losses.py:39-101) is cut out of the reference file with `ast` at generation time and executed with `jax.numpy`
bound to numpy (float32) and a numpy `lax.scan(reverse=True)` / `stop_gradient`.  Nothing of the reference is copied into the
repository: only the inputs and the outputs it computed are stored in `tests/golden/gae.npz`.

    python tools/make_golden_gae.py        # needs /root/reference; output is committed
"""
import ast
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/track_mjx/agent/mlp_ppo/losses.py"


def reference_compute_gae():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_gae")
    text = ast.get_source_segment(src, fn)
    jnp = types.SimpleNamespace(
        ndarray=np.ndarray, concatenate=np.concatenate, expand_dims=np.expand_dims, zeros_like=np.zeros_like, add=np.add)

    def scan(f, init, xs, length=None, reverse=False):
        n = length if length is not None else len(xs[0])
        order = range(n - 1, -1, -1) if reverse else range(n)
        carry, ys = init, [None] * n
        for t in order:
            carry, y = f(carry, tuple(x[t] for x in xs))
            ys[t] = y
        return carry, np.stack(ys).astype(np.float32)

    jax = types.SimpleNamespace(lax=types.SimpleNamespace(scan=scan, stop_gradient=lambda x: x))
    ns = {"jnp": jnp, "jax": jax}
    exec(compile(text, REF, "exec"), ns)
    return ns["compute_gae"]


def make_inputs(rng, T, B):
    f = np.float32
    termination = (rng.random((T, B)) < 0.08).astype(f)
    truncation = ((rng.random((T, B)) < 0.05) & (termination == 0)).astype(f)
    rewards = rng.normal(0.5, 1.0, (T, B)).astype(f)
    values = rng.normal(2.0, 3.0, (T, B)).astype(f)
    bootstrap = rng.normal(2.0, 3.0, (B,)).astype(f)
    return truncation, termination, rewards, values, bootstrap


def main():
    gae = reference_compute_gae()
    rng = np.random.default_rng(7)
    out = {}
    for i, (T, B, lam, disc) in enumerate([(20, 64, 0.95, 0.99), (1, 5, 1.0, 0.9), (7, 33, 0.0, 0.97), (50, 3, 0.8, 1.0)]):
        tr, te, r, v, b = make_inputs(rng, T, B)
        vs, adv = gae(truncation=tr, termination=te, rewards=r, values=v, bootstrap_value=b, lambda_=np.float32(lam), discount=np.float32(disc))
        out.update({f"c{i}_truncation": tr, f"c{i}_termination": te, f"c{i}_rewards": r, f"c{i}_values": v, f"c{i}_bootstrap": b,
                    f"c{i}_lambda": np.float32(lam), f"c{i}_discount": np.float32(disc),
                    f"c{i}_vs": np.asarray(vs, np.float32), f"c{i}_advantages": np.asarray(adv, np.float32)})
    path = os.path.join(ROOT, "tests", "golden", "gae.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("c0")})


if __name__ == "__main__":
    main()
