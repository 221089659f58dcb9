"""Golden vectors for the intention network, produced by the REFERENCE'S OWN module text.

`track_mjx/agent/mlp_ppo/intention_network.py` (Encoder / Decoder / reparameterize / IntentionNetwork, :14-142) is a flax.linen
module tree; flax, jax and brax are not installable here, so the UNMODIFIED source text of the file is executed at generation time
with small numpy stand-ins for the framework primitives it calls:

  * `flax.linen.Module` / `@nn.compact` / `setup`: attribute-defined sub-modules and name scoping over a nested parameter dict
    (`{"params": {"encoder": {"hidden_0": {"kernel", "bias"}, "LayerNorm_0": {"scale", "bias"}, ...}, "decoder": {...}}}`, the tree
    flax builds for this module: unnamed `nn.LayerNorm()` instances are auto-named `LayerNorm_<i>` in creation order);
  * `nn.Dense` = `x @ kernel + bias`, `nn.LayerNorm` = flax 0.10 `_compute_stats` / `_normalize` with its defaults (epsilon 1e-6,
    use_fast_variance: var = max(0, E[x^2] - E[x]^2)), `nn.silu` = `x * sigmoid(x)` -- restated, float32;
  * `jax.random.split` / `random.normal`: the key object carries the latent noise, which is therefore an INPUT of the vectors;
  * `running_statistics.normalize` (`agent/masked_running_statistics.py:217-236`): `(obs - mean) / std`, restated.

What is under test is the reference's wiring: which slice of the observation feeds the encoder, the layer order Dense -> SiLU ->
LayerNorm, the un-activated last decoder layer, the (mean | logvar) heads, z = mean + eps exp(logvar / 2), the concatenation
[z | egocentric obs].  Nothing of the reference is copied into the repository; inputs are regenerated from seeds by the tests
(`policy.golden_case`), only the outputs the reference computed are stored in `tests/golden/policy.npz`.

    python tools/make_golden_policy.py        # needs /root/reference; output is committed
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/track_mjx/agent/mlp_ppo/intention_network.py"

f32 = np.float32
_scope = []          # stack of parameter sub-dicts
_ln_count = []       # per compact call: number of unnamed LayerNorm created so far


def _compact(fn):
    def wrapped(self, *a, **k):
        top = _scope[-1]
        _scope.append(top[self._name] if getattr(self, "_name", None) else top)
        _ln_count.append(0)
        try:
            return fn(self, *a, **k)
        finally:
            _scope.pop()
            _ln_count.pop()
    return wrapped


class Module:
    """dataclass-like flax Module: class annotations with defaults become constructor kwargs; sub-modules assigned in `setup` are
    scoped by their attribute name."""

    def __init__(self, **kw):
        ann = {}
        for c in reversed(type(self).__mro__):
            ann.update(getattr(c, "__annotations__", {}))
        for name in ann:
            if name in kw:
                object.__setattr__(self, name, kw[name])
            elif hasattr(type(self), name):
                object.__setattr__(self, name, getattr(type(self), name))
            else:
                raise TypeError(f"missing field {name}")
        if "name" in kw:
            object.__setattr__(self, "_name", kw["name"])
        if hasattr(self, "setup"):
            self.setup()

    def __setattr__(self, k, v):
        if isinstance(v, Module):
            object.__setattr__(v, "_name", k)
        object.__setattr__(self, k, v)

    def apply(self, variables, *a, **k):
        _scope.append(variables["params"])
        _ln_count.append(0)
        try:
            return self(*a, **k)
        finally:
            _scope.pop()
            _ln_count.pop()


class Dense:
    def __init__(self, features, name=None, kernel_init=None, use_bias=True):
        self.features, self.name, self.use_bias = features, name, use_bias

    def __call__(self, x):
        p = _scope[-1][self.name]
        assert p["kernel"].shape[1] == self.features
        y = (x.astype(f32) @ p["kernel"].astype(f32)).astype(f32)
        return (y + p["bias"].astype(f32)).astype(f32) if self.use_bias else y


class LayerNorm:
    def __init__(self, epsilon=1e-6):
        self.eps = f32(epsilon)
        self.name = f"LayerNorm_{_ln_count[-1]}"
        _ln_count[-1] += 1

    def __call__(self, x):
        p = _scope[-1][self.name]
        x = x.astype(f32)
        mean = x.mean(-1, keepdims=True, dtype=f32)
        mean2 = (x * x).mean(-1, keepdims=True, dtype=f32)
        var = np.maximum(f32(0), mean2 - mean * mean)
        y = (x - mean) * (f32(1) / np.sqrt(var + self.eps)).astype(f32)
        return (y * p["scale"].astype(f32) + p["bias"].astype(f32)).astype(f32)


class Key:
    def __init__(self, eps):
        self.eps = eps


def reference_module():
    nn = types.SimpleNamespace(Module=Module, compact=_compact, Dense=Dense, LayerNorm=LayerNorm,
                               silu=lambda x: (x / (f32(1) + np.exp(-x))).astype(f32))
    random = types.SimpleNamespace(normal=lambda key, shape: key.eps.reshape(shape), split=lambda key, n=2: [key] * n, PRNGKey=lambda seed: None)
    jax = types.SimpleNamespace(random=random, numpy=None, nn=types.SimpleNamespace(initializers=types.SimpleNamespace(lecun_uniform=lambda: None)),
                                Array=np.ndarray)
    np_like = types.SimpleNamespace(**{k: getattr(np, k) for k in ("ndarray", "exp", "concatenate")}, zeros=lambda shape: np.zeros(shape, f32))
    jax.numpy = np_like
    stubs = {
        "brax": types.ModuleType("brax"), "brax.training": types.ModuleType("brax.training"),
        "brax.training.networks": types.SimpleNamespace(ActivationFn=object, Initializer=object, FeedForwardNetwork=lambda **k: types.SimpleNamespace(**k)),
        "brax.training.types": types.SimpleNamespace(PRNGKey=object, PreprocessObservationFn=object, identity_observation_preprocessor=lambda o, p: o),
        "jax": jax, "jax.numpy": np_like, "jax.random": random, "flax": types.SimpleNamespace(linen=nn), "flax.linen": nn,
    }
    stubs["brax.training"].networks, stubs["brax.training"].types = stubs["brax.training.networks"], stubs["brax.training.types"]
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mod = types.ModuleType("ref_intention_network")
        exec(compile(open(REF).read(), REF, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def flax_tree(cfg, p):
    """policy.init_params naming -> the nested dict flax would hold for IntentionNetwork."""
    enc, dec = {}, {}
    for i in range(len(cfg.encoder_layers)):
        enc[f"hidden_{i}"] = {"kernel": p[f"encoder/hidden_{i}/kernel"], "bias": p[f"encoder/hidden_{i}/bias"]}
        enc[f"LayerNorm_{i}"] = {"scale": p[f"encoder/LayerNorm_{i}/scale"], "bias": p[f"encoder/LayerNorm_{i}/bias"]}
    for h in ("fc2_mean", "fc2_logvar"):
        enc[h] = {"kernel": p[f"encoder/{h}/kernel"], "bias": p[f"encoder/{h}/bias"]}
    nd = len(cfg.decoder_layers)
    for i in range(nd + 1):
        dec[f"hidden_{i}"] = {"kernel": p[f"decoder/hidden_{i}/kernel"], "bias": p[f"decoder/hidden_{i}/bias"]}
        if i < nd:
            dec[f"LayerNorm_{i}"] = {"scale": p[f"decoder/LayerNorm_{i}/scale"], "bias": p[f"decoder/LayerNorm_{i}/bias"]}
    return {"params": {"encoder": enc, "decoder": dec}}


def main():
    from track_mjx_b200 import policy as P

    ref = reference_module()
    out = {}
    for name, cfg, seed, rows in P.GOLDEN_CASES:
        p, obs, eps = P.golden_case(cfg, seed, rows)
        net = ref.make_intention_policy(2 * cfg.action_size, latent_size=cfg.latent_size, total_obs_size=cfg.obs_size,
                                        reference_obs_size=cfg.reference_obs_size,
                                        preprocess_observations_fn=lambda o, mean_std: ((o - mean_std[0]) / mean_std[1]).astype(f32),
                                        encoder_hidden_layer_sizes=tuple(cfg.encoder_layers), decoder_hidden_layer_sizes=tuple(cfg.decoder_layers))
        tree = flax_tree(cfg, p)
        norm = (p["norm/mean"], p["norm/std"])
        logits, mean, logvar = net.apply(norm, tree, obs, Key(eps))
        det_logits, _, _ = net.apply(norm, tree, obs, Key(eps), deterministic=True)
        _, _, _, acts = net.apply(norm, tree, obs, Key(eps), get_activation=True)
        out[f"{name}/logits"], out[f"{name}/latent_mean"], out[f"{name}/latent_logvar"] = logits, mean, logvar
        out[f"{name}/deterministic_logits"] = det_logits
        out[f"{name}/intention"] = acts["intention"]
        out[f"{name}/encoder_layer_0"] = acts["encoder"]["layer_0"]
        out[f"{name}/decoder_layer_last"] = acts["decoder"][f"layer_{len(cfg.decoder_layers) - 1}"]
        print(name, logits.shape, float(np.abs(logits).max()), float(np.abs(mean).max()))
    path = os.path.join(ROOT, "tests", "golden", "policy.npz")
    np.savez_compressed(path, **{k: np.asarray(v, f32) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
