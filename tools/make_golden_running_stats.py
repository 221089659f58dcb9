"""Golden vectors for the observation normaliser update, produced by the REFERENCE'S OWN `running_statistics.update`.

`track_mjx/agent/masked_running_statistics.py` is plain jax.numpy apart from its imports (brax types for annotations, flax's
struct.dataclass, jax.tree_util, lax.psum).  The unmodified module text is executed with those names bound to numpy-backed
stand-ins (arrays are the only pytree leaves used here; psum is the identity without a pmap axis) and fed observation batches;
inputs, the state before and the state after are stored in `tests/golden/running_stats.npz`.

    python tools/make_golden_running_stats.py        # needs /root/reference; output is committed
"""
import dataclasses
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/track_mjx/agent/masked_running_statistics.py"


def load_reference_module():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.ndarray = np.ndarray
    jnp.array = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    jnp.zeros = lambda shape, dtype=np.float32: np.zeros(shape, dtype)
    jnp.ones = lambda shape, dtype=np.float32: np.ones(shape, dtype)
    jnp.float32, jnp.float64 = np.float32, np.float64
    # jnp.sum takes any iterable of axes; jax promotes float32 + int32 -> float32 (numpy would go to float64), so the integer
    # step increment is handed back as a float32 scalar (exact below 2^24 samples per update)
    jnp.sum = lambda a, axis=None: np.sum(a, axis=tuple(axis) if isinstance(axis, range) else axis)
    jnp.prod = lambda a: np.float32(np.prod(a))
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.config = types.SimpleNamespace(jax_enable_x64=False)
    jax.tree_util = types.SimpleNamespace(tree_map=lambda f, *xs: f(*xs), tree_structure=lambda x: "leaf", tree_leaves=lambda x: [x])
    jax.lax = types.SimpleNamespace(psum=lambda x, axis_name=None: x)
    flax = types.ModuleType("flax")
    flax.struct = types.SimpleNamespace(dataclass=dataclasses.dataclass)
    acme = types.ModuleType("brax.training.acme")
    acme.types = types.SimpleNamespace(Nest=object, NestedArray=object)
    saved = {k: sys.modules.get(k) for k in ("jax", "jax.numpy", "flax", "brax", "brax.training", "brax.training.acme")}
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "flax": flax, "brax": types.ModuleType("brax"),
                        "brax.training": types.ModuleType("brax.training"), "brax.training.acme": acme})
    try:
        mod = types.ModuleType("ref_running_statistics")
        exec(compile(open(REF).read(), REF, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    ref = load_reference_module()
    rng = np.random.default_rng(3)
    out = {}
    D = 696
    scale = rng.uniform(0.05, 3.0, D).astype(np.float32)
    shift = rng.normal(0, 1.0, D).astype(np.float32)
    state = ref.init_state(np.zeros(D, np.float32))
    for i, shape in enumerate([(4, 5), (10, 16), (200,)]):
        batch = (rng.normal(size=shape + (D,)).astype(np.float32) * scale + shift).astype(np.float32)
        batch[..., 10] = 2.5        # a constant feature: std is clipped at std_min_value
        new = ref.update(state, batch)
        out.update({f"u{i}_batch": batch.reshape(-1, D), f"u{i}_count0": np.float32(state.count), f"u{i}_mean0": state.mean,
                    f"u{i}_sv0": state.summed_variance, f"u{i}_std0": state.std, f"u{i}_count1": np.float32(new.count),
                    f"u{i}_mean1": np.asarray(new.mean, np.float32), f"u{i}_sv1": np.asarray(new.summed_variance, np.float32),
                    f"u{i}_std1": np.asarray(new.std, np.float32)})
        state = new
    x = rng.normal(size=(7, D)).astype(np.float32)
    out["norm_in"] = x
    out["norm_out"] = np.asarray(ref.normalize(x, state), np.float32)
    path = os.path.join(ROOT, "tests", "golden", "running_stats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "counts", [float(out[f"u{i}_count1"]) for i in range(3)], "std range", float(state.std.min()), float(state.std.max()))


if __name__ == "__main__":
    main()
