"""Golden values of the KL-weight schedule, produced by the REFERENCE'S OWN `create_ramp_schedule` text (losses.py:248-290, cut out
with `ast`, annotations dropped, executed with `jax.numpy` bound to numpy float32).  Output: tests/golden/ramp_schedule.npz.

    python tools/make_golden_ramp_schedule.py        # needs /root/reference; output is committed
"""
import ast
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/track_mjx/agent/mlp_ppo/losses.py"


def main():
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "create_ramp_schedule")
    fn.returns = None
    for a in fn.args.args:
        a.annotation = None
    jnp = types.SimpleNamespace(asarray=lambda x, dtype=None: np.asarray(x, dtype=dtype), float32=np.float32, clip=np.clip, where=np.where,
                                pi=np.pi, cos=np.cos, sin=np.sin)
    ns = {"jnp": jnp}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    make = ns["create_ramp_schedule"]
    steps = np.array([0, 1, 2, 5, 10, 44, 45, 99, 100, 101, 500, 1000, 5000], np.int64)
    cases = [dict(max_value=0.1, min_value=0.0001, ramp_steps=1000, warmup_steps=0, schedule="linear", period=45),
             dict(max_value=1.0, min_value=0.001, ramp_steps=100, warmup_steps=5, schedule="linear", period=45),
             dict(max_value=0.1, min_value=0.0001, ramp_steps=1000, warmup_steps=0, schedule="cosine", period=45),
             dict(max_value=0.5, min_value=0.01, ramp_steps=1000, warmup_steps=0, schedule="sine", period=30)]
    out = {"steps": steps, "n_cases": np.int64(len(cases))}
    for i, kw in enumerate(cases):
        f = make(**kw)
        out[f"c{i}_values"] = np.array([np.float32(f(int(s))) for s in steps], np.float32)
        for k, v in kw.items():
            out[f"c{i}_{k}"] = np.array(v)
        print(kw, out[f"c{i}_values"][:6])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ramp_schedule.npz"), **out)


if __name__ == "__main__":
    main()
