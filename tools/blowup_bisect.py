"""Why does the rodent blow up under N(0,1) actions?    CPU only, fp64 oracle.

Runs 64 envs from reset for 20 control steps (200 substeps) under the bench's action law at several solver settings and prints, per setting, the fraction
of envs with NaN / |qvel| > 1e4, the first dof to exceed 1e3 rad/s and the substep at which it happens.  Finding (profiles/r2_blowup_bisect.txt): the
state stays bounded (max |qvel| ~ 150 rad/s, no NaN) when the constraint solve is converged (CG 50/50 or Newton 20/20) and diverges within 2 control
steps only with the shipped CG 5/5: the explosion is the truncated solver under saturated +-1 controls (lumbar / cervical tendon actuators of +-20 N
on gram-scale vertebrae push joints ~1 rad past their limits in one substep; five CG iterations over 187 rows do not resolve the limit forces), not the
model compile, the integrator or the implicit damping.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from track_mjx_b200 import clips as CL  # noqa: E402
from track_mjx_b200 import config  # noqa: E402
from track_mjx_b200.walker import Rodent  # noqa: E402


def main():
    w = Rodent(torque_actuators=True, rescale_factor=0.9)
    cl = CL.make_synthetic_clips(w.sections, 2)
    n, nsub = 64, 200
    dof_names = w.dof_names if hasattr(w, "dof_names") else None

    def run(scale, label, **over):
        args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
        args.update(over, physics_steps_per_control_step=1)
        o = Oracle(w.blob, config.make_task_config(w, config.RewardConfig(), **args), cl, dtype=np.float64)
        b = o.alloc(n, debug=False)
        common.put(b, common.init_buffers(b, cl, seed=0))
        o.forward(b, 0)
        rng = np.random.default_rng(42)
        first = None
        peak = 0.0
        for s in range(nsub):
            if s % 10 == 0:
                act = scale * rng.normal(size=(n, w.nu))
            o.step(b, act)
            v = np.nan_to_num(np.abs(b["qvel"]), nan=np.inf)
            if first is None and v.max() > 1e3:
                e, d = np.unravel_index(int(np.argmax(v)), v.shape)
                first = (s, int(d))
            fin = v[np.isfinite(v).all(1)]
            peak = max(peak, float(fin.max()) if fin.size else 0.0)
        bad = (~np.isfinite(b["qvel"]).all(1)) | (np.abs(np.nan_to_num(b["qvel"])).max(1) > 1e4)
        nm = f" ({dof_names[first[1]]})" if (first and dof_names) else ""
        print(f"{label:28s} action scale {scale:4.2f}: diverged envs {bad.mean():5.2f}  first |qvel|>1e3 at substep/dof "
              f"{first}{nm}  peak finite |qvel| {peak:9.3e}", flush=True)

    for sc in (0.0, 0.1, 0.3, 1.0):
        run(sc, "cg 5/5 (shipped)")
    for it in (8, 10, 20, 50):
        run(1.0, f"cg {it}/{it}", iterations=it, ls_iterations=it)
    run(1.0, "cg 5 / ls 50", iterations=5, ls_iterations=50)
    run(1.0, "cg 50 / ls 5", iterations=50, ls_iterations=5)
    run(1.0, "newton 10/10 (configs[4])", solver="newton", iterations=10, ls_iterations=10)
    run(1.0, "newton 20/20", solver="newton", iterations=20, ls_iterations=20)


if __name__ == "__main__":
    main()
