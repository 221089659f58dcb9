"""Edge sizes through the public env API on one GPU: 1, 13, 15 and 65536 environments step without error, stay finite, and the
first 4096 rows of a 65536-env batch equal a 4096-env batch with the same per-env inputs (results do not depend on the grid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common
from track_mjx_b200 import clips as clipmod, config
from track_mjx_b200.env import Stepper
from track_mjx_b200.walker import Rodent
from track_mjx_b200 import _lib as L

w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 3)
args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
cfg = config.make_task_config(w, config.RewardConfig(), **args)
ref = None
big_init = None
for n in (65536, 4096, 1, 13, 15):
    g = Stepper(w.blob, cfg, cl, n, 0)
    if big_init is None:
        host = {k: (v.cpu().numpy()) for k, v in g.buf.items()}
        big_init = common.init_buffers(host, cl, seed=1)
    init = {k: v[:n] for k, v in big_init.items()}     # env e gets the same inputs whatever the batch size
    common.put(g.buf, init)
    g.forward(L.TMJX_F_SNAPSHOT)
    act = torch.from_numpy(np.random.default_rng(5).normal(size=(65536, w.nu)).astype(np.float32)[:n].copy()).cuda()
    for _ in range(3):
        g.step(act, L.TMJX_F_AUTORESET)
    torch.cuda.synchronize()
    out = {k: g.buf[k].cpu().numpy() for k in ("qpos", "obs", "reward", "done")}
    assert all(np.isfinite(v).all() for v in out.values()), n
    if n == 65536:
        ref = out
    else:
        same = all(np.array_equal(out[k], ref[k][:n]) for k in out)
        print(f"{n}-env batch bitwise equal to the first {n} rows of the 65536-env batch:", same)
        assert same
    print("n", n, "ok  mean reward", float(out["reward"].mean()), "done frac", float(out["done"].mean()))
    g.close()
