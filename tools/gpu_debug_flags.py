"""Debug: which envs flip a termination flag between the CUDA path and the fp32 oracle, and how close to the threshold."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from oracle.oracle import Oracle
from track_mjx_b200 import _lib as L, clips as clipmod, config
from track_mjx_b200.env import Stepper
from track_mjx_b200.walker import Rodent
from test_gpu_parity import make_cfg, rollout_states

w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 2)
n = 1024
st = rollout_states(w, cl, n, 6, 0.3, seed=4)
cfg = make_cfg(w, physics_steps_per_control_step=1)
o32, o64 = Oracle(w.blob, cfg, cl, dtype=np.float32), Oracle(w.blob, cfg, cl, dtype=np.float64)
g = Stepper(w.blob, cfg, cl, n, 0, debug=True)
a, b = o32.alloc(n), o64.alloc(n)
common.put(a, st); common.put(b, st); common.put(g.buf, st)
act = (0.3 * np.random.default_rng(9).normal(size=(n, w.nu))).astype(np.float32)
o32.step(a, act); o64.step(b, act); g.step(torch.from_numpy(act).cuda())
gb = common.get(g.buf)
m = config.METRIC_NAMES
for name in ("too_far", "bad_pose", "bad_quat", "fall", "nan", "done"):
    i = m.index(name)
    bad = np.nonzero(gb["metrics"][:, i] != a["metrics"][:, i])[0]
    print(name, "gpu!=o32:", bad.tolist(), " o32!=o64:", np.nonzero(a["metrics"][:, i] != b["metrics"][:, i])[0].tolist())
qd = m.index("quat_distance")
bad = np.nonzero(gb["metrics"][:, m.index("bad_quat")] != a["metrics"][:, m.index("bad_quat")])[0]
for e in bad[:8]:
    print("env", e, "quat_distance gpu/o32/o64", gb["metrics"][e, qd], a["metrics"][e, qd], b["metrics"][e, qd],
          "qpos err gpu-o32", np.abs(gb["qpos"][e] - a["qpos"][e]).max(), "o32-o64", np.abs(a["qpos"][e] - b["qpos"][e]).max(),
          "qvel max", np.abs(a["qvel"][e]).max(), "qvel err", np.abs(gb["qvel"][e] - a["qvel"][e]).max(),
          "in quat", st["qpos"][e, 3:7], "nan in", np.isnan(st["qpos"][e]).any())
for k in ("qpos", "qvel", "dbg_qacc", "dbg_qacc_smooth", "dbg_efc_force"):
    eg = np.abs(gb[k].astype(np.float64) - a[k]).max(1); en = np.abs(a[k].astype(np.float64) - b[k]).max(1)
    sc = np.abs(b[k]).max(1) + 1e-30
    print(k, "rel err per env: gpu-o32 median %.2e max %.2e | o32-o64 median %.2e max %.2e" % (np.nanmedian(eg / sc), np.nanmax(eg / sc), np.nanmedian(en / sc), np.nanmax(en / sc)))
