"""Critical-path breakdown of one environment's control step: clock64() cycles per phase, measured by warp 0 of block 0 of the
14-warp CG kernel in a development build (tools/build_dev_lib.sh pt -DTMJX_PHASE_TIMING -> libtmjx_pt.so; the product library carries no timers).

    bash tools/build_dev_lib.sh pt -DTMJX_PHASE_TIMING && TMJX_LIB_PATH=track-mjx_b200/csrc/libtmjx_pt.so python tools/gpu_phase_timing.py [warps alive]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("TMJX_LIB_PATH", os.path.join(ROOT, "track-mjx_b200", "csrc", "libtmjx_pt.so"))
os.environ["TMJX_ENVS_PER_BLOCK"] = "14"
if len(sys.argv) > 1:
    os.environ["TMJX_DEBUG_WARPS"] = sys.argv[1]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
from track_mjx_b200 import _lib as L  # noqa: E402
from track_mjx_b200 import clips as clipmod, config  # noqa: E402
from track_mjx_b200.env import Stepper  # noqa: E402
from track_mjx_b200.walker import Rodent  # noqa: E402

NAMES = {0: "substep barrier", 1: "kinematics", 2: "com_pos (com, cinert, cdof)", 3: "com_vel + rne (bias)", 4: "passive + actuation",
         5: "crb + build M", 6: "M * warmstart", 7: "factor M and M + h D", 8: "park Euler factor head (global)", 9: "qacc_smooth solve",
         10: "contacts + make_constraint (+ J qvel)", 11: "solve_cg total", 12: "euler (unpark + solve + integrate)",
         13: "kernel prologue + substeps (whole)", 14: "epilogue (reward, obs, write-back)",
         16: "  cg: warm-start choice (2 x J x, costs)", 17: "  cg: init (J^T f, solve)", 18: "  cg: iteration head (norms, vput)",
         19: "  cg: J search", 20: "  cg: line search", 21: "  cg: state update", 22: "  cg: update_constraint (J^T f)",
         23: "  cg: update_gradient (solve)", 24: "  cg: Polak-Ribiere"}

w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 1)
args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
cfg = config.make_task_config(w, config.RewardConfig(), **args)
nenv = 148 * 14
g = Stepper(w.blob, cfg, cl, nenv, 0)
lib = L.load()
lib.tmjx_debug_phase_times.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
hb = {k: np.zeros(tuple(v.shape), np.float32 if v.dtype == torch.float32 else np.int32) for k, v in g.buf.items()}
common.put(g.buf, common.init_buffers(hb, cl, seed=1))
g.forward(2)
acts = [0.1 * torch.randn(nenv, 38, device="cuda") for _ in range(8)]
for i in range(3):
    g.step(acts[i], 1)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
assert lib.tmjx_debug_phase_times(buf, 1) == 0
K = 10
for i in range(K):
    g.step(acts[i % 8], 1)
torch.cuda.synchronize()
assert lib.tmjx_debug_phase_times(buf, 0) == 0
t = np.array(list(buf), np.float64) / K
nsub = cfg.physics_steps_per_control_step
total = t[13] + t[14]
print(f"warps alive per block: {os.environ.get('TMJX_DEBUG_WARPS', '14')};  cycles per control step (warp 0 of block 0): {total:,.0f}  "
      f"= {total / 1.965e9 * 1e3:.3f} ms at 1.965 GHz;  per substep {t[13] / nsub:,.0f}")
for i in sorted(NAMES):
    if i in (13,):
        continue
    per = t[i] / (1 if i == 14 else nsub)
    print(f"  {NAMES[i]:48s} {per:10,.0f} cycles {'per step   ' if i == 14 else 'per substep'}  {100 * t[i] / total:5.1f} %")
