"""Stage-by-stage comparison of the CUDA path against the CPU oracle (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from track_mjx_b200 import clips as clipmod, config  # noqa: E402
from track_mjx_b200.env import Stepper  # noqa: E402
from track_mjx_b200.walker import Rodent  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 2)
args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
cfg = config.make_task_config(w, config.RewardConfig(), **args)
o32 = Oracle(w.blob, cfg, cl, dtype=np.float32)
o64 = Oracle(w.blob, cfg, cl, dtype=np.float64)
g = Stepper(w.blob, cfg, cl, n, 0, debug=True)
print("dims", g.dims)
b32, b64 = o32.alloc(n), o64.alloc(n)
init = common.init_buffers(b32, cl, seed=0)
for b in (b32, b64, g.buf):
    common.put(b, init)
o32.forward(b32); o64.forward(b64); g.forward(); torch.cuda.synchronize()
keys = ["qpos", "xpos", "xquat", "qfrc_actuator", "qacc_warmstart", "obs", "dbg_subtree_com", "dbg_qfrc_bias", "dbg_qacc_smooth",
        "dbg_contact_dist", "dbg_efc_force", "dbg_qfrc_constraint", "dbg_qacc", "cur_frame"]


def report(tag, gb, a, b):
    print(f"--- {tag}")
    for k in keys + ["qvel", "act", "time", "reward", "done", "metrics"]:
        if k not in gb:
            continue
        x, y, z = gb[k], a[k], b[k]
        e_g32 = common.err(x, y); e_g64 = common.err(x, z); e_3264 = common.err(y, z)
        print(f"{k:22s} gpu-vs-o32 abs {e_g32[0]:.3e} rel {e_g32[1]:.3e} | gpu-vs-o64 rel {e_g64[1]:.3e} | o32-vs-o64 rel {e_3264[1]:.3e}")


report("forward", common.get(g.buf), b32, b64)
rng = np.random.default_rng(42)
for s in range(nsteps):
    act = rng.normal(size=(n, g.dims["nu"])).astype(np.float32)
    # single-step parity: all three start from the fp32 oracle's state
    st = common.get(b32, common.STATE_KEYS)
    common.put(g.buf, st); common.put(b64, st)
    o32.step(b32, act); o64.step(b64, act); g.step(torch.from_numpy(act).cuda()); torch.cuda.synchronize()
    gb = common.get(g.buf)
    report(f"step {s}", gb, b32, b64)
    print("done gpu/o32 mismatches:", int((gb["done"] != b32["done"]).sum()), "cur_frame mismatches:", int((gb["cur_frame"] != b32["cur_frame"]).sum()),
          "dones:", int(b32["done"].sum()))
