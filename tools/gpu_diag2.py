"""Calibration run: sub-step and control-step parity on contact-rich states + first timing (GPU box)."""
import os
import sys
import time

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

n = 64
w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 2)


def mk(nf):
    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    args["physics_steps_per_control_step"] = nf
    return config.make_task_config(w, config.RewardConfig(), **args)


keys = ["qpos", "qvel", "act", "xpos", "qfrc_actuator", "qacc_warmstart", "obs", "dbg_qfrc_bias", "dbg_qacc_smooth", "dbg_contact_dist",
        "dbg_efc_force", "dbg_qfrc_constraint", "dbg_qacc", "reward", "metrics"]


def report(tag, gb, a, b):
    print(f"--- {tag}")
    for k in keys:
        e_g32 = common.err(gb[k], a[k]); e_g64 = common.err(gb[k], b[k]); e_3264 = common.err(a[k], b[k])
        print(f"{k:22s} gpu-vs-o32 abs {e_g32[0]:.3e} rel {e_g32[1]:.3e} | gpu-vs-o64 rel {e_g64[1]:.3e} | o32-vs-o64 rel {e_3264[1]:.3e}")
    print("done mismatches:", int((gb["done"] != a["done"]).sum()), "cur_frame mismatches:", int((gb["cur_frame"] != a["cur_frame"]).sum()),
          "dones:", int(a["done"].sum()), "active contacts/env:", float((a["dbg_contact_dist"] < 0).sum(1).mean()),
          "active limits/env:", float((a["dbg_efc_force"][:, :67] > 0).sum(1).mean()))


for scale in (0.0, 0.1):
    cfg10, cfg1 = mk(10), mk(1)
    roll = Oracle(w.blob, cfg10, cl, dtype=np.float32)
    rb = roll.alloc(n)
    common.put(rb, common.init_buffers(rb, cl, seed=0))
    roll.forward(rb)
    rng = np.random.default_rng(7)
    o32_1, o64_1 = Oracle(w.blob, cfg1, cl, dtype=np.float32), Oracle(w.blob, cfg1, cl, dtype=np.float64)
    o32_10, o64_10 = Oracle(w.blob, cfg10, cl, dtype=np.float32), Oracle(w.blob, cfg10, cl, dtype=np.float64)
    g1, g10 = Stepper(w.blob, cfg1, cl, n, 0, debug=True), Stepper(w.blob, cfg10, cl, n, 0, debug=True)
    for s in range(13):
        act = (scale * rng.normal(size=(n, 38))).astype(np.float32)
        if s in (3, 7, 12):
            st = common.get(rb, common.STATE_KEYS)
            for (o32, o64, g, tag) in ((o32_1, o64_1, g1, "substep"), (o32_10, o64_10, g10, "ctrlstep")):
                a, b = o32.alloc(n), o64.alloc(n)
                common.put(a, st); common.put(b, st); common.put(g.buf, st)
                o32.step(a, act); o64.step(b, act); g.step(torch.from_numpy(act).cuda()); torch.cuda.synchronize()
                report(f"scale {scale} after {s} ctrl steps: {tag}", common.get(g.buf), a, b)
        roll.step(rb, act)
    g1.close(); g10.close()

# ---- first timing
for nenv in (4096, 16384):
    cfg10 = mk(10)
    g = Stepper(w.blob, cfg10, cl, nenv, 0)
    hb = {k: np.zeros(tuple(v.shape), np.float32 if v.dtype == torch.float32 else np.int32) for k, v in g.buf.items()}
    common.put(g.buf, common.init_buffers(hb, cl, seed=1))
    g.forward(2)
    for scale in (0.0, 1.0):
        acts = [torch.randn(nenv, 38, device="cuda") * scale for _ in range(8)]
        for i in range(3):
            g.step(acts[i], 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 20
        for i in range(K):
            g.step(acts[i % 8], 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"n_env {nenv} action scale {scale}: {ms:.3f} ms/step -> {nenv / ms * 1e3:.0f} env-steps/s; nan frac {float(g.buf['metrics'][:, 14].mean()):.3f} done frac {float(g.buf['done'].mean()):.3f}")
    print("fp32 peak TFLOP/s", g.fp32_peak_tflops())
    g.close()
