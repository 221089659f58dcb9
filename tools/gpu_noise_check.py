"""GPU and fp32-oracle errors against the fp64 oracle after one physics substep and one 10-substep control step (256 envs,
N(0,1) actions): shows that the env-to-env spread of a full control step is fp32 noise amplified by the unconverged CG (the fp32
oracle has the same error distribution), not a property of the CUDA path.  Run on a GPU box: python tools/gpu_noise_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common
from oracle.oracle import Oracle
from track_mjx_b200 import clips as clipmod, config
from track_mjx_b200.env import Stepper
from track_mjx_b200.walker import Rodent
w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 2)
args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
for nf in (1, 10):
    a2 = dict(args); a2["physics_steps_per_control_step"] = nf
    cfg = config.make_task_config(w, config.RewardConfig(), **a2)
    n = 256
    o32, o64 = Oracle(w.blob, cfg, cl, dtype=np.float32), Oracle(w.blob, cfg, cl, dtype=np.float64)
    g = Stepper(w.blob, cfg, cl, n, 0)
    b32, b64 = o32.alloc(n, debug=False), o64.alloc(n, debug=False)
    init = common.init_buffers(b32, cl, seed=3)
    for b in (b32, b64, g.buf): common.put(b, init)
    o32.forward(b32); o64.forward(b64); g.forward()
    act = np.random.default_rng(0).normal(size=(n, g.dims["nu"])).astype(np.float32)
    o32.step(b32, act); o64.step(b64, act); g.step(torch.from_numpy(act).cuda()); torch.cuda.synchronize()
    gb = common.get(g.buf)
    eg = np.abs(gb["qpos"] - b64["qpos"]).max(axis=1); eo = np.abs(b32["qpos"] - b64["qpos"]).max(axis=1)
    q = [50, 90, 99, 100]
    print(f"n_frames={nf}: |gpu - oracle64| percentiles {np.percentile(eg, q).round(5).tolist()}   |oracle32 - oracle64| {np.percentile(eo, q).round(5).tolist()}")
    print("   envs with gpu err > 0.01:", int((eg > 0.01).sum()), " oracle32 err > 0.01:", int((eo > 0.01).sum()))
