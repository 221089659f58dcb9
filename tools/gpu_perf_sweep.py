"""Timing sweep on the GPU box: step time vs number of envs for the kernel variants selected by env knobs.

    python tools/gpu_perf_sweep.py [n_env ...]      (TMJX_ENVS_PER_BLOCK=4|7, TMJX_NO_GEN=1 select variants)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from track_mjx_b200 import clips as clipmod, config  # noqa: E402
from track_mjx_b200.env import Stepper  # noqa: E402
from track_mjx_b200.walker import Rodent  # noqa: E402

w = Rodent(torque_actuators=True)
cl = clipmod.make_synthetic_clips(w.sections, 1)
args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
cfg = config.make_task_config(w, config.RewardConfig(), **args)
sizes = [int(x) for x in sys.argv[1:]] or [148, 1776, 2072, 4096, 8192]
tag = f"epb={os.environ.get('TMJX_ENVS_PER_BLOCK', 'auto')} spill={os.environ.get('TMJX_L2_SPILL', '1')} nogen={os.environ.get('TMJX_NO_GEN', '0')}"
for nenv in sizes:
    g = Stepper(w.blob, cfg, cl, nenv, 0)
    hb = {k: np.zeros(tuple(v.shape), np.float32 if v.dtype == torch.float32 else np.int32) for k, v in g.buf.items()}
    common.put(g.buf, common.init_buffers(hb, cl, seed=1))
    g.forward(2)
    acts = [torch.randn(nenv, 38, device="cuda") for _ in range(8)]
    for i in range(3):
        g.step(acts[i], 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for i in range(K):
        g.step(acts[i % 8], 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"[{tag}] n_env {nenv:6d}: {ms:8.3f} ms/step -> {nenv / ms * 1e3:10.0f} env-steps/s  (done frac {float(g.buf['done'].mean()):.3f})", flush=True)
    g.close()
