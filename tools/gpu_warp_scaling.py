"""Round time of the step kernel against the number of resident warps per SM (timing experiment, TMJX_DEBUG_WARPS).

One lock-step round over 148 x W environments is launched with k = 1 .. W of the block's W warps alive (the other warps'
environments are skipped, so results are invalid -- only the time of the k live warps per SM is measured).  A latency-bound kernel
shows a flat curve (more warps are free), a throughput-bound one a line through the origin.  Run on a GPU box:
    python tools/gpu_warp_scaling.py > gpurun_out/warp_scaling.txt
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch

    import common
    from track_mjx_b200 import clips as clipmod, config
    from track_mjx_b200.env import Stepper
    from track_mjx_b200.walker import Rodent

    w = Rodent(torque_actuators=True)
    cl = clipmod.make_synthetic_clips(w.sections, 1)
    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    cfg = config.make_task_config(w, config.RewardConfig(), **args)
    epb = int(os.environ["TMJX_ENVS_PER_BLOCK"])
    nenv = 148 * epb
    g = Stepper(w.blob, cfg, cl, nenv, 0)
    hb = {k: np.zeros(tuple(v.shape), np.float32 if v.dtype == torch.float32 else np.int32) for k, v in g.buf.items()}
    common.put(g.buf, common.init_buffers(hb, cl, seed=1))
    g.forward(2)
    acts = [0.1 * torch.randn(nenv, 38, device="cuda") for _ in range(8)]
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
    k = int(os.environ.get("TMJX_DEBUG_WARPS", "0")) or epb
    ms = e0.elapsed_time(e1) / K
    print(f"block of {epb:2d} warps, {k:2d} alive: {ms:7.3f} ms per round -> {ms / k * 1e3:7.1f} us per warp-round, "
          f"{148 * k / ms * 1e3 / 1e6:6.3f} M env-steps/s at this residency", flush=True)
else:
    for epb in (14, 16):
        for k in (1, 2, 4, 6, 8, 10, 12, 14, 16):
            if k > epb:
                continue
            env = dict(os.environ, TMJX_ENVS_PER_BLOCK=str(epb), TMJX_DEBUG_WARPS=str(k))
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
