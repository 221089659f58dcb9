"""Device-timed acting policy and value network: the fused chain launch (csrc/tmjx_chain.cuh) against the per-layer launches.
    python tools/gpu_policy_bench.py [n_env]   -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, ValueNetwork, init_params, init_value_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = IntentionNetworkConfig()
p = init_params(cfg, seed=0)
hidden = (512, 512, 512, 512, 512, 256)
vp = init_value_params(cfg.obs_size, hidden, seed=1)
obs = torch.randn(n, cfg.obs_size, device="cuda")
ez, ea = torch.randn(n, cfg.latent_size, device="cuda"), torch.randn(n, cfg.action_size, device="cuda")
flops_pol = 2 * (470 * 1024 + 1024 * 512 + 3 * 512 * 512 + 512 * 120 + 286 * 512 + 2 * 512 * 512 + 512 * 256 + 256 * 256 + 256 * 76)
flops_val = 2 * (696 * 512 + 4 * 512 * 512 + 512 * 256 + 256)
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, iters=30, flush_l2=False):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


res = {"n_env": n}
for fused in (1, 0):
    os.environ["TMJX_POLICY_FUSED"] = str(fused)
    pol = IntentionPolicy(cfg, p, max_env=n)
    val = ValueNetwork(cfg.obs_size, vp, max_env=n, hidden_layers=hidden)
    tag = "fused" if fused else "per_layer"
    t = timed(lambda: pol.act(obs, ez, ea))
    tv = timed(lambda: val.apply(obs))
    res[tag] = {"launches_per_act": pol.launches_per_act, "act_ms": t, "act_tflops": flops_pol * n / t / 1e9, "act_ms_l2_flushed": timed(lambda: pol.act(obs, ez, ea), 10, True),
                "value_ms": tv, "value_tflops": flops_val * n / tv / 1e9}
    pol.close(); val.close()
print(json.dumps(res))
