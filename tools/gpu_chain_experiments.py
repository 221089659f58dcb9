"""A/B knobs of the fused policy launch (csrc/tmjx_chain.cuh), one device-timed number each (16384 envs, median of 20 launches):
thread-block clusters with TMA multicast of the weight slices, de-phased odd CTAs, and the work knock-outs that size the kernel's parts
(TMJX_CHAIN_DBG bits: 1 no MMAs, 2 no epilogue work, 4 no operand loads, 8 no prologue, 16 no latent / action heads, 32 no output stores,
64 no SiLU, 128 no TMEM loads -- results are invalid with any bit set).
    python tools/gpu_chain_experiments.py > profiles/rNN_chain_experiments.txt"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch, numpy as np
sys.path.insert(0, %r)
from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, init_params
n = 16384
cfg = IntentionNetworkConfig(); pol = IntentionPolicy(cfg, init_params(cfg, seed=0), max_env=n)
obs = torch.randn(n, cfg.obs_size, device="cuda"); ez = torch.randn(n, cfg.latent_size, device="cuda"); ea = torch.randn(n, cfg.action_size, device="cuda")
for _ in range(5): pol.act(obs, ez, ea)
torch.cuda.synchronize()
ms = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pol.act(obs, ez, ea); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print("%%.4f" %% float(np.median(ms)))
''' % ROOT


def run(env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    return out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else "failed: " + out.stderr[-200:]


print("fused policy launch, 16384 envs, ms per act (median of 20)")
print("default                                   ", run({}))
print("per-layer launches (TMJX_POLICY_FUSED=0)  ", run({"TMJX_POLICY_FUSED": 0}))
for c in (2, 4):
    print(f"TMJX_CHAIN_CLUSTER={c} (weight-slice multicast)", run({"TMJX_CHAIN_CLUSTER": c}))
for ns in (4000, 12000):
    print(f"TMJX_CHAIN_STAGGER_NS={ns:<6d}              ", run({"TMJX_CHAIN_STAGGER_NS": ns}))
for d, what in ((1, "no MMAs"), (2, "no epilogue work"), (4, "no operand loads"), (3, "loads only"), (6, "MMAs only"), (7, "handshakes + prologue + heads"),
                (15, "handshakes + heads"), (31, "handshakes only"), (24, "no prologue, no heads"), (32, "no output stores"), (64, "no SiLU"), (96, "no stores, no SiLU")):
    print(f"TMJX_CHAIN_DBG={d:<3d} ({what})".ljust(42), run({"TMJX_CHAIN_DBG": d}))
