"""Timeline of one CTA of the fused policy launch (csrc/tmjx_chain.cuh): clock64() marks per layer.
    python tools/gpu_chain_trace.py [n_env]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from track_mjx_b200 import _lib as L  # noqa: E402
from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, init_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = IntentionNetworkConfig()
pol = IntentionPolicy(cfg, init_params(cfg, seed=0), max_env=n)
lib = L.load()
lib.tmjx_policy_chain_trace.argtypes = [C.c_int, C.POINTER(C.c_longlong)]
obs = torch.randn(n, cfg.obs_size, device="cuda")
ez, ea = torch.randn(n, cfg.latent_size, device="cuda"), torch.randn(n, cfg.action_size, device="cuda")
for _ in range(3):
    pol.act(obs, ez, ea)
torch.cuda.synchronize()
assert lib.tmjx_policy_chain_trace(1, None) == 0
pol.act(obs, ez, ea)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 14))()
assert lib.tmjx_policy_chain_trace(0, buf) == 0
t = np.array(list(buf), np.int64).reshape(14, 8)
t0 = t[0, 0]
us = lambda c: (c - t0) / 1965.0
names = ["enc 470-1024", "enc 1024-512", "enc 512-512", "enc 512-512", "enc 512-512", "head 512-120", "dec 286-512", "dec 512-512", "dec 512-512", "dec 512-256",
         "dec 256-256", "logits 256-76"]
print("TMJX_CHAIN_DBG =", os.environ.get("TMJX_CHAIN_DBG", "0"))
print("microseconds since the first layer's input was ready (CTA 0, 1.965 GHz)")
print(f"{'layer':14s} {'ready':>8s} {'1st MMA':>8s} {'MMAs issued':>11s} {'tfull c0':>9s} {'tfull last':>10s} {'pass1 done':>10s} {'layer done':>10s} | {'MMA span':>8s} {'tail':>6s}")
for l, nm in enumerate(names):
    r = t[l]
    print(f"{nm:14s} {us(r[0]):8.2f} {us(r[1]):8.2f} {us(r[2]):11.2f} {us(r[3]):9.2f} {us(r[4]):10.2f} {us(r[5]):10.2f} {us(r[6]):10.2f} | {us(r[4]) - us(r[1]):8.2f} {us(r[6]) - us(r[4]):6.2f}")
