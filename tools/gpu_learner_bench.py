"""Device timings of the learner-side kernels (GAE, observation-normaliser update) at BASELINE configs[2] sizes.

    python tools/gpu_learner_bench.py > gpurun_out/learner_bench.json

Loss head: 20 x 16384 transitions, 38 actions, 60 latents (all inputs read and all gradients written once = 1.9 KB per row;
the timing includes the wrapper's output allocations).  Normaliser: one training step's observations (unroll 20 x 16384 envs x 696 floats = 912 MB, larger than L2); algorithmic
traffic = one read of the batch.  GAE: [20, 16384] x 4 inputs + 2 outputs.  CUDA events on the current stream, 3 warm-ups.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_mjx_b200.learner import Adam, RunningStatistics, compute_gae, ppo_loss_head  # noqa: E402


QUICK = "--quick" in sys.argv        # under ncu: one warm-up, two timed calls


def timed(fn, reps=10):
    reps = 2 if QUICK else reps
    for _ in range(1 if QUICK else 3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    T, B, D = 20, 16384, 696
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    x = torch.randn(T * B, D, device="cuda") * 2 + 1
    st = RunningStatistics(D)
    ms = timed(lambda: st.update(x))
    gb = x.numel() * 4 / 1e9
    out = {"running_stats": {"rows": T * B, "D": D, "ms": ms, "algorithmic_GB": gb, "GBps": gb / (ms * 1e-3)}}
    tr = torch.zeros(T, B, device="cuda")
    te = (torch.rand(T, B, device="cuda") < 0.02).float()
    r, v, bv = torch.rand(T, B, device="cuda"), torch.randn(T, B, device="cuda"), torch.randn(B, device="cuda")
    ms = timed(lambda: compute_gae(tr, te, r, v, bv, 0.95, 0.95), reps=50)
    gb = (6 * T * B + B) * 4 / 1e9
    out["gae"] = {"T": T, "B": B, "ms": ms, "algorithmic_GB": gb, "GBps": gb / (ms * 1e-3)}
    A, Lz = 38, 60
    g = lambda *shape: torch.randn(*shape, device="cuda")
    logits, mu, lv, raw, eps, blp = g(T, B, 2 * A) * 0.5, g(T, B, Lz), g(T, B, Lz) * 0.5 - 1, g(T, B, A), g(T, B, A), g(T, B) * 0.3 - 40
    disc = 1 - te
    ms = timed(lambda: ppo_loss_head(logits, mu, lv, v, bv, r, disc, tr, raw, blp, eps), reps=20)
    # every input once (2A + A + A + 2L + 6 floats per row) + every output once (2A + 2L + 3)
    gb = T * B * (4 * A + 2 * Lz + 6 + 2 * A + 2 * Lz + 3) * 4 / 1e9
    out["ppo_loss_head"] = {"T": T, "B": B, "A": A, "L": Lz, "ms": ms, "algorithmic_GB": gb, "GBps": gb / (ms * 1e-3), "launches": 6, "two_pass": bool(int(os.environ.get("TMJX_PPO_TWO_PASS", "0")))}
    # optimiser step over the intention network's parameter count (2.6 M: L2-resident) and over 256 Mi parameters (HBM-bound)
    for name, n in (("adam_2p6M", 2_600_000), ("adam_256M", 1 << 28)):
        if QUICK and n > 1 << 24:
            continue
        prm, grd = torch.zeros(n, device="cuda"), torch.randn(n, device="cuda")
        o = Adam(prm)
        ms = timed(lambda: o.step(grd), reps=20)
        gb = n * 32 / 1e9          # g twice (norm pass + update), p / mu / nu read and written
        out[name] = {"n": n, "ms": ms, "algorithmic_GB": gb, "GBps": gb / (ms * 1e-3), "launches": 2}
        del prm, grd, o
    # value network forward over the rollout (baseline of the loss): 696 -> 1024 -> 1024 -> 1, TF32 tcgen05 GEMMs
    from track_mjx_b200.policy import ValueNetwork, init_value_params

    rows = B if QUICK else T * B
    net = ValueNetwork(D, init_value_params(D), max_env=B)
    xo = x[:rows]
    ms = timed(lambda: net.apply(xo), reps=5)
    fl = rows * 2.0 * (D * 1024 + 1024 * 1024 + 1024)
    out["value_network"] = {"rows": rows, "ms": ms, "algorithmic_TFLOP": fl / 1e12, "TFLOPs": fl / 1e12 / (ms * 1e-3)}
    critic = (512, 512, 512, 512, 512, 256)                  # critic_layer_sizes of config/rodent-full-clips.yaml:54 (the shipped value network)
    net2 = ValueNetwork(D, init_value_params(D, critic), max_env=B, hidden_layers=critic)
    ms = timed(lambda: net2.apply(xo), reps=5)
    fl = rows * 2.0 * (D * 512 + 4 * 512 * 512 + 512 * 256 + 256)
    out["value_network_shipped_critic"] = {"rows": rows, "hidden": list(critic), "ms": ms, "algorithmic_TFLOP": fl / 1e12, "TFLOPs": fl / 1e12 / (ms * 1e-3)}
    out["peaks"] = {k: peaks.get(k) for k in ("hbm_gbs", "gpu_name")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
