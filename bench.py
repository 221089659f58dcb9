#!/usr/bin/env python
"""Headline benchmark: rodent-tracking env-steps/sec (device-timed), BASELINE.json `configs[1]`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one control step of the wrapped tracking env over one batch of 4096 envs per GPU
(10 physics substeps + reward + termination + observation + fused episode/auto-reset), random N(0,1)
actions.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

  value     whole-job env-steps/s, actions already resident in HBM, CUDA events around each step launch,
            L2 flushed between timed iterations, max over ranks;
  e2e       the same through the Python env API with HOST buffers: every step copies the actions from
            pinned host memory and reads obs / reward / done back to pinned host memory;
  roofline  FP32-pipe (the binding roofline of this path) and HBM fractions of the step kernel;
  cpu_baseline  the CPU oracle (a port of the reference algorithm, `oracle/`) on a bounded sample.

`--impl reference` times the reference's CPU path.  The reference itself (JAX + MuJoCo-MJX + Brax) cannot be
installed in this image (no wheels, no network), so that arm runs the oracle port on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "rodent-tracking env-steps/sec (device-timed)"
UNIT = "env-steps/s"
ENVS_PER_GPU = 4096
# BASELINE.json configs: the default bench line is configs[1]; the others are selectable parity / scaling cases
WORKLOADS = {
    "tracking": dict(envs=4096, n_clips=1, policy=False, env_args={}, w_alg=None,
                     name="rodent single-clip tracking, 4096 envs per B200, full step + reward + obs + fused auto-reset, fp32 (BASELINE configs[1])"),
    "intention": dict(envs=16384, n_clips=842, policy=True, env_args={}, w_alg=None,
                      name="rodent-mc-intention multi-clip, 842 synthetic clips, 16384 envs per B200, in-loop intention-network policy "
                           "inference on tcgen05 (tf32) + full step (BASELINE configs[2])"),
    "contact": dict(envs=4096, n_clips=1, policy=False,
                    env_args=dict(solver="newton", iterations=10, ls_iterations=10, physics_steps_per_control_step=20), w_alg=None,
                    name="contact-heavy rodent tracking: Newton solver, 10 iterations / 10 line-search iterations, 20 physics steps per "
                         "control step, 4096 envs per B200 (BASELINE configs[4], 32768 envs on 8 GPUs)"),
}
# algorithmic work per env-step at the default config (formulas in DESIGN.md "Roofline")
B_ALG = 15348.0            # bytes of unavoidable HBM traffic per env-step
W_ALG = 3.775e6            # structure-exploiting FLOPs per env-step (10 x 375 kFLOP + 25 kFLOP)


def workload_config(n_gpus, workload="tracking"):
    from track_mjx_b200 import config

    w = WORKLOADS[workload]
    ea = dict(config.DEFAULT_ENV_ARGS)
    ea.update(w["env_args"])
    return {
        "workload": w["name"],
        "envs_per_gpu": w["envs"], "global_envs": w["envs"] * n_gpus, "n_clips": w["n_clips"], "clip_length": 250,
        "physics_steps_per_control_step": ea["physics_steps_per_control_step"], "solver": ea["solver"], "iterations": ea["iterations"],
        "ls_iterations": ea["ls_iterations"],
        "actions": ("tanh-normal samples of the seed-0 LeCun-uniform intention network on the env's own observations" if w["policy"]
                    else "N(0,1) per step (clipped to ctrlrange by the actuator model); under the shipped CG 5/5 solve this law makes ~3 % of the env-steps end in NaN (nan_frac; profiles/r2_blowup_bisect.txt), which the reference's NaN => done rule and the auto-reset absorb"),
        "l2": "flushed between timed iterations",
        "parallelism": f"env-sharded x{n_gpus}, no data-path collective",
    }


def build_env_pieces(n_clips=1):
    from track_mjx_b200 import clips as clipmod, config
    from track_mjx_b200.walker import Rodent

    walker = Rodent(torque_actuators=True, rescale_factor=0.9)
    clips = clipmod.make_synthetic_clips(walker.sections, n_clips)
    return walker, clips, config


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm, all host threads, bounded sample per step."""
    if rank != 0:
        return
    import numpy as np

    import common
    from oracle.oracle import Oracle

    walker, clips, config = build_env_pieces(1)
    env_args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    cfg = config.make_task_config(walker, config.RewardConfig(), **env_args)
    cores = os.cpu_count() or 1
    orc = Oracle(walker.blob, cfg, clips, dtype=np.float32, nthreads=cores)
    sample = min(ENVS_PER_GPU, max(64, 4 * cores))
    buf = orc.alloc(sample, debug=False)
    common.put(buf, common.init_buffers(buf, clips, seed=0))
    from track_mjx_b200 import _lib as L

    orc.forward(buf, L.TMJX_F_SNAPSHOT)
    rng = np.random.default_rng(42)
    acts = [rng.normal(size=(sample, walker.nu)).astype(np.float32) for _ in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        orc.step(buf, acts[i], L.TMJX_F_AUTORESET)
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.step(buf, acts[args.warmup + i], L.TMJX_F_AUTORESET)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of {ENVS_PER_GPU} envs per step x {args.steps} control steps, fp32 oracle port (dense MJX-style algebra), OpenMP over envs"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (JAX/MJX/Brax) not installable in this image; this is the CPU oracle port of its algorithm",
    }
    emit(out)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    from track_mjx_b200.env import MultiClipTracking, wrap
    from track_mjx_b200.sharding import Shard, max_over_ranks, reduce_episode_stats

    wl = WORKLOADS[args.workload]
    ENVS_PER_GPU = wl["envs"]
    shard = Shard(rank, world, ENVS_PER_GPU * world)   # weak scaling: fixed envs per GPU, contiguous global env ids per rank
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    walker, clips, config = build_env_pieces(wl["n_clips"])
    env_args = dict(config.DEFAULT_ENV_ARGS)
    env_args.update(wl["env_args"])
    env = wrap(MultiClipTracking(clips, walker, config.RewardConfig(), num_envs=ENVS_PER_GPU, device=local_rank, **env_args))
    state = env.reset(shard.seed(1000))
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    n_act = min(K + W, 64)
    policy = None
    if wl["policy"]:
        from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, init_params

        pcfg = IntentionNetworkConfig(obs_size=env.observation_size, reference_obs_size=env.stepper.dims["reference_obs_size"],
                                      action_size=env.action_size)
        policy = IntentionPolicy(pcfg, init_params(pcfg, seed=0), max_env=ENVS_PER_GPU, device=local_rank)
        eps = [(torch.randn(ENVS_PER_GPU, pcfg.latent_size, device=dev, generator=gen),
                torch.randn(ENVS_PER_GPU, pcfg.action_size, device=dev, generator=gen)) for _ in range(min(n_act, 8))]
        acts = None
    else:
        acts = [args.action_scale * torch.randn(ENVS_PER_GPU, env.action_size, device=dev, generator=gen) for _ in range(n_act)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def next_action(i, st):
        """random actions, or the in-loop policy on the current observations (acting half of ppo.py:333-340)"""
        if policy is None:
            return acts[i % n_act]
        ez, ea = eps[i % len(eps)]
        return policy.act(st.obs, ez, ea)[0]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident arm
    for i in range(W):
        state = env.step(state, next_action(i, state))
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    rew_sum = torch.zeros((), device=dev)
    done_sum = torch.zeros((), device=dev)
    nan_sum = torch.zeros((), device=dev)
    for i in range(K):
        flush.zero_()
        evs[i][0].record()
        a_i = next_action(W + i, state)
        evs[i][1].record()
        state = env.step(state, a_i)
        evs[i][2].record()
        rew_sum += state.reward.sum()
        done_sum += state.done.sum()
        nan_sum += state.metrics["nan"].sum()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(c) for a, _, c in evs)
    kern_ms = sum(b.elapsed_time(c) for _, b, c in evs)      # the step kernel alone
    pol_ms = sum(a.elapsed_time(b) for a, b, _ in evs)       # the policy kernels (0 launches for random actions)
    clocks = sampler.stop() if rank == 0 else None
    max_ms = max_over_ranks(dev_ms, dev, shard)                      # device time, max over ranks
    stats = reduce_episode_stats(rew_sum, done_sum, K, shard, nan_sum=nan_sum)   # SUM over NVLink: the only collective of the env path
    value = ENVS_PER_GPU * world * K / (max_ms * 1e-3)

    # ---- end-to-end arm: host buffers in, host buffers out, every step
    obs_dim = env.observation_size
    h_act = [torch.randn(ENVS_PER_GPU, env.action_size).pin_memory() for _ in range(4)]
    d_act = torch.empty(ENVS_PER_GPU, env.action_size, device=dev)
    if policy is not None:
        h_ez = [torch.randn(ENVS_PER_GPU, pcfg.latent_size).pin_memory() for _ in range(4)]
        h_ea = [torch.randn(ENVS_PER_GPU, pcfg.action_size).pin_memory() for _ in range(4)]
        d_ez = torch.empty(ENVS_PER_GPU, pcfg.latent_size, device=dev)
        d_ea = torch.empty(ENVS_PER_GPU, pcfg.action_size, device=dev)
    h_obs = torch.empty(ENVS_PER_GPU, obs_dim).pin_memory()
    h_rew = torch.empty(ENVS_PER_GPU).pin_memory()
    h_done = torch.empty(ENVS_PER_GPU).pin_memory()

    def e2e_step(i, st):
        if policy is None:
            # host actions in, host obs / reward / done out through the env's host-buffer entry point: the batch is cut at lock-step round
            # boundaries so that the device->host copy of the first part overlaps the step kernel of the second (env.step_host)
            return env.step_host(st, h_act[i % 4], h_obs, h_rew, h_done)
        # the policy consumes the device-resident observation; its Gaussian noise comes from pinned host memory, the host reads the metrics
        d_ez.copy_(h_ez[i % 4], non_blocking=True)
        d_ea.copy_(h_ea[i % 4], non_blocking=True)
        a = policy.act(st.obs, d_ez, d_ea)[0]
        st = env.step(st, a)
        h_rew.copy_(st.reward, non_blocking=True)
        h_done.copy_(st.done, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller needs the host results before the next action
        return st

    for i in range(W):
        state = e2e_step(i, state)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        state = e2e_step(i, state)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_value = ENVS_PER_GPU * world * K / max_over_ranks(e2e_s, dev, shard)
    h2d = ENVS_PER_GPU * (env.action_size if policy is None else pcfg.latent_size + pcfg.action_size) * 4
    d2h = ENVS_PER_GPU * ((obs_dim if policy is None else 0) + 2) * 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the step kernel (the only kernel of a step), from the same CUDA-event durations
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    fp32_peak = env.stepper.fp32_peak_tflops()   # FMA microbenchmark on this box (no fp32 figure in MEASURED_PEAKS.json)
    kern_rate = ENVS_PER_GPU * K / (kern_ms * 1e-3)   # env-steps/s of the step kernel on rank 0
    per_gpu_rate = kern_rate
    nf, its = env_args["physics_steps_per_control_step"], env_args["iterations"]
    # canonical FLOPs per env-step: per substep 205 kFLOP outside the solver + 34 kFLOP per CG iteration; Newton adds
    # the Hessian assembly + one sparse factorisation per iteration (~2 x 21 kFLOP + 25 kFLOP)
    w_sub = 205e3 + its * (34e3 + (67e3 if env_args["solver"] == "newton" else 0.0))
    W_ALG_WL = nf * w_sub + 25e3
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "step_kernel_dram_bytes.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {
        "bound": "fp32", "achieved": per_gpu_rate * W_ALG_WL / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": per_gpu_rate * W_ALG_WL / 1e12 / fp32_peak if fp32_peak > 0 else None, "traffic": traffic if args.workload == "tracking" else None,
        "peak_source": "measured FP32 FMA microbenchmark (tmjx_fp32_peak_tflops) in this run",
        "kernel": "tmjx_env_kernel<true>", "kernel_ms": kern_ms / K, "algorithmic_flops_per_launch": W_ALG_WL * ENVS_PER_GPU,
        "algorithmic_bytes_per_launch": B_ALG * ENVS_PER_GPU,
        "hbm": {"achieved": per_gpu_rate * B_ALG / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": per_gpu_rate * B_ALG / 1e9 / hbm_peak,
                "peak_source": hbm_src},
    }

    # counters of the same kernel from the committed ncu capture (static evidence, not measured in this run)
    try:
        import glob

        cap = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_kernel_ncu_full.json")))[-1]
        cj = json.load(open(cap))
        pick = {"issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "fma_pipe_active_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                "registers_per_thread": "launch__registers_per_thread", "warp_instructions": "smsp__inst_executed.sum"}
        roofline["ncu_counters"] = {k: float(cj[v][0]) for k, v in pick.items() if v in cj}
        roofline["ncu_counters"]["source"] = os.path.relpath(cap, ROOT)
    except (IndexError, OSError, ValueError, KeyError):
        pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "tracking":
        cpu_baseline = cpu_baseline_leg(walker, clips, config, full=args.full_cpu_baseline)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, args.workload), "clocks": clocks,
        "gpu_launches": K * (1 + (policy.launches_per_act if policy is not None else 0)),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "policy": None if policy is None else {
            "ms_per_step": pol_ms / K, "launches_per_step": policy.launches_per_act, "kernel": ("mlp_chain_kernel: one persistent TMA + tcgen05 kind::tf32 launch per act (csrc/tmjx_chain.cuh)" if policy.launches_per_act == 1
                       else "linear_tf32_tma_kernel (TMA + tcgen05 kind::tf32, TMEM accumulators) + row kernels"),
            "flops_per_env_step": 5.48e6, "achieved_tflops": 5.48e6 * ENVS_PER_GPU * K / (pol_ms * 1e-3) / 1e12},
        "wall_s_timed_region": wall,
        "episode_stats": stats,
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


def run_ppo(args, rank, world, local_rank):
    """BASELINE configs[3]: rodent-mc-intention PPO training, 65536 envs sharded over the GPUs of the job, NCCL gradient all-reduce.
    One "step" = one training step of the reference (ppo.py:320-395): unroll_length env steps on every env with the intention
    network in the loop, the normaliser update and num_updates_per_batch x num_minibatches minibatch updates (forward, loss head,
    backward, all-reduce, Adam).  value = env-steps/s INCLUDING the update; strong scaling (the 65536 envs are split over the ranks)."""
    import torch

    from track_mjx_b200.env import MultiClipTracking, wrap
    from track_mjx_b200.policy import IntentionNetworkConfig
    from track_mjx_b200.ppo import PPO, PPOConfig
    from track_mjx_b200.sharding import Shard, max_over_ranks

    total_envs = args.ppo_envs
    if total_envs % (world * 16):
        raise SystemExit("--ppo-envs must be divisible by 16 x the number of GPUs")
    B = total_envs // world
    shard = Shard(rank, world, total_envs)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    walker, clips, config = build_env_pieces(args.ppo_clips)
    env = wrap(MultiClipTracking(clips, walker, config.RewardConfig(), num_envs=B, device=local_rank, **dict(config.DEFAULT_ENV_ARGS)))
    pcfg = IntentionNetworkConfig(obs_size=env.observation_size, reference_obs_size=env.stepper.dims["reference_obs_size"], action_size=env.action_size)
    ppo = PPO(env, pcfg, PPOConfig())
    ppo.reset(shard.seed(1000))
    K, W = args.steps, max(1, min(args.warmup, 3))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(W):
        ppo.training_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ppo.timing = True
    phases = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        losses = ppo.training_step()
        torch.cuda.synchronize()
        for k, v in ppo.phase_ms().items():
            phases[k] = phases.get(k, 0.0) + v
    e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    max_ms = max_over_ranks(dev_ms, dev, shard)
    ppo.timing = False
    # data-parallel invariant: every rank holds the same parameters after the same number of updates (pmean'd gradients, same seeds)
    params_identical = None
    if world > 1:
        cs = torch.stack([ppo.trainer.params.double().sum(), ppo.trainer.params.double().abs().sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        params_identical = bool(torch.equal(lo, hi))
    # the same steps with the gradient all-reduce switched off: the difference is what the collective costs when it is NOT hidden
    exposed_ms = None
    if world > 1:
        ppo.all_reduce = False
        barrier()
        e0.record()
        for _ in range(K):
            ppo.training_step()
        e1.record()
        barrier()
        exposed_ms = (dev_ms - e0.elapsed_time(e1)) / K
        ppo.all_reduce = True
    env_steps = B * ppo.T * world
    value = env_steps * K / (max_ms * 1e-3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    c = ppo.cfg
    n_mb = c.num_updates_per_batch * c.num_minibatches
    rows = ppo.T * ppo.Bm
    flops_mb = rows * 3 * (5.48e6 + 3.1e6)          # forward + dgrad + wgrad of the policy (5.48 MFLOP / row) and the value net (3.1)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 env / tf32 tensor-core GEMMs with fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"rodent-mc-intention PPO training, {total_envs} envs sharded over {world} B200, NCCL gradient all-reduce (BASELINE configs[3])",
                   "global_envs": total_envs, "envs_per_gpu": B, "n_clips": args.ppo_clips, "unroll_length": ppo.T, "num_minibatches": c.num_minibatches,
                   "num_updates_per_batch": c.num_updates_per_batch, "minibatch_rows_per_gpu": rows,
                   "networks": "intention encoder 470-1024-512x4-(60|60), decoder 286-512x3-256x2-76, critic 696-512x5-256-1 (rodent-full-clips.yaml:50-57)",
                   "parallelism": f"data-parallel x{world}: env shard + minibatch shard per GPU, SUM all-reduce of the flat gradient per minibatch",
                   "l2": "working set (rollout 3.8 GB at 65536 envs) far exceeds L2; no flush"},
        "clocks": clocks,
        "phases_ms_per_step": {k: v / K for k, v in phases.items()},
        "learner": {"minibatch_updates_per_step": n_mb, "ms_per_minibatch": phases.get("sgd", 0.0) / K / n_mb,
                    "gemm_tflops_tf32": flops_mb * n_mb / (phases.get("sgd", 1e-9) / K * 1e-3) / 1e12,
                    "gradient_bytes": int(ppo.trainer.n_params) * 4,
                    "allreduce_exposed_ms_per_step": exposed_ms,
                    "allreduce_share_of_step": None if exposed_ms is None else exposed_ms / (max_ms / K),
                    "params_identical_across_ranks": params_identical},
        "gpu_launches": None, "wall_s_timed_region": wall, "losses_last_minibatch": [float(x) for x in losses.cpu()],
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline_leg(walker, clips, config, full=False):
    """The oracle port on the host cores of this box, bounded to ~10-20 s; `full`: BASELINE configs[0] at its stated length
    (64 envs, random actions, 1000 control steps; about two minutes on 16 cores)."""
    import numpy as np

    import common
    from oracle.oracle import Oracle
    from track_mjx_b200 import _lib as L

    env_args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    cfg = config.make_task_config(walker, config.RewardConfig(), **env_args)
    cores = os.cpu_count() or 1
    orc = Oracle(walker.blob, cfg, clips, dtype=np.float32, nthreads=cores)
    sample = 64 if full else min(ENVS_PER_GPU, max(64, 4 * cores))
    max_steps, max_s = (1000, 1e9) if full else (50, 10.0)
    buf = orc.alloc(sample, debug=False)
    common.put(buf, common.init_buffers(buf, clips, seed=0))
    orc.forward(buf, L.TMJX_F_SNAPSHOT)
    rng = np.random.default_rng(42)
    orc.step(buf, rng.normal(size=(sample, walker.nu)).astype(np.float32), L.TMJX_F_AUTORESET)
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < max_s and steps < max_steps:
        orc.step(buf, rng.normal(size=(sample, walker.nu)).astype(np.float32), L.TMJX_F_AUTORESET)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} of {ENVS_PER_GPU} envs x {steps} control steps, fp32 CPU oracle (port of the MJX algorithm, dense), OpenMP over envs"}


_JSON_FD = None


def emit(obj):
    """The one JSON line goes to the REAL stdout; everything else a library prints there (NCCL's version banner is written
    straight to fd 1 by the C library) has been routed to stderr by main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="tracking", choices=sorted(WORKLOADS) + ["ppo"],
                    help="tracking = BASELINE configs[1] (the bench line); intention = configs[2]; ppo = configs[3]; contact = configs[4]")
    ap.add_argument("--action-scale", type=float, default=1.0, help="std of the random actions (default 1.0 = the N(0,1) law of SURVEY 8d)")
    ap.add_argument("--full-cpu-baseline", action="store_true", help="cpu_baseline = BASELINE configs[0] at its stated length: 64 envs x 1000 steps")
    ap.add_argument("--ppo-envs", type=int, default=65536, help="global number of envs of the ppo workload (split over the GPUs)")
    ap.add_argument("--ppo-clips", type=int, default=842)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "ppo":
        run_ppo(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
