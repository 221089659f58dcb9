"""GPU parity of the tcgen05 intention-network inference (csrc/tmjx_policy.cu) against a plain PyTorch fp32 restatement of
the reference network (intention_network.py:14-142, ppo_networks.py:34-100).

Tolerances: the Dense layers run as TF32 tensor-core MMAs with fp32 accumulation (the arithmetic class XLA uses for fp32
matmuls on NVIDIA GPUs by default); the hardware truncates the fp32 operands to 10 mantissa bits, i.e. a relative error
of <= 2^-10 per product and ~1e-3 after a 512..1024-long reduction of O(1) terms, renormalised by every LayerNorm.
A single layer is checked to 4e-3 (abs, on O(1) outputs), the 11-layer stack to 2e-2 on logits / latents.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def torch_reference(cfg, p, obs, eps_z, eps_a, deterministic=False):
    """fp32 restatement of IntentionNetwork.__call__ + NormalTanhDistribution (allow_tf32 off)."""
    F = torch.nn.functional
    t = lambda k: torch.from_numpy(p[k]).to(obs.device)
    x = (obs - t("norm/mean")) / t("norm/std")
    h = x[:, : cfg.reference_obs_size]
    for i, n in enumerate(cfg.encoder_layers):
        h = h @ t(f"encoder/hidden_{i}/kernel") + t(f"encoder/hidden_{i}/bias")
        h = F.silu(h)
        h = F.layer_norm(h, (n,), t(f"encoder/LayerNorm_{i}/scale"), t(f"encoder/LayerNorm_{i}/bias"), eps=1e-6)
    mean = h @ t("encoder/fc2_mean/kernel") + t("encoder/fc2_mean/bias")
    logvar = h @ t("encoder/fc2_logvar/kernel") + t("encoder/fc2_logvar/bias")
    z = mean if deterministic else mean + eps_z * torch.exp(0.5 * logvar)
    h = torch.cat([z, x[:, cfg.reference_obs_size:]], dim=-1)
    nd = len(cfg.decoder_layers)
    for i, n in enumerate(cfg.decoder_layers):
        h = h @ t(f"decoder/hidden_{i}/kernel") + t(f"decoder/hidden_{i}/bias")
        h = F.silu(h)
        h = F.layer_norm(h, (n,), t(f"decoder/LayerNorm_{i}/scale"), t(f"decoder/LayerNorm_{i}/bias"), eps=1e-6)
    logits = h @ t(f"decoder/hidden_{nd}/kernel") + t(f"decoder/hidden_{nd}/bias")
    loc, raw_scale = logits[:, : cfg.action_size], logits[:, cfg.action_size:]
    scale = F.softplus(raw_scale) + 0.001
    raw = loc if deterministic else loc + scale * eps_a
    logp = (-0.5 * ((raw - loc) / scale) ** 2 - torch.log(scale) - 0.5 * np.log(2 * np.pi)
            - 2.0 * (np.log(2.0) - raw - F.softplus(-2.0 * raw))).sum(-1)
    return torch.tanh(raw), dict(raw_action=raw, log_prob=logp, logits=logits, latent_mean=mean, latent_logvar=logvar)


@pytest.fixture(scope="module")
def setup():
    from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, init_params

    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = IntentionNetworkConfig()
    p = init_params(cfg, seed=0)
    rng = np.random.default_rng(1)
    # non-trivial normaliser, biases and LayerNorm affine so that every parameter is exercised
    p["norm/mean"] = rng.normal(0, 0.1, cfg.obs_size).astype(np.float32)
    p["norm/std"] = rng.uniform(0.5, 2.0, cfg.obs_size).astype(np.float32)
    for k in list(p):
        if k.endswith("/bias"):
            p[k] = rng.normal(0, 0.05, p[k].shape).astype(np.float32)
        if k.endswith("/scale"):
            p[k] = rng.uniform(0.8, 1.2, p[k].shape).astype(np.float32)
    pol = IntentionPolicy(cfg, p, max_env=1000)
    yield cfg, p, pol
    pol.close()


@pytest.mark.parametrize("which,n", [(0, 1000), (1, 128), (5, 77), (6, 300), (11, 129)])
def test_single_layer_matches_fp32(setup, which, n):
    cfg, p, pol = setup
    dev = pol.device
    enc = list(cfg.encoder_layers)
    dec = list(cfg.decoder_layers)
    ins = [cfg.reference_obs_size] + enc[:-1] + [enc[-1]] + [cfg.latent_size + cfg.obs_size - cfg.reference_obs_size] + dec
    outs = enc + [2 * cfg.latent_size] + dec + [2 * cfg.action_size]
    k, no = ins[which], outs[which]
    kpad, npad = (k + 31) // 32 * 32, (no + 127) // 128 * 128
    g = torch.Generator(device="cpu").manual_seed(which)
    x = torch.zeros(n, kpad, device=dev)
    x[:, :k] = torch.randn(n, k, generator=g).to(dev)
    y = torch.full((n, npad), float("nan"), device=dev)
    pol.linear(which, x, y)
    torch.cuda.synchronize()
    t = lambda key: torch.from_numpy(p[key]).to(dev)
    ne = len(enc)
    if which < ne:
        name, ln = f"encoder/hidden_{which}", f"encoder/LayerNorm_{which}"
    elif which == ne:
        name, ln = None, None
    else:
        i = which - ne - 1
        name, ln = f"decoder/hidden_{i}", (f"decoder/LayerNorm_{i}" if i < len(dec) else None)
    if name is None:
        W = torch.cat([t("encoder/fc2_mean/kernel"), t("encoder/fc2_logvar/kernel")], 1)
        b = torch.cat([t("encoder/fc2_mean/bias"), t("encoder/fc2_logvar/bias")])
    else:
        W, b = t(name + "/kernel"), t(name + "/bias")
    ref = x[:, :k] @ W + b
    if ln is not None:
        ref = torch.nn.functional.layer_norm(torch.nn.functional.silu(ref), (no,), t(ln + "/scale"), t(ln + "/bias"), eps=1e-6)
    err = (y[:, :no] - ref).abs().max().item()
    assert torch.isfinite(y[:, :no]).all()
    assert err < 4e-3, err


@pytest.mark.parametrize("deterministic", [False, True])
def test_act_matches_fp32_reference(setup, deterministic):
    cfg, p, pol = setup
    n = 777
    g = torch.Generator(device="cpu").manual_seed(5)
    obs = torch.randn(n, cfg.obs_size, generator=g).to(pol.device)
    ez = torch.randn(n, cfg.latent_size, generator=g).to(pol.device)
    ea = torch.randn(n, cfg.action_size, generator=g).to(pol.device)
    act, ex = pol.act(obs, ez, ea, deterministic=deterministic)
    torch.cuda.synchronize()
    ract, rex = torch_reference(cfg, p, obs, ez, ea, deterministic)
    assert (ex["latent_mean"] - rex["latent_mean"]).abs().max().item() < 2e-2
    assert (ex["latent_logvar"] - rex["latent_logvar"]).abs().max().item() < 2e-2
    assert (ex["logits"] - rex["logits"]).abs().max().item() < 2e-2
    assert (act - ract).abs().max().item() < 2e-2
    # log_prob evaluated at the kernel's own sample must agree with the closed form on the kernel's logits (fp32 path)
    loc, rs = ex["logits"][:, : cfg.action_size], ex["logits"][:, cfg.action_size:]
    scale = torch.nn.functional.softplus(rs) + 0.001
    raw = ex["raw_action"]
    lp = (-0.5 * ((raw - loc) / scale) ** 2 - torch.log(scale) - 0.5 * np.log(2 * np.pi)
          - 2.0 * (np.log(2.0) - raw - torch.nn.functional.softplus(-2.0 * raw))).sum(-1)
    assert (ex["log_prob"] - lp).abs().max().item() < 1e-3 * max(1.0, lp.abs().max().item())
    assert (act.abs() <= 1).all()


def test_zero_copy_rollout_matches_step_by_step(walker, clips2):
    """Rollout.generate (obs written by the step kernel straight into the [T+1, B, obs] buffer, policy outputs into slot t)
    equals acting step by step through env.step / policy.act on the same noise -- bit for bit (deterministic kernels)."""
    from track_mjx_b200 import config
    from track_mjx_b200.env import MultiClipTracking, wrap
    from track_mjx_b200.policy import IntentionNetworkConfig, IntentionPolicy, init_params
    from track_mjx_b200.rollout import Rollout

    n, T = 96, 5
    pcfg = IntentionNetworkConfig()
    params = init_params(pcfg, seed=3)

    def make():
        env = wrap(MultiClipTracking(clips2, walker, config.RewardConfig(), num_envs=n, device=0, **config.DEFAULT_ENV_ARGS))
        return env, env.reset(7), IntentionPolicy(pcfg, params, max_env=n)

    g = torch.Generator(device="cpu").manual_seed(11)
    ez = torch.randn(T, n, pcfg.latent_size, generator=g).cuda()
    ea = torch.randn(T, n, pcfg.action_size, generator=g).cuda()

    env, st, pol = make()
    ro = Rollout(env, pol, T)
    st_end, tr = ro.generate(st, eps=(ez, ea))
    torch.cuda.synchronize()

    env2, st2, pol2 = make()
    obs_seq, rew_seq, done_seq, act_seq = [st2.obs.clone()], [], [], []
    for t in range(T):
        a, _ = pol2.act(st2.obs, ez[t], ea[t])
        act_seq.append(a.clone())
        st2 = env2.step(st2, a)
        obs_seq.append(st2.obs.clone()); rew_seq.append(st2.reward.clone()); done_seq.append(st2.done.clone())
    torch.cuda.synchronize()
    assert torch.equal(tr.observation, torch.stack(obs_seq[:-1])) and torch.equal(tr.next_observation, torch.stack(obs_seq[1:]))
    assert torch.equal(tr.action, torch.stack(act_seq))
    assert torch.equal(tr.reward, torch.stack(rew_seq)) and torch.equal(1.0 - tr.discount, torch.stack(done_seq))
    assert torch.equal(st_end.obs, st2.obs) and torch.equal(st_end.pipeline_state.qpos, st2.pipeline_state.qpos)
    assert tr.extras["state_extras"]["truncation"].shape == (T, n)
    assert torch.isfinite(tr.extras["policy_extras"]["log_prob"]).all()


@pytest.mark.parametrize("variant", ["1", "2"])
def test_older_gemm_kernels_agree_with_the_tma_kernel(setup, variant, monkeypatch):
    """TMJX_POLICY_V1=1 (128 x 128 block-synchronous cp.async kernel) and =2 (warp-specialised cp.async producers) stay in the
    library as A/B references for the TMA kernel.  Per layer they agree to 1e-6 (the TMA epilogue uses a 2-ulp SiLU); over the
    11-layer stack a 1e-6 difference in an activation crosses TF32 truncation boundaries of the next layer's operands and the
    two runs decorrelate down to the TF32 noise floor of this network (~3e-3 on the logits, measured), which is what the
    tolerance allows; the fp32-reference tests above bound the absolute error of each."""
    from track_mjx_b200.policy import IntentionPolicy

    cfg, p, pol = setup
    n = 300
    g = torch.Generator(device="cpu").manual_seed(9)
    obs = torch.randn(n, cfg.obs_size, generator=g).to(pol.device)
    a0, e0 = pol.act(obs, deterministic=True)
    a0, l0 = a0.clone(), e0["logits"].clone()
    monkeypatch.setenv("TMJX_POLICY_V1", variant)
    alt = IntentionPolicy(cfg, p, max_env=n)
    a1, e1 = alt.act(obs, deterministic=True)
    torch.cuda.synchronize()
    assert (e1["logits"] - l0).abs().max().item() < 1e-2
    assert (a1 - a0).abs().max().item() < 1e-2
    # one layer: same operands, same products
    x = torch.zeros(n, 512, device=pol.device)
    x[:, :512] = torch.randn(n, 512, generator=g).to(pol.device)
    y0, y1 = torch.zeros(n, 512, device=pol.device), torch.zeros(n, 512, device=pol.device)
    pol.linear(2, x, y0)
    alt.linear(2, x, y1)
    torch.cuda.synchronize()
    assert (y1 - y0).abs().max().item() < 1e-5
    alt.close()


@pytest.mark.gpu
def test_value_network_matches_fp32_reference_and_feeds_the_loss_head():
    """`ValueNetwork.apply` (normalise, Dense + swish x 2, Dense to 1 on the TF32 tensor-core GEMM) against a float64 torch
    evaluation of the same MLP: 1e-2 absolute on O(1) values (TF32 operand truncation over K = 696 / 1024 / 1024, the bound of the
    single-layer test above accumulated over three layers).  Then the learner chain on a small rollout: baseline / bootstrap from the
    value network -> `ppo_loss_head` -> finite loss terms and gradient seeds of the right shapes."""
    import torch

    from track_mjx_b200.learner import ppo_loss_head
    from track_mjx_b200.policy import ValueNetwork, init_value_params

    rng = np.random.default_rng(3)
    obs_size, T, B, A, Lz = 696, 5, 300, 38, 60
    params = init_value_params(obs_size, (1024, 1024), seed=1)
    params["norm/mean"] = rng.normal(0, 0.5, obs_size).astype(np.float32)
    params["norm/std"] = rng.uniform(0.5, 2.0, obs_size).astype(np.float32)
    for i in range(3):
        params[f"hidden_{i}/bias"] = rng.normal(0, 0.1, params[f"hidden_{i}/bias"].shape).astype(np.float32)
    net = ValueNetwork(obs_size, params, max_env=1024)           # 1500 rows -> two chunks
    obs = torch.from_numpy((rng.normal(size=(T, B, obs_size)) * 1.5 + 0.3).astype(np.float32)).cuda()
    got = net.apply(obs)
    torch.cuda.synchronize()
    assert got.shape == (T, B)
    x = (obs.double().cpu() - torch.from_numpy(params["norm/mean"]).double()) / torch.from_numpy(params["norm/std"]).double()
    for i in range(3):
        x = x @ torch.from_numpy(params[f"hidden_{i}/kernel"]).double() + torch.from_numpy(params[f"hidden_{i}/bias"]).double()
        if i < 2:
            x = torch.nn.functional.silu(x)
    want = x.squeeze(-1)
    err = (got.double().cpu() - want).abs().max().item()
    assert want.abs().max().item() > 0.3 and err < 1e-2, (err, want.abs().max().item())
    boot = net.apply(obs[-1])
    assert torch.equal(boot, got[-1])                            # same rows, same kernel: bitwise
    g = lambda *s: torch.from_numpy(rng.normal(size=s).astype(np.float32)).cuda()
    out = ppo_loss_head(g(T, B, 2 * A) * 0.5, g(T, B, Lz), g(T, B, Lz) * 0.5 - 1, got, boot, g(T, B), torch.ones(T, B, device="cuda"),
                        torch.zeros(T, B, device="cuda"), g(T, B, A), g(T, B) - 40, g(T, B, A))
    torch.cuda.synchronize()
    assert all(np.isfinite(float(out[k])) for k in ("total_loss", "policy_loss", "v_loss", "kl_latent_loss", "entropy_loss"))
    assert out["d_baseline"].shape == (T, B) and out["d_logits"].shape == (T, B, 2 * A) and torch.isfinite(out["d_logits"]).all()
    net.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 130, 777, 1000])
def test_fused_chain_kernel_matches_the_per_layer_launches(setup, n, monkeypatch):
    """The one-launch policy (`mlp_chain_kernel`, csrc/tmjx_chain.cuh: TMEM-resident SiLU + LayerNorm epilogue, latent and action rows
    in the same launch) against the per-layer launches it replaces (TMJX_POLICY_FUSED=0).  Both run the same tcgen05 TF32 MMAs over the
    same K order, so the pre-activations are bit-identical; the LayerNorm sums are taken in a different order (thread-sequential + four
    column groups instead of a warp shuffle tree): 1e-6 relative per layer, which flips TF32 operand roundings (2^-11 relative) of the
    next layer here and there: a few 1e-3 absolute on O(1) logits after eleven layers (raw actions add scale x eps with |eps| up to 4), inside the budget
    against fp32 (2e-2).  Row counts cover a single partial CTA, a ragged tail and max_env."""
    from track_mjx_b200.policy import IntentionPolicy

    cfg, p, pol = setup
    assert pol.launches_per_act == 1
    monkeypatch.setenv("TMJX_POLICY_FUSED", "0")
    ref = IntentionPolicy(cfg, p, max_env=1000)
    assert ref.launches_per_act == 25
    g = torch.Generator(device="cpu").manual_seed(11 + n)
    obs = torch.randn(n, cfg.obs_size, generator=g).to(pol.device)
    ez = torch.randn(n, cfg.latent_size, generator=g).to(pol.device)
    ea = torch.randn(n, cfg.action_size, generator=g).to(pol.device)
    for det in (False, True):
        a0, e0 = ref.act(obs, ez, ea, deterministic=det)
        a1, e1 = pol.act(obs, ez, ea, deterministic=det)
        torch.cuda.synchronize()
        for k in ("latent_mean", "latent_logvar", "logits", "raw_action"):
            assert torch.isfinite(e1[k]).all()
            assert (e1[k] - e0[k]).abs().max().item() < (2e-2 if k == "raw_action" else 1e-2), (k, det, (e1[k] - e0[k]).abs().max().item())
        assert (a1 - a0).abs().max().item() < 1e-2
        assert (e1["log_prob"] - e0["log_prob"]).abs().max().item() < 2e-2 * max(1.0, e0["log_prob"].abs().max().item())
    # the fused launch is deterministic and independent of the row's position in the batch
    a_fwd = pol.act(obs, ez, ea)[0].clone()           # (act returns views into the policy's own buffers)
    a_rev = pol.act(obs.flip(0).contiguous(), ez.flip(0).contiguous(), ea.flip(0).contiguous())[0].clone()
    torch.cuda.synchronize()
    assert torch.equal(a_fwd.flip(0), a_rev)
    ref.close()


@pytest.mark.gpu
def test_fused_value_chain_matches_the_per_layer_launches(monkeypatch):
    from track_mjx_b200.policy import ValueNetwork, init_value_params

    rng = np.random.default_rng(5)
    obs_size, hidden = 696, (512, 512, 512, 512, 512, 256)          # the shipped critic (rodent-full-clips.yaml:54)
    params = init_value_params(obs_size, hidden, seed=2)
    params["norm/mean"] = rng.normal(0, 0.5, obs_size).astype(np.float32)
    params["norm/std"] = rng.uniform(0.5, 2.0, obs_size).astype(np.float32)
    for i in range(len(hidden) + 1):
        params[f"hidden_{i}/bias"] = rng.normal(0, 0.1, params[f"hidden_{i}/bias"].shape).astype(np.float32)
    obs = torch.from_numpy((rng.normal(size=(1111, obs_size)) * 1.5 + 0.3).astype(np.float32)).cuda()
    fused = ValueNetwork(obs_size, params, max_env=2048, hidden_layers=hidden)
    monkeypatch.setenv("TMJX_POLICY_FUSED", "0")
    ref = ValueNetwork(obs_size, params, max_env=2048, hidden_layers=hidden)
    a, b = fused.apply(obs).clone(), ref.apply(obs).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and b.abs().max().item() > 0.05
    assert (a - b).abs().max().item() < 1e-5 * max(1.0, b.abs().max().item())     # no LayerNorm: same MMAs, same SiLU -> (almost) bitwise
    fused.close(); ref.close()
