"""CUDA forward + backward of the intention network and the value network (csrc/tmjx_train.cuh, through the C ABI) against the
float64 layer-by-layer restatement oracle/mlp_grad.py (itself equal to torch autograd to 1e-9, tests/test_mlp_grad.py).

Tolerance: every Dense layer, dgrad and wgrad is a TF32 tensor-core GEMM (10-bit operand mantissas, fp32 accumulation), the class
XLA uses for fp32 matmuls on NVIDIA GPUs; a gradient crosses up to 11 of them plus SiLU' / LayerNorm backward in fp32 with
MUFU-based exponentials.  Per tensor the error is held to 2e-2 of the tensor's largest entry (measured: a few 1e-3)."""
import numpy as np
import pytest
import torch

from oracle import mlp_grad as G
from track_mjx_b200 import policy as P
from track_mjx_b200.learner import Adam, Trainer

pytestmark = pytest.mark.gpu

VALUE_LAYERS = (512, 512, 512, 512, 512, 256)          # critic_layer_sizes of config/rodent-full-clips.yaml:54


def perturbed(cfg, seed):
    """Random-init parameters with non-trivial biases / LayerNorm parameters / normaliser (zero biases would hide bias-gradient bugs)."""
    rng = np.random.default_rng(seed)
    p = P.init_params(cfg, seed)
    for k in p:
        if k.endswith("/bias"):
            p[k] = (0.1 * rng.normal(size=p[k].shape)).astype(np.float32)
        if k.endswith("/scale"):
            p[k] = (1.0 + 0.2 * rng.normal(size=p[k].shape)).astype(np.float32)
    p["norm/mean"] = (0.3 * rng.normal(size=cfg.obs_size)).astype(np.float32)
    p["norm/std"] = (0.5 + rng.uniform(size=cfg.obs_size)).astype(np.float32)
    v = P.init_value_params(cfg.obs_size, VALUE_LAYERS, seed + 1)
    for k in v:
        if k.endswith("/bias"):
            v[k] = (0.1 * rng.normal(size=v[k].shape)).astype(np.float32)
    v["norm/mean"], v["norm/std"] = p["norm/mean"], p["norm/std"]
    return p, v


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def split_flat(cfg, flat, value_layers):
    """Flat [policy | value] vector -> dicts keyed like policy.init_params / init_value_params."""
    out, off = {}, 0

    def take(name, shape):
        nonlocal off
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n

    take("norm/mean", (cfg.obs_size,)); take("norm/std", (cfg.obs_size,))
    k = cfg.reference_obs_size
    for i, n in enumerate(cfg.encoder_layers):
        take(f"encoder/hidden_{i}/kernel", (k, n)); take(f"encoder/hidden_{i}/bias", (n,))
        take(f"encoder/LayerNorm_{i}/scale", (n,)); take(f"encoder/LayerNorm_{i}/bias", (n,))
        k = n
    take("encoder/fc2_mean/kernel", (k, cfg.latent_size)); take("encoder/fc2_mean/bias", (cfg.latent_size,))
    take("encoder/fc2_logvar/kernel", (k, cfg.latent_size)); take("encoder/fc2_logvar/bias", (cfg.latent_size,))
    k = cfg.latent_size + cfg.obs_size - cfg.reference_obs_size
    for i, n in enumerate(cfg.decoder_layers):
        take(f"decoder/hidden_{i}/kernel", (k, n)); take(f"decoder/hidden_{i}/bias", (n,))
        take(f"decoder/LayerNorm_{i}/scale", (n,)); take(f"decoder/LayerNorm_{i}/bias", (n,))
        k = n
    nd = len(cfg.decoder_layers)
    take(f"decoder/hidden_{nd}/kernel", (k, 2 * cfg.action_size)); take(f"decoder/hidden_{nd}/bias", (2 * cfg.action_size,))
    pol, out = out, {}
    take("norm/mean", (cfg.obs_size,)); take("norm/std", (cfg.obs_size,))
    k = cfg.obs_size
    for i, n in enumerate(list(value_layers) + [1]):
        take(f"hidden_{i}/kernel", (k, n)); take(f"hidden_{i}/bias", (n,))
        k = n
    assert off == flat.size
    return pol, out


@pytest.mark.parametrize("rows", [512, 200])
def test_policy_and_value_gradients_match_the_float64_restatement(rows):
    cfg = P.IntentionNetworkConfig()
    p, v = perturbed(cfg, 3)
    tr = Trainer(cfg, p, v, VALUE_LAYERS, max_rows=512)
    rng = np.random.default_rng(7)
    obs = rng.normal(size=(rows, cfg.obs_size)).astype(np.float32)
    eps = rng.normal(size=(rows, cfg.latent_size)).astype(np.float32)
    d_logits = (rng.normal(size=(rows, 2 * cfg.action_size)) / rows).astype(np.float32)
    d_mean = (rng.normal(size=(rows, cfg.latent_size)) / rows).astype(np.float32)
    d_logvar = (rng.normal(size=(rows, cfg.latent_size)) / rows).astype(np.float32)
    d_value = (rng.normal(size=rows) / rows).astype(np.float32)
    cu = lambda a: torch.from_numpy(a).cuda()

    logits, mean, logvar = tr.policy_forward(cu(obs), cu(eps))
    value = tr.value_forward(cu(obs))
    tr.value_backward(cu(d_value))
    tr.policy_backward(cu(d_logits), cu(d_mean), cu(d_logvar))
    torch.cuda.synchronize()

    p64 = {k: a.astype(np.float64) for k, a in p.items()}
    v64 = {k: a.astype(np.float64) for k, a in v.items()}
    r_logits, r_mean, r_logvar, caches = G.intention_fwd(cfg, p64, obs.astype(np.float64), eps.astype(np.float64))
    r_value, vc = G.value_fwd(v64, obs.astype(np.float64), len(VALUE_LAYERS))
    assert rel(logits.cpu().numpy(), r_logits) < 2e-2 and rel(mean.cpu().numpy(), r_mean) < 2e-2 and rel(logvar.cpu().numpy(), r_logvar) < 2e-2
    assert rel(value.cpu().numpy(), r_value) < 2e-2
    gp = G.intention_bwd(cfg, p64, caches, d_logits.astype(np.float64), d_mean.astype(np.float64), d_logvar.astype(np.float64))
    gv = G.value_bwd(v64, vc, d_value.astype(np.float64), len(VALUE_LAYERS))
    got_p, got_v = split_flat(cfg, tr.grads.cpu().numpy(), VALUE_LAYERS)
    worst = {}
    for k, ref in gp.items():
        worst[k] = rel(got_p[k], ref)
    for k, ref in gv.items():
        worst["value/" + k] = rel(got_v[k], ref)
    bad = {k: e for k, e in worst.items() if not e < 2e-2}
    assert not bad, bad
    assert not got_p["norm/mean"].any() and not got_p["norm/std"].any() and not got_v["norm/mean"].any()     # the normaliser is not trained
    # bitwise reproducible: a second pass over the same minibatch gives the same gradient bits
    g1 = tr.grads.clone()
    tr.policy_forward(cu(obs), cu(eps)); tr.value_forward(cu(obs))
    tr.value_backward(cu(d_value)); tr.policy_backward(cu(d_logits), cu(d_mean), cu(d_logvar))
    assert torch.equal(g1, tr.grads)
    tr.close()


def test_optimiser_step_reaches_the_acting_policy():
    """after an optimiser / normaliser update the acting policy must see the new parameters (tmjx_policy_set_params),
    and the training forward must agree with the acting forward on the same parameters."""
    cfg = P.IntentionNetworkConfig()
    p, v = perturbed(cfg, 5)
    rows = 256
    tr = Trainer(cfg, p, v, VALUE_LAYERS, max_rows=rows)
    pol = P.IntentionPolicy(cfg, p, max_env=rows)
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randn(rows, cfg.obs_size, device="cuda", generator=g)
    ez = torch.randn(rows, cfg.latent_size, device="cuda", generator=g)
    ea = torch.randn(rows, cfg.action_size, device="cuda", generator=g)
    _, ex0 = pol.act(obs, ez, ea)
    logits0 = ex0["logits"].clone()
    lt, _, _ = tr.policy_forward(obs, ez)
    # same parameters, same TF32 MMAs; the acting launch applies LayerNorm through folded weights (csrc/tmjx_chain.cuh), so its TF32 operand
    # roundings meet s g instead of LN(s): a few 1e-3 of the largest logit after eleven layers, well inside the 2e-2 budget against fp32
    assert rel(lt.cpu().numpy(), logits0.cpu().numpy()) < 5e-3
    # one optimiser step on a random gradient, then hand the parameters to the actor
    opt = Adam(tr.params, learning_rate=1e-2)
    tr.grads.copy_(torch.randn(tr.n_params, device="cuda", generator=g))
    tr.grads[: 2 * cfg.obs_size].zero_()
    before = tr.params.clone()
    opt.step(tr.grads, all_reduce=False)
    assert not torch.equal(before, tr.params)
    tr.set_normalizer(torch.full((cfg.obs_size,), 0.1, device="cuda"), torch.full((cfg.obs_size,), 2.0, device="cuda"))
    tr.sync()
    pol.set_params(tr.params[: tr.n_policy])
    _, ex1 = pol.act(obs, ez, ea)
    assert rel(ex1["logits"].cpu().numpy(), logits0.cpu().numpy()) > 1e-2          # the actor really changed
    lt1, _, _ = tr.policy_forward(obs, ez)
    assert rel(lt1.cpu().numpy(), ex1["logits"].cpu().numpy()) < 5e-3              # and agrees with the learner's forward (folded-LayerNorm rounding, see above)
    # a policy created from scratch with the updated parameters gives the same bits as the refreshed one
    pol_p, _ = split_flat(cfg, tr.params.cpu().numpy(), VALUE_LAYERS)
    fresh = P.IntentionPolicy(cfg, pol_p, max_env=rows)
    _, ex2 = fresh.act(obs, ez, ea)
    assert torch.equal(ex2["logits"], ex1["logits"])
    tr.close(); pol.close(); fresh.close()


def test_mn_major_wgrad_equals_the_transposing_form(monkeypatch):
    """The wgrad GEMMs read x and dH through MN-major tcgen05 operands (32-byte-atom swizzle, transpose bits in the instruction
    descriptor; csrc/tmjx_policy.cu `make_desc_sw128_mn`) instead of transposing both into K-major copies first.  Same MMAs over the same
    K order and the same split-K plane order: the parameter gradients must agree with the transposing form (TMJX_WGRAD_MN=0) to rounding,
    and the two per-operand bisecting modes as well."""
    cfg = P.IntentionNetworkConfig()
    p, v = perturbed(cfg, 9)
    rows = 384
    rng = np.random.default_rng(11)
    cu = lambda a: torch.from_numpy(a).cuda()
    obs, eps = rng.normal(size=(rows, cfg.obs_size)).astype(np.float32), rng.normal(size=(rows, cfg.latent_size)).astype(np.float32)
    d_logits = (rng.normal(size=(rows, 2 * cfg.action_size)) / rows).astype(np.float32)
    d_mean = (rng.normal(size=(rows, cfg.latent_size)) / rows).astype(np.float32)
    d_logvar = (rng.normal(size=(rows, cfg.latent_size)) / rows).astype(np.float32)
    d_value = (rng.normal(size=rows) / rows).astype(np.float32)
    grads = {}
    for mode in ("0", "1", "2", "3"):
        monkeypatch.setenv("TMJX_WGRAD_MN", mode)
        tr = Trainer(cfg, p, v, VALUE_LAYERS, max_rows=rows)
        tr.policy_forward(cu(obs), cu(eps)); tr.value_forward(cu(obs))
        tr.value_backward(cu(d_value)); tr.policy_backward(cu(d_logits), cu(d_mean), cu(d_logvar))
        torch.cuda.synchronize()
        grads[mode] = tr.grads.clone()
        tr.close()
    ref = grads["0"]
    assert ref.abs().max().item() > 1e-3
    for mode in ("1", "2", "3"):
        assert (grads[mode] - ref).abs().max().item() <= 1e-6 * ref.abs().max().item(), mode
