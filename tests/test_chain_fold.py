"""The algebra behind the fused policy launch's LayerNorm handling (csrc/tmjx_chain.cuh, `fold_ln_kernel` in csrc/tmjx_policy.cu), on the CPU.

A hidden layer of the intention network is Dense -> SiLU -> LayerNorm (reference intention_network.py:34-47; flax LayerNorm with fast variance,
epsilon 1e-6).  The fused launch never materialises the LayerNorm output: it stores s = SiLU(h), keeps the row's (mean, rstd), and lets the
consumer layer apply the normalisation through its operands,
    LN(s) W + b = rstd * (s @ (g[:, None] * W)) - (rstd * mean) * (g @ W) + (beta @ W + b).
This test pins that identity in float64 (exact up to rounding) and bounds the float32 difference between the two evaluation orders."""
import numpy as np


def layer_norm(s, g, beta, eps=1e-6):
    mean = s.mean(-1, keepdims=True)
    var = np.maximum(0.0, (s * s).mean(-1, keepdims=True) - mean * mean)          # use_fast_variance
    return (s - mean) / np.sqrt(var + eps) * g + beta


def folded(s, g, beta, W, b, eps=1e-6):
    mean = s.mean(-1, keepdims=True)
    var = np.maximum(0.0, (s * s).mean(-1, keepdims=True) - mean * mean)
    rstd = 1.0 / np.sqrt(var + eps)
    Wf, cvec, bf = g[:, None] * W, g @ W, beta @ W + b                            # what fold_ln_kernel precomputes per parameter update
    return rstd * (s @ Wf) - (rstd * mean) * cvec + bf


def test_folded_layernorm_is_the_same_function():
    rng = np.random.default_rng(0)
    for k, n in ((512, 512), (1024, 512), (256, 76), (512, 120)):
        h = rng.normal(size=(64, k)) * 1.5 + 0.2
        s = h / (1.0 + np.exp(-h))                                                # SiLU outputs: mean > 0, the unfavourable case for cancellation
        g, beta = rng.uniform(0.5, 1.5, k), rng.normal(0, 0.2, k)
        W, b = rng.normal(size=(k, n)) / np.sqrt(k), rng.normal(0, 0.1, n)
        want = layer_norm(s, g, beta) @ W + b
        got = folded(s, g, beta, W, b)
        assert np.abs(got - want).max() < 1e-11 * max(1.0, np.abs(want).max())
        # float32 evaluation of both orders: the folded form loses at most a few ulps more (mild cancellation between rstd (s W') and
        # rstd mean cvec: |mean| < std for SiLU outputs), far below the TF32 operand rounding (2^-11) both paths carry on the GPU
        f = lambda a: a.astype(np.float32)
        want32 = layer_norm(f(s), f(g), f(beta)) @ f(W) + f(b)
        got32 = folded(f(s), f(g), f(beta), f(W), f(b))
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got32 - want).max() < 2e-5 * scale and np.abs(want32 - want).max() < 2e-5 * scale


def test_padding_columns_do_not_disturb_the_statistics():
    """Columns n .. npad - 1 of a padded layer are SiLU(0 + 0) = 0: they add nothing to the row sums the kernel accumulates, and the
    statistics divide by the true width n."""
    rng = np.random.default_rng(1)
    s = rng.normal(size=(8, 120))
    sp = np.concatenate([s, np.zeros((8, 8))], -1)
    assert np.allclose(sp.sum(-1) / 120, s.mean(-1)) and np.allclose((sp * sp).sum(-1) / 120, (s * s).mean(-1))
