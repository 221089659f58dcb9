"""Optimiser step and KL-weight schedule.

Schedule: `learner.create_ramp_schedule` against values computed by the reference's own function text (tests/golden/
ramp_schedule.npz, tools/make_golden_ramp_schedule.py), 1e-6 relative.  Adam: the restatement of optax's published algorithm
(oracle/optimizer.py, parity unpinned against optax itself) is cross-checked against torch.optim.Adam on the CPU (float64, no
clipping: identical formula), and on a GPU the `tmjx_adam_step` kernels are compared with the restatement over several steps
with and without the global-norm clip triggering: 2e-6 relative on parameters and moments (float32 elementwise arithmetic, same
operation order up to fused multiply-adds; the norm is accumulated in double on both sides), plus 1e-7 of the largest entry on the
first moment, whose two terms can cancel."""
import os

import numpy as np
import pytest

from oracle import optimizer as opt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ramp_schedule.npz")


def test_ramp_schedule_matches_reference_outputs():
    from track_mjx_b200.learner import create_ramp_schedule

    g = np.load(GOLD)
    steps = g["steps"]
    for i in range(int(g["n_cases"])):
        kw = dict(max_value=float(g[f"c{i}_max_value"]), min_value=float(g[f"c{i}_min_value"]), ramp_steps=int(g[f"c{i}_ramp_steps"]),
                  warmup_steps=int(g[f"c{i}_warmup_steps"]), schedule=str(g[f"c{i}_schedule"]), period=int(g[f"c{i}_period"]))
        fn = create_ramp_schedule(**kw)
        got = np.array([fn(s) for s in steps], np.float32)
        assert np.allclose(got, g[f"c{i}_values"], rtol=1e-6, atol=1e-9), (kw, got, g[f"c{i}_values"])
    with pytest.raises(ValueError):
        create_ramp_schedule(schedule="step")


def test_adam_restatement_matches_torch_adam():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=257)
    tp = torch.tensor(p0.copy(), requires_grad=True)
    topt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, mu, nu, count = p0.copy(), np.zeros(257), np.zeros(257), 0
    for _ in range(5):
        g = rng.normal(size=257) * 3
        tp.grad = torch.tensor(g.copy())
        topt.step()
        p, mu, nu, count, _ = opt.adam_step(p, g, mu, nu, count, learning_rate=1e-3, max_grad_norm=0.0, dtype=np.float64)
        assert np.allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-12)
    # clipping: a gradient of norm 50 is rescaled to norm 10 before the moments see it
    g = rng.normal(size=257)
    g *= 50 / np.linalg.norm(g)
    _, mu1, _, _, norm = opt.adam_step(p, g, np.zeros(257), np.zeros(257), 0, max_grad_norm=10.0, dtype=np.float64)
    assert np.isclose(norm, 50.0) and np.isclose(np.linalg.norm(mu1 / 0.1), 10.0)


@pytest.mark.gpu
def test_cuda_adam_matches_restatement():
    torch = pytest.importorskip("torch")
    from track_mjx_b200.learner import Adam

    rng = np.random.default_rng(1)
    for n in (1, 1000, 2_600_003):
        p = rng.normal(size=n).astype(np.float32)
        dev = torch.from_numpy(p.copy()).cuda()
        a = Adam(dev, learning_rate=1e-3, max_grad_norm=10.0)
        mu, nu, count = np.zeros(n, np.float32), np.zeros(n, np.float32), 0
        for k in range(4):
            g = (rng.normal(size=n) * (0.001 if k % 2 else 1.0)).astype(np.float32)    # alternately above / below the clip norm
            if n == 1:
                g = np.array([20.0 if k % 2 == 0 else 0.5], np.float32)
            a.step(torch.from_numpy(g).cuda())
            p, mu, nu, count, norm = opt.adam_step(p, g, mu, nu, count, learning_rate=1e-3, max_grad_norm=10.0)
            torch.cuda.synchronize()
            assert np.isclose(float(a.grad_norm.item()), norm, rtol=1e-6)
            assert np.allclose(a.mu.cpu().numpy(), mu, rtol=2e-6, atol=1e-7 * np.abs(mu).max())   # b1 mu + (1 - b1) g can cancel
            assert np.allclose(a.nu.cpu().numpy(), nu, rtol=2e-6, atol=1e-20)
            assert np.allclose(dev.cpu().numpy(), p, rtol=2e-6, atol=2e-7)
        assert a.count == 4
    # determinism
    g = torch.from_numpy(rng.normal(size=2_600_003).astype(np.float32)).cuda()
    x, y = torch.zeros(2_600_003, device="cuda"), torch.zeros(2_600_003, device="cuda")
    Adam(x).step(g); Adam(y).step(g)
    assert torch.equal(x, y)


def test_shuffle_minibatches_matches_reference_convert_data():
    """`learner.shuffle_minibatches` on time-major `[T, B, ...]` tensors against a numpy transcription of `convert_data`
    (ppo.py:305-310: permute axis 0 of `[B, T, ...]`, reshape to `[num_minibatches, -1, T, ...]`): exact, including nested extras."""
    torch = pytest.importorskip("torch")
    from track_mjx_b200.learner import shuffle_minibatches

    rng = np.random.default_rng(2)
    T, B, nm = 5, 24, 4
    data = {"observation": rng.normal(size=(T, B, 7)).astype(np.float32), "reward": rng.normal(size=(T, B)).astype(np.float32),
            "extras": {"policy_extras": {"raw_action": rng.normal(size=(T, B, 3)).astype(np.float32)}}}
    perm = rng.permutation(B)

    def convert_data(x_bt):                                     # the reference's layout and operations
        x = x_bt[perm]
        return x.reshape((nm, -1) + x.shape[1:])

    to_t = lambda d: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in d.items()}
    got = shuffle_minibatches(to_t(data), torch.from_numpy(perm), nm)
    for leaf, want_src in ((got["observation"], data["observation"]), (got["reward"], data["reward"]),
                           (got["extras"]["policy_extras"]["raw_action"], data["extras"]["policy_extras"]["raw_action"])):
        want = convert_data(np.swapaxes(want_src, 0, 1))         # [nm, B / nm, T, ...]
        assert leaf.shape[:3] == (nm, T, B // nm)
        assert np.array_equal(np.swapaxes(leaf.numpy(), 1, 2), want)
    with pytest.raises(ValueError):
        shuffle_minibatches(to_t(data), torch.from_numpy(perm), 5)
