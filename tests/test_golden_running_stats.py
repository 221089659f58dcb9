"""Observation-normaliser statistics: the CPU restatement (oracle/running_stats.py) against golden vectors computed by the
reference's own `running_statistics.update` / `normalize` (masked_running_statistics.py, executed by
tools/make_golden_running_stats.py), and -- on a GPU -- the CUDA kernels behind `learner.RunningStatistics` against it.

Tolerances.  Counts are exact.  The kernel reads the batch once and evaluates sum((x-m)(x-m')) as M2 + n (xbar-m)(xbar-m')
(identical in exact arithmetic; the second term is evaluated as a sum of squares, see stats_mean_kernel in tmjx_policy.cu); the reference's float32 evaluation of the left-hand side is itself up to
2.4e-5 (relative, std) away from the float64 value on the first update, where the running mean starts at 0 far from the data
(checked below), so the CUDA result is compared (a) with the reference's float32 golden output at 1e-4 relative and (b) with
the float64 evaluation of the same statistics at 2e-6 relative, which the reference's own output does not meet."""
import os

import numpy as np
import pytest

from oracle import running_stats as rs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "running_stats.npz")


def test_restatement_matches_reference_outputs():
    g = np.load(GOLD)
    for i in range(3):
        c, m, sv, std = rs.update(g[f"u{i}_count0"], g[f"u{i}_mean0"], g[f"u{i}_sv0"], g[f"u{i}_batch"])
        assert c == g[f"u{i}_count1"]
        assert np.allclose(m, g[f"u{i}_mean1"], rtol=1e-6, atol=1e-7)
        assert np.allclose(sv, g[f"u{i}_sv1"], rtol=1e-5, atol=1e-6)
        assert np.allclose(std, g[f"u{i}_std1"], rtol=1e-5, atol=1e-9)
    x0 = g["u0_batch"].astype(np.float64)              # the reference's own float32 error on the first update (see the header)
    err = np.delete(np.abs(g["u0_std1"] - x0.std(axis=0)) / np.maximum(x0.std(axis=0), 1e-3), 10)   # 10: the constant feature
    assert 1e-5 < err.max() < 1e-4
    assert g["u2_std1"][10] == np.float32(1e-6)          # the constant feature sits on std_min_value
    assert np.allclose(rs.normalize(g["norm_in"], g["u2_mean1"], g["u2_std1"]), g["norm_out"], rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_cuda_running_statistics_match_restatement_and_golden():
    torch = pytest.importorskip("torch")
    from track_mjx_b200.learner import RunningStatistics

    g = np.load(GOLD)
    st = RunningStatistics(696)
    for i in range(3):
        st.update(torch.from_numpy(g[f"u{i}_batch"]).cuda())
        torch.cuda.synchronize()
        assert float(st.count.item()) == float(g[f"u{i}_count1"])
        assert np.allclose(st.mean.cpu().numpy(), g[f"u{i}_mean1"], rtol=1e-5, atol=1e-6)
        assert np.allclose(st.summed_variance.cpu().numpy(), g[f"u{i}_sv1"], rtol=2e-4, atol=1e-4)
        assert np.allclose(st.std.cpu().numpy(), g[f"u{i}_std1"], rtol=1e-4, atol=1e-9)
        seen = np.concatenate([g[f"u{j}_batch"] for j in range(i + 1)]).astype(np.float64)
        std64 = np.clip(seen.std(axis=0), 1e-6, 1e6)
        assert np.allclose(st.mean.cpu().numpy(), seen.mean(axis=0), rtol=2e-6, atol=2e-6)
        assert np.allclose(st.std.cpu().numpy(), std64, rtol=2e-6, atol=1e-9)
    # full size: 20 x 16384 observations (one PPO training step of BASELINE configs[2]) in two halves == in one piece
    rng = np.random.default_rng(0)
    x = torch.from_numpy((rng.normal(size=(40000, 696)) * 2 + 1).astype(np.float32)).cuda()
    a, b = RunningStatistics(696), RunningStatistics(696)
    a.update(x)
    b.update(x[:15000]); b.update(x[15000:])
    torch.cuda.synchronize()
    assert float(a.count.item()) == float(b.count.item()) == 40000.0
    assert torch.allclose(a.mean, b.mean, rtol=1e-5, atol=1e-6) and torch.allclose(a.std, b.std, rtol=1e-5, atol=1e-7)
    # against the float64 statistics; the float32 restatement is looser here because numpy adds the 40000 rows of an axis-0
    # sum one after the other (relative error ~ eps sqrt(N)), the kernel merges 192 slab moments
    x64 = x.cpu().numpy().astype(np.float64)
    assert np.allclose(a.mean.cpu().numpy(), x64.mean(axis=0), rtol=2e-6, atol=2e-6)
    assert np.allclose(a.std.cpu().numpy(), x64.std(axis=0), rtol=5e-6)
    c, m, sv, std = rs.update(0.0, np.zeros(696, np.float32), np.zeros(696, np.float32), x.cpu().numpy())
    assert np.allclose(a.mean.cpu().numpy(), m, rtol=1e-4, atol=1e-5) and np.allclose(a.std.cpu().numpy(), std, rtol=1e-4)
    # D not a multiple of 4: the scalar-load kernel; ragged row count; second update moves the mean
    y = torch.from_numpy((rng.normal(size=(3001, 13)) * 0.3 - 4).astype(np.float32)).cuda()
    s13 = RunningStatistics(13)
    s13.update(y[:1000]); s13.update(y[1000:] + 2.0)
    y64 = np.concatenate([y[:1000].cpu().numpy().astype(np.float64), (y[1000:] + 2.0).cpu().numpy().astype(np.float64)])
    assert float(s13.count.item()) == 3001.0
    assert np.allclose(s13.mean.cpu().numpy(), y64.mean(axis=0), rtol=2e-6, atol=2e-6)
    assert np.allclose(s13.std.cpu().numpy(), y64.std(axis=0), rtol=5e-6)
    # determinism: fixed-order combination of the per-block partial sums
    a2 = RunningStatistics(696)
    a2.update(x)
    assert torch.equal(a2.mean, a.mean) and torch.equal(a2.summed_variance, a.summed_variance)
