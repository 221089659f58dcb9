"""GAE: the CPU restatement (oracle/gae.py) against golden vectors computed by the reference's own `compute_gae`
(losses.py:39-101, executed by tools/make_golden_gae.py), and -- on a GPU -- the CUDA kernel `tmjx_gae` against the restatement.
Float32, same operation order, no fused multiply-add on either side: bit-exact."""
import os

import numpy as np
import pytest

from oracle.gae import compute_gae

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gae.npz")


def cases():
    g = np.load(GOLD)
    n = len({k.split("_")[0] for k in g.files})
    for i in range(n):
        yield {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}


def test_restatement_matches_reference_outputs_bit_exactly():
    seen = 0
    for c in cases():
        vs, adv = compute_gae(c["truncation"], c["termination"], c["rewards"], c["values"], c["bootstrap"], c["lambda"], c["discount"])
        assert np.array_equal(vs, c["vs"]) and np.array_equal(adv, c["advantages"])
        seen += 1
    assert seen == 4


def test_gae_properties():
    """Size-independent checks: lambda = 0 gives the one-step TD target; a truncated step has zero advantage and vs = V."""
    rng = np.random.default_rng(0)
    T, B = 12, 40
    r, v = rng.normal(size=(T, B)).astype(np.float32), rng.normal(size=(T, B)).astype(np.float32)
    b = rng.normal(size=B).astype(np.float32)
    z = np.zeros((T, B), np.float32)
    vs, adv = compute_gae(z, z, r, v, b, 0.0, 0.9)
    v1 = np.concatenate([v[1:], b[None]])
    assert np.allclose(vs, r + np.float32(0.9) * v1, atol=1e-6)
    tr = z.copy()
    tr[5] = 1.0
    vs2, adv2 = compute_gae(tr, z, r, v, b, 0.95, 0.99)
    assert np.array_equal(adv2[5], np.zeros(B, np.float32)) and np.array_equal(vs2[5], v[5])


@pytest.mark.gpu
@pytest.mark.parametrize("T,B", [(20, 4096), (1, 7), (33, 1000)])
def test_cuda_gae_matches_restatement_bit_exactly(T, B):
    torch = pytest.importorskip("torch")
    from track_mjx_b200.learner import compute_gae as gpu_gae

    rng = np.random.default_rng(T * 1000 + B)
    f = np.float32
    te = (rng.random((T, B)) < 0.1).astype(f)
    tr = ((rng.random((T, B)) < 0.05) & (te == 0)).astype(f)
    r, v, b = rng.normal(size=(T, B)).astype(f), rng.normal(2, 3, (T, B)).astype(f), rng.normal(2, 3, B).astype(f)
    vs, adv = compute_gae(tr, te, r, v, b, 0.95, 0.99)
    d = lambda a: torch.from_numpy(a).cuda()
    gvs, gadv = gpu_gae(d(tr), d(te), d(r), d(v), d(b), 0.95, 0.99)
    torch.cuda.synchronize()
    assert np.array_equal(gvs.cpu().numpy(), vs) and np.array_equal(gadv.cpu().numpy(), adv)
    for c in cases():
        gvs, gadv = gpu_gae(d(c["truncation"]), d(c["termination"]), d(c["rewards"]), d(c["values"]), d(c["bootstrap"]),
                            float(c["lambda"]), float(c["discount"]))
        assert np.array_equal(gvs.cpu().numpy(), c["vs"]) and np.array_equal(gadv.cpu().numpy(), c["advantages"])
