"""CPU validation of the generated (tree-specialised) solve schedule: the IR that tools/gen_tree_kernels.py prints as
CUDA is executed by its numpy interpreter and compared with a dense solve, and the committed header is checked to be
in sync with the walker's dof tree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tree_kernels as G  # noqa: E402


def _check(tree, seed):
    rng = np.random.default_rng(seed)
    M = G.random_tree_spd(tree, rng)
    L = G.factor_ref(tree, G.sparse_from_dense(tree, M))
    b = rng.normal(size=tree.nv)
    x = G.run_ir(G.build_solve_ir(tree), tree, L.astype(np.float32), b.astype(np.float32))
    ref = np.linalg.solve(M, b)
    assert np.abs(x - ref).max() / np.abs(ref).max() < 2e-4, np.abs(x - ref).max() / np.abs(ref).max()


def test_generated_solve_matches_dense_rodent():
    t = G.rodent_tree()
    assert (t.nv, t.nM, t.maxdepth) == (73, 1119, 35)
    for seed in range(3):
        _check(t, seed)


def test_generated_solve_random_trees():
    rng = np.random.default_rng(5)
    for nv in (1, 7, 33, 70, 96):
        parent = [-1] + [int(rng.integers(max(0, i - 6), i)) for i in range(1, nv)]
        _check(G.Tree(parent), nv)


def _check_factor(tree, seed):
    rng = np.random.default_rng(seed)
    M = G.random_tree_spd(tree, rng)
    M2 = M + np.diag(rng.uniform(0.0, 0.3, tree.nv))
    s1, s2 = G.sparse_from_dense(tree, M), G.sparse_from_dense(tree, M2)
    L1, L2 = G.run_factor_ir(G.build_factor_ir(tree), tree, s1.astype(np.float32), s2.astype(np.float32))
    for got, ref in ((L1, G.factor_ref(tree, s1)), (L2, G.factor_ref(tree, s2))):
        assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-5


def test_generated_factor_matches_reference_rodent():
    t = G.rodent_tree()
    for seed in range(2):
        _check_factor(t, seed)


def test_generated_factor_random_trees():
    rng = np.random.default_rng(11)
    for nv in (1, 5, 20, 40, 64):
        parent = [-1] + [int(rng.integers(max(0, i - 3), i)) for i in range(1, nv)]
        _check_factor(G.Tree(parent), nv)


def test_generated_mul_m_matches_dense():
    rng = np.random.default_rng(3)
    for tree in (G.rodent_tree(), G.Tree([-1] + [int(rng.integers(max(0, i - 4), i)) for i in range(1, 50)])):
        M = G.random_tree_spd(tree, rng)
        x = rng.normal(size=tree.nv)
        y = G.run_mulm_ir(G.build_mulm_ir(tree), tree, G.sparse_from_dense(tree, M), x.astype(np.float32))
        ref = M @ x
        assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-5


def test_committed_header_is_current():
    assert open(G.OUT).read() == G.emit_cuda(G.rodent_tree())
