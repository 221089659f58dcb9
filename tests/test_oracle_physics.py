"""Physics invariants that pin the CPU oracle's MJX restatement (no golden vectors exist upstream).

Each check compares the oracle with an INDEPENDENT formulation written in numpy (`mjcf.py`'s Jacobian-based
mass matrix) or with a conservation law, in fp64 so that the tolerance is the algorithm's, not round-off.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from track_mjx_b200 import clips as clipmod, config, mjcf, model_blob

G = 9.81


def make_oracle(walker, cl, dtype=np.float64, sections=None, **over):
    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    args.update(over)
    cfg = config.make_task_config(walker, config.RewardConfig(), **args)
    blob = walker.blob if sections is None else model_blob.pack_sections(sections)
    return Oracle(blob, cfg, cl, dtype=dtype)


def random_state(walker, rng, n, z=0.5, vel=1.0):
    sec = walker.sections
    lo, hi = sec["jnt_range"].reshape(-1, 2)[1:].T
    qpos = np.zeros((n, walker.nq))
    qpos[:, :3] = rng.normal(0, 0.1, (n, 3)) + [0, 0, z]
    q = rng.normal(size=(n, 4))
    qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qpos[:, 7:] = lo + (hi - lo) * rng.uniform(0.3, 0.7, (n, walker.nq - 7))
    qvel = rng.normal(0, vel, (n, walker.nv))
    return qpos, qvel


def test_mass_matrix_matches_jacobian_formulation(walker, clips2):
    """crb + make_m (composite rigid body) == sum_b J_b^T diag(m, I) J_b + armature (mjcf.mass_matrix)."""
    o = make_oracle(walker, clips2)
    model = model_blob.as_model(walker.sections)
    rng = np.random.default_rng(1)
    n = 3
    buf = o.alloc(n)
    qpos, _ = random_state(walker, rng, n)
    buf["qpos"][:] = qpos
    o.forward(buf)
    for e in range(n):
        m_ref = mjcf.mass_matrix(model, buf["qpos"][e])
        m_orc = buf["dbg_qM"][e].reshape(walker.nv, walker.nv)
        assert np.allclose(m_orc, m_orc.T, atol=1e-15)
        # the blob stores fp32-rounded constants: agreement to ~1e-7 relative
        assert np.abs(m_orc - m_ref).max() <= 2e-6 * np.abs(m_ref).max()
        assert np.linalg.eigvalsh(m_orc).min() > 0


def test_gravity_bias_matches_jacobian_formulation(walker, clips2):
    """With qvel = 0, qfrc_bias = -sum_b J_b^T (m_b g)  (rne reduces to gravity compensation)."""
    o = make_oracle(walker, clips2)
    model = model_blob.as_model(walker.sections)
    rng = np.random.default_rng(2)
    buf = o.alloc(2)
    qpos, _ = random_state(walker, rng, 2)
    buf["qpos"][:] = qpos
    o.forward(buf)
    for e in range(2):
        jac, _, _ = mjcf.body_jacobians(model, buf["qpos"][e])
        ref = np.zeros(walker.nv)
        for b in range(1, walker.nbody):
            ref += jac[b, 0:3].T @ (model["body_mass"][b] * np.array([0, 0, G]))
        assert np.abs(buf["dbg_qfrc_bias"][e] - ref).max() <= 1e-6 * np.abs(ref).max() + 1e-12


def _free_flight_energy(walker, cl, dt, nsub, seed=3):
    sec = {k: v.copy() for k, v in walker.sections.items()}
    sec["dof_damping"][:] = 0
    sec["jnt_stiffness"][:] = 0
    sec["jnt_range"] = np.tile(np.array([-100.0, 100.0], np.float32), walker.njnt)  # no limit can activate
    o = make_oracle(walker, cl, sections=sec, mj_model_timestep=dt, physics_steps_per_control_step=1)
    rng = np.random.default_rng(seed)
    buf = o.alloc(1)
    qpos, qvel = random_state(walker, rng, 1, z=5.0, vel=2.0)
    buf["qpos"][:] = qpos
    o.forward(buf)                      # zeroes act / warmstart, normalises the quaternion
    buf["qvel"][:] = qvel
    mass = float(sec["body_mass"].sum())
    energies = []
    zero = np.zeros((1, walker.nu))
    for _ in range(nsub):
        v = buf["qvel"][0].copy()
        o.step(buf, zero)               # one substep; debug taps describe the state BEFORE integration
        m = buf["dbg_qM"][0].reshape(walker.nv, walker.nv)
        energies.append(0.5 * v @ m @ v + mass * G * buf["dbg_subtree_com"][0, 2])
        assert buf["dbg_contact_dist"].min() > 0.5
    return np.array(energies)


def test_energy_conservation_free_flight(walker, clips2):
    """No damping / springs / limits / contacts / actuation: E = KE + PE drifts O(dt) under semi-implicit Euler.

    This ties together kinematics, com_pos, crb (M), com_vel + rne (Coriolis, gravity), the solve and _advance:
    a sign or frame error in any of them breaks conservation at O(1)."""
    e1 = _free_flight_energy(walker, clips2, 2e-4, 200)
    e2 = _free_flight_energy(walker, clips2, 1e-4, 400)
    scale = np.abs(e1[0] - 0.26 * G * 5.0) + 1e-3      # kinetic part
    d1, d2 = abs(e1[-1] - e1[0]), abs(e2[-1] - e2[0])
    assert d1 < 2e-2 * scale, (d1, scale)
    assert d2 < 0.75 * d1 + 1e-9, (d1, d2)             # first-order: halving dt roughly halves the drift


def test_settles_on_floor_with_weight_supported(walker, clips2):
    """Static equilibrium on the floor: with the joints stiffened (springs towards the initial pose, so the
    passive model can stand on its paws) and zero control, the pyramidal contact forces converge to the weight.

    Exercises plane-capsule/ellipsoid collision, contact Jacobians, impedance / reference acceleration, the CG
    solver with warm start and the Euler update end to end."""
    cl = clips2
    q0 = np.concatenate([cl.position[0, 0], cl.quaternion[0, 0], cl.joints[0, 0]])
    sec = {k: v.copy() for k, v in walker.sections.items()}
    sec["jnt_stiffness"][1:] = 2.0
    sec["dof_damping"][6:] = 0.02
    sec["qpos_spring"][7:] = q0[7:]
    o = make_oracle(walker, clips2, sections=sec)
    buf = o.alloc(1)
    buf["qpos"][0] = q0
    o.forward(buf)
    buf["qpos"][0, 2] -= buf["dbg_contact_dist"].min() - 1e-3   # lowest paw 1 mm above the floor
    o.forward(buf)
    zero = np.zeros((1, walker.nu))
    for _ in range(100):
        o.step(buf, zero)
    assert not np.isnan(buf["qpos"]).any()
    nlim = walker.nefc - 4 * walker.ncon
    normal = buf["dbg_efc_force"][0, nlim:].sum()        # pyramid rows: sum of forces == normal force
    weight = float(sec["body_mass"].sum()) * G
    assert np.abs(buf["qvel"]).max() < 0.02
    assert -2e-3 < buf["dbg_contact_dist"].min() < 0      # soft contact: ~1 mm penetration
    assert abs(normal - weight) < 0.01 * weight, (normal, weight)


def test_newton_and_converged_cg_agree(walker, clips2):
    """Both solvers minimise the same convex cost: run to convergence they give the same trajectory."""
    import common

    outs = []
    for solver, it in (("cg", 100), ("newton", 50)):
        o = make_oracle(walker, clips2, solver=solver, iterations=it, ls_iterations=50)
        buf = o.alloc(4)
        common.put(buf, common.init_buffers(buf, clips2, seed=0))
        o.forward(buf)
        for _ in range(8):
            o.step(buf, np.zeros((4, walker.nu)))
        outs.append(buf["qvel"].copy())
    assert np.abs(outs[0] - outs[1]).max() < 0.05 * np.abs(outs[0]).max()


def test_fp32_oracle_tracks_fp64_over_one_step(walker, clips2):
    """First control step after reset (free fall, no active constraints yet, zero control): the fp32 restatement
    stays within 1e-4 relative of fp64.  (Once contacts are active the 5-iteration CG path amplifies round-off by
    orders of magnitude within one control step -- see DESIGN.md "Tolerances" -- so longer horizons are compared
    statistically, not per element.)"""
    o64, o32 = make_oracle(walker, clips2), make_oracle(walker, clips2, dtype=np.float32)
    import common

    n = 16
    b64, b32 = o64.alloc(n), o32.alloc(n)
    init = common.init_buffers(b32, clips2, seed=5)
    common.put(b64, init)
    common.put(b32, init)
    o64.forward(b64)
    o32.forward(b32)
    act = np.zeros((n, walker.nu))
    o64.step(b64, act)
    o32.step(b32, act)
    for k in ("qpos", "qvel", "xpos"):
        assert common.err(b32[k], b64[k])[1] < 1e-4, k
