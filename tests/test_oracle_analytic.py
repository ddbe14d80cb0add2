"""Pins the CPU oracle (oracle/rigid_oracle.hpp) with the reference's own *analytic* test assertions.

The reference stores no golden vectors for the rigid step (its tests consult MuJoCo live), so
the restatement is pinned by re-expressing the closed-form assertions of the reference test-suite:
  tests/test_rigid_physics.py:2285-2317  test_mass_mat          (L^T D L == M)
  tests/test_rigid_physics.py:903-1030   test_pendulum_links_acc (closed-form pendulum acceleration)
  tests/test_rigid_physics.py:3468-3494  test_gravity            (free fall under per-scene gravity)
  tests/test_rigid_physics.py:1387-1483  test_position_control   (PD via control_dofs_position == explicit torque)
  tests/test_rigid_physics.py:2123-2191  test_contact_forces     (net contact force == weight at rest)
  tests/test_rigid_physics.py:413-500    test_box_plane_dynamics (box settles on the plane, no drift)
"""

import numpy as np
import pytest

from conftest import TOL_SINGLE
from helpers import PENDULUM_XML, gs, host_model, make_franka, make_go2, oracle_for, scene_from_morphs, write_tmp


def _rand_qpos(model, rng, scale=0.5):
    q = np.array(model.qpos0, dtype=np.float64)
    lim = np.asarray(model.dofs_limit).reshape(-1, 2)
    for j in range(model.n_joints):
        t, qs, ds = int(model.joints_type[j]), int(model.joints_q_start[j]), int(model.joints_dof_start[j])
        if t in (gs.JOINT_TYPE.REVOLUTE, gs.JOINT_TYPE.PRISMATIC):
            lo, hi = lim[ds]
            lo, hi = (max(lo, -2.0), min(hi, 2.0))
            q[qs] = lo + (hi - lo) * rng.random()
        elif t == gs.JOINT_TYPE.FREE:
            quat = rng.normal(size=4)
            q[qs + 3: qs + 7] = quat / np.linalg.norm(quat)
    return q


@pytest.mark.parametrize("make", [make_franka, make_go2], ids=["franka", "go2"])
@pytest.mark.parametrize("double", [True, False], ids=["f64", "f32"])
def test_mass_matrix_factorisation(make, double):
    """test_mass_mat: the L^T D L factor reproduces M; M is symmetric positive definite."""
    scene, solver, model = host_model(make)
    orc = oracle_for(solver, model, 1, double=double)
    rng = np.random.default_rng(0)
    nD = model.n_dofs
    for _ in range(5):
        orc.set("qpos", _rand_qpos(model, rng))
        orc.forward_dynamics(0)
        M = orc.get("M", nD * nD, 0).reshape(nD, nD)
        L = np.tril(orc.get("L", nD * nD, 0).reshape(nD, nD))
        Dinv = orc.get("Dinv", nD, 0)
        assert np.allclose(M, M.T, atol=1e-12 if double else 1e-6)
        assert np.linalg.eigvalsh(M).min() > 0
        rec = L.T @ np.diag(1.0 / Dinv) @ L
        tol = 1e-9 if double else TOL_SINGLE
        np.testing.assert_allclose(rec, M, atol=tol * max(1.0, np.abs(M).max()), rtol=tol)


def test_mass_matrix_matches_kinetic_energy():
    """M from CRBA equals the Jacobian form sum_l J_l^T I_l J_l, checked through 0.5 v^T M v = sum of link kinetic energies."""
    scene, solver, model = host_model(make_go2)
    orc = oracle_for(solver, model, 1, double=True)
    rng = np.random.default_rng(1)
    nD, nL = model.n_dofs, model.n_links
    orc.set("qpos", _rand_qpos(model, rng))
    v = rng.normal(size=nD)
    orc.set("vel", v)
    orc.forward_dynamics(0)
    M = orc.get("M", nD * nD, 0).reshape(nD, nD) - np.diag(model.dofs_armature)
    # approximate_implicitfast adds dt*(damping + kv) to diag(M) (forward_dynamics.py:528-541): remove it again
    M = M - np.diag(model.dofs_damping * model.dt)
    ke_M = 0.5 * v @ M @ v
    # link kinetic energy from the c-frame spatial velocity and the c-frame inertia (cinr) of each link
    cd_vel, cd_ang = orc.get("cd_vel", (nL, 3), 0), orc.get("cd_ang", (nL, 3), 0)
    i_pos = orc.get("i_pos", (nL, 3), 0)
    lquat = orc.get("links_quat", (nL, 4), 0)
    ke = 0.0
    from genesis_b200.model import geom_np as gn

    for l in range(nL):
        m = model.links_inertial_mass[l]
        vc = cd_vel[l] + np.cross(cd_ang[l], i_pos[l])  # velocity of the link COM
        R = gn.quat_to_R(gn.quat_mul(lquat[l], model.links_inertial_quat[l]))
        Iw = R @ model.links_inertial_i[l].reshape(3, 3) @ R.T
        ke += 0.5 * m * vc @ vc + 0.5 * cd_ang[l] @ Iw @ cd_ang[l]
    assert ke_M == pytest.approx(ke, rel=1e-9)


def test_pendulum_closed_form_acceleration():
    """test_pendulum_links_acc: theta_ddot = -m g l sin(theta) / (I + m l^2)."""
    path = write_tmp(PENDULUM_XML, ".xml")
    scene, _ = scene_from_morphs([gs.morphs.MJCF(file=path, default_armature=0.0)], enable_collision=False)
    model = scene.rigid_solver.build_model()
    orc = oracle_for(scene.rigid_solver, model, 1, double=True)
    for theta in (0.1, 0.3, -1.2, 2.5):
        orc.set("qpos", [theta])
        orc.forward_dynamics(0)
        expect = -9.81 * 0.5 * np.sin(theta) / (0.01 + 0.25)
        assert orc.get("acc_smooth", 1, 0)[0] == pytest.approx(expect, rel=1e-12)


def test_pendulum_energy_conservation():
    """Semi-implicit Euler keeps the pendulum's energy bounded (symplectic): |dE|/E stays below 1% over 5 s."""
    path = write_tmp(PENDULUM_XML, ".xml")
    scene, _ = scene_from_morphs([gs.morphs.MJCF(file=path, default_armature=0.0)], enable_collision=False, dt=0.001)
    model = scene.rigid_solver.build_model()
    orc = oracle_for(scene.rigid_solver, model, 1, double=True)
    orc.set("qpos", [1.0])

    def energy():
        th, w = orc.qpos()[0, 0], orc.vel()[0, 0]
        return 0.5 * (0.01 + 0.25) * w * w - 9.81 * 0.5 * np.cos(th)

    e0 = energy()
    for _ in range(50):
        orc.step(100)
        assert abs(energy() - e0) < 0.01 * abs(e0)


@pytest.mark.parametrize("gravity", [(0.0, 0.0, -9.81), (0.0, 0.0, 3.0), (1.0, -2.0, -5.0)])
def test_free_fall(gravity):
    """test_gravity: a free box far above the plane accelerates with g; positions follow semi-implicit Euler exactly."""
    scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0.0, 0.0, 50.0))], dt=0.01,
                                 gravity=gravity)
    model = scene.rigid_solver.build_model()
    orc = oracle_for(scene.rigid_solver, model, 1, double=True)
    K = 100
    orc.step(K)
    g = np.asarray(gravity)
    np.testing.assert_allclose(orc.vel()[0, :3], g * K * 0.01, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.qpos()[0, :3], np.array([0, 0, 50.0]) + g * 0.01**2 * K * (K + 1) / 2, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.qpos()[0, 3:7], [1, 0, 0, 0], atol=1e-12)


def test_position_control_equals_explicit_torque():
    """test_position_control: control_dofs_position(q*) gives the same step as control_dofs_force(kp (q*-q) - kv v)."""
    scene, solver, model = host_model(make_franka)
    a = oracle_for(solver, model, 1, double=True)
    b = oracle_for(solver, model, 1, double=True)
    rng = np.random.default_rng(3)
    q = _rand_qpos(model, rng)
    v = 0.3 * rng.normal(size=model.n_dofs)
    tgt = q + 0.1 * rng.normal(size=model.n_dofs)
    for o in (a, b):
        o.set("qpos", q)
        o.set("vel", v)
    a.set("ctrl_pos", tgt)
    a.set("ctrl_mode", np.full(model.n_dofs, int(gs.CTRL_MODE.POSITION)))
    bias = np.asarray(model.dofs_act_bias).reshape(-1, 3)
    dpos = q - np.asarray(model.qpos0)
    tau = model.dofs_kp * (tgt - dpos) + bias[:, 0] + (model.dofs_kp + bias[:, 1]) * dpos + bias[:, 2] * v
    fr = np.asarray(model.dofs_force_range).reshape(-1, 2)
    tau = np.clip(tau, fr[:, 0], fr[:, 1])
    b.set("ctrl_force", tau)
    b.set("ctrl_mode", np.full(model.n_dofs, int(gs.CTRL_MODE.FORCE)))
    a.forward_dynamics(0)
    b.forward_dynamics(0)
    np.testing.assert_allclose(a.get("qf_applied", model.n_dofs, 0), tau, rtol=1e-12, atol=1e-12)
    # identical applied torque; the POSITION arm additionally carries dt*kv on diag(M) (approximate implicitfast)
    np.testing.assert_allclose(a.get("qf_applied", model.n_dofs, 0), b.get("qf_applied", model.n_dofs, 0), rtol=1e-12, atol=1e-12)


def test_position_control_tracks_target():
    """test_position_control (tracking part): the PD-controlled Franka converges to the commanded posture."""
    scene, solver, model = host_model(make_franka)
    orc = oracle_for(solver, model, 1, double=False)
    q0 = np.array([0.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0, 0.02, 0.02])
    orc.set("qpos", q0)
    orc.set("ctrl_pos", q0)
    orc.set("ctrl_mode", np.full(9, int(gs.CTRL_MODE.POSITION)))
    orc.step(300)
    assert np.abs(orc.qpos()[0, :7] - q0[:7]).max() < 1e-2
    assert np.abs(orc.vel()[0]).max() < 1e-2
    # fingers: the tendon actuator is a general affine actuator (gain 0.0157, bias (0, -100, -10)),
    # so the static equilibrium is q = gain * ctrl / 100, not the commanded 0.02
    np.testing.assert_allclose(orc.qpos()[0, 7:], 0.01568627451 * 0.02 / 100.0, atol=1e-6)


@pytest.mark.parametrize("double", [True, False], ids=["f64", "f32"])
def test_box_rests_on_plane_with_weight_as_contact_force(double):
    """test_contact_forces / test_box_plane_dynamics: after settling the net contact force equals the weight."""
    size, rho = 0.2, 600.0
    scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(size,) * 3, pos=(0.0, 0.0, 0.5 * size + 0.002))], dt=0.01)
    model = scene.rigid_solver.build_model()
    orc = oracle_for(scene.rigid_solver, model, 1, double=double)
    orc.step(300)
    weight = rho * size**3 * 9.81
    f = orc.get("links_contact_force", (model.n_links, 3), 0)
    assert f[1, 2] == pytest.approx(weight, rel=1e-5 if double else 5e-4)
    assert np.abs(f[1, :2]).max() < 1e-6 * weight if double else 1e-3 * weight
    np.testing.assert_allclose(f[0], -f[1], atol=1e-9 if double else 1e-3)
    assert np.abs(orc.vel()[0]).max() < 1e-4
    c = orc.contacts(0)
    assert len(c["geom_a"]) == 4 and set(c["geom_a"]) == {0} and set(c["geom_b"]) == {1}
    assert np.allclose(c["normal"], [0, 0, -1]) or np.allclose(c["normal"], [0, 0, 1])
    # resting height: penetration small and positive, box centre ~ half size above the plane
    assert 0.0 < c["penetration"].max() < 2e-3
    assert orc.qpos()[0, 2] == pytest.approx(0.5 * size, abs=2e-3)


def test_go2_stands_on_its_feet():
    """Go2 under PD control settles on four foot contacts (foot spheres vs plane -> 1 contact per foot, Appendix A)."""
    scene, solver, model = host_model(make_go2)
    orc = oracle_for(solver, model, 1, double=False)
    from genesis_b200.bench_scenes import GO2_CTRL, GO2_Q0

    orc.set("qpos", GO2_Q0)
    ctrl = np.zeros(18)
    ctrl[6:] = GO2_CTRL
    mode = np.full(18, int(gs.CTRL_MODE.FORCE))
    mode[6:] = int(gs.CTRL_MODE.POSITION)
    orc.set("ctrl_pos", ctrl)
    orc.set("ctrl_mode", mode)
    orc.step(200)
    c = orc.contacts(0)
    assert len(c["geom_a"]) == 4
    assert set(c["geom_a"]) == {0}
    foot_geoms = set(int(g) for g in c["geom_b"])
    assert len(foot_geoms) == 4 and all(model.geoms_type[g] == gs.GEOM_TYPE.SPHERE for g in foot_geoms)
    total_mass = model.links_inertial_mass[1:].sum()
    f = orc.get("links_contact_force", (model.n_links, 3), 0)
    assert f[1:, 2].sum() == pytest.approx(total_mass * 9.81, rel=2e-2)
    assert 0.2 < orc.qpos()[0, 2] < 0.45


def test_f32_oracle_tracks_f64_oracle():
    """Conditioning check used to set the GPU-parity tolerance: fp32 vs fp64 restatement over 100 controlled steps."""
    scene, solver, model = host_model(make_franka)
    q0 = np.array([0.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0, 0.02, 0.02])
    outs = []
    for double in (True, False):
        orc = oracle_for(solver, model, 1, double=double)
        orc.set("qpos", q0)
        orc.set("vel", [0.1, -0.1, 0.2, 0.1, -0.2, 0.1, 0.1, 0, 0])
        orc.set("ctrl_pos", q0)
        orc.set("ctrl_mode", np.full(9, int(gs.CTRL_MODE.POSITION)))
        orc.step(100)
        outs.append((orc.qpos()[0], orc.vel()[0]))
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-4
    assert np.abs(outs[0][1] - outs[1][1]).max() < 1e-3


def test_capsule_analytic_contacts_reference_expectations():
    """tests/test_rigid_physics_analytical_vs_gjk.py:587-615 (test_capsule_analytical_accuracy): two vertical capsules
    (r = 0.1, half length 0.25) 0.15 apart -> penetration 0.05, normal (-1, 0, 0); plus the sphere-capsule closed form."""
    from helpers import scene_from_morphs, write_tmp

    xml = """
    <mujoco><worldbody>
      <body name="a" pos="0 0 0"><freejoint/><geom type="capsule" size="0.1 0.25"/></body>
      <body name="b" pos="0.15 0 0"><freejoint/><geom type="capsule" size="0.1 0.25"/></body>
      <body name="s" pos="-0.17 0 0.1"><freejoint/><geom type="sphere" size="0.1"/></body>
    </worldbody></mujoco>"""
    scene, _ = scene_from_morphs([gs.morphs.MJCF(file=write_tmp(xml, ".xml"))], gravity=(0.0, 0.0, 0.0))
    solver = scene.rigid_solver
    m = solver.build_model()
    orc = oracle_for(solver, m, 1)
    orc.step(1)
    c = orc.contacts(0)
    pairs = {}
    for i, (a, b) in enumerate(zip(c["geom_a"], c["geom_b"])):
        pairs.setdefault((int(a), int(b)), i)  # first contact of the pair = the unperturbed detection
    i_cc = pairs[(0, 1)]
    assert abs(c["penetration"][i_cc] - 0.05) < 1e-5
    np.testing.assert_allclose(c["normal"][i_cc], (-1.0, 0.0, 0.0), atol=1e-6)
    np.testing.assert_allclose(c["position"][i_cc][:2], (0.075, 0.0), atol=1e-6)
    i_sc = pairs[(2, 0)]  # geoms of a pair are ordered by type: the sphere is geom A
    assert abs(c["penetration"][i_sc] - 0.03) < 1e-5
    np.testing.assert_allclose(c["normal"][i_sc], (-1.0, 0.0, 0.0), atol=1e-6)
    np.testing.assert_allclose(c["position"][i_sc], (-0.085, 0.0, 0.1), atol=1e-6)


def test_weld_equality_static_balance():
    """func_equality_weld (constraint/solver.py:984-1173): a free body welded to the world settles (soft constraint), its six
    weld rows carry exactly its weight, and a payload welded to an arm link keeps the anchor offset of the reference pose."""
    from helpers import make_weld

    scene, solver, m = host_model(make_weld)
    assert list(m.eq_type) == [int(gs.EQUALITY_TYPE.WELD)] * 2 and m.n_dofs == 14
    orc = oracle_for(solver, m, 1, double=True)
    orc.step(400)
    assert int(orc.get("n_constraints", 1, 0)[0]) == 12
    acc = orc.get("acc", m.n_dofs, 0)
    qfc = orc.get("qf_constraint", m.n_dofs, 0)
    # dofs 8..13: the body welded to the world (0.4 kg): at rest, constraint force = weight on the z translation
    assert np.abs(acc[8:]).max() < 1e-4  # the solver's own stopping tolerance
    assert qfc[10] == pytest.approx(0.4 * 9.81, rel=1e-5) and np.abs(qfc[8:10]).max() < 1e-5
    # the arm swings with the welded payload; the weld keeps the payload at its anchor
    lp = orc.get("links_pos", (m.n_links, 3), 0)
    lq = orc.get("links_quat", (m.n_links, 4), 0)

    def rot(q, v):
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        return R @ v

    d = m.eq_data.reshape(-1, 11)[0]
    g1 = lp[2] + rot(lq[2], d[3:6])
    g2 = lp[3] + rot(lq[3], d[0:3])
    assert np.abs(g1 - g2).max() < 5e-3  # soft constraint: millimetres of give while the arm is still swinging


def test_noslip_suppresses_creep_on_an_incline():
    """constraint/noslip.py (options/solvers.py:443-447 "suppress slip/drift"): a box on a plane under gravity tilted well
    inside the friction cone creeps with the soft pyramidal friction rows (about 1 mm/s) and stops with the noslip pass."""
    vel = {}
    for ns in (0, 5):
        scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 0.05))], dt=0.01,
                                     gravity=(2.0, 0.0, -9.81), noslip_iterations=ns)
        solver = scene.rigid_solver
        m = solver.build_model()
        orc = oracle_for(solver, m, 1, double=True)
        orc.step(200)
        vel[ns] = abs(orc.vel()[0][0])
        assert int(orc.get("n_contacts", 1, 0)[0]) == 4
    assert vel[0] > 5e-4 and vel[5] < 1e-2 * vel[0]


def test_mass_shift_changes_free_fall_contact_force_not_acceleration():
    """links_state.mass_shift (forward_kinematics.py:377, 437): free fall is mass independent, the resting contact force and
    the mass matrix are not: a 0.6 kg box with a +0.4 kg shift weighs 1.0 kg."""
    scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 0.05))], dt=0.01)
    solver = scene.rigid_solver
    m = solver.build_model()
    orc = oracle_for(solver, m, 1, double=True)
    shift = np.zeros(m.n_links)
    shift[1] = 0.4
    orc.set("mass_shift", shift)
    orc.step(150)
    M = orc.get("M", m.n_dofs * m.n_dofs, 0).reshape(m.n_dofs, m.n_dofs)
    assert M[0, 0] == pytest.approx(1.0 + float(m.dofs_armature[0]), rel=1e-9)
    f = orc.get("links_contact_force", (m.n_links, 3), 0)
    assert f[1, 2] == pytest.approx(1.0 * 9.81, rel=2e-3)
    # a COM shift along x tilts nothing for a free box at rest on its face but moves the subtree COM
    orc2 = oracle_for(solver, m, 1, double=True)
    cs = np.zeros((m.n_links, 3))
    cs[1, 0] = 0.02
    orc2.set("com_shift", cs.reshape(-1))
    orc2.step(1)
    assert orc2.get("root_COM", (m.n_links, 3), 0)[1, 0] == pytest.approx(0.02, abs=1e-6)


def test_exact_penetration_fallback_minimum_translation():
    """Row a17 (narrowphase.py:726-733 + gjk.py:420-446 result convention): for a deep overlap the contact of the pair is the
    minimum-translation one.  Two axis-aligned 0.1 cubes offset by (0.03, 0.06, 0.02) overlap by 0.07 / 0.04 / 0.08 along
    x / y / z: penetration 0.04 along y, normal from B to A = (0, -1, 0).  On random deep overlaps the fallback never reports
    a larger penetration than MPR (it is the minimum over all directions) and reports a smaller one for some."""
    xml = """
    <mujoco><option gravity="0 0 0"/><worldbody>
      <body name="a" pos="0 0 0"><freejoint/><geom type="box" size="0.05 0.05 0.05"/></body>
      <body name="b" pos="{p}" euler="{e}"><freejoint/><geom type="{t}" size="{s}"/></body>
    </worldbody></mujoco>"""

    def contact0(p, e, t, s, fallback):
        scene, _ = scene_from_morphs([gs.morphs.MJCF(file=write_tmp(xml.format(p=p, e=e, t=t, s=s), ".xml"))], dt=0.01)
        solver = scene.rigid_solver
        orc = oracle_for(solver, solver.build_model(), 1, double=True, gjk_fallback=fallback)
        orc.step(1)
        c = orc.contacts(0)
        return (c["penetration"][0], c["normal"][0], c["position"][0]) if len(c["penetration"]) else None

    pen, nrm, pos = contact0("0.03 0.06 0.02", "0 0 0", "box", "0.05 0.05 0.05", 1)
    assert pen == pytest.approx(0.04, abs=1e-6)
    np.testing.assert_allclose(nrm, (0.0, -1.0, 0.0), atol=1e-6)
    assert pos[1] == pytest.approx(0.03, abs=1e-6)  # midway between the two overlapping faces (y = 0.05 and y = 0.01)
    rng = np.random.default_rng(0)
    smaller = 0
    for trial in range(12):
        t, s = [("capsule", "0.03 0.08"), ("box", "0.03 0.06 0.02"), ("cylinder", "0.03 0.05")][trial % 3]
        p = "%f %f %f" % tuple(rng.uniform(-0.07, 0.07, 3))
        e = "%f %f %f" % tuple(rng.uniform(-90, 90, 3))
        a, b = contact0(p, e, t, s, 0), contact0(p, e, t, s, 1)
        if a is None or b is None:
            continue
        assert b[0] <= a[0] + 1e-5, f"trial {trial}: exact {b[0]} > MPR {a[0]}"
        assert abs(np.linalg.norm(b[1]) - 1.0) < 1e-6
        smaller += b[0] < a[0] - 1e-4
    assert smaller >= 3


@pytest.mark.parametrize("solver_name", ["Newton", "CG"])
@pytest.mark.parametrize("integrator_name", ["implicitfast", "Euler", "approximate_implicitfast"])
def test_box_box_dynamics_reference_expectations(solver_name, integrator_name):
    """The reference's test_box_box_dynamics (tests/test_rigid_physics.py:1031-1053), its assertions applied to the oracle: a
    0.4 m cube dropped with a small tilt onto another one comes to rest on it -- |qvel| <= 1e-2 from step 101 on and the upper
    cube's height is 0.6 +- 2e-3 after 110 steps -- for CG / Newton and Euler / implicitfast (box-box MPR contacts, the
    multi-contact probes and the exact-penetration fallback on the first touch)."""
    from helpers import box_box_qpos, make_box_box

    B = 6
    scene, solver, m = host_model(make_box_box, constraint_solver=getattr(gs.constraint_solver, solver_name),
                                  integrator=getattr(gs.integrator, integrator_name))
    orc = oracle_for(solver, m, B, threads=3)
    orc.set("qpos", box_box_qpos(B, seed=1))
    orc.step(101)
    for _ in range(9):
        orc.step(1)
        assert np.abs(orc.vel()).max() <= 1e-2
    q = orc.qpos()
    np.testing.assert_allclose(q[:, 9], 0.6, atol=2e-3)
    np.testing.assert_allclose(q[:, 2], 0.2, atol=2e-3)


def test_double_pendulum_link_accelerations_reference_closed_form():
    """The reference's test_double_pendulum_links_acc (tests/test_rigid_physics.py:955-1020; CG, Euler, fixed links kept), its
    closed-form assertions applied to the oracle for 100 steps from a random state: angular accelerations of the links, the
    spatial linear acceleration carried in the COM frame (`cacc_lin`), and the classical acceleration of the point masses
    (the `get_links_acc` formula) in the rotating frames.  Poses and velocities are the ones the accelerations were computed
    from (the state before the step: what the reference's MuJoCo-compatible test fixture reads back)."""
    from helpers import make_double_pendulum

    scene, solver, m = host_model(make_double_pendulum, constraint_solver=gs.constraint_solver.CG, integrator=gs.integrator.Euler)
    assert (m.n_links, m.n_dofs) == (5, 2)
    orc = oracle_for(solver, m, 1, double=True)
    rng = np.random.default_rng(0)
    orc.set("qpos", rng.random(2))
    orc.set("vel", rng.random(2))
    orc.step(1)
    L, sel, tol = m.n_links, [0, 2, 4], 1e-9
    for _ in range(100):
        theta, theta_dot = orc.qpos()[0].copy(), orc.vel()[0].copy()
        pos, com = orc.get("links_pos", (L, 3), 0)[sel], orc.get("root_COM", (L, 3), 0)[-1]
        ang, vel = orc.get("cd_ang", (L, 3), 0)[sel], orc.get("cd_vel", (L, 3), 0)[sel]
        orc.step(1)
        tdd = orc.get("acc", 2, 0)
        acc_ang = orc.get("cacc_ang", (L, 3), 0)
        cacc = orc.get("cacc_lin", (L, 3), 0)[sel]
        np.testing.assert_allclose(acc_ang[0], 0, atol=tol)
        np.testing.assert_allclose(acc_ang[1], [tdd[0], 0, 0], atol=tol)
        np.testing.assert_allclose(acc_ang[-1], [tdd.sum(), 0, 0], atol=tol)
        np.testing.assert_allclose(cacc[1], np.cross(acc_ang[2], com), atol=tol)
        spatial = cacc + np.cross(acc_ang[sel], pos - com)
        np.testing.assert_allclose(spatial[0], 0, atol=tol)
        tw = theta.cumsum()
        R = np.array([[np.ones(2), np.zeros(2), np.zeros(2)], [np.zeros(2), np.cos(tw), np.sin(tw)], [np.zeros(2), -np.sin(tw), np.cos(tw)]])
        local = np.matmul(np.moveaxis(R, 2, 0), spatial[1:, :, None])[..., 0]
        np.testing.assert_allclose(local[0], [0, -tdd[0], 0], atol=tol)
        np.testing.assert_allclose(local[1], R[..., 1] @ (R[..., 0].T @ np.array([0, -tdd[0], theta_dot[0] * theta_dot[1]])) + np.array([0, -tdd.sum(), 0]), atol=tol)
        classical = spatial + np.cross(ang, vel + np.cross(ang, pos - com))
        np.testing.assert_allclose(classical[0], 0, atol=tol)
        local = np.matmul(np.moveaxis(R, 2, 0), classical[1:, :, None])[..., 0]
        np.testing.assert_allclose(local[0], [0, -tdd[0], -(theta_dot[0] ** 2)], atol=tol)
        np.testing.assert_allclose(local[1], R[..., 1] @ classical[1] + np.array([0, -tdd.sum(), -(theta_dot.sum() ** 2)]), atol=tol)


def test_link_velocity_reference_expectations():
    """The reference's test_link_velocity (tests/test_rigid_physics.py:826-896, `two_aligned_hinges`): subtree COM, link
    origins and the COM-frame spatial velocities `cd_vel` / `cd_ang` for its special cases and its random configuration, plus
    the link-origin velocity formula of `get_links_vel`."""
    from helpers import make_two_hinges

    scene, solver, m = host_model(make_two_hinges, constraint_solver=gs.constraint_solver.CG, integrator=gs.integrator.Euler)
    orc = oracle_for(solver, m, 1, double=True)
    L, tol = m.n_links, 1e-9

    def state(q, v):
        orc.set("qpos", q)
        orc.set("vel", v)
        orc.forward_dynamics(0)
        return [orc.get(k, (L, 3), 0) for k in ("cd_vel", "cd_ang", "root_COM", "links_pos")]

    cvel, _, _, _ = state([0.0, 0.0], [0.0, 1.0])
    np.testing.assert_allclose(cvel, 0, atol=tol)
    cvel, _, _, _ = state([0.0, 0.0], [1.0, 0.0])
    np.testing.assert_allclose(cvel, [[0.0, 0.5, 0.0]] * 2, atol=tol)
    cvel, _, com, pos = state([0.0, np.pi / 2.0], [0.0, 1.2])
    np.testing.assert_allclose(com[0], [0.375, 0.125, 0.0], atol=tol)
    np.testing.assert_allclose(pos[1], [0.5, 0.0, 0.0], atol=tol)  # = xanchor of joint 1
    np.testing.assert_allclose(cvel[0], 0, atol=tol)
    np.testing.assert_allclose(cvel[1], [-1.2 * (0.125 - 0.0), 1.2 * (0.375 - 0.5), 0.0], atol=tol)
    # random configuration
    t0, t1 = -0.7, 0.2
    cvel, cang, com, pos = state([t0, t1], [3.0, 13.0])
    xanchor = pos[1]
    np.testing.assert_allclose(xanchor, [0.5 * np.cos(t0), 0.5 * np.sin(t0), 0.0], atol=tol)
    com0 = np.array([0.25 * np.cos(t0), 0.25 * np.sin(t0), 0.0])
    com1 = np.array([0.5 * np.cos(t0) + 0.25 * np.cos(t0 + t1), 0.5 * np.sin(t0) + 0.25 * np.sin(t0 + t1), 0.0])
    COM = com[0]
    np.testing.assert_allclose(COM, 0.5 * (com0 + com1), atol=tol)
    w0, w1 = cang[0, 2], cang[1, 2]
    assert w0 == pytest.approx(3.0, abs=tol) and w1 - w0 == pytest.approx(13.0, abs=tol)
    np.testing.assert_allclose(cvel[0], w0 * np.array([-COM[1], COM[0], 0.0]), atol=tol)
    np.testing.assert_allclose(cvel[1], cvel[0] + (w1 - w0) * np.array([xanchor[1] - COM[1], COM[0] - xanchor[0], 0.0]), atol=tol)
    np.testing.assert_allclose(pos[0], 0, atol=tol)
    xvel = cvel + np.cross(cang, pos - com)  # get_links_vel (abd/accessor.py): velocity of the link origin
    np.testing.assert_allclose(xvel[0], 0, atol=tol)
    np.testing.assert_allclose(xvel[1], w0 * np.array([-xanchor[1], xanchor[0], 0.0]), atol=tol)


@pytest.mark.parametrize("solver_name", ["CG", "Newton"])
@pytest.mark.parametrize("integrator_name", ["implicitfast", "Euler"])
def test_frictionloss_reference_expectation(solver_name, integrator_name):
    """The reference's test_frictionloss (tests/test_rigid_physics.py:509-520, `hinge_slide`: joint friction loss 0.08 / 0.3, damping,
    a 200 N/m spring; MuJoCo's default time step and no default armature like its fixture): started with qvel = (0.7, -0.9)
    the mechanism is at rest after 2000 steps, |qvel| <= 1e-2.  (The step-by-step comparison with MuJoCo of the same test
    cannot run here.)  Also pins the link masses MuJoCo's compiler infers from the capsule geoms (density 1000)."""
    from helpers import make_hinge_slide

    scene, solver, m = host_model(make_hinge_slide, constraint_solver=getattr(gs.constraint_solver, solver_name),
                                  integrator=getattr(gs.integrator, integrator_name))
    vol = lambda r, half: np.pi * r * r * 2 * half + 4.0 / 3.0 * np.pi * r**3
    np.testing.assert_allclose(np.asarray(m.links_inertial_mass)[1:], [1000 * vol(0.02, 0.05), 1000 * vol(0.015, 0.1)], rtol=1e-9)
    assert list(m.dofs_armature) == [0.0, 0.0]
    orc = oracle_for(solver, m, 1, double=True)
    orc.set("vel", [0.7, -0.9])
    orc.step(2000)
    assert np.abs(orc.vel()).max() <= 1e-2
    assert orc.get("n_constraints", 1, 0)[0] == 2  # the two friction-loss rows
