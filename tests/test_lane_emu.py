"""CPU tier: arithmetic of the lane-per-environment step kernel (genesis_b200/csrc/gsr_lane.cuh) against the oracle.

The lane kernel maps one thread to one environment and uses no warp intrinsics inside a phase, so its device functions
compile verbatim for the host (tests/lane_emu, test infrastructure).  These tests run that host build environment by
environment on the same seeded inputs as the oracle, with the tolerances of tests/test_gpu_parity.py.  What they do NOT
cover is the device-only part (interleaved shared-memory addressing, cooperative record copies, launch geometry): that is
the job of the `-m gpu` suite run with GSR_GROUP_SIZE=1.
"""

import numpy as np
import pytest

from conftest import TOL_SINGLE
from emu import LaneEmu
from helpers import gs, host_model, make_franka, make_go2
from helpers import make_franka_cluster_r1 as make_franka_grasp
from helpers import oracle_for as _oracle_for


def oracle_for(*args, **kw):
    """The lane-per-environment kernel is frozen at its round-1 state: after MPR it runs EPA seeded with MPR's final
    portal, where the reference -- and now the oracle and the lane-group kernels -- re-run the pair through GJK + EPA.  The oracle
    keeps that round-1 variant as gjk_fallback = 2 so that this emulation tier still checks the lane kernel's arithmetic."""
    kw.setdefault("gjk_fallback", 2)
    return _oracle_for(*args, **kw)
from genesis_b200.bench_scenes import FRANKA_Q0, GO2_CTRL, GO2_Q0

TOL_SOLVER = 5e-4


def pair(make, B, sets, **kw):
    scene, solver, m = host_model(make, **kw)
    orc = oracle_for(solver, m, B, threads=4)
    emu = LaneEmu(m, B, solver.model_options())
    for name, val in sets(m, B).items():
        orc.set(name, val)
        emu.set(name, val)
    return m, orc, emu


def franka_sets(controlled=True, randomized=True):
    def f(m, B):
        rng = np.random.default_rng(0)
        q = np.tile(np.array(FRANKA_Q0), (B, 1))
        out = {"qpos": q}
        if randomized:
            v = 0.2 * np.clip(rng.normal(size=(B, 9)), -1, 1)
            v[:, 7:] = 0
            out["vel"] = v
        if controlled:
            out["ctrl_pos"] = q.copy()
            out["ctrl_mode"] = np.full((B, 9), int(gs.CTRL_MODE.POSITION))
        return out
    return f


def go2_sets(m, B):
    rng = np.random.default_rng(0)
    lim = np.asarray(m.dofs_limit).reshape(-1, 2)[6:]
    q = np.tile(np.array(GO2_Q0), (B, 1))
    q[:, 7:] = lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((B, 12))
    ctrl = np.zeros((B, 18))
    ctrl[:, 6:] = GO2_CTRL
    mode = np.full((B, 18), int(gs.CTRL_MODE.FORCE))
    mode[:, 6:] = int(gs.CTRL_MODE.POSITION)
    return {"qpos": q, "ctrl_pos": ctrl, "ctrl_mode": mode}


def close(a, b, tol, what):
    scale = max(1.0, np.abs(b).max())
    err = np.abs(a - b).max()
    assert err <= tol * scale, f"{what}: max|diff| {err:.3e} > {tol:.1e} * {scale:.3e}"


def check_all_quantities(m, orc, emu):
    pre = {"acc_smooth", "qf_bias", "qf_applied", "qf_passive", "mass_mat", "mass_Dinv"}
    for gname, oname in [("qpos", "qpos"), ("vel", "vel"), ("acc", "acc"), ("acc_smooth", "acc_smooth"), ("qf_bias", "qf_bias"),
                         ("qf_applied", "qf_applied"), ("qf_passive", "qf_passive"), ("force", "force"), ("mass_mat", "M"),
                         ("mass_Dinv", "Dinv"), ("qf_constraint", "qf_constraint"), ("dofs_pos", "dofs_pos")]:
        a = emu.rows(gname)
        close(a, orc.get(oname, a.shape[1]), TOL_SINGLE if gname in pre else TOL_SOLVER, gname)
    for name, comps in [("links_pos", 3), ("links_quat", 4), ("root_COM", 3), ("cd_vel", 3), ("cd_ang", 3), ("cacc_lin", 3),
                        ("cacc_ang", 3), ("links_contact_force", 3)]:
        a = emu.rows(name)
        close(a, orc.get(name, (m.n_links, comps)).reshape(a.shape), TOL_SOLVER, name)
    for name in ("geoms_pos", "aabb_min", "aabb_max") if m.n_geoms else ():
        a = emu.rows(name)
        close(a, orc.get(name, (m.n_geoms, 3)).reshape(a.shape), TOL_SOLVER if name == "geoms_pos" else TOL_SINGLE, name)
    for name in ("cdof_ang", "cdof_vel", "cdofd_ang", "cdofd_vel"):
        a = emu.rows(name)
        close(a, orc.get(name, (m.n_dofs, 3)).reshape(a.shape), TOL_SOLVER, name)


def contacts_equal(orc, emu, B):
    n_e = emu.field("n_contacts")[:, 0]
    n_o = orc.get("n_contacts", 1)[:, 0].astype(int)
    assert (n_e == n_o).all()
    for b in range(B):
        c = orc.contacts(b)
        n = int(n_o[b])
        assert (emu.field("contact_geom_a")[b, :n] == c["geom_a"]).all() and (emu.field("contact_geom_b")[b, :n] == c["geom_b"]).all()
        if n:
            assert np.abs(emu.field("contact_pos")[b, :n] - c["position"]).max() <= 1e-4
            assert np.abs(emu.field("contact_normal")[b, :n] - c["normal"]).max() <= 1e-4
            assert np.abs(emu.field("contact_penetration")[b, :n] - c["penetration"]).max() <= 1e-4


@pytest.mark.parametrize("variant", ["random", "fall"])
def test_lane_franka_single_step_all_quantities(variant):
    m, orc, emu = pair(make_franka, 4, franka_sets(controlled=variant != "fall", randomized=variant == "random"))
    orc.step(1)
    emu.step(1)
    check_all_quantities(m, orc, emu)


def test_lane_go2_single_step_and_trajectory():
    B = 6
    m, orc, emu = pair(make_go2, B, go2_sets)
    orc.step(1)
    emu.step(1)
    check_all_quantities(m, orc, emu)
    contacts_equal(orc, emu, B)
    orc.step(39)
    emu.step(39)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    assert int(emu.field("errno_").max()) == 0
    # rows of a Go2 on its feet exceed the lane kernel's on-chip capacity: the SPILL path is what ran
    assert int(emu.field("n_constraints").max()) > 8


def test_lane_franka_controlled_trajectory():
    m, orc, emu = pair(make_franka, 4, franka_sets())
    orc.step(50)
    emu.step(50)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-4 and np.abs(emu.rows("vel") - orc.vel()).max() <= 1e-3


def test_lane_franka_fall_self_collision_contacts():
    B = 3

    def sets(m, B):
        rng = np.random.default_rng(1)
        q = np.tile(np.array(FRANKA_Q0), (B, 1)) + 0.05 * rng.normal(size=(B, 9))
        q[:, 7:] = 0.02
        return {"qpos": q}

    m, orc, emu = pair(make_franka, B, sets, self_collision=True)
    orc.step(80)
    emu.step(80)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    assert int(orc.get("n_contacts", 1).sum()) > 0
    contacts_equal(orc, emu, B)


def test_lane_grasp_cluster_box_mesh_contacts():
    B = 2

    def sets(m, B):
        q = np.tile(np.asarray(m.qpos0, dtype=float), (B, 1))
        q[:, :9] = [-0.0, 0.45, 0.0, -2.2, 0.0, 2.65, 0.78, 0.04, 0.04]
        ctrl = np.zeros((B, m.n_dofs))
        ctrl[:, :9] = q[:, :9]
        ctrl[:, 7:9] = 0.0
        mode = np.full((B, m.n_dofs), int(gs.CTRL_MODE.FORCE))
        mode[:, :9] = int(gs.CTRL_MODE.POSITION)
        return {"qpos": q, "ctrl_pos": ctrl, "ctrl_mode": mode}

    m, orc, emu = pair(make_franka_grasp, B, sets)
    orc.step(30)
    emu.step(30)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    assert int(orc.get("n_contacts", 1).sum()) >= 16 * B
    contacts_equal(orc, emu, B)


@pytest.mark.parametrize("opts", [dict(constraint_solver="CG"), dict(integrator="Euler"), dict(integrator="implicitfast")])
def test_lane_solver_and_integrator_options(opts):
    kw = {}
    if "constraint_solver" in opts:
        kw["constraint_solver"] = gs.constraint_solver.CG
    if "integrator" in opts:
        kw["integrator"] = getattr(gs.integrator, opts["integrator"])
    B = 3
    m, orc, emu = pair(make_go2, B, go2_sets, **kw)
    orc.step(20)
    emu.step(20)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2


def test_lane_capsule_analytic_pairs():
    """capsule-capsule / sphere-capsule closed forms (collider/capsule_contact.py) and plane-capsule support contacts."""
    from helpers import make_capsules

    B = 2
    m, orc, emu = pair(make_capsules, B, lambda m, B: {})
    orc.step(40)
    emu.step(40)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    types = {(int(m.geoms_type[a]), int(m.geoms_type[b])) for a, b in zip(orc.contacts(0)["geom_a"], orc.contacts(0)["geom_b"])}
    assert (4, 4) in types and ((1, 4) in types or (0, 1) in types)
    contacts_equal(orc, emu, B)


def test_lane_weld_equality():
    from helpers import make_weld

    B = 2
    m, orc, emu = pair(make_weld, B, lambda m, B: {})
    orc.step(1)
    emu.step(1)
    check_all_quantities(m, orc, emu)
    orc.step(99)
    emu.step(99)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-4 and np.abs(emu.rows("vel") - orc.vel()).max() <= 1e-3
    assert (emu.field("n_constraints_equality")[:, 0] == 12).all()


def test_lane_noslip_post_solver():
    """constraint/noslip.py: the kernel keeps the rows M^-1 J^T and updates J^T f incrementally, the oracle materialises
    A = J M^-1 J^T like the reference -- two formulations of the same projected Gauss-Seidel."""
    B = 3
    m, orc, emu = pair(make_go2, B, go2_sets, noslip_iterations=5)
    scene, solver, _ = host_model(make_go2)
    ref = oracle_for(solver, m, B, threads=3, noslip_iterations=0)
    for name, val in go2_sets(m, B).items():
        ref.set(name, val)
    orc.step(50)
    emu.step(50)
    ref.step(50)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    assert np.abs(orc.qpos() - ref.qpos()).max() > 5e-3  # the pass is not a no-op on this scene


def test_lane_mass_and_com_shift():
    """Per-environment mass / COM shifts (domain randomisation): different shifts per environment, Go2 trajectories."""
    B = 3

    def sets(m, B):
        out = go2_sets(m, B)
        rng = np.random.default_rng(3)
        out["mass_shift"] = np.abs(rng.normal(size=(B, m.n_links))) * 0.5
        out["mass_shift"][:, 0] = 0.0
        out["com_shift"] = 0.02 * rng.normal(size=(B, m.n_links * 3))
        return out

    scene, solver, m = host_model(make_go2)
    orc = oracle_for(solver, m, B, threads=3)
    emu = LaneEmu(m, B, solver.model_options())
    for name, val in sets(m, B).items():
        orc.set(name, val)
        emu.set(name, val)
    emu.field("has_shift")[:] = 1
    ref = oracle_for(solver, m, B, threads=3)
    for name, val in go2_sets(m, B).items():
        ref.set(name, val)
    orc.step(1)
    emu.step(1)
    check_all_quantities(m, orc, emu)
    orc.step(39)
    emu.step(39)
    ref.step(40)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2
    assert np.abs(orc.qpos() - ref.qpos()).max() > 1e-3  # the shifts matter


def _np_quat_mul(u, v):
    return np.array([u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3], u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
                     u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1], u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0]])


def _np_rotate(v, q):
    qv = np.array([0.0, *v])
    qc = np.array([q[0], -q[1], -q[2], -q[3]])
    return _np_quat_mul(_np_quat_mul(q, qv), qc)[1:]


def test_lane_dynamic_weld_constraint():
    """add_weld_constraint (constraint/solver.py:1293-1340): a box welded in flight to a second, spinning box."""
    from helpers import scene_from_morphs

    B = 2
    scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 1.0)),
                                  gs.morphs.Box(size=(0.08, 0.08, 0.08), pos=(0.2, 0.05, 1.1))], dt=0.01)
    solver = scene.rigid_solver
    m = solver.build_model()
    orc = oracle_for(solver, m, B, threads=2)
    emu = LaneEmu(m, B, solver.model_options())
    v = np.zeros((B, m.n_dofs))
    v[:, 3:6] = [1.0, -2.0, 0.5]
    v[:, 6] = 0.3
    orc.set("vel", v)
    emu.set("vel", v)
    orc.step(5)
    emu.step(5)
    l1, l2 = 1, 2
    for b in range(B):
        orc.set("add_weld", [l1, l2], env=b)
        p1, q1 = emu.field("links_pos")[b, l1].astype(float), emu.field("links_quat")[b, l1].astype(float)
        p2, q2 = emu.field("links_pos")[b, l2].astype(float), emu.field("links_quat")[b, l2].astype(float)
        q1c, q2c = q1 * [1, -1, -1, -1], q2 * [1, -1, -1, -1]
        rec = np.zeros(20, np.float32)
        rec[:2] = np.array([l1, l2], np.int32).view(np.float32)
        rec[2:5] = _np_rotate(p1 - p2, q2c)
        rec[8:12] = _np_quat_mul(q1c, q2)
        rec[12] = 1.0
        rec[13:20] = [2 * 0.01, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]
        emu.field("dyn_eq")[b, 0] = rec
        emu.field("n_dyn_eq")[b, 0] = 1
    d0 = np.linalg.norm(emu.field("links_pos")[0, l1] - emu.field("links_pos")[0, l2])
    orc.step(40)
    emu.step(40)
    assert (emu.field("n_constraints_equality")[:, 0] == 6).all()
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-4 and np.abs(emu.rows("vel") - orc.vel()).max() <= 1e-3
    d1 = np.linalg.norm(emu.field("links_pos")[0, l1] - emu.field("links_pos")[0, l2])
    assert abs(d1 - d0) < 5e-3  # the welded boxes keep their distance while tumbling (soft constraint: millimetres of give)


def test_lane_exact_penetration_fallback():
    """Deeply interpenetrating convex pairs: the EPA fallback after MPR (gsr_epa.cuh) against the oracle's own implementation."""
    from helpers import make_jumble

    B = 2
    m, orc, emu = pair(make_jumble, B, lambda m, B: {})
    scene, solver, _ = host_model(make_jumble)
    ref = oracle_for(solver, m, B, threads=2, gjk_fallback=0)
    orc.step(1)
    emu.step(1)
    ref.step(1)
    contacts_equal(orc, emu, B)
    # the fallback is what decided these contacts: MPR alone reports different ones
    c1, c0 = orc.contacts(0), ref.contacts(0)
    assert len(c1["penetration"]) != len(c0["penetration"]) or np.abs(c1["penetration"] - c0["penetration"]).max() > 1e-4
    # the bodies fly apart at > 10 m/s; compare until they land on each other again (later the fp32 paths bifurcate, as the
    # fp32 and fp64 oracles do between themselves)
    orc.step(19)
    emu.step(19)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-4 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2


def test_lane_box_box_first_touch_and_settling():
    """The reference's box_box scene (tests/test_rigid_physics.py:82-97, 1031-1053): free fall, first touch (exact-penetration
    fallback, un-cached pair), multi-contact settling of a cube on a cube against the oracle."""
    from helpers import box_box_qpos, make_box_box

    B = 3
    q0 = box_box_qpos(B, seed=3)
    m, orc, emu = pair(make_box_box, B, lambda m, B: {"qpos": q0})
    for k in range(3):
        orc.step(10)
        emu.step(10)
        contacts_equal(orc, emu, B)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-3 and np.abs(emu.rows("vel") - orc.vel()).max() <= 2e-2


def test_lane_double_pendulum_link_accelerations():
    """The reference's double pendulum (tests/test_rigid_physics.py:230-297, fixed links kept, CG + Euler): every quantity of a
    step and the link accelerations along a 50-step swing against the oracle (which satisfies the reference's closed-form
    assertions, tests/test_oracle_analytic.py)."""
    from helpers import make_double_pendulum

    B = 3
    rng = np.random.default_rng(1)
    q0, v0 = rng.random((B, 2)), rng.random((B, 2))
    m, orc, emu = pair(make_double_pendulum, B, lambda m, B: {"qpos": q0, "vel": v0}, constraint_solver=gs.constraint_solver.CG,
                       integrator=gs.integrator.Euler)
    orc.step(1)
    emu.step(1)
    check_all_quantities(m, orc, emu)
    orc.step(49)
    emu.step(49)
    assert np.abs(emu.rows("qpos") - orc.qpos()).max() <= 1e-4 and np.abs(emu.rows("vel") - orc.vel()).max() <= 1e-3
    for name in ("cacc_lin", "cacc_ang"):
        a = emu.rows(name)
        close(a, orc.get(name, (m.n_links, 3)).reshape(a.shape), TOL_SOLVER, name)


def test_lane_all_fixed_scene_steps():
    """The reference's test_all_fixed (tests/test_rigid_physics.py:2093-2115): a scene whose only entity is a fixed box (no dofs,
    no qs) builds and steps; the box stays where it is (oracle and lane kernel)."""
    from helpers import scene_from_morphs

    scene, _ = scene_from_morphs([gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0.0, 0.0, 0.0), fixed=True)], dt=0.01)
    solver = scene.rigid_solver
    m = solver.build_model()
    assert (m.n_links, m.n_dofs, m.n_qs, m.n_geoms) == (1, 0, 0, 1)
    orc = oracle_for(solver, m, 2)
    emu = LaneEmu(m, 2, solver.model_options())
    orc.step(3)
    emu.step(3)
    assert np.abs(orc.get("links_pos", (1, 3))).max() == 0.0 and np.abs(emu.field("links_pos")).max() == 0.0
    assert int(emu.field("errno_").max()) == 0
