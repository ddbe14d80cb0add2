"""Parity tests proper: the CUDA path (through the C ABI / ctypes, genesis_b200.RigidSolver) against the CPU oracle on the
same seeded inputs.  Tolerances (stated here, used below):

  * single step, every quantity computed before the constraint solver: |a - b| <= TOL_SINGLE * max(1, max|b|) (the
    reference's own 5e-5); quantities downstream of the iterative solver (acc, vel, qpos, forces), whose stopping rule
    `improvement < tolerance * meaninertia * nD` (solver.py:2851) leaves that much slack:   |a - b| <= 5e-4 * max(1, max|b|)
  * trajectories, contact-free (Franka PD, K=50):     max|dqpos| <= 1e-4, max|dqvel| <= 1e-3
  * trajectories with contacts (Go2 K=100, Franka fall K=150): max|dqpos| <= 1e-3, max|dqvel| <= 2e-2,
    contact-set agreement: Jaccard index of the (env, geom_a, geom_b) sets >= 0.98
The fp32-vs-fp64 conditioning of the same trajectories is measured in tests/test_oracle_analytic.py.
"""

import numpy as np
import pytest
import torch

from conftest import TOL_SINGLE
from helpers import gs, make_capsules, make_jumble, make_weld, make_franka, make_franka_grasp, make_go2, oracle_for, scene_from_morphs

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _diagnostics_on(monkeypatch):
    """The parity tests read the intermediate quantities (mass_mat, qf_*, acc_smooth, geoms_pos) of the step that produced the
    state: cfg_diagnostics = 1 writes them every step (the product default is 0: on demand, see test_lazy_diagnostics)."""
    monkeypatch.setenv("GSR_DIAGNOSTICS", "1")


TRAJ_QPOS_FREE, TRAJ_QVEL_FREE = 1e-4, 1e-3
TRAJ_QPOS_CONTACT, TRAJ_QVEL_CONTACT = 1e-3, 2e-2
JACCARD_MIN = 0.98


def mirror_oracle(scene, double=False, **opts):
    """Oracle with the same model, options, qpos, velocities and controls as the built CUDA scene."""
    s = scene.rigid_solver
    orc = oracle_for(s, s.model, s._B, double=double, threads=8, **opts)
    for name in ("qpos", "vel", "ctrl_pos", "ctrl_vel", "ctrl_force", "ctrl_mode"):
        orc.set(name, s.rows(name).cpu().numpy())
    return orc


def gpu_field(scene, name):
    return scene.rigid_solver.rows(name).cpu().numpy().astype(np.float64)


def contact_set(scene_or_orc, B):
    out = set()
    if hasattr(scene_or_orc, "rigid_solver"):
        s = scene_or_orc.rigid_solver
        n = s.field("n_contacts")[:, 0].cpu().numpy()
        n_max = int(n.max()) if len(n) else 0
        ga = s.field("contact_geom_a")[:, :n_max].cpu().numpy()
        gb = s.field("contact_geom_b")[:, :n_max].cpu().numpy()
        for b in range(B):
            out |= {(b, int(ga[b, c]), int(gb[b, c])) for c in range(n[b])}
    else:
        for b in range(B):
            c = scene_or_orc.contacts(b)
            out |= {(b, int(a), int(bb)) for a, bb in zip(c["geom_a"], c["geom_b"])}
    return out


def jaccard(a, b):
    return 1.0 if not (a or b) else len(a & b) / len(a | b)


def assert_close(a, b, tol, what):
    scale = max(1.0, np.abs(b).max())
    err = np.abs(a - b).max()
    assert err <= tol * scale, f"{what}: max|diff| {err:.3e} > {tol:.1e} * {scale:.3e}"


SINGLE_STEP_FIELDS = [("qpos", "qpos"), ("vel", "vel"), ("acc", "acc"), ("acc_smooth", "acc_smooth"), ("qf_bias", "qf_bias"),
                      ("qf_applied", "qf_applied"), ("qf_passive", "qf_passive"), ("force", "force"), ("mass_mat", "M"),
                      ("mass_Dinv", "Dinv"), ("qf_constraint", "qf_constraint"), ("dofs_pos", "dofs_pos")]
LINK_FIELDS = [("links_pos", 3), ("links_quat", 4), ("root_COM", 3), ("cd_vel", 3), ("cd_ang", 3), ("cacc_lin", 3), ("cacc_ang", 3),
               ("links_contact_force", 3)]


TOL_SOLVER = 5e-4
PRE_SOLVER = {"acc_smooth", "qf_bias", "qf_applied", "qf_passive", "mass_mat", "mass_Dinv"}


def check_single_step(scene, orc):
    s = scene.rigid_solver
    m = s.model
    for gname, oname in SINGLE_STEP_FIELDS:
        a = gpu_field(scene, gname)
        assert_close(a, orc.get(oname, a.shape[1]), TOL_SINGLE if gname in PRE_SOLVER else TOL_SOLVER, gname)
    for name, comps in LINK_FIELDS:
        a = gpu_field(scene, name)
        assert_close(a, orc.get(name, (m.n_links, comps)).reshape(a.shape), TOL_SOLVER, name)
    for name in ("geoms_pos", "aabb_min", "aabb_max") if m.n_geoms else ():
        a = gpu_field(scene, name)
        assert_close(a, orc.get(name, (m.n_geoms, 3)).reshape(a.shape), TOL_SOLVER if name == "geoms_pos" else TOL_SINGLE, name)
    for name in ("cdof_ang", "cdof_vel", "cdofd_ang", "cdofd_vel"):
        a = gpu_field(scene, name)
        assert_close(a, orc.get(name, (m.n_dofs, 3)).reshape(a.shape), TOL_SOLVER, name)


@pytest.mark.parametrize("variant", ["random", "free", "fall"])
def test_franka_single_step_all_quantities(variant):
    scene, _ = make_franka(32, randomized=variant == "random", controlled=variant != "fall")
    orc = mirror_oracle(scene)
    scene.step()
    orc.step(1)
    check_single_step(scene, orc)


def test_go2_single_step_all_quantities():
    scene, _ = make_go2(32)
    orc = mirror_oracle(scene)
    scene.step()
    orc.step(1)
    check_single_step(scene, orc)
    assert contact_set(scene, 32) == contact_set(orc, 32)


def test_franka_controlled_trajectory():
    B, K = 64, 50
    scene, _ = make_franka(B, randomized=True)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_FREE
    assert int(scene.rigid_solver.field("errno_").max()) == 0


def test_franka_fall_trajectory_and_contacts():
    B, K = 32, 150
    scene, _ = make_franka(B, controlled=False)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_CONTACT
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_CONTACT
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert len(b) > 0 and jaccard(a, b) >= JACCARD_MIN


@pytest.mark.parametrize("self_collision", [True, False])
def test_go2_trajectory_and_contacts(self_collision):
    B, K = 64, 100
    scene, _ = make_go2(B, enable_self_collision=self_collision)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    dv = np.abs(gpu_field(scene, "vel") - orc.vel()).max(axis=1)
    # contact-rich locomotion is chaotic: every env must stay within the stated tolerance except at most 2% outliers
    # whose contact set flipped (those are caught by the Jaccard bound below)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.98, f"qpos outliers: {np.sort(dq)[-5:]}"
    assert (dv <= TRAJ_QVEL_CONTACT).mean() >= 0.98, f"qvel outliers: {np.sort(dv)[-5:]}"
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert len(b) >= B and jaccard(a, b) >= JACCARD_MIN
    assert int(scene.rigid_solver.field("errno_").max()) == 0


def test_contact_records_match():
    """Contact position / normal / penetration / force agree record by record (canonical order) for resting Go2."""
    B, K = 16, 60
    scene, robot = make_go2(B, randomized=False)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    con = robot.get_contacts()
    n = scene.rigid_solver.field("n_contacts")[:, 0].cpu().numpy()
    for b in range(B):
        c = orc.contacts(b)
        assert n[b] == len(c["geom_a"]) > 0
        np.testing.assert_array_equal(con["geom_b"][b, : n[b]].cpu().numpy(), c["geom_b"])
        np.testing.assert_allclose(con["position"][b, : n[b]].cpu().numpy(), c["position"], atol=1e-4)
        np.testing.assert_allclose(con["normal"][b, : n[b]].cpu().numpy(), c["normal"], atol=1e-4)
        np.testing.assert_allclose(con["penetration"][b, : n[b]].cpu().numpy(), c["penetration"], atol=1e-4)
        f = c["force"]
        np.testing.assert_allclose(con["force"][b, : n[b]].cpu().numpy(), f, atol=2e-2 * max(1.0, np.abs(f).max()))


@pytest.mark.parametrize("shape", ["box", "sphere", "stack", "cylinder_capsule"])
def test_primitives_on_plane(shape):
    if shape == "cylinder_capsule":  # gs.morphs.Cylinder (mesh-backed support field) and gs.morphs.Capsule (analytic pairs) over a sphere
        morphs = [gs.morphs.Plane(), gs.morphs.Cylinder(radius=0.06, height=0.3, pos=(0, 0, 0.2), quat=(0.9, 0.3, 0.2, 0.1)),
                  gs.morphs.Capsule(radius=0.05, length=0.2, pos=(0.5, 0, 0.2), quat=(0.8, 0.1, 0.55, 0.1)),
                  gs.morphs.Sphere(radius=0.05, pos=(0.5, 0.02, 0.4))]
    elif shape == "box":
        morphs = [gs.morphs.Plane(), gs.morphs.Box(size=(0.2, 0.3, 0.1), pos=(0, 0, 0.3), quat=(0.9, 0.1, 0.3, 0.2))]
    elif shape == "sphere":
        morphs = [gs.morphs.Plane(), gs.morphs.Sphere(radius=0.1, pos=(0, 0, 0.3))]
    else:
        morphs = [gs.morphs.Plane(), gs.morphs.Box(size=(0.2, 0.2, 0.2), pos=(0, 0, 0.1)),
                  gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0.02, 0.01, 0.28))]
    scene, ents = scene_from_morphs(morphs, dt=0.01)
    scene.build(n_envs=4)
    orc = mirror_oracle(scene)
    K = 60 if shape == "cylinder_capsule" else 120  # tumbling round bodies: a shorter horizon keeps fp32 round-off below the tolerance
    for _ in range(K):
        scene.step()
    orc.step(K)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_CONTACT
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_CONTACT
    assert contact_set(scene, 4) == contact_set(orc, 4)
    # the body came to rest on the plane and the contact force carries its weight
    m = scene.rigid_solver.model
    f = gpu_field(scene, "links_contact_force").reshape(4, m.n_links, 3)
    if shape not in ("stack", "cylinder_capsule"):
        assert f[0, 1, 2] == pytest.approx(m.links_inertial_mass[1] * 9.81, rel=5e-3)


def test_franka_grasp_cluster_mpr_contacts():
    """Config 4 (multi-geom MPR contacts): Franka over a cluster of 4 cm boxes, short horizon."""
    B, K = 16, 40
    scene, franka, boxes = make_franka_grasp(B)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"{np.sort(dq)}"
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert len(b) >= 4 * B and jaccard(a, b) >= 0.95


def test_weld_equality():
    """func_equality_weld (constraint/solver.py:984-1173): arm + welded payload, body welded to the world."""
    B, K = 8, 100
    scene, _ = make_weld(B)
    orc = mirror_oracle(scene)
    scene.step()
    orc.step(1)
    check_single_step(scene, orc)
    for _ in range(K - 1):
        scene.step()
    orc.step(K - 1)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_FREE
    assert (scene.rigid_solver.field("n_constraints_equality")[:, 0] == 12).all()


def test_exact_penetration_fallback_epa():
    """Row a17: deeply interpenetrating box / capsule / cylinder-mesh pairs; the EPA fallback after MPR (gsr_epa.cuh, trigger of
    narrowphase.py:726-733) decides the first contacts.  Contact records against the oracle's own implementation, and against
    the oracle with the fallback switched off to show that it is what ran."""
    B = 8
    scene, _ = make_jumble(B)
    orc = mirror_oracle(scene)
    ref = mirror_oracle(scene, gjk_fallback=0)
    scene.step()
    orc.step(1)
    ref.step(1)
    s = scene.rigid_solver
    n = s.field("n_contacts")[:, 0].cpu().numpy()
    assert (n == orc.get("n_contacts", 1)[:, 0]).all() and n.min() >= 4
    for b in range(B):
        c = orc.contacts(b)
        k = int(n[b])
        assert (s.field("contact_geom_a")[b, :k].cpu().numpy() == c["geom_a"]).all()
        np.testing.assert_allclose(s.field("contact_penetration")[b, :k].cpu().numpy(), c["penetration"], atol=1e-4)
        np.testing.assert_allclose(s.field("contact_normal")[b, :k].cpu().numpy(), c["normal"], atol=1e-3)
        np.testing.assert_allclose(s.field("contact_pos")[b, :k].cpu().numpy(), c["position"], atol=1e-4)
    c1, c0 = orc.contacts(0), ref.contacts(0)
    assert len(c1["penetration"]) != len(c0["penetration"]) or np.abs(c1["penetration"] - c0["penetration"]).max() > 1e-4
    for _ in range(19):
        scene.step()
    orc.step(19)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE


def test_dynamic_weld_constraint_api():
    """add_weld_constraint / delete_weld_constraint per environment (rigid_solver.py:2591-2598, solver.py:1293-1390)."""
    B = 6
    scene, ents = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 1.0)),
                                     gs.morphs.Box(size=(0.08, 0.08, 0.08), pos=(0.2, 0.05, 1.1))], dt=0.01)
    scene.build(n_envs=B)
    s = scene.rigid_solver
    v = torch.zeros((B, s.model.n_dofs), device=gs.device)
    v[:, 3:6] = torch.tensor([1.0, -2.0, 0.5])
    v[:, 6] = 0.3
    s.set_dofs_velocity(v)
    orc = mirror_oracle(scene)
    for _ in range(5):
        scene.step()
    orc.step(5)
    welded = [0, 2, 3, 5]
    s.add_weld_constraint(1, 2, envs_idx=welded)
    for b in welded:
        orc.set("add_weld", [1, 2], env=b)
    w = s.get_weld_constraints()
    assert w["n"].tolist() == [1, 0, 1, 1, 0, 1] and int(w["link_a"][0, 0]) == 1 and int(w["link_b"][0, 0]) == 2
    for _ in range(30):
        scene.step()
    orc.step(30)
    assert s.field("n_constraints_equality")[:, 0].tolist() == [6, 0, 6, 6, 0, 6]
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_FREE
    s.delete_weld_constraint(1, 2, envs_idx=[2, 3])
    for b in (2, 3):
        orc.set("clear_welds", [0.0], env=b)
    assert s.get_weld_constraints()["n"].tolist() == [1, 0, 0, 0, 0, 1]
    for _ in range(10):
        scene.step()
    orc.step(10)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE
    with pytest.raises(gs.GenesisException):
        for _ in range(9):
            s.add_weld_constraint(1, 2, envs_idx=[0])


def test_mass_and_com_shift_domain_randomisation():
    """set_mass_shift / set_COM_shift per environment (links_state.mass_shift / i_pos_shift, forward_kinematics.py:377-437)."""
    B, K = 16, 40
    scene, robot = make_go2(B)
    s = scene.rigid_solver
    m = s.model
    g = torch.Generator(device="cpu").manual_seed(3)
    ms = 0.5 * torch.rand((B, m.n_links - 1), generator=g)
    cs = 0.02 * torch.randn((B, m.n_links - 1, 3), generator=g)
    robot.set_mass_shift(ms.to(gs.device))
    robot.set_COM_shift(cs.to(gs.device))
    assert torch.allclose(robot.get_mass_shift().cpu(), ms) and torch.allclose(robot.get_COM_shift().cpu(), cs)
    orc = mirror_oracle(scene)
    full_ms = np.zeros((B, m.n_links)); full_ms[:, 1:] = ms.numpy()
    full_cs = np.zeros((B, m.n_links, 3)); full_cs[:, 1:] = cs.numpy()
    orc.set("mass_shift", full_ms)
    orc.set("com_shift", full_cs.reshape(B, -1))
    scene.step()
    orc.step(1)
    check_single_step(scene, orc)
    for _ in range(K - 1):
        scene.step()
    orc.step(K - 1)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"{np.sort(dq)}"
    # mass matrix of env 0: translational block = total mass including the shifts
    M = s.get_mass_mat()[0].cpu().numpy()
    total = float(np.asarray(m.links_inertial_mass)[1:].sum() + ms[0].sum())
    assert M[0, 0] == pytest.approx(total + float(m.dofs_armature[0]), rel=1e-4)


def test_noslip_post_solver():
    """constraint/noslip.py after the Newton solve (noslip_iterations=5): Go2 trajectories and a box on a tilted-gravity plane."""
    B, K = 16, 50
    scene, _ = make_go2(B, noslip_iterations=5)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"{np.sort(dq)}"
    assert jaccard(contact_set(scene, B), contact_set(orc, B)) >= 0.95
    scene, _ = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 0.05))], dt=0.01,
                                 gravity=(2.0, 0.0, -9.81), noslip_iterations=5)
    scene.build(n_envs=4)
    orc = mirror_oracle(scene)
    for _ in range(200):
        scene.step()
    orc.step(200)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_CONTACT
    assert np.abs(gpu_field(scene, "vel")[:, 0]).max() < 5e-5  # creeps at ~1 mm/s without the pass (tests/test_oracle_analytic.py)


def test_capsule_analytic_pairs():
    """Analytic capsule-capsule / sphere-capsule contacts (collider/capsule_contact.py) and plane-capsule support contacts."""
    B, K = 8, 60
    scene, _ = make_capsules(B)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_CONTACT
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_CONTACT
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert a == b and len(b) >= 3 * B


@pytest.mark.parametrize("opts", [dict(constraint_solver=gs.constraint_solver.CG), dict(integrator=gs.integrator.Euler),
                                  dict(integrator=gs.integrator.implicitfast)], ids=["cg", "euler", "implicitfast"])
def test_solver_and_integrator_variants(opts):
    B, K = 16, 40
    scene, _ = make_go2(B, **opts)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"{np.sort(dq)}"
    assert jaccard(contact_set(scene, B), contact_set(orc, B)) >= 0.95


def test_external_force_and_masked_fk():
    scene, ents = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 2.0))], dt=0.01)
    scene.build(n_envs=3)
    s = scene.rigid_solver
    f = torch.tensor([[[1.0, 0.0, 5.886]]], device=gs.device).repeat(3, 1, 1)  # mass 0.6 kg: weight 5.886 N
    s.apply_links_external_force(f, links_idx=[1])
    scene.step()
    acc = s.field("acc").cpu().numpy()
    np.testing.assert_allclose(acc[:, :3], [[1.0 / 0.6, 0.0, 0.0]] * 3, atol=1e-4)
    # the external force is cleared after the step (rigid_solver.py:2603-2612)
    scene.step()
    np.testing.assert_allclose(s.field("acc").cpu().numpy()[:, 2], -9.81, atol=1e-4)
    # masked set_qpos only touches the selected env
    before = s.get_links_pos().clone()
    s.set_qpos(torch.tensor([[0.5, 0.5, 1.0, 1, 0, 0, 0]], device=gs.device), envs_idx=[1])
    after = s.get_links_pos()
    assert torch.equal(before[0], after[0]) and torch.equal(before[2], after[2])
    np.testing.assert_allclose(after[1, 1].cpu().numpy(), [0.5, 0.5, 1.0], atol=1e-6)


@pytest.mark.parametrize("solver_name,integrator_name", [("Newton", "approximate_implicitfast"), ("CG", "Euler"), ("Newton", "implicitfast")])
def test_box_box_dynamics_reference_expectations(solver_name, integrator_name):
    """The reference's own test_box_box_dynamics (tests/test_rigid_physics.py:1031-1053) on the CUDA path: 64 tilted cubes
    dropped onto cubes come to rest (|qvel| <= 1e-2 after step 100, upper cube at z = 0.6 +- 2e-3), and the first 30 steps
    (free fall, first touch with the exact-penetration fallback, multi-contact settling) follow the fp32 oracle."""
    from helpers import box_box_qpos, make_box_box

    B = 64
    scene, ent = make_box_box(B, constraint_solver=getattr(gs.constraint_solver, solver_name), integrator=getattr(gs.integrator, integrator_name))
    ent.set_qpos(torch.tensor(box_box_qpos(B, seed=2), dtype=torch.float32, device=gs.device))
    orc = mirror_oracle(scene)
    for _ in range(30):
        scene.step()
    orc.step(30)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.95, f"qpos outliers: {np.sort(dq)[-5:]}"
    assert jaccard(contact_set(scene, B), contact_set(orc, B)) >= JACCARD_MIN
    for i in range(30, 110):
        scene.step()
        if i > 100:
            assert float(ent.get_dofs_velocity().abs().max()) <= 1e-2
    q = ent.get_qpos().cpu().numpy()
    np.testing.assert_allclose(q[:, 9], 0.6, atol=2e-3)
    np.testing.assert_allclose(q[:, 2], 0.2, atol=2e-3)


def test_box_box_detection_contact_records_and_stack():
    """RigidOptions.box_box_detection (scope row f4, box_contact.py:100-785): (1) a pile of boxes at arbitrary orientations -- face /
    vertex, face / edge and edge / edge contacts -- gives the oracle's contact records one by one (the oracle itself is pinned to
    the reference's func_box_box_contact by tests/golden/ref_boxbox_*.npz); (2) the reference's stacked cubes follow the oracle through the first touch
    and stay stacked (upper cube at z = 0.6 +- 2e-3)."""
    import ref_cases
    from helpers import box_box_qpos, make_box_box, write_tmp

    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, box_box_detection=True))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(ref_cases.BOXBOX_PILE_XML, ".xml")))
    B = 8
    scene.build(n_envs=B)
    s = scene.rigid_solver
    rng = np.random.default_rng(3)
    q = np.tile(np.asarray(s.model.qpos0, np.float32), (B, 1))
    q[:, 0::7][:, :5] += rng.uniform(-0.01, 0.01, (B, 5)).astype(np.float32)  # jitter x of every box: different contact sets per env
    s.set_qpos(torch.tensor(q, device=gs.device))
    orc = mirror_oracle(scene)
    scene.step()
    orc.step(1)
    con = ent.get_contacts()
    n = s.field("n_contacts")[:, 0].cpu().numpy()
    n_bb = 0
    for b in range(B):
        c = orc.contacts(b)
        assert n[b] == len(c["geom_a"]) > 0
        np.testing.assert_array_equal(con["geom_a"][b, : n[b]].cpu().numpy(), c["geom_a"])
        np.testing.assert_array_equal(con["geom_b"][b, : n[b]].cpu().numpy(), c["geom_b"])
        np.testing.assert_allclose(con["position"][b, : n[b]].cpu().numpy(), c["position"], atol=1e-5)
        np.testing.assert_allclose(con["normal"][b, : n[b]].cpu().numpy(), c["normal"], atol=1e-5)
        np.testing.assert_allclose(con["penetration"][b, : n[b]].cpu().numpy(), c["penetration"], atol=1e-5)
        n_bb += int((np.asarray(c["geom_a"]) > 0).sum())
    assert n_bb >= 4 * B  # box-box contacts, not only plane contacts
    assert int(s.field("errno_").max()) == 0

    B = 32
    scene, ent = make_box_box(B, box_box_detection=True)
    ent.set_qpos(torch.tensor(box_box_qpos(B, seed=5), dtype=torch.float32, device=gs.device))
    orc = mirror_oracle(scene)
    for _ in range(25):
        scene.step()
    orc.step(25)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.95, f"qpos outliers: {np.sort(dq)[-5:]}"
    # the stack holds: with five of the eight face-face candidates kept per pair the reference's algorithm leaves the upper cube
    # rocking at ~0.1 rad/s for good (fp32 oracle: max |qvel| 0.10-0.16 between steps 100 and 400), so no rest criterion here
    for i in range(25, 200):
        scene.step()
    q = ent.get_qpos().cpu().numpy()
    np.testing.assert_allclose(q[:, 9], 0.6, atol=2e-3)
    np.testing.assert_allclose(q[:, 2], 0.2, atol=2e-3)
    assert float(ent.get_dofs_velocity().abs().max()) <= 0.5


def test_gjk_first_collision_mode():
    """RigidOptions.use_gjk_collision=True (scope row a17, the `gjk=` argument of the reference's benchmark factories): every convex
    pair through the GJK distance query + EPA instead of MPR.  The grasp cluster (fingertip boxes / pads against cubes, a stacked
    pair) follows the oracle, whose GJK-first path is pinned by tests/golden/ref_jumble_gjk.npz and ref_box_box_gjk.npz."""
    B, K = 16, 30
    scene, franka, boxes = make_franka_grasp(B, use_gjk_collision=True)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"qpos outliers: {np.sort(dq)[-5:]}"
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert len(b) >= 4 * B and jaccard(a, b) >= 0.95, (len(a), len(b), jaccard(a, b))
    assert int(scene.rigid_solver.field("errno_").max()) == 0


def test_go2_on_height_field_terrain():
    """Height-field terrain (scope row f4, narrowphase.py:346-517): Go2 robots with randomised joints dropped on a bumpy height field
    follow the oracle (itself pinned to the reference's func_contact_mpr_terrain by tests/golden/ref_terrain_primitives.npz): same
    contact sets, trajectories within the contact tolerances, nobody falls through the ground."""
    import ref_cases

    B, K = 32, 80
    hf = ref_cases.terrain_height_field(rows=24, cols=24, seed=5)
    scene, robot = make_go2(B, seed=9, terrain=(hf, 0.25, 0.004, (-3.0, -3.0, 0.0)))
    s = scene.rigid_solver
    q = s.field("qpos").clone()
    g = torch.Generator(device="cpu").manual_seed(1)
    q[:, 0:2] += (1.5 * (torch.rand((B, 2), generator=g) - 0.5)).to(gs.device)   # different cells under every robot
    q[:, 2] += 0.25                                                            # above the highest bump
    s.set_qpos(q)
    orc = mirror_oracle(scene)
    for _ in range(K):
        scene.step()
    orc.step(K)
    dq = np.abs(gpu_field(scene, "qpos") - orc.qpos()).max(axis=1)
    assert (dq <= TRAJ_QPOS_CONTACT).mean() >= 0.9, f"qpos outliers: {np.sort(dq)[-5:]}"
    a, b = contact_set(scene, B), contact_set(orc, B)
    assert len(b) >= 2 * B and jaccard(a, b) >= 0.95, (len(a), len(b), jaccard(a, b))
    assert int(s.field("errno_").max()) == 0
    z = gpu_field(scene, "qpos")[:, 2]
    assert (z > 0.05).all() and (z < 1.0).all()  # the trunk rests on the ground (heights 0.05 .. 0.35 m), it has not fallen through


def test_double_pendulum_link_accelerations():
    """The reference's double pendulum (tests/test_rigid_physics.py:230-297, 955-1020: massless arms, fixed links kept, CG +
    Euler) on the CUDA path: a 50-step swing and the link accelerations (`get_links_acc_ang`, `cacc_lin`) against the oracle,
    which satisfies the reference's closed-form assertions (tests/test_oracle_analytic.py)."""
    from helpers import make_double_pendulum

    B = 16
    scene, ent = make_double_pendulum(B, constraint_solver=gs.constraint_solver.CG, integrator=gs.integrator.Euler)
    rng = np.random.default_rng(1)
    ent.set_qpos(torch.tensor(rng.random((B, 2)), dtype=torch.float32, device=gs.device))
    ent.set_dofs_velocity(torch.tensor(rng.random((B, 2)), dtype=torch.float32, device=gs.device))
    orc = mirror_oracle(scene)
    for _ in range(50):
        scene.step()
    orc.step(50)
    assert np.abs(gpu_field(scene, "qpos") - orc.qpos()).max() <= TRAJ_QPOS_FREE
    assert np.abs(gpu_field(scene, "vel") - orc.vel()).max() <= TRAJ_QVEL_FREE
    nL = scene.rigid_solver.model.n_links
    assert nL == 5
    for name in ("cacc_lin", "cacc_ang"):
        a = gpu_field(scene, name)
        assert_close(a, orc.get(name, (nL, 3)).reshape(a.shape), 5e-4, name)
    acc_ang = scene.rigid_solver.get_links_acc_ang().cpu().numpy()
    acc = gpu_field(scene, "acc")
    np.testing.assert_allclose(acc_ang[:, 1, 0], acc[:, 0], atol=1e-4)  # reference: acc_ang[1] = (theta_ddot_0, 0, 0)
    np.testing.assert_allclose(acc_ang[:, -1, 0], acc[:, 0] + acc[:, 1], atol=1e-4)


@pytest.mark.parametrize("model", ["franka", "go2", "grasp", "franka_cg_euler"])
def test_workspace_layouts_give_identical_steps(model, monkeypatch):
    """The shared-memory workspace has two layouts: with cfg_diagnostics = 1 (what most parity tests run) the smooth force terms and
    D^-1 stay live until the epilogue; with the default cfg_diagnostics = 0 they live in the phase scratch that later phases recycle
    (gsr_host.h).  Both must produce bit-identical steps, so the parity established with diagnostics on carries over."""
    def make():
        if model == "go2":
            return make_go2(64, seed=4)[0]
        if model == "grasp":
            return make_franka_grasp(32)[0]
        if model == "franka_cg_euler":
            return make_franka(64, randomized=True, seed=2, constraint_solver=gs.constraint_solver.CG, integrator=gs.integrator.Euler)[0]
        return make_franka(128, randomized=True, seed=2)[0]

    monkeypatch.setenv("GSR_NO_SPEC", "1")  # both in the generic library: a model-specialised copy inlines differently (other FMA contractions)
    scene_a = make()
    monkeypatch.setenv("GSR_DIAGNOSTICS", "0")
    scene_b = make()
    sa, sb = scene_a.rigid_solver, scene_b.rigid_solver
    assert sa.tuning["diagnostics"] == 1 and sb.tuning["diagnostics"] == 0
    assert sb.launch_info["smem_bytes_per_env"] <= sa.launch_info["smem_bytes_per_env"]
    for _ in range(30):
        scene_a.step(); scene_b.step()
    for k in ("qpos", "vel", "acc", "force", "qf_applied", "qf_constraint", "n_contacts", "contact_pos", "links_pos", "links_contact_force", "cacc_lin"):
        assert torch.equal(sa.field(k), sb.field(k)), k


def test_lazy_diagnostics_match_per_step_diagnostics(monkeypatch):
    """cfg_diagnostics = 0 (the default): get_mass_mat / geoms poses / qf_* are recomputed on demand for the current state by
    k_diag.  They must equal what a cfg_diagnostics = 1 scene writes during the step that starts from the same state."""
    monkeypatch.setenv("GSR_NO_SPEC", "1")        # one binary for both scenes (bit-exact comparison)
    scene_a, fa = make_go2(16, seed=3)            # diagnostics on (fixture)
    monkeypatch.setenv("GSR_DIAGNOSTICS", "0")
    scene_b, fb = make_go2(16, seed=3)
    sa, sb = scene_a.rigid_solver, scene_b.rigid_solver
    assert sa.tuning["diagnostics"] == 1 and sb.tuning["diagnostics"] == 0
    for _ in range(5):
        scene_a.step(); scene_b.step()
    assert torch.equal(sa.field("qpos"), sb.field("qpos"))
    Mb = sb.get_mass_mat()                                   # lazy: smooth dynamics of the current state
    gp, gq = sb.get_geoms_pos(), sb.get_geoms_quat()
    lazy = {k: sb.field(k).clone() for k in ("qf_bias", "qf_passive", "acc_smooth", "mass_Dinv")}
    assert torch.equal(gp, sa.field("geoms_pos")) and torch.equal(gq, sa.field("geoms_quat"))  # both post-integration poses
    scene_a.step()                                           # eager: written by the step that starts from that state
    assert torch.equal(Mb, sa.get_mass_mat())
    for k, v in lazy.items():
        assert torch.equal(v, sa.field(k)), k
    assert sb.kernel_launches()["k_diag"] >= 1 and sa.kernel_launches()["k_diag"] == 0
