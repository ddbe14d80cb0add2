"""Drop-in API surface: argument conventions and return shapes of the reference's accessors
(rigid_entity.py:1641-1720, 3372-3650, 3919-4013; kinematic_solver.py:940-1003)."""

import numpy as np
import pytest
import torch

from helpers import gs, make_franka, make_go2

pytestmark = pytest.mark.gpu


def test_batched_accessor_shapes_and_semantics():
    B = 8
    scene, franka = make_franka(B)
    assert franka.n_dofs == 9 and franka.n_qs == 9
    q = franka.get_qpos()
    assert q.shape == (B, 9) and q.dtype == torch.float32 and q.device.type == "cuda"
    # dofs_state.pos of hinge/slide joints is qpos - qpos0 (forward_kinematics.py:689-725); panda qpos0 == 0
    assert torch.allclose(franka.get_dofs_position(), q)
    assert franka.get_dofs_velocity().shape == (B, 9)
    assert franka.get_links_pos().shape == (B, franka.n_links, 3)
    assert franka.get_links_quat().shape == (B, franka.n_links, 4)
    assert franka.get_dofs_position(dofs_idx_local=[0, 3], envs_idx=[1, 2, 5]).shape == (3, 2)
    # getters return copies
    q[:] = 123.0
    assert not torch.equal(franka.get_qpos(), q)
    # masked control write
    franka.control_dofs_position(torch.full((2, 2), 0.3, device=gs.device), dofs_idx_local=[1, 2], envs_idx=[0, 4])
    cp = scene.rigid_solver.field("ctrl_pos")
    assert torch.allclose(cp[[0, 4]][:, [1, 2]], torch.full((2, 2), 0.3, device=gs.device))
    assert float(cp[1, 1]) == 0.0
    lo, hi = franka.get_dofs_limit()
    assert lo.shape == (9,) and (lo < hi).all()


def test_unbatched_scene_shapes():
    scene, franka = make_franka(0)
    scene.step()
    assert franka.get_qpos().shape == (9,)
    assert franka.get_links_pos().shape == (franka.n_links, 3)
    assert scene.rigid_solver.get_mass_mat().shape == (9, 9)


def test_reset_restores_initial_state_and_clears_warmstart():
    B = 16
    scene, robot = make_go2(B, randomized=False)
    s = scene.rigid_solver
    q0 = s.field("qpos").clone()
    for _ in range(30):
        scene.step()
    assert not torch.equal(s.field("qpos"), q0)
    robot.set_qpos(q0)
    robot.set_dofs_velocity(torch.zeros(B, 18, device=gs.device))
    assert torch.equal(s.field("qpos"), q0)
    assert int(s.field("is_warmstart").max()) == 0 and float(s.field("qacc_ws").abs().max()) == 0.0
    # after an identical reset the trajectory repeats bit-exactly
    ref = []
    for _ in range(10):
        scene.step()
    ref = s.field("qpos").clone()
    robot.set_qpos(q0)
    robot.set_dofs_velocity(torch.zeros(B, 18, device=gs.device))
    for _ in range(10):
        scene.step()
    assert torch.equal(s.field("qpos"), ref)


def test_get_contacts_dictionary():
    B = 8
    scene, robot = make_go2(B, randomized=False)
    for _ in range(60):
        scene.step()
    c = robot.get_contacts()
    n = int(scene.rigid_solver.n_contacts.max())
    assert n >= 4
    for k in ("link_a", "link_b", "geom_a", "geom_b", "penetration"):
        assert c[k].shape == (B, n)
    for k in ("position", "normal", "force"):
        assert c[k].shape == (B, n, 3)
    assert (c["geom_a"][:, :4] == 0).all()  # plane is geom 0 and type-ordered first
    assert torch.allclose(c["normal"][:, 0].norm(dim=-1), torch.ones(B, device=gs.device), atol=1e-5)


def test_errno_contact_overflow_is_reported():
    from helpers import scene_from_morphs

    morphs = [gs.morphs.Plane(), gs.morphs.Box(size=(0.2, 0.2, 0.2), pos=(0, 0, 0.1)), gs.morphs.Box(size=(0.2, 0.2, 0.2), pos=(0.5, 0, 0.1))]
    scene, _ = scene_from_morphs(morphs, dt=0.01, max_collision_pairs=1)  # 5 contact slots, two resting boxes need 8
    scene.build(n_envs=4)
    with pytest.raises(gs.GenesisException, match="contact pairs|candidate"):
        for _ in range(30):
            scene.step()


def test_links_acc_pendulum_closed_form():
    """tests/test_rigid_physics.py:903-958 (test_pendulum_links_acc): a pendulum released horizontally; get_links_acc returns
    the classical acceleration of the link frame origin (zero at the hinge) and get_links_acc_ang the angular one."""
    from helpers import PENDULUM_XML, scene_from_morphs, write_tmp

    scene, (pend,) = scene_from_morphs([gs.morphs.MJCF(file=write_tmp(PENDULUM_XML, ".xml"))], dt=0.01)
    scene.build(n_envs=2)
    pend.set_dofs_position(torch.full((2, 1), np.pi / 2, device=gs.device))
    scene.step()
    s = scene.rigid_solver
    # point mass-like rod: I = 0.01 + m l^2 about the hinge (+ armature), torque = m g l at horizontal
    m_, l_, I_ = 1.0, 0.5, 0.01
    arm = float(s.model.dofs_armature[0])
    alpha = m_ * 9.81 * l_ / (I_ + m_ * l_ * l_ + arm)
    acc_ang = pend.get_links_acc_ang()
    assert acc_ang.shape == (2, s.model.n_links, 3)
    assert abs(abs(float(acc_ang[0, -1, 1])) - alpha) < 2e-2 * alpha
    acc = pend.get_links_acc()
    assert acc.shape == (2, s.model.n_links, 3)
    # the link frame origin sits on the hinge axis; like the reference, the getter mixes the pre-integration spatial
    # acceleration with the post-integration pose / velocity, which leaves an O(dt) residual
    assert float(acc[0, -1].norm()) < 2e-2


def test_external_torque_frictionloss_and_friction_ratio():
    from helpers import scene_from_morphs

    scene, ents = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 1.0))], dt=0.01)
    scene.build(n_envs=3)
    s, box = scene.rigid_solver, ents[1]
    inertia = 0.6 * (0.1**2 + 0.1**2) / 12.0
    box.apply_links_external_torque(torch.tensor([[[0.0, 0.0, 1e-3]]], device=gs.device).repeat(3, 1, 1))
    scene.step()
    acc = s.field("acc").cpu().numpy()
    np.testing.assert_allclose(acc[:, 5], 1e-3 / inertia, rtol=1e-3)
    np.testing.assert_allclose(acc[:, 2], -9.81, atol=1e-4)
    # friction ratio scales the contact friction of the env it is set for (contact.py:202-211)
    # (the contact takes the larger of the two scaled coefficients, so both geoms of the pair are scaled)
    box.set_friction_ratio(torch.tensor([[0.5]], device=gs.device), envs_idx=[1])
    ents[0].set_friction_ratio(torch.tensor([[0.5]], device=gs.device), envs_idx=[1])
    for _ in range(80):
        scene.step()
    mu = s.field("contact_friction")[:, 0].cpu().numpy()
    assert mu[1] == pytest.approx(0.5 * mu[0]) and mu[2] == pytest.approx(mu[0])
    # friction loss adds one row per dof with a positive value (solver.py:1236-1286)
    assert int(s.field("n_constraints_frictionloss").max()) == 0
    box.set_dofs_frictionloss(torch.full((6,), 0.01, device=gs.device))
    scene.step()
    assert (s.field("n_constraints_frictionloss")[:, 0] == 6).all()
    assert torch.allclose(box.get_dofs_frictionloss(), torch.full((6,), 0.01, device=gs.device))


def test_collider_and_constraint_solver_facades():
    """`scene.rigid_solver.collider.get_contacts(as_tensor, to_torch, keep_batch_dim)` and the ConstraintSolver attributes the
    reference's tests read (collider.py:810-913, constraint/solver.py:27-112)."""
    B = 4
    scene, robot = make_go2(B, randomized=False)
    for _ in range(60):
        scene.step()
    s = scene.rigid_solver
    c = s.collider.get_contacts(as_tensor=True, to_torch=True)
    n = s.collider.n_contacts
    assert c["geom_a"].shape == (B, int(n.max()))
    ragged = s.collider.get_contacts(as_tensor=False, to_torch=False)
    assert isinstance(ragged["position"], tuple) and len(ragged["position"]) == B
    for b in range(B):
        assert ragged["position"][b].shape == (int(n[b]), 3) and isinstance(ragged["position"][b], np.ndarray)
    cs = s.constraint_solver
    assert cs.qacc_ws.shape == (B, 18)
    assert (cs.n_constraints >= 4 * n).all() and (cs.n_constraints_equality == 0).all()
    assert (cs.n_iterations >= 1).all()


def test_link_velocity_reference_expectations():
    """The reference's test_link_velocity (tests/test_rigid_physics.py:826-896, `two_aligned_hinges`) through the accessors:
    set_qpos / set_dofs_velocity refresh the kinematics, get_links_pos / get_links_vel / get_links_ang / get_links_root_COM
    return what the reference asserts (5e-5, its fp32 tolerance)."""
    from helpers import make_two_hinges

    B, tol = 2, 5e-5
    scene, ent = make_two_hinges(B, constraint_solver=gs.constraint_solver.CG, integrator=gs.integrator.Euler)
    s = scene.rigid_solver

    def state(q, v):
        ent.set_qpos(torch.tensor([q] * B, dtype=torch.float32, device=gs.device))
        ent.set_dofs_velocity(torch.tensor([v] * B, dtype=torch.float32, device=gs.device))
        return [x[0].cpu().numpy().astype(np.float64) for x in (s.field("cd_vel"), s.get_links_ang(), s.get_links_root_COM(), s.get_links_pos(), s.get_links_vel())]

    cvel = state([0.0, 0.0], [0.0, 1.0])[0]
    np.testing.assert_allclose(cvel, 0, atol=tol)
    cvel = state([0.0, 0.0], [1.0, 0.0])[0]
    np.testing.assert_allclose(cvel, [[0.0, 0.5, 0.0]] * 2, atol=tol)
    cvel, _, com, pos, _ = state([0.0, np.pi / 2.0], [0.0, 1.2])
    np.testing.assert_allclose(com[0], [0.375, 0.125, 0.0], atol=tol)
    np.testing.assert_allclose(pos[1], [0.5, 0.0, 0.0], atol=tol)
    np.testing.assert_allclose(cvel[1], [-1.2 * 0.125, 1.2 * (0.375 - 0.5), 0.0], atol=tol)
    t0, t1 = -0.7, 0.2
    cvel, cang, com, pos, xvel = state([t0, t1], [3.0, 13.0])
    xanchor, COM = pos[1], com[0]
    np.testing.assert_allclose(xanchor, [0.5 * np.cos(t0), 0.5 * np.sin(t0), 0.0], atol=tol)
    com0 = np.array([0.25 * np.cos(t0), 0.25 * np.sin(t0), 0.0])
    com1 = np.array([0.5 * np.cos(t0) + 0.25 * np.cos(t0 + t1), 0.5 * np.sin(t0) + 0.25 * np.sin(t0 + t1), 0.0])
    np.testing.assert_allclose(COM, 0.5 * (com0 + com1), atol=tol)
    w0, w1 = cang[0, 2], cang[1, 2]
    assert abs(w0 - 3.0) <= tol and abs(w1 - w0 - 13.0) <= 2e-4  # 16.0 in fp32
    np.testing.assert_allclose(cvel[0], w0 * np.array([-COM[1], COM[0], 0.0]), atol=2e-4)
    np.testing.assert_allclose(cvel[1], cvel[0] + (w1 - w0) * np.array([xanchor[1] - COM[1], COM[0] - xanchor[0], 0.0]), atol=2e-4)
    np.testing.assert_allclose(xvel[0], 0, atol=2e-4)
    np.testing.assert_allclose(xvel[1], w0 * np.array([-xanchor[1], xanchor[0], 0.0]), atol=2e-4)


def test_external_force_reference_frames():
    """apply_links_external_force(ref=..., local=...) builds the c-frame wrench of abd/misc.py:692-712: the default reference is
    the LINK ORIGIN (torque arm = links_pos - root_COM), 'link_com' uses the link's COM, 'root_com' no arm at all; local=True
    rotates the force by the link's inertial-frame orientation.  Checked on a Franka link whose COM is offset from its origin."""
    from helpers import make_franka

    scene, franka = make_franka(2)
    s = scene.rigid_solver
    l = 5
    F = torch.tensor([1.0, -2.0, 0.5], device=gs.device)
    pos, root = s.field("links_pos")[:, l], s.field("root_COM")[:, l]
    com = s._get_links_com(torch.tensor([l], device=gs.device), None)[:, 0]
    assert (com - pos).norm(dim=-1).min() > 1e-3  # the COM really is offset
    for ref, arm in (("link_origin", pos - root), ("link_com", com - root), ("root_com", torch.zeros_like(pos))):
        s.field("cfrc_applied_vel").zero_(); s.field("cfrc_applied_ang").zero_()
        s.apply_links_external_force(F, links_idx=[l], ref=ref)
        assert torch.allclose(s.field("cfrc_applied_vel")[:, l], -F.expand(2, 3))
        assert torch.allclose(s.field("cfrc_applied_ang")[:, l], -torch.cross(arm, F.expand(2, 3), dim=-1), atol=1e-6), ref
    # default == link_origin (the reference's default), and the entity forwarder takes the same keywords
    s.field("cfrc_applied_ang").zero_()
    franka.apply_links_external_force(F, links_idx_local=[l - 1])
    assert torch.allclose(s.field("cfrc_applied_ang")[:, l], -torch.cross(pos - root, F.expand(2, 3), dim=-1), atol=1e-6)
    with pytest.raises(ValueError):
        s.apply_links_external_force(F, links_idx=[l], ref="root_com", local=True)


def test_set_pos_on_fixed_base_moves_the_base_not_the_joints():
    """RigidEntity.set_pos / set_quat on a fixed-base arm move the (fixed) root link and leave the joint coordinates alone; on a
    free body they write qpos[0:7] (rigid_entity.py set_pos -> set_base_links_pos)."""
    from helpers import make_franka, scene_from_morphs

    scene, franka = make_franka(2)
    s = scene.rigid_solver
    q_before = franka.get_qpos().clone()
    hand_before = s.get_links_pos()[:, -3].clone()
    franka.set_pos(torch.tensor([0.5, -0.25, 0.1], device=gs.device))
    assert torch.equal(franka.get_qpos(), q_before)
    base = s.get_links_pos()[:, franka._info.link_start]
    assert torch.allclose(base, torch.tensor([0.5, -0.25, 0.1], device=gs.device).expand(2, 3))
    assert torch.allclose(s.get_links_pos()[:, -3] - hand_before, torch.tensor([0.5, -0.25, 0.1], device=gs.device).expand(2, 3), atol=1e-6)
    scene2, (plane, box) = scene_from_morphs([gs.morphs.Plane(), gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0, 0, 1.0))], dt=0.01)
    scene2.build(n_envs=2)
    box.set_pos(torch.tensor([1.0, 2.0, 3.0], device=gs.device))
    assert torch.allclose(box.get_qpos()[:, :3], torch.tensor([1.0, 2.0, 3.0], device=gs.device).expand(2, 3))


def _accessor_step(scene, franka, state0, mask, qpos0, vel0, stiff, damp):
    """The body of the reference's `franka_accessors` benchmark step (tests/test_rigid_benchmarks.py:377-394), verbatim."""
    scene.step()
    franka.get_ang()
    franka.get_vel()
    franka.get_dofs_position()
    franka.get_dofs_velocity()
    franka.get_links_pos()
    franka.get_links_quat()
    franka.get_links_vel()
    franka.get_contacts()
    scene.rigid_solver.set_state(0, state0, envs_idx=mask, partial=True)
    franka.control_dofs_position(qpos0)
    franka.set_dofs_stiffness(stiff)
    franka.set_dofs_damping(damp)
    franka.set_dofs_velocity(vel0, envs_idx=mask, skip_forward=True)
    franka.set_qpos(qpos0, envs_idx=mask, zero_velocity=False, skip_forward=True)


def test_franka_accessors_reference_step_body_runs_and_resets(monkeypatch):
    """RL-loop boundary (scope row f3): the reference's franka_accessors step body runs unmodified -- RigidSolverState from
    get_state(), set_state(f, state, envs_idx=<bool mask>, partial=True), masked setters with skip_forward=True (the forward
    kinematics of the reset environments are deferred to the next step) -- and the reset environments then replay the trajectory
    of the initial state bit for bit, while the others are untouched."""
    from helpers import make_franka

    # bit-for-bit replay: both scenes keep the reference's eager kinematics (with a model-specialised library copy the lazy and
    # the eager order run two inlined copies of the kinematics, equal to round-off only: test_kinematics_policy_defaults)
    monkeypatch.setenv("GSR_LAZY_FK", "0")
    B = 64
    scene, franka = make_franka(B, randomized=True, seed=5)
    s = scene.rigid_solver
    qpos0 = torch.tensor([0, 0, 0, -1.0, 0, 1.0, 0, 0.02, 0.02], device=gs.device).repeat(B, 1)
    vel0 = s.field("vel").clone()
    state0 = s.get_state()
    assert type(state0).__name__ == "RigidSolverState"
    for k in ("qpos", "dofs_vel", "dofs_acc", "links_pos", "links_quat", "i_pos_shift", "mass_shift", "friction_ratio"):
        assert getattr(state0, k).shape[0] == B
    assert torch.equal(state0.qpos, s.field("qpos")) and torch.equal(state0.links_quat, s.field("links_quat"))
    mask = torch.zeros(B, dtype=torch.bool, device=gs.device)
    mask[::7] = True
    stiff, damp = franka.get_dofs_stiffness(), franka.get_dofs_damping()
    # reference trajectory of the initial state: 6 plain steps of an identical scene
    scene_ref, _ = make_franka(B, randomized=True, seed=5)
    traj = []
    for _ in range(6):
        scene_ref.step()
        traj.append(scene_ref.rigid_solver.field("qpos").clone())
    launches0 = s.kernel_launches()
    for _ in range(4):
        _accessor_step(scene, franka, state0, mask, qpos0, vel0, stiff, damp)
    launches = s.kernel_launches()
    assert launches["k_set_state"] - launches0["k_set_state"] == 4       # one fused launch per masked reset
    assert launches["k_gather_fields"] - launches0["k_gather_fields"] == 4  # the eight getters share one gather per step
    # skip_forward / partial: no eager FK per setter -- one deferred, masked k_forward right before the next step instead
    assert 3 <= launches["k_forward"] - launches0["k_forward"] <= 4
    # the masked envs were reset after step 4: two more steps replay steps 1-2 of the initial state; the others are at step 6
    scene.step(); scene.step()
    q = s.field("qpos")
    assert torch.equal(q[mask], traj[1][mask]), "reset environments do not replay the initial trajectory"
    assert torch.equal(q[~mask], traj[5][~mask]), "environments outside the mask were disturbed"
    assert int(s.field("errno_").max()) == 0


def test_set_state_full_and_skip_forward_semantics():
    from helpers import make_go2

    B = 8
    scene, robot = make_go2(B, seed=1)
    s = scene.rigid_solver
    st = s.get_state()
    for _ in range(10):
        scene.step()
    after10 = s.field("qpos").clone()
    s.set_state(0, st, envs_idx=[1, 3])               # partial=False: contacts cleared, FK run for the two envs
    assert int(s.field("n_contacts")[1, 0]) == 0 and int(s.field("n_contacts")[0, 0]) >= 0
    assert torch.equal(s.field("qpos")[1], st.qpos[1]) and torch.equal(s.field("qpos")[0], after10[0])
    assert torch.allclose(s.field("links_pos")[1, 1], st.qpos[1, :3])  # base link pose == free-joint qpos: FK was run
    assert int(s.field("ctrl_mode")[1].min()) == int(gs.CTRL_MODE.FORCE) and int(s.field("fk_dirty")[1, 0]) == 0
    # skip_forward=True defers the kinematics to the next step and gives the same step result as the eager path
    scene_b, robot_b = make_go2(B, seed=1)
    q = st.qpos.clone(); q[:, 2] += 0.05
    robot.set_qpos(q, skip_forward=True); robot_b.set_qpos(q)
    assert int(s.field("fk_dirty").min()) == 1
    s.field("ctrl_mode").copy_(scene_b.rigid_solver.field("ctrl_mode")); s.field("vel").copy_(scene_b.rigid_solver.field("vel"))
    scene.step(); scene_b.step()
    assert torch.equal(s.field("qpos"), scene_b.rigid_solver.field("qpos"))
    assert int(s.field("fk_dirty").max()) == 0


def test_batch_dofs_info_per_environment_gains():
    """RigidOptions(batch_dofs_info=True) (array_class.py:1347-1364): kp / kv / armature per environment.  Env 0 gets the gains of
    the reference's test_position_control (tests/test_rigid_physics.py:1417-1440); it must step exactly like an unbatched scene
    built with those gains, env 1 like the default scene."""
    from helpers import make_franka

    KP = torch.tensor([4500.0, 4500.0, 3500.0, 3500.0, 2000.0, 2000.0, 2000.0, 100.0, 100.0], device=gs.device)
    KD = torch.tensor([450.0, 450.0, 350.0, 350.0, 200.0, 200.0, 200.0, 10.0, 10.0], device=gs.device)
    tgt = torch.tensor([0.69, -0.11, -0.72, -2.73, -0.15, 2.64, 0.89, 0.04, 0.04], device=gs.device)
    scene, franka = make_franka(2, batch_dofs_info=True, disable_constraint=True)
    franka.set_dofs_kp(KP, envs_idx=[0]); franka.set_dofs_kv(KD, envs_idx=[0])
    franka.control_dofs_position(tgt)
    ref_a, fa = make_franka(1, disable_constraint=True); fa.set_dofs_kp(KP); fa.set_dofs_kv(KD); fa.control_dofs_position(tgt)
    ref_b, fb = make_franka(1, disable_constraint=True); fb.control_dofs_position(tgt)
    assert scene.rigid_solver.dofs_info.armature.to_numpy().shape == (9, 2)
    for _ in range(20):
        scene.step(); ref_a.step(); ref_b.step()
    q = scene.rigid_solver.field("qpos")
    assert torch.equal(q[0], ref_a.rigid_solver.field("qpos")[0])
    assert torch.equal(q[1], ref_b.rigid_solver.field("qpos")[0])
    assert not torch.equal(q[0], q[1])
    assert torch.allclose(franka.get_dofs_kp(envs_idx=[0])[0], KP) and int(scene.rigid_solver.field("n_constraints").max()) == 0
