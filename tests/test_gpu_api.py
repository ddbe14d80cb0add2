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
