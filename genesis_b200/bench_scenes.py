"""Scene factories of the benchmark configs, mirroring the reference's
tests/test_rigid_benchmarks.py make_franka (:320-396) and make_go2 (:219-261)."""

import numpy as np
import torch

import genesis_b200 as gs

FRANKA_Q0 = [0.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0, 0.02, 0.02]
GO2_Q0 = [0.0, 0.0, 0.42, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5]
GO2_CTRL = [0.0, 0.0, 0.0, 0.0, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5]


def make_franka(n_envs, self_collision=False, randomized=False, controlled=True, build=True, seed=0, **rigid_kwargs):
    """Config 2 ("franka_free" / "franka_random"): plane + panda.xml, position control at q0."""
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, enable_self_collision=self_collision, **rigid_kwargs))
    scene.add_entity(gs.morphs.Plane())
    franka = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml"))
    if not build:
        return scene, franka
    scene.build(n_envs=n_envs)
    q0 = torch.tensor(FRANKA_Q0, dtype=torch.float32, device=gs.device)
    if n_envs > 0:
        q0 = q0.repeat(n_envs, 1)
    franka.set_qpos(q0)
    if controlled:
        franka.control_dofs_position(q0)
    if randomized and n_envs > 0:
        g = torch.Generator(device="cpu").manual_seed(seed)
        vel0 = 0.2 * torch.clip(torch.randn((n_envs, 9), generator=g), -1.0, 1.0)
        vel0[:, 7:] = 0.0
        franka.set_dofs_velocity(vel0.to(gs.device))
    return scene, franka


def make_go2(n_envs, randomized=True, build=True, seed=0, **rigid_kwargs):
    """Config 3: plane + go2.urdf, joints uniform within limits, PD targets (test_rigid_benchmarks.py:240-256)."""
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    scene.add_entity(gs.morphs.Plane())
    robot = scene.add_entity(gs.morphs.URDF(file="urdf/go2/urdf/go2.urdf"))
    if not build:
        return scene, robot
    scene.build(n_envs=n_envs)
    robot.control_dofs_position(torch.tensor(GO2_CTRL, dtype=torch.float32, device=gs.device), dofs_idx_local=slice(6, None))
    B = max(1, n_envs)
    q = torch.tensor([GO2_Q0], dtype=torch.float32).repeat(B, 1)
    if randomized:
        lo, hi = robot.get_dofs_limit()
        lo, hi = lo.cpu()[6:], hi.cpu()[6:]
        g = torch.Generator(device="cpu").manual_seed(seed)
        q[:, 7:] = lo + (hi - lo) * torch.rand((B, 12), generator=g)
    robot.set_qpos(q.to(gs.device) if n_envs > 0 else q[0].to(gs.device))
    return scene, robot


def make_franka_grasp(n_envs, n_boxes=4, build=True, **rigid_kwargs):
    """Config 4: plane + panda + a cluster of 4 cm boxes near (0.65, 0, 0.02) (pattern of tests/test_rigid_physics.py:2123-2191)."""
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    scene.add_entity(gs.morphs.Plane())
    franka = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml"))
    boxes = []
    for i in range(n_boxes):
        dx, dy = (i % 2) * 0.06 - 0.03, (i // 2) * 0.06 - 0.03
        boxes.append(scene.add_entity(gs.morphs.Box(size=(0.04, 0.04, 0.04), pos=(0.65 + dx, dy, 0.02))))
    if not build:
        return scene, franka, boxes
    scene.build(n_envs=n_envs)
    q = torch.tensor([-0.0, 0.45, 0.0, -2.2, 0.0, 2.65, 0.78, 0.04, 0.04], dtype=torch.float32, device=gs.device)
    if n_envs > 0:
        q = q.repeat(n_envs, 1)
    franka.set_qpos(q)
    franka.control_dofs_position(q)
    return scene, franka, boxes
