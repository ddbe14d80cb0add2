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


def make_go2(n_envs, randomized=True, build=True, seed=0, terrain=None, **rigid_kwargs):
    """Config 3: plane + go2.urdf, joints uniform within limits, PD targets (test_rigid_benchmarks.py:240-256).
    terrain = (height_field, horizontal_scale, vertical_scale, pos) puts the robot on a height field instead of the plane."""
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    if terrain is None:
        scene.add_entity(gs.morphs.Plane())
    else:
        scene.add_entity(gs.morphs.Terrain(height_field=terrain[0], horizontal_scale=terrain[1], vertical_scale=terrain[2], pos=terrain[3]))
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


# joint angles that put the hand at (0.65, 0, 0.135) / (0.65, 0, 0.30) pointing down (quat (0, 1, 0, 0)): the inverse-kinematics targets of
# the reference's test_contact_forces (tests/test_rigid_physics.py:2160-2186), solved once offline with the model's own FK
GRASP_Q_DOWN = [-0.5161, 0.8579, 0.5578, -1.7541, -0.6506, 2.4193, 1.1756]
GRASP_Q_LIFT = [-0.6480, 0.5776, 0.7431, -1.7814, -0.4656, 2.1753, 1.0716]
GRASP_BOXES = [(0.65, 0.0, 0.02),      # the cube between the fingers
               (0.74, 0.0, 0.02), (0.74, 0.0, 0.06),  # a stacked pair behind it: box-box contacts
               (0.74, 0.05, 0.02)]     # and a neighbour on the plane


def make_franka_grasp(n_envs, n_boxes=4, build=True, close_fingers=True, **rigid_kwargs):
    """Config 4: plane + panda + a cluster of 4 cm boxes near (0.65, 0, 0.02); the hand hovers over the first box with its fingers
    straddling it and closes them (pattern of tests/test_rigid_physics.py:2123-2191: hold, grasp): finger-box contacts through the
    convex-convex narrow phase (fingertip boxes and pads against the cube), box-box contacts in the stacked pair, plane-box corners."""
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    scene.add_entity(gs.morphs.Plane())
    franka = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml"))
    boxes = []
    for i in range(n_boxes):
        boxes.append(scene.add_entity(gs.morphs.Box(size=(0.04, 0.04, 0.04), pos=GRASP_BOXES[i % len(GRASP_BOXES)])))
    if not build:
        return scene, franka, boxes
    scene.build(n_envs=n_envs)
    q = torch.tensor(GRASP_Q_DOWN + [0.04, 0.04], dtype=torch.float32, device=gs.device)
    tgt = torch.tensor(GRASP_Q_DOWN + ([0.0, 0.0] if close_fingers else [0.04, 0.04]), dtype=torch.float32, device=gs.device)
    if n_envs > 0:
        q, tgt = q.repeat(n_envs, 1), tgt.repeat(n_envs, 1)
    franka.set_qpos(q)
    franka.control_dofs_position(tgt)
    return scene, franka, boxes
