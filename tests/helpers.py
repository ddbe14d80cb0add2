"""Shared helpers of the test-suite: host-only model construction and oracle wrappers."""

import os
import tempfile

import numpy as np

import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka, make_franka_grasp, make_go2
from oracle.pyoracle import Oracle

gs.init(backend=gs.gpu, seed=0)


def host_model(make, **kw):
    """Compile the model of a benchmark scene on the host (no CUDA needed)."""
    out = make(0, build=False, **kw)
    scene = out[0]
    solver = scene.rigid_solver
    model = solver.build_model()
    return scene, solver, model


def oracle_for(solver, model, n_envs=1, double=False, threads=1, **opt_overrides):
    opts = solver.model_options()
    opts.update(opt_overrides)
    return Oracle(model, n_envs, double=double, options=opts, threads=threads)


def scene_from_morphs(morphs, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(**rigid_kwargs))
    ents = [scene.add_entity(m) for m in morphs]
    return scene, ents


def write_tmp(text, suffix):
    fd, path = tempfile.mkstemp(suffix=suffix)
    with os.fdopen(fd, "w") as f:
        f.write(text)
    return path


PENDULUM_XML = """
<mujoco model="pendulum">
  <worldbody>
    <body name="arm" pos="0 0 1">
      <joint name="hinge" type="hinge" axis="0 1 0" pos="0 0 0"/>
      <inertial pos="0 0 -0.5" mass="1.0" diaginertia="0.01 0.01 0.01"/>
    </body>
  </worldbody>
</mujoco>
"""

# two free capsules crossing over each other and a sphere on top: analytic capsule-capsule / sphere-capsule pairs
# (collider/capsule_contact.py) plus plane-capsule support contacts
CAPSULES_XML = """
<mujoco model="capsules">
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <body name="ca" pos="0 0 0.06" euler="0 90 0">
      <freejoint/>
      <geom type="capsule" size="0.05 0.15"/>
    </body>
    <body name="cb" pos="0.02 0 0.17" euler="90 0 0">
      <freejoint/>
      <geom type="capsule" size="0.04 0.12"/>
    </body>
    <body name="sp" pos="0.1 0.01 0.15">
      <freejoint/>
      <geom type="sphere" size="0.04"/>
    </body>
  </worldbody>
</mujoco>
"""


def make_capsules(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(CAPSULES_XML, ".xml")))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent

# a free box welded (pose-locked) to the last link of a two-link arm, plus a box welded to the world:
# func_equality_weld (constraint/solver.py:984-1173), 6 rows per weld
WELD_XML = """
<mujoco model="weld">
  <option gravity="0 0 -9.81"/>
  <worldbody>
    <body name="l1" pos="0 0 1.0">
      <joint name="j1" type="hinge" axis="0 1 0" damping="0.5"/>
      <inertial pos="0.25 0 0" mass="1.0" diaginertia="0.01 0.02 0.02"/>
      <body name="l2" pos="0.5 0 0">
        <joint name="j2" type="hinge" axis="0 1 0" damping="0.5"/>
        <inertial pos="0.2 0 0" mass="0.5" diaginertia="0.005 0.01 0.01"/>
      </body>
    </body>
    <body name="payload" pos="0.95 0.05 1.0" euler="0 0 30">
      <freejoint/>
      <inertial pos="0 0 0" mass="0.3" diaginertia="0.002 0.002 0.002"/>
    </body>
    <body name="anchored" pos="0 0.5 0.5">
      <freejoint/>
      <inertial pos="0 0 0" mass="0.4" diaginertia="0.003 0.003 0.003"/>
    </body>
  </worldbody>
  <equality>
    <weld body1="l2" body2="payload" anchor="0.02 0 0"/>
    <weld body1="anchored"/>
  </equality>
</mujoco>
"""


def make_weld(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(WELD_XML, ".xml")))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent

# deeply interpenetrating convex pairs (box-box, box-capsule, box-cylinder mesh): the exact-penetration fallback after MPR
# (narrowphase.py:726-733, gjk.py / epa.py) decides the first contacts
JUMBLE_XML = """
<mujoco model="jumble">
  <option gravity="0 0 -9.81"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <body name="a" pos="0 0 0.3"><freejoint/><geom type="box" size="0.05 0.05 0.05"/></body>
    <body name="b" pos="0.03 0.06 0.32" euler="10 25 40"><freejoint/><geom type="box" size="0.03 0.06 0.02"/></body>
    <body name="c" pos="-0.04 -0.03 0.36" euler="70 -20 15"><freejoint/><geom type="capsule" size="0.03 0.08"/></body>
    <body name="d" pos="0.02 -0.05 0.24" euler="-35 50 5"><freejoint/><geom type="cylinder" size="0.03 0.05"/></body>
  </worldbody>
</mujoco>
"""


def make_jumble(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(JUMBLE_XML, ".xml")))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent

# the reference's `box_box` fixture (tests/test_rigid_physics.py:82-97): two 0.4 m cubes over a plane
BOX_BOX_XML = """
<mujoco model="one_box">
  <option timestep="0.01"/>
  <default><geom contype="1" conaffinity="1" condim="3" friction="1. 0.5 0.5"/></default>
  <worldbody>
    <geom type="plane" name="floor" pos="0. 0. 0." size="40. 40. 40."/>
    <body name="box1" pos="0. 0. 0.2"><geom type="box" size="0.2 0.2 0.2" pos="0. 0. 0."/><joint name="root1" type="free"/></body>
    <body name="box2" pos="0. 0. 0.8"><geom type="box" size="0.2 0.2 0.2" pos="0. 0. 0."/><joint name="root2" type="free"/></body>
  </worldbody>
</mujoco>
"""


def make_box_box(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(BOX_BOX_XML, ".xml")))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent


def box_box_qpos(n, seed=0):
    """Initial states of the reference's test_box_box_dynamics (tests/test_rigid_physics.py:1038-1045): the lower cube at
    rest, the upper one dropped from 0.65-0.75 m with a small tilt and an arbitrary yaw."""
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 14))
    q[:, :7] = (0, 0, 0.2, 1, 0, 0, 0)
    for b in range(n):
        rx, ry, rz = 0.15 * rng.random(), 0.15 * rng.random(), np.pi * rng.random()
        # intrinsic x-y-z Euler angles -> quaternion (w, x, y, z), as gu.xyz_to_quat
        cx, sx, cy, sy, cz, sz = np.cos(rx / 2), np.sin(rx / 2), np.cos(ry / 2), np.sin(ry / 2), np.cos(rz / 2), np.sin(rz / 2)
        qx, qy, qz_ = np.array([cx, sx, 0, 0]), np.array([cy, 0, sy, 0]), np.array([cz, 0, 0, sz])

        def mul(a, c):
            return np.array([a[0] * c[0] - a[1] * c[1] - a[2] * c[2] - a[3] * c[3], a[0] * c[1] + a[1] * c[0] + a[2] * c[3] - a[3] * c[2],
                             a[0] * c[2] - a[1] * c[3] + a[2] * c[0] + a[3] * c[1], a[0] * c[3] + a[1] * c[2] - a[2] * c[1] + a[3] * c[0]])
        q[b, 7:10] = (0, 0, 0.65 + 0.1 * rng.random())
        q[b, 10:14] = mul(mul(qx, qy), qz_)
    return q


def multi_pendulum_urdf(n=2, joint_damping=0.0, joint_friction=0.0):
    """The reference's `_build_multi_pendulum` fixture (tests/test_rigid_physics.py:230-292): n massless 1 m arms hinged about x,
    each carrying a 1 kg point mass at its tip."""
    out = ['<robot name="multi_pendulum">', '<link name="base"/>']
    parent = "base"
    for i in range(n):
        out.append(f'<joint name="PendulumJoint_{i}" type="continuous"><origin xyz="0 0 0" rpy="0 0 0"/><axis xyz="1 0 0"/>'
                   f'<parent link="{parent}"/><child link="PendulumArm_{i}"/><limit effort="{100.0 * (n - i)}" velocity="30.0"/>'
                   f'<dynamics damping="{joint_damping}" friction="{joint_friction}"/></joint>')
        out.append(f'<link name="PendulumArm_{i}"><inertial><origin xyz="0 0 0" rpy="0 0 0"/><mass value="0.0"/>'
                   '<inertia ixx="0.0" ixy="0.0" ixz="0.0" iyy="0.0" iyz="0.0" izz="0.0"/></inertial></link>')
        out.append(f'<joint name="PendulumMassJoint_{i}" type="fixed"><origin xyz="0 0 1.0" rpy="0 0 0"/>'
                   f'<parent link="PendulumArm_{i}"/><child link="PendulumMass_{i}"/></joint>')
        out.append(f'<link name="PendulumMass_{i}"><inertial><origin xyz="0 0 0" rpy="0 0 0"/><mass value="1.0"/>'
                   '<inertia ixx="1e-12" ixy="0.0" ixz="0.0" iyy="1e-12" iyz="0.0" izz="1e-12"/></inertial></link>')
        parent = f"PendulumMass_{i}"
    out.append("</robot>")
    return "\n".join(out)


def make_double_pendulum(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.URDF(file=write_tmp(multi_pendulum_urdf(2), ".urdf"), merge_fixed_links=False, fixed=True))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent

# the reference's `two_aligned_hinges` fixture (tests/test_rigid_physics.py:170-182)
TWO_HINGES_XML = """
<mujoco model="two_aligned_hinges">
  <option timestep="0.05"/>
  <default><geom contype="1" conaffinity="1" condim="3"/></default>
  <worldbody>
    <body name="body0">
      <geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.05"/>
      <joint type="hinge" name="joint0" axis="0 0 1"/>
      <body name="body1" pos="0.5 0 0">
        <geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.05"/>
        <joint type="hinge" name="joint1" axis="0 0 1"/>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


def make_two_hinges(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.05, **rigid_kwargs))
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(TWO_HINGES_XML, ".xml")))
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent

# the reference's `hinge_slide` fixture (tests/test_rigid_physics.py:351-365): joint friction loss, damping and a spring
HINGE_SLIDE_XML = """
<mujoco model="hinge_slide">
  <default><joint damping="0.01"/></default>
  <worldbody>
    <body name="pendulum" pos="0.15 0.0 0.0">
      <joint name="hinge" type="hinge" axis="0 1 0" frictionloss="0.08"/>
      <geom name="geom1" type="capsule" size="0.02" fromto="0.0 0.0 0.0 0.1 0.0 0.0"/>
      <body name="link1" pos="0.1 0.0 0.0">
        <joint name="slide" type="slide" axis="1 0 0" frictionloss="0.3" stiffness="200.0"/>
        <geom name="geom2" type="capsule" size="0.015" fromto="-0.1 0.0 0.0 0.1 0.0 0.0"/>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


def make_hinge_slide(n_envs, build=True, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.002, **rigid_kwargs))  # MuJoCo's default timestep
    ent = scene.add_entity(gs.morphs.MJCF(file=write_tmp(HINGE_SLIDE_XML, ".xml"), default_armature=None))  # as tests/utils.py:596
    if build:
        scene.build(n_envs=n_envs)
    return scene, ent


import torch  # noqa: E402


def make_franka_cluster_r1(n_envs, n_boxes=4, build=True, **rigid_kwargs):
    """Round-1 box cluster (arm hovering over four boxes, fingers open): kept for the frozen lane-kernel emulation tests."""
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
