"""Expectations of the reference's own test-suite that need neither MuJoCo nor external assets, replayed on the CUDA path with the
reference's scenes, step counts and tolerances (file:line of the original next to each test; /root/reference/tests/test_rigid_physics.py
unless noted).  They complement the fixture-based parity tests: these are the behaviours a Genesis user would check first."""
import math

import numpy as np
import pytest
import torch

from helpers import gs, write_tmp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("integrator", [gs.integrator.Euler, gs.integrator.approximate_implicitfast])
def test_energy_analytical_and_conservation(integrator):
    """:5719-5795 -- two spheres dropped from 0.5 m at dt = 1 ms: kinetic / potential energy follow semi-implicit Euler exactly during
    the free fall; with nearly undamped contact parameters the bouncing sphere keeps its energy within 1 % over 400 steps, the
    default (damped) one loses energy."""
    g, dt, h0, radius, n_steps = 9.81, 0.001, 0.5, 0.1, 400
    undamped = [10.0, 0.001, 0.9, 0.95, 0.001, 0.5, 2.0]
    scene = gs.Scene(sim_options=gs.options.SimOptions(dt=dt, gravity=(0, 0, -g)), rigid_options=gs.options.RigidOptions(integrator=integrator))
    plane = scene.add_entity(gs.morphs.Plane())
    sphere_a = scene.add_entity(gs.morphs.Sphere(radius=radius, pos=(0, 0, h0)))
    sphere_b = scene.add_entity(gs.morphs.Sphere(radius=radius, pos=(0.5, 0, h0)))
    scene.build()
    plane.geoms[0].set_sol_params(undamped)
    sphere_a.geoms[0].set_sol_params(undamped)
    s = scene.rigid_solver
    mass = float(sphere_a.get_links_inertial_mass()[0])
    inertia = 0.4 * mass * radius * radius

    def energies(ent):
        v = ent.get_dofs_velocity().double()
        ke = 0.5 * mass * float(v[:3] @ v[:3]) + 0.5 * inertia * float(v[3:] @ v[3:])
        return ke, mass * g * float(ent.get_pos()[2])

    te_initial = mass * g * h0
    ke_a, pe_a, ke_b, pe_b = [], [], [], []
    impact_step = -1
    for i in range(n_steps):
        scene.step()
        ka, pa = energies(sphere_a); kb, pb = energies(sphere_b)
        ke_a.append(ka); pe_a.append(pa); ke_b.append(kb); pe_b.append(pb)
        if impact_step < 0 and int(s.field("n_contacts").sum()) > 0:
            impact_step = i
    assert impact_step > 0
    for i in range(impact_step):
        n = i + 1
        expected_ke = 0.5 * mass * (n * g * dt) ** 2
        expected_pe = mass * g * (h0 - g * dt**2 * n * (n + 1) / 2)
        # fp32 state: 2e-5 relative on the energies (the reference's `tol` is 5e-5 in fp32)
        assert ke_a[i] == pytest.approx(expected_ke, rel=5e-5, abs=5e-6) and pe_a[i] == pytest.approx(expected_pe, rel=5e-5)
        assert ke_b[i] == pytest.approx(expected_ke, rel=5e-5, abs=5e-6) and pe_b[i] == pytest.approx(expected_pe, rel=5e-5)
    assert ke_a[-1] + pe_a[-1] == pytest.approx(te_initial, rel=0.01)   # undamped: conserved after bouncing
    assert ke_b[-1] + pe_b[-1] < te_initial                             # damped: strictly decreased


def test_contype_conaffinity():
    """:4269-4337 -- four stacked boxes with contype / conaffinity 3, 2, 1, 0: box1 rests on the plane, box2 on box1, box3 falls
    through box2 (no common bit) onto box1 next to it in height, box4 collides with nothing and stays in free fall."""
    gravity = (0.0, 0.0, -10.0)
    scene = gs.Scene(sim_options=gs.options.SimOptions(gravity=gravity))
    scene.add_entity(gs.morphs.Plane(pos=(0.0, 0.0, 0.0)))
    boxes = [scene.add_entity(gs.morphs.Box(size=(0.5, 0.5, 0.5), pos=(0.0, 0.0, 0.5 * (k + 1)), contype=c, conaffinity=c))
             for k, c in enumerate((3, 2, 1, 0))]
    scene.build()
    for _ in range(80):
        scene.step()
    box1, box2, box3, box4 = boxes
    np.testing.assert_allclose(box1.get_pos().cpu().numpy(), (0.0, 0.0, 0.25), atol=5e-4)
    np.testing.assert_allclose(box2.get_pos().cpu().numpy(), (0.0, 0.0, 0.75), atol=2e-3)
    np.testing.assert_allclose(box2.get_pos().cpu().numpy(), box3.get_pos().cpu().numpy(), atol=2e-3)
    acc = scene.rigid_solver.get_links_acc(slice(box4.link_start, box4.link_start + box4.n_links))
    np.testing.assert_allclose(acc.cpu().numpy().reshape(3), gravity, atol=5e-5)


ELLIPSOID_XML = """
<mujoco model="ellipsoid">
  <worldbody>
    <body name="obj" pos="0 0 0.0">
      <joint name="root" type="free"/>
      <geom type="ellipsoid" size="0.05 0.05 0.02"/>
    </body>
  </worldbody>
</mujoco>
"""


def test_ellipsoid():
    """:4570-4610 -- an ellipsoid (semi-axes 5 x 5 x 2 cm) dropped from 0.2 m spinning at up to 20 rad/s: its AABB matches the
    semi-axes, it comes to rest flat on the plane (|qvel| <= 5e-3, bottom of the AABB within 5 mm below the plane, roll = pitch = 0)."""
    scene = gs.Scene(sim_options=gs.options.SimOptions(dt=0.02))
    scene.add_entity(gs.morphs.Plane())
    entity = scene.add_entity(gs.morphs.MJCF(file=write_tmp(ELLIPSOID_XML, ".xml"), pos=(0, 0, 0.2)))
    scene.build()
    rng = np.random.default_rng(0)
    entity.set_dofs_velocity(torch.tensor(20 * rng.random(3), dtype=torch.float32, device=gs.device), dofs_idx_local=slice(3, 6))
    entity.set_dofs_kv(0.002, dofs_idx_local=slice(3, 6))
    entity.control_dofs_velocity(0.0, dofs_idx_local=slice(3, 6))
    aabb = entity.get_AABB().cpu().numpy()
    np.testing.assert_allclose(aabb[1] - aabb[0], (0.10, 0.10, 0.04), atol=1e-3)
    for _ in range(100):
        scene.step()
    np.testing.assert_allclose(entity.get_dofs_velocity().cpu().numpy(), 0, atol=5e-3)
    assert -0.005 < float(entity.get_AABB()[0, 2]) < 0.0
    w, x, y, z = (float(v) for v in entity.get_quat())
    roll = math.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = math.asin(max(-1.0, min(1.0, 2 * (w * y - z * x))))
    assert abs(roll) <= 5e-3 and abs(pitch) <= 5e-3


def test_set_sol_params_sanitised():
    """:1661-1692 (the geom part) -- set_sol_params clamps: time constant >= 2 dt, impedances in [1e-4, 1 - 1e-4], width >= 0, power >= 1."""
    scene = gs.Scene(sim_options=gs.options.SimOptions(dt=0.01))
    robot = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml", pos=(0.0, 0.4, 0.1)))
    scene.build(n_envs=2)
    for geom in robot.geoms[:3]:
        sp = geom.sol_params + 1.0
        geom.set_sol_params(sp)
        assert not torch.allclose(geom.sol_params, sp)  # clamped
        geom.set_sol_params([0.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0])
        np.testing.assert_allclose(geom.sol_params.cpu().numpy(), [2.0e-02, 0.5, 1e-4, 1e-4, 0.0, 1e-4, 1.0], rtol=1e-6)
    scene.step()
    assert int(scene.rigid_solver.field("errno_").max()) == 0
