"""Host-side model compiler (genesis_b200/model) against the facts SURVEY.md section 8 derives from the reference assets."""

import numpy as np
import pytest

from helpers import gs, host_model, make_franka, make_franka_grasp, make_go2


def test_franka_counts_and_actuators():
    scene, solver, m = host_model(make_franka)
    # panda.xml: plane + link0..7 + hand + 2 fingers; 7 hinge + 2 slide; 1 joint equality
    assert (m.n_links, m.n_dofs, m.n_qs, m.n_geoms, m.n_equalities) == (12, 9, 9, 24, 1)
    assert m.n_pairs == 22  # plane x every non-fixed geom (link0 is fixed-vs-fixed), self-collision off
    assert list(m.joints_type[:7]) == [gs.JOINT_TYPE.REVOLUTE] * 7 and list(m.joints_type[7:9]) == [gs.JOINT_TYPE.PRISMATIC] * 2
    np.testing.assert_allclose(m.dofs_armature, 0.1)
    np.testing.assert_allclose(m.dofs_damping, 1.0)
    bias = np.asarray(m.dofs_act_bias).reshape(-1, 3)
    # panda.xml:264-277 general actuators: gainprm[0] = kp, biasprm = (0, -kp, -kd); fingers share the tendon actuator
    np.testing.assert_allclose(m.dofs_kp[:7], [4500, 4500, 3500, 3500, 2000, 2000, 2000])
    np.testing.assert_allclose(bias[:7, 1], -np.asarray(m.dofs_kp[:7]))
    np.testing.assert_allclose(bias[:7, 2], [-450, -450, -350, -350, -200, -200, -200])
    np.testing.assert_allclose(m.dofs_kp[7:], 0.01568627451, rtol=1e-6)
    np.testing.assert_allclose(bias[7:, 1:], [[-100, -10]] * 2)
    assert m.eq_type[0] == gs.EQUALITY_TYPE.JOINT
    # timeconst >= 2 dt after sanitising (rigid_solver.py:187-218)
    assert (np.asarray(m.geoms_sol_params)[:, 0] >= 2 * m.dt - 1e-12).all()
    assert m.meaninertia > 0 and np.isfinite(m.links_invweight).all() and (m.dofs_invweight > 0).all()


def test_go2_counts():
    scene, solver, m = host_model(make_go2)
    # go2.urdf with merged fixed links: base + 4 x (hip, thigh, calf) + plane; free base + 12 hinges; 27 + 1 geoms
    assert (m.n_links, m.n_dofs, m.n_qs, m.n_geoms) == (14, 18, 19, 28)
    assert m.joints_type[0] == gs.JOINT_TYPE.FREE
    np.testing.assert_allclose(m.dofs_kp[6:], 100.0)  # default URDF gains (utils/geom.py:2116-2121)
    np.testing.assert_allclose(np.asarray(m.dofs_act_bias).reshape(-1, 3)[6:, 2], -10.0)
    np.testing.assert_allclose(m.dofs_armature[6:], 0.1)
    types = list(m.geoms_type)
    assert types.count(gs.GEOM_TYPE.SPHERE) == 5 and types.count(gs.GEOM_TYPE.BOX) == 5 and types.count(gs.GEOM_TYPE.CYLINDER) == 17
    total = m.links_inertial_mass[1:].sum()
    assert 14.0 < total < 16.5  # Go2 weighs ~15 kg
    # self-collision disabled => only the 27 plane pairs remain
    _, _, m2 = host_model(make_go2, enable_self_collision=False)
    assert m2.n_pairs == 27


def test_pair_table_is_consistent():
    for make in (make_franka, make_go2, make_franka_grasp):
        _, _, m = host_model(make)
        pairs = np.asarray(m.pairs).reshape(-1, 2)
        assert (pairs[:, 0] < pairs[:, 1]).all()
        for p, (a, b) in enumerate(pairs):
            assert m.pair_idx[a, b] == p
            assert m.geoms_link[a] != m.geoms_link[b]
            assert not (m.geoms_is_fixed[a] and m.geoms_is_fixed[b])
        assert (m.pair_idx >= 0).sum() == len(pairs)
        assert m.max_contact_pairs == min(150, m.n_pairs) * 5


def test_support_field_matches_brute_force():
    """support_field.py:22-91: each cell of the 180x180 table is the argmax vertex for that direction."""
    _, _, m = host_model(make_go2)
    g = int(np.nonzero(np.asarray(m.geoms_type) == gs.GEOM_TYPE.CYLINDER)[0][0])
    v = np.asarray(m.verts)[m.geoms_vert_start[g]: m.geoms_vert_end[g]]
    s = int(m.geoms_support_start[g])
    rng = np.random.default_rng(0)
    for _ in range(50):
        i, j = rng.integers(0, 180, 2)
        theta, phi = i / 180 * 2 * np.pi - np.pi, j / 180 * np.pi
        d = np.array([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)])
        cell = np.asarray(m.support_v)[s + i * 180 + j]
        assert cell @ d == pytest.approx((v @ d).max(), abs=1e-12)


def test_unknown_option_is_rejected():
    with pytest.raises(TypeError):
        gs.options.RigidOptions(not_an_option=1)
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(use_hibernation=True))
    scene.add_entity(gs.morphs.Plane())
    with pytest.raises(gs.GenesisException):
        scene.rigid_solver.build_model()


def test_mjcf_inertia_inferred_from_geoms_like_mujoco():
    """Bodies without <inertial>: mass / COM from the geoms at MuJoCo's default density 1000 kg/m^3, per-geom `density` and
    `mass` attributes honoured (the reference reads body_mass / body_ipos from the MuJoCo compiler, utils/mjcf.py:240-243);
    the material density of the entity plays no role for them."""
    from helpers import scene_from_morphs, write_tmp

    xml = """
    <mujoco><worldbody>
      <body name="a" pos="0 0 1"><freejoint/><geom type="box" size="0.1 0.2 0.3"/></body>
      <body name="b" pos="1 0 1"><freejoint/><geom type="sphere" size="0.1" density="200"/><geom type="box" size="0.1 0.1 0.1" pos="0.5 0 0" mass="2.0"/></body>
      <body name="c" pos="2 0 1"><freejoint/><inertial pos="0 0 0" mass="3.0" diaginertia="0.1 0.1 0.1"/><geom type="sphere" size="0.1"/></body>
    </worldbody></mujoco>"""
    scene, _ = scene_from_morphs([gs.morphs.MJCF(file=write_tmp(xml, ".xml"))], dt=0.01)
    m = scene.rigid_solver.build_model()
    mass = np.asarray(m.links_inertial_mass)
    names = list(scene.rigid_solver._entities[0]._info.link_names)
    ia, ib, ic = names.index("a"), names.index("b"), names.index("c")
    assert mass[ia] == pytest.approx(1000 * 0.2 * 0.4 * 0.6, rel=1e-9)
    m_sphere = 200 * 4.0 / 3.0 * np.pi * 0.1**3
    assert mass[ib] == pytest.approx(m_sphere + 2.0, rel=1e-9)
    ipos = np.asarray(m.links_inertial_pos).reshape(-1, 3)
    assert ipos[ib, 0] == pytest.approx(2.0 * 0.5 / (m_sphere + 2.0), rel=1e-9)
    assert mass[ic] == pytest.approx(3.0, rel=1e-9)


def test_mjcf_morph_pose_offsets_the_real_base_links():
    """rigid_entity.py:623-651 then 744-767: the virtual MJCF 'world' link is dropped BEFORE morph.pos / morph.quat offset the base
    links, so a fixed-base arm placed at (1, 2, 0.5) has its link0 there, and a free body's init_qpos carries the offset."""
    from helpers import BOX_BOX_XML, write_tmp

    scene = gs.Scene()
    scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml", pos=(1.0, 2.0, 0.5), quat=(0.70710678, 0.0, 0.0, 0.70710678)))
    m = scene.rigid_solver.build_model()
    np.testing.assert_allclose(m.links_pos[0], [1.0, 2.0, 0.5])
    np.testing.assert_allclose(m.links_quat[0], [0.70710678, 0, 0, 0.70710678], atol=1e-7)
    np.testing.assert_allclose(m.links_pos[1], [0.0, 0.0, 0.333])  # children stay relative to their parent
    scene = gs.Scene()
    scene.add_entity(gs.morphs.MJCF(file=write_tmp(BOX_BOX_XML, ".xml"), pos=(1.0, 0.0, 0.5)))
    m = scene.rigid_solver.build_model()
    np.testing.assert_allclose(m.qpos0[:3], [1.0, 0.0, 0.7])
    np.testing.assert_allclose(m.qpos0[7:10], [1.0, 0.0, 1.3])
