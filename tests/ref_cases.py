"""Scenes and inputs of the reference-generated fixtures (tests/golden/ref_*.npz).

Shared by tools/make_ref_golden.py (which EXECUTES THE REFERENCE'S OWN KERNEL SOURCES on these inputs through
oracle/qd_shim and records what they compute) and by tests/test_ref_golden.py / tests/test_gpu_ref_golden.py (which replay the
same inputs through the fp64 oracle and through the CUDA path).  Each case returns (solver, model, init, K, options) where
`init` maps oracle / accessor field names to (B, n) arrays and `options` are RigidOptions keyword overrides.
"""

import numpy as np

import genesis_b200 as gs
from helpers import (BOX_BOX_XML, CAPSULES_XML, HINGE_SLIDE_XML, JUMBLE_XML, PENDULUM_XML, TWO_HINGES_XML, WELD_XML, box_box_qpos, host_model,
                     make_franka, make_go2, multi_pendulum_urdf, write_tmp)
from genesis_b200.bench_scenes import FRANKA_Q0, GO2_CTRL, GO2_Q0

POSITION, VELOCITY, FORCE = 0, 1, 2

# a three-link chain hanging from a BALL joint, its tip tied to the world by a CONNECT equality and its second link to a
# free box by a second one: func_equality_connect (constraint/solver.py:881-937), spherical joints in FK / integration
CONNECT_XML = """
<mujoco model="connect">
  <option gravity="0 0 -9.81"/>
  <worldbody>
    <body name="a" pos="0 0 1.0">
      <joint name="ball" type="ball" damping="0.05"/>
      <inertial pos="0.15 0 0" mass="0.8" diaginertia="0.004 0.01 0.01"/>
      <body name="b" pos="0.3 0 0">
        <joint name="h1" type="hinge" axis="0 1 0" damping="0.02"/>
        <inertial pos="0.15 0 0" mass="0.5" diaginertia="0.002 0.006 0.006"/>
        <body name="c" pos="0.3 0 0">
          <joint name="h2" type="hinge" axis="0 0 1" damping="0.02"/>
          <inertial pos="0.1 0 0" mass="0.3" diaginertia="0.001 0.003 0.003"/>
        </body>
      </body>
    </body>
    <body name="payload" pos="0.3 0.1 0.8">
      <freejoint/>
      <inertial pos="0 0 0" mass="0.4" diaginertia="0.002 0.002 0.002"/>
    </body>
  </worldbody>
  <equality>
    <connect body1="c" anchor="0.2 0 0"/>
    <connect body1="b" body2="payload" anchor="0.05 0.1 -0.2"/>
  </equality>
</mujoco>
"""


def _scene(xml=None, urdf=None, morphs=None, dt=0.01, urdf_kw=None, mjcf_kw=None, **opts):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=dt, **opts))
    if xml is not None:
        scene.add_entity(gs.morphs.MJCF(file=write_tmp(xml, ".xml"), **(mjcf_kw or {})))
    if urdf is not None:
        scene.add_entity(gs.morphs.URDF(file=write_tmp(urdf, ".urdf"), **(urdf_kw or {})))
    for mo in morphs or ():
        scene.add_entity(mo)
    solver = scene.rigid_solver
    return solver, solver.build_model()


def _rng(seed):
    return np.random.default_rng(seed)


def case_pendulum():
    s, m = _scene(xml=PENDULUM_XML)
    return s, m, dict(qpos=np.array([[0.7], [-2.1]]), vel=np.array([[0.3], [1.5]])), 5, {}


def case_franka_pd():
    _, s, m = host_model(make_franka)
    r = _rng(1)
    B = 2
    q = np.tile(np.array(FRANKA_Q0), (B, 1)) + 0.2 * r.uniform(-1, 1, (B, 9))
    q[:, 7:] = [[0.02, 0.03], [0.01, 0.035]]
    v = 0.5 * r.uniform(-1, 1, (B, 9))
    ctrl = np.tile(np.array(FRANKA_Q0), (B, 1))
    return s, m, dict(qpos=q, vel=v, ctrl_pos=ctrl, ctrl_mode=np.full((B, 9), POSITION)), 4, {}


def case_franka_limits_force():
    """Joint-limit rows (solver.py:1177-1233) and FORCE / VELOCITY control modes: joints driven into their limits."""
    _, s, m = host_model(make_franka)
    lim = np.asarray(m.dofs_limit).reshape(-1, 2)
    q = np.array([FRANKA_Q0], float)
    q[0, 0] = lim[0, 1] + 0.01          # beyond the upper limit
    q[0, 3] = lim[3, 0] - 0.02          # beyond the lower limit
    q[0, 7] = lim[7, 1] + 0.001
    mode = np.full((1, 9), FORCE)
    mode[0, 1] = VELOCITY
    f = np.array([[5.0, 0, -3.0, 2.0, 1.0, 0.5, 0.1, 2.0, -2.0]])
    cv = np.zeros((1, 9)); cv[0, 1] = 0.4
    return s, m, dict(qpos=q, vel=np.zeros((1, 9)), ctrl_force=f, ctrl_vel=cv, ctrl_mode=mode), 4, {}


def case_franka_fall_selfcollision():
    """Uncontrolled arm collapsing onto itself and the plane: mesh-mesh MPR with multi-contact, plane-mesh support contacts."""
    _, s, m = host_model(make_franka, self_collision=True)
    q = np.array([[-1.913, -0.924, 0.657, -1.265, -0.688, 0.017, -0.299, 0.011, 0.04],     # hand meshes touching link 5 (11 contacts)
                  [-1.846, -0.107, 2.852, -3.024, -0.748, 1.243, -0.547, 0.035, 0.018]])   # link 2 against link 6 (1 cm deep)
    return s, m, dict(qpos=q, vel=np.array([[0.0, 2.0, 0, -3.0, 0, 1.0, 0, 0, 0], [0.0] * 9])), 3, dict(enable_self_collision=True)


def case_go2_drop():
    _, s, m = host_model(make_go2)
    r = _rng(2)
    B = 2
    lim = np.asarray(m.dofs_limit).reshape(-1, 2)[6:]
    q = np.tile(np.array(GO2_Q0), (B, 1))
    q[:, 2] = [0.30, 0.27]
    q[1, 3:7] = np.array([0.96, 0.2, -0.15, 0.1]) / np.linalg.norm([0.96, 0.2, -0.15, 0.1])
    q[:, 7:] = lim[:, 0] + (lim[:, 1] - lim[:, 0]) * (0.3 + 0.4 * r.random((B, 12)))
    ctrl = np.zeros((B, 18)); ctrl[:, 6:] = GO2_CTRL
    mode = np.full((B, 18), FORCE); mode[:, 6:] = POSITION
    v = np.zeros((B, 18)); v[:, 2] = -0.5; v[1, 3:6] = [0.5, -0.3, 0.2]
    return s, m, dict(qpos=q, vel=v, ctrl_pos=ctrl, ctrl_mode=mode), 6, {}


def case_primitives_plane():
    """plane-box corners (box_contact.py:25-96), plane-sphere, sphere-box MPR, one contact per sphere pair."""
    s, m = _scene(morphs=[gs.morphs.Plane(), gs.morphs.Box(size=(0.2, 0.3, 0.1), pos=(0, 0, 0.049), quat=(0.98, 0.05, 0.1, 0.15)),
                          gs.morphs.Sphere(radius=0.07, pos=(0.03, 0.02, 0.16)), gs.morphs.Sphere(radius=0.05, pos=(0.6, 0, 0.045))])
    q = np.asarray(m.qpos0, float)[None].copy()
    q[0, 3:7] /= np.linalg.norm(q[0, 3:7])
    return s, m, dict(qpos=q, vel=np.zeros((1, m.n_dofs))), 5, {}


def case_box_box_mpr():
    """The reference's box_box fixture (tests/test_rigid_physics.py:82-97): box-box through MPR (box_box_detection off)."""
    s, m = _scene(xml=BOX_BOX_XML)
    q = box_box_qpos(2, seed=3)
    q[:, 9] = [0.595, 0.59]   # start in contact
    return s, m, dict(qpos=q, vel=np.zeros((2, 12))), 5, {}


def case_capsules():
    """Analytic capsule-capsule / sphere-capsule pairs (capsule_contact.py) and plane-capsule support contacts, started in contact."""
    s, m = _scene(xml=CAPSULES_XML)
    q = np.asarray(m.qpos0, float)[None].copy()
    q[0, 2] = 0.052           # capsule a: its lower end ~3 mm into the plane ...
    q[0, 3:7] = [np.cos(0.76), 0.0, np.sin(0.76), 0.0]  # ... and 3 degrees off the horizontal: a capsule lying exactly flat has a
    # degenerate support point along -z (fp32 and fp64 pick different ends of the axis), which is conditioning, not parity
    q[0, 7 + 2] = 0.13        # capsule b: 8 mm into capsule a
    q[0, 14 + 2] = 0.135      # sphere: 3 mm into capsule a
    return s, m, dict(qpos=q, vel=np.zeros((1, m.n_dofs))), 5, {}


def case_jumble_deep():
    """Deeply interpenetrating convex pairs: MPR reports a deep overlap and the pair is re-run through GJK + EPA
    (narrowphase.py:726-733, gjk.py:170-446, epa.py)."""
    s, m = _scene(xml=JUMBLE_XML)
    return s, m, dict(qpos=np.asarray(m.qpos0, float)[None].copy(), vel=np.zeros((1, m.n_dofs))), 3, {}


def case_weld():
    s, m = _scene(xml=WELD_XML)
    q = np.asarray(m.qpos0, float)[None].copy()
    q[0, 0] = 0.3
    return s, m, dict(qpos=q, vel=0.2 * _rng(5).uniform(-1, 1, (1, m.n_dofs))), 5, {}


def case_connect_ball():
    s, m = _scene(xml=CONNECT_XML)
    q = np.asarray(m.qpos0, float)[None].copy()
    qb = np.array([0.95, 0.1, 0.25, -0.1]); q[0, 0:4] = qb / np.linalg.norm(qb)
    q[0, 4:6] = [0.4, -0.3]
    return s, m, dict(qpos=q, vel=0.5 * _rng(6).uniform(-1, 1, (1, m.n_dofs))), 5, {}


def case_frictionloss():
    """The reference's hinge_slide fixture: friction-loss rows (solver.py:1236-1286), joint spring and damping."""
    s, m = _scene(xml=HINGE_SLIDE_XML, dt=0.002, mjcf_kw=dict(default_armature=None))
    return s, m, dict(qpos=np.array([[0.4, 0.02], [0.0, 0.0]]), vel=np.array([[1.0, -0.5], [1e-4, 0.0]])), 8, {}


def case_two_hinges():
    s, m = _scene(xml=TWO_HINGES_XML, dt=0.05)
    return s, m, dict(qpos=np.array([[0.5, -1.0]]), vel=np.array([[1.0, 2.0]])), 4, {}


def case_double_pendulum():
    s, m = _scene(urdf=multi_pendulum_urdf(2, joint_damping=0.1), urdf_kw=dict(merge_fixed_links=False, fixed=True))
    return s, m, dict(qpos=np.array([[0.9, -0.4]]), vel=np.array([[0.5, 1.5]])), 5, {}


def _box_variant(**opts):
    s, m = _scene(xml=BOX_BOX_XML, **opts)
    q = box_box_qpos(1, seed=4)
    q[:, 9] = 0.597
    return s, m, dict(qpos=q, vel=np.zeros((1, 12))), 4, opts


def case_box_box_cg():
    return _box_variant(constraint_solver=gs.constraint_solver.CG)


def case_box_box_euler():
    return _box_variant(integrator=gs.integrator.Euler)


def case_box_box_implicitfast():
    return _box_variant(integrator=gs.integrator.implicitfast)


def case_box_noslip():
    s, m = _scene(morphs=[gs.morphs.Plane(), gs.morphs.Box(size=(0.2, 0.2, 0.2), pos=(0, 0, 0.0995))], gravity=(1.5, 0.5, -9.81),
                  noslip_iterations=5)
    return s, m, dict(qpos=np.asarray(m.qpos0, float)[None].copy(), vel=np.zeros((1, 6))), 5, dict(noslip_iterations=5)


# box-box pairs through the reference's func_box_box_contact (RigidOptions.box_box_detection, box_contact.py:100-785):
# boxes of different sizes thrown together at arbitrary orientations -- face / vertex, face / edge and edge / edge contacts
BOXBOX_PILE_XML = """
<mujoco model="boxbox_pile">
  <option gravity="0 0 -9.81"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <body name="a" pos="0 0 0.30"><freejoint/><geom type="box" size="0.10 0.08 0.05"/></body>
    <body name="b" pos="0.09 0.05 0.385" euler="7 41 -13"><freejoint/><geom type="box" size="0.06 0.07 0.04"/></body>
    <body name="c" pos="-0.07 -0.04 0.39" euler="40 35 20"><freejoint/><geom type="box" size="0.05 0.05 0.05"/></body>
    <body name="d" pos="0.0 0.11 0.29" euler="52 -9 31"><freejoint/><geom type="box" size="0.03 0.09 0.03"/></body>
    <body name="e" pos="0.02 -0.02 0.21" euler="5 -8 60"><freejoint/><geom type="box" size="0.07 0.07 0.04"/></body>
  </worldbody>
</mujoco>
"""


def case_boxbox_stack():
    """The reference's box_box fixture with box_box_detection=True: the upper cube rests on the lower one face to face (aligned: every
    edge axis degenerate; yawed: eight candidate points, five kept), or touches it with an edge / a corner first (tilted)."""
    s, m = _scene(xml=BOX_BOX_XML, box_box_detection=True)
    q = box_box_qpos(4, seed=7)
    q[0, 10:14] = (1, 0, 0, 0)                                  # aligned
    q[1, 10:14] = (np.cos(0.35), 0, 0, np.sin(0.35))            # yawed by 0.7 rad, flat
    q[:, 9] = [0.598, 0.597, 0.615, 0.62]                       # all four start in contact
    return s, m, dict(qpos=q, vel=np.zeros((4, 12))), 4, dict(box_box_detection=True)


def case_boxbox_pile():
    s, m = _scene(xml=BOXBOX_PILE_XML, box_box_detection=True)
    return s, m, dict(qpos=np.asarray(m.qpos0, float)[None].copy(), vel=np.zeros((1, m.n_dofs))), 4, dict(box_box_detection=True)


def terrain_height_field(rows=16, cols=14, seed=11):
    """A bumpy height field in units of vertical_scale = 0.005 m: smooth hills plus integer noise (deterministic)."""
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    hills = 30.0 * np.sin(0.55 * r) * np.cos(0.4 * c) + 12.0 * np.cos(0.9 * r + 0.3 * c)
    return np.round(hills + _rng(seed).integers(-4, 5, (rows, cols))).astype(np.int32) + 40


def terrain_morphs():
    """Height-field terrain (narrowphase.py:346-517 func_contact_mpr_terrain) under a box, a sphere and a Go2-sized capsule."""
    return [gs.morphs.Terrain(height_field=terrain_height_field(), horizontal_scale=0.25, vertical_scale=0.005, pos=(-1.0, -1.2, 0.0)),
            gs.morphs.Box(size=(0.3, 0.22, 0.12), pos=(0.1, 0.0, 0.5), quat=(0.96, 0.1, -0.2, 0.15)),
            gs.morphs.Sphere(radius=0.09, pos=(0.8, 0.6, 0.5)), gs.morphs.Box(size=(0.5, 0.08, 0.08), pos=(-0.3, 0.9, 0.5), quat=(0.9, 0.3, 0.1, -0.3))]


def case_terrain_primitives():
    s, m = _scene(morphs=terrain_morphs())
    q = np.asarray(m.qpos0, float)[None].repeat(2, 0).copy()
    for k in range(3):
        q[:, 7 * k + 3: 7 * k + 7] /= np.linalg.norm(q[:, 7 * k + 3: 7 * k + 7], axis=1, keepdims=True)
    q[1, 0::7][:3] += [0.13, -0.21, 0.17]; q[1, 1::7][:3] += [0.06, 0.11, -0.14]   # a second set of cells
    q[:, 2::7] = [[0.30, 0.345, 0.165], [0.315, 0.335, 0.175]]                      # about a centimetre above the ground
    return s, m, dict(qpos=q, vel=np.zeros((2, m.n_dofs))), 10, {}


def case_jumble_gjk():
    """RigidOptions.use_gjk_collision: every convex pair through the GJK distance query + EPA, no MPR (narrowphase.py:604-607, 736-850;
    gjk.py:170-446): boxes, a capsule and a cylinder touching and overlapping, multi-contact perturbations included."""
    s, m = _scene(xml=JUMBLE_XML, use_gjk_collision=True)
    return s, m, dict(qpos=np.asarray(m.qpos0, float)[None].copy(), vel=np.zeros((1, m.n_dofs))), 3, dict(use_gjk_collision=True)


def case_box_box_gjk():
    """The reference's box_box fixture in GJK mode (test_many_boxes_dynamics parametrises gjk_collision, tests/test_rigid_physics.py:1059-1066)."""
    s, m = _scene(xml=BOX_BOX_XML, use_gjk_collision=True)
    q = box_box_qpos(2, seed=3)
    q[:, 9] = [0.595, 0.59]
    return s, m, dict(qpos=q, vel=np.zeros((2, 12))), 4, dict(use_gjk_collision=True)


def case_go2_drop_sap():
    """The same drop with the sweep-and-prune broad phase of the CPU backend (broadphase.py:142-396): the insertion sort is warm-started
    from the previous step, the candidate pairs come out in sweep order."""
    s, m, init, K, opts = case_go2_drop()
    return s, m, init, K, dict(opts, _broadphase="sap")


def case_jumble_deep_sap():
    s, m, init, K, opts = case_jumble_deep()
    return s, m, init, K, dict(opts, _broadphase="sap")


CASES = {name[5:]: fn for name, fn in list(globals().items()) if name.startswith("case_")}

# per-step quantities recorded by the reference run, with the oracle field (orc.get name, shape per env) they pin and the scope
# row of SURVEY.md section 8a each one covers.  "pre" = read after the constraint solve, before integration (the quantities of
# the step's own configuration); "post" = after kernel_step_2 (integration + FK).
PRE_FIELDS = [
    # (fixture key, oracle name, kind) -- kind: "dof" (nD), "link3", "link4", "mat" (nD x nD), "geom3", "geom4"
    ("mass_mat", "M", "mat"), ("mass_mat_D_inv", "Dinv", "dof"), ("qf_bias", "qf_bias", "dof"), ("qf_passive", "qf_passive", "dof"),
    ("qf_applied", "qf_applied", "dof"), ("acc_smooth", "acc_smooth", "dof"), ("force", "force", "dof"),
    ("qf_constraint", "qf_constraint", "dof"), ("acc", "acc", "dof"), ("aabb_min", "aabb_min", "geom3"), ("aabb_max", "aabb_max", "geom3"),
    ("links_contact_force", "links_contact_force", "link3"),
]
POST_FIELDS = [
    ("qpos", "qpos", "q"), ("vel", "vel", "dof"), ("dofs_pos", "dofs_pos", "dof"), ("links_pos", "links_pos", "link3"), ("links_quat", "links_quat", "link4"),
    ("root_COM", "root_COM", "link3"), ("cd_vel", "cd_vel", "link3"), ("cd_ang", "cd_ang", "link3"), ("cdof_ang", "cdof_ang", "dof3"),
    ("cdof_vel", "cdof_vel", "dof3"), ("cdofd_ang", "cdofd_ang", "dof3"), ("cdofd_vel", "cdofd_vel", "dof3"), ("cacc_lin", "cacc_lin", "link3"),
    ("cacc_ang", "cacc_ang", "link3"), ("geoms_pos", "geoms_pos", "geom3"), ("geoms_quat", "geoms_quat", "geom4"),
]

# Round 1 ran EPA from MPR's final portal where the reference re-runs the pair through GJK + EPA (narrowphase.py:726-733); the
# oracle and the CUDA path now restate gjk.py / epa.py themselves (scope row a17), so no case needs a looser bound any more.
GJK_PATH_TOL = {}
