"""Size-independent properties at BASELINE.json's full sizes (30000 Franka envs, 4096 Go2 envs), where the oracle is too
slow to be the checker: batch-position invariance, run-to-run determinism, errno cleanliness, physical sanity, and the
equality of the host-buffer plugin call (gsr_step_host) with the device-resident path."""

import numpy as np
import pytest
import torch

from helpers import gs, make_franka, make_go2, oracle_for

pytestmark = pytest.mark.gpu


def test_franka_30000_batch_invariance_and_oracle_sample():
    B, K, R = 30000, 30, 50
    scene, franka = make_franka(B, randomized=True)
    s = scene.rigid_solver
    # make env b a replica of env (b % R): same initial velocity
    v = s.field("vel")
    v.copy_(v[:R].repeat(B // R, 1))
    s._forward()
    orc = oracle_for(s, s.model, R, threads=8)
    for name in ("qpos", "vel", "ctrl_pos", "ctrl_vel", "ctrl_force", "ctrl_mode"):
        orc.set(name, s.rows(name)[:R].cpu().numpy())
    for _ in range(K):
        scene.step()
    q = s.field("qpos")
    assert torch.equal(q, q[:R].repeat(B // R, 1)), "a replica env diverged from its source env (bit-exact expected)"
    assert int(s.field("errno_").max()) == 0
    orc.step(K)
    assert np.abs(q[:R].cpu().numpy() - orc.qpos()).max() <= 1e-4


def test_go2_4096_determinism_and_sanity():
    B, K = 4096, 100
    outs = []
    for _ in range(2):
        scene, robot = make_go2(B, seed=7)
        for _ in range(K):
            scene.step()
        s = scene.rigid_solver
        outs.append((s.field("qpos").clone(), s.field("vel").clone(), s.field("n_contacts").clone()))
        assert int(s.field("errno_").max()) == 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    q = outs[0][0]
    assert torch.isfinite(q).all()
    assert torch.allclose(q[:, 3:7].norm(dim=1), torch.ones(B, device=q.device), atol=1e-4)  # unit base quaternion
    assert (q[:, 2] > 0.0).all() and (q[:, 2] < 0.6).all()  # nobody fell through the plane or flew away
    assert (outs[0][2] > 0).float().mean() > 0.95  # nearly every robot touches the ground after 1 s


@pytest.mark.parametrize("buffers,host_io", [("pinned", 2), ("pinned", 1), ("pageable", 2), ("pinned", 0), ("pageable", 0)])
def test_step_host_equals_device_path(buffers, host_io):
    """gsr_step_host equals the device-resident path bit for bit in all its variants: k_step reading the targets from and
    writing the results to dense staging (cfg_host_io = 1), results posted straight into page-locked caller buffers (2, the
    default; pageable buffers fall back to 1), and the staged copies with k_scatter / k_gather launches (0)."""
    import os

    B, K = 256, 10
    scene_a, fa = make_franka(B, randomized=True)
    os.environ["GSR_HOST_IO"] = str(host_io)
    try:
        scene_b, fb = make_franka(B, randomized=True)
    finally:
        del os.environ["GSR_HOST_IO"]
    sa, sb = scene_a.rigid_solver, scene_b.rigid_solver
    ctrl0 = np.tile(np.array([[0.1, 0.0, 0.0, -1.2, 0.0, 1.1, 0.0, 0.02, 0.02]], dtype=np.float32), (B, 1))
    if buffers == "pinned":
        keep = [torch.empty((B, 9), dtype=torch.float32).pin_memory() for _ in range(3)]
        ctrl, qpos, vel = (t.numpy() for t in keep)
        ctrl[:] = ctrl0
        qpos[:] = 0
        vel[:] = 0
    else:
        ctrl, qpos, vel = ctrl0.copy(), np.zeros((B, 9), np.float32), np.zeros((B, 9), np.float32)
    fa.control_dofs_position(torch.from_numpy(ctrl0).to(gs.device))
    for _ in range(K):
        scene_a.step()
        sb.step_host(ctrl, qpos, vel)
    assert np.array_equal(sa.field("qpos").cpu().numpy(), qpos)
    assert np.array_equal(sa.field("vel").cpu().numpy(), vel)
    assert np.array_equal(sb.field("ctrl_pos").cpu().numpy(), ctrl0)  # the targets also landed in the device record
    launches = sb.kernel_launches()
    assert sb.launch_count() > 0 and launches["k_step"] >= K
    fused = host_io > 0 and sb.launch_info["group_size"] != 1  # the lane-per-environment kernel (GSR_GROUP_SIZE=1) always stages
    assert (launches["k_scatter"] == 0) == fused and (launches["k_gather"] == 0) == fused


def test_grasp_8192_determinism_launch_order_and_sanity():
    """BASELINE.json configs[3] at full size: Franka over a cluster of four boxes, 8192 envs.  Two runs agree bit for bit
    although the cost-ordered launch permutes the environments differently over time (reorder period 16 vs off), nothing
    overflows, and the boxes stay on the plane."""
    import os

    from helpers import make_franka_grasp

    B, K = 8192, 60
    outs = []
    for period in ("16", "0"):
        os.environ["GSR_REORDER_PERIOD"] = period
        try:
            scene, franka, boxes = make_franka_grasp(B)
        finally:
            os.environ.pop("GSR_REORDER_PERIOD", None)
        s = scene.rigid_solver
        # make the batch heterogeneous: shift the arm target of every env a little
        q = torch.tensor([-0.0, 0.45, 0.0, -2.2, 0.0, 2.65, 0.78, 0.04, 0.04], device=gs.device).repeat(B, 1)
        q[:, 1] += 0.2 * (torch.arange(B, device=gs.device) % 17) / 17.0
        franka.control_dofs_position(q)
        for _ in range(K):
            scene.step()
        assert int(s.field("errno_").max()) == 0
        outs.append((s.field("qpos").clone(), s.field("n_contacts").clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    qpos = outs[0][0]
    assert torch.isfinite(qpos).all()
    z = qpos[:, 9 + 2::7][:, :4]  # z of the four free boxes
    assert (z > 0.015).all() and (z < 0.2).all()
    assert float(outs[0][1].float().mean()) > 20.0  # four boxes x four corner contacts + finger-box and box-box contacts


def test_step_host_after_device_steps_keeps_adaptive_rows_alive(monkeypatch):
    """Regression for the use-after-free of round 1 (gsr_api.cu: the first gsr_step_host after >= 16 device steps freed the row
    histogram of the adaptive on-chip capacity).  Grasp scene with one box (a few dozen rows: within reach of the adaptation,
    gsr_api.cu adapt_rows; the generic library, the model-specialised copies fix the capacity): 40 device steps (two reorder
    periods: the histogram exists and the capacity has been raised once), then 64 host-buffer steps; the capacity must still
    follow the workload and the results must equal a scene that only ever took device steps."""
    from helpers import make_franka_grasp

    monkeypatch.setenv("GSR_NO_SPEC", "1")
    B = 2048
    scene_a, fa, _ = make_franka_grasp(B, n_boxes=1)
    scene_b, fb, _ = make_franka_grasp(B, n_boxes=1)
    sa, sb = scene_a.rigid_solver, scene_b.rigid_solver
    rows0 = sb.launch_info["smem_rows"]
    for _ in range(40):
        scene_a.step(); scene_b.step()
    keep = [torch.empty((B, n), dtype=torch.float32).pin_memory() for n in (sb.model.n_dofs, sb.model.n_qs, sb.model.n_dofs)]
    ctrl, qpos, vel = (t.numpy() for t in keep)
    ctrl[:] = sb.field("ctrl_pos").cpu().numpy()
    for _ in range(64):
        scene_a.step()
        sb.step_host(ctrl, qpos, vel)
    assert np.array_equal(sa.field("qpos").cpu().numpy(), qpos) and np.array_equal(sa.field("vel").cpu().numpy(), vel)
    assert int(sb.field("errno_").max()) == 0
    rows = sb.field("n_constraints").float()
    print("rows: default", rows0, "mean", float(rows.mean()), "max", float(rows.max()), "now", sb.launch_info)
    if rows0 + 8 <= float(rows.min()) and float(rows.max()) <= 88:
        # every environment needs more rows than the default capacity and the need is within reach (GSR_ROWS_GROW_MAX = 96):
        # the adaptation must have happened and survived
        assert sb.launch_info["smem_rows"] > rows0, (rows0, sb.launch_info)
    assert sb.launch_info == sa.launch_info


def test_adaptive_rows_stay_at_full_residency_when_out_of_reach(monkeypatch):
    """The full grasp cluster carries about 200 rows per environment: raising the on-chip capacity that far costs three quarters
    of the residency (measured 1.48 M against 2.44 M env-steps/s, profiles/README.md), so the capacity stays at its default."""
    from helpers import make_franka_grasp

    monkeypatch.setenv("GSR_NO_SPEC", "1")
    scene, _, _ = make_franka_grasp(1024)
    s = scene.rigid_solver
    rows0 = s.launch_info["smem_rows"]
    for _ in range(72):
        scene.step()
    assert float(s.field("n_constraints").float().mean()) > 96
    assert s.launch_info["smem_rows"] == rows0
    assert int(s.field("errno_").max()) == 0


@pytest.mark.parametrize("model", ["go2", "franka"])
def test_lazy_kinematics_equals_eager(model, monkeypatch):
    """Adaptive kinematics (RigidSolver.substep, opt-in with GSR_LAZY_FK=1): a loop that only steps runs `lazy` -- every step recomputes the kinematic cache
    on chip from qpos / vel and stores the state block only -- while a loop that reads link poses stays `eager` (the reference's
    FK at the end of kernel_step_2).  Same arithmetic either way: trajectories, contacts and the link poses an accessor finally
    asks for are bit-identical, and mode switches in the middle of a run change nothing."""
    from genesis_b200 import _lib

    make = (lambda: make_go2(256, seed=11)) if model == "go2" else (lambda: make_franka(256, self_collision=True, randomized=True, controlled=False, seed=11))
    monkeypatch.setenv("GSR_NO_SPEC", "1")  # the generic library shares one copy of the kinematics between both call sites: bit-exact
    monkeypatch.setenv("GSR_LAZY_FK", "0")
    scene_e, robot_e = make()
    monkeypatch.setenv("GSR_LAZY_FK", "1")
    scene_l, robot_l = make()
    se, sl = scene_e.rigid_solver, scene_l.rigid_solver
    for k in range(40):
        scene_e.step(); scene_l.step()
        if k in (9, 25):  # an accessor in the middle of the run: one forward pass, then eager for the following step
            assert torch.equal(robot_l.get_links_pos(), robot_e.get_links_pos())
    assert sl._eager_fk is False and sl._L.gsr_kinematics_valid(sl._handle) == 0, "the stepping-only loop did not run lazy"
    assert se._eager_fk is True and se._L.gsr_kinematics_valid(se._handle) == 1
    assert torch.equal(se.field("qpos"), sl.field("qpos")) and torch.equal(se.field("vel"), sl.field("vel"))
    assert torch.equal(se.field("n_contacts"), sl.field("n_contacts"))
    assert torch.equal(se.field("contact_pos"), sl.field("contact_pos"))
    for name in ("links_pos", "links_quat", "cd_vel", "cd_ang", "root_COM", "dofs_pos", "cdof_ang", "cinr_I"):
        assert torch.equal(se.field(name), sl.field(name)), name
    assert sl.kernel_launches()["k_forward"] > se.kernel_launches()["k_forward"]


def test_kinematics_policy_defaults(monkeypatch):
    """Default kinematics policy (RigidSolver._apply_kinematics_policy): a model-specialised library copy steps lazily -- measured
    faster there, profiles/README.md item 25 --, the generic library eagerly like the reference; GSR_LAZY_FK forces either.  An
    accessor that needs link poses gets them either way (one forward pass, then one eager step), and the specialised copy's lazy
    and eager trajectories agree to round-off (its kinematics are inlined at both call sites: same source, two code copies)."""
    monkeypatch.delenv("GSR_LAZY_FK", raising=False)
    scene_l, robot_l = make_franka(256, randomized=True, seed=3)
    monkeypatch.setenv("GSR_LAZY_FK", "0")
    scene_e, robot_e = make_franka(256, randomized=True, seed=3)
    monkeypatch.delenv("GSR_LAZY_FK")
    monkeypatch.setenv("GSR_NO_SPEC", "1")
    scene_g, robot_g = make_franka(256, randomized=True, seed=3)
    monkeypatch.delenv("GSR_NO_SPEC")
    sl, se, sg = scene_l.rigid_solver, scene_e.rigid_solver, scene_g.rigid_solver
    assert sl.specialised and se.specialised and not sg.specialised
    for _ in range(40):
        scene_l.step(); scene_e.step(); scene_g.step()
    assert sl._eager_fk is False and sl._L.gsr_kinematics_valid(sl._handle) == 0, "the specialised copy did not step lazily by default"
    assert se._eager_fk is True and se._L.gsr_kinematics_valid(se._handle) == 1
    assert sg._eager_fk is True and sg._L.gsr_kinematics_valid(sg._handle) == 1, "the generic library must keep the reference's eager order"
    np.testing.assert_allclose(sl.field("qpos").cpu().numpy(), se.field("qpos").cpu().numpy(), atol=2e-6)
    np.testing.assert_allclose(robot_l.get_links_pos().cpu().numpy(), robot_e.get_links_pos().cpu().numpy(), atol=2e-6)
    np.testing.assert_allclose(robot_l.get_links_pos().cpu().numpy(), robot_g.get_links_pos().cpu().numpy(), atol=1e-5)
    scene_l.step()
    assert sl._eager_fk is True, "the step after an accessor that needed link poses leaves them in the records"
    scene_l.step()
    assert sl._eager_fk is False
    assert int(sl.field("errno_").max()) == 0


@pytest.mark.parametrize("model", ["franka", "go2"])
def test_specialised_library_is_used_and_matches(model, monkeypatch):
    """Model-specialised builds (tools/specialize.py, built by __graft_entry__.build() for the two benchmark models): the benchmark
    scene picks the copy compiled for it (layouts, counts and options as literals, kinematics / math helpers inlined), any other
    option set runs the generic library.  Same source, other inlining => other FMA contractions: the two binaries agree to round-off
    (not bit for bit), and the specialised one is compared with the fp32 oracle directly, like the generic one in the parity suite."""
    make = (lambda: make_go2(128, seed=3)) if model == "go2" else (lambda: make_franka(256, randomized=True, seed=3))
    scene_s, _ = make()
    monkeypatch.setenv("GSR_NO_SPEC", "1")
    scene_g, _ = make()
    monkeypatch.delenv("GSR_NO_SPEC")
    ss, sg = scene_s.rigid_solver, scene_g.rigid_solver
    assert ss.specialised and ss._L.gsr_is_specialised() == 1, "the specialised copy was not built / not found (run __graft_entry__.build())"
    assert not sg.specialised and sg._L.gsr_is_specialised() == 0
    orc = oracle_for(ss, ss.model, ss._B, threads=8)
    for f in ("qpos", "vel", "ctrl_pos", "ctrl_vel", "ctrl_force", "ctrl_mode"):
        orc.set(f, ss.rows(f).cpu().numpy())
    K = 40
    for _ in range(K):
        scene_s.step(); scene_g.step()
    orc.step(K)
    qs, qg = ss.field("qpos").cpu().numpy(), sg.field("qpos").cpu().numpy()
    tol = 1e-3 if model == "go2" else 1e-5  # Go2: contacts (the parity suite's trajectory tolerance); Franka: smooth dynamics
    assert (np.abs(qs - qg).max(axis=1) <= tol).mean() >= 0.98, np.sort(np.abs(qs - qg).max(axis=1))[-5:]
    assert (np.abs(qs - orc.qpos()).max(axis=1) <= max(tol, 1e-4)).mean() >= 0.98, np.sort(np.abs(qs - orc.qpos()).max(axis=1))[-5:]
    nc_s, nc_g = ss.field("n_contacts")[:, 0].cpu().numpy(), sg.field("n_contacts")[:, 0].cpu().numpy()
    assert (nc_s == nc_g).mean() >= 0.95
    other, _ = (make_go2(16, seed=3, noslip_iterations=2) if model == "go2" else make_franka(16, randomized=True, seed=3, iterations=40))
    assert not other.rigid_solver.specialised  # another option set: the generic library
    other.step()
