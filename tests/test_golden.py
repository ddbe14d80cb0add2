"""Committed trajectories (tests/golden/*.npz, written by tools/make_golden.py from the fp64 oracle) pin (a) the fp32 and
fp64 oracle builds on the CPU and (b) the CUDA path on the GPU without running the oracle.  They are regression vectors of
the restatement, NOT reference outputs (the reference cannot run here; see DESIGN.md 'Oracle')."""

import os

import numpy as np
import pytest

from helpers import gs, make_franka, make_go2, make_weld, oracle_for, host_model

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# name: (scene factory, factory kwargs, |dqpos| tolerance, |dqvel| tolerance)
CASES = {"franka_random": (make_franka, {}, 1e-4, 1e-3), "go2_random": (make_go2, {}, 1e-3, 2e-2),
         # (the capsule scene is left out on purpose: a sphere balanced on a capsule amplifies fp32 rounding to centimetres in
         # 20 steps, so fp64 vectors cannot pin an fp32 path there; it is covered against the fp32 oracle instead)
         "weld": (make_weld, {}, 1e-4, 1e-3),
         "go2_noslip": (make_go2, dict(noslip_iterations=5), 1e-3, 2e-2)}


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("double", [True, False], ids=["f64", "f32"])
def test_oracle_reproduces_golden(name, double):
    make, kw, tq, tv = CASES[name]
    g = _load(name)
    scene, solver, model = host_model(make, **kw)
    B = g["init_qpos"].shape[0]
    orc = oracle_for(solver, model, B, double=double)
    for k in ("qpos", "vel", "ctrl_pos", "ctrl_mode"):
        orc.set(k, g["init_" + k])
    done = 0
    for i, step in enumerate(g["steps"]):
        orc.step(int(step) - done)
        done = int(step)
        if double:
            np.testing.assert_allclose(orc.qpos(), g["qpos"][i], atol=1e-9)
            np.testing.assert_array_equal(orc.get("n_contacts", 1)[:, 0], g["n_contacts"][i])
        else:
            dq = np.abs(orc.qpos() - g["qpos"][i]).max(axis=1)
            assert (dq <= tq).mean() >= 0.85, f"step {step}: {np.sort(dq)}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(name):
    import torch

    make, kw, tq, tv = CASES[name]
    g = _load(name)
    B = g["init_qpos"].shape[0]
    scene = make(B, **kw)[0]
    s = scene.rigid_solver
    dev = s.device
    s.set_qpos(torch.tensor(g["init_qpos"], dtype=torch.float32, device=dev), skip_forward=True)
    s.set_dofs_velocity(torch.tensor(g["init_vel"], dtype=torch.float32, device=dev))
    s.field("ctrl_pos")[:] = torch.tensor(g["init_ctrl_pos"], dtype=torch.float32, device=dev)
    s.field("ctrl_mode")[:] = torch.tensor(g["init_ctrl_mode"], dtype=torch.int32, device=dev)
    done = 0
    for i, step in enumerate(g["steps"]):
        for _ in range(int(step) - done):
            scene.step()
        done = int(step)
        dq = np.abs(s.field("qpos").cpu().numpy() - g["qpos"][i]).max(axis=1)
        dv = np.abs(s.field("vel").cpu().numpy() - g["vel"][i]).max(axis=1)
        # fp32 CUDA vs fp64 golden: every env within the stated trajectory tolerance except chaotic outliers (<= 15 %)
        assert (dq <= tq).mean() >= 0.85, f"step {step}: {np.sort(dq)}"
        assert (dv <= tv).mean() >= 0.85, f"step {step}: {np.sort(dv)}"
