"""The CUDA path against fixtures produced by executing the REFERENCE'S OWN kernel sources (tests/golden/ref_*.npz, generator
tools/make_ref_golden.py, inputs tests/ref_cases.py): identical initial states, K steps, every recorded quantity compared.

Tolerances (fp32 CUDA vs the reference's algorithm evaluated in fp64), as in tests/test_gpu_parity.py:
  * before the constraint solver (mass matrix, D^-1, bias / passive / applied forces, smooth acceleration, AABBs): 5e-5 * max(1, |ref|)
    (the reference's own TOL_SINGLE)
  * downstream of the solver and accumulated over the K <= 8 steps (acc, forces, vel, qpos, FK of the new state): 5e-4 * max(1, |ref|)
  * contacts: same count per env and step, same (geom_a, geom_b) multiset.
Cases on the reference's GJK re-run path (ref_cases.GJK_PATH_TOL, scope row a17) are compared on their first step with 2e-3.
"""

import os

import numpy as np
import pytest
import torch

import ref_cases
from conftest import TOL_SINGLE

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_SOLVER = 5e-4
POST_SOLVER = {"qf_constraint", "acc", "links_contact_force", "force"}
GPU_FIELD = {"mass_mat_D_inv": "mass_Dinv"}


@pytest.fixture(autouse=True)
def _diagnostics_on(monkeypatch):
    monkeypatch.setenv("GSR_DIAGNOSTICS", "1")


def available():
    return sorted(n for n in ref_cases.CASES if os.path.exists(os.path.join(GOLDEN, f"ref_{n}.npz")))


def close(a, b, tol, what):
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= tol * scale, f"{what}: max|cuda - reference| = {err:.3e} > {tol:.0e} * {scale:.3e}"


@pytest.mark.parametrize("name", available())
def test_cuda_matches_reference_fixture(name):
    import genesis_b200 as gs

    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    solver, m, init, K, opts = ref_cases.CASES[name]()
    B = len(init["qpos"])
    scene = solver._scene
    scene.build(n_envs=B)
    s = scene.rigid_solver
    dev = gs.device
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a), dtype=dt, device=dev)
    for key, field in (("ctrl_pos", "ctrl_pos"), ("ctrl_vel", "ctrl_vel"), ("ctrl_force", "ctrl_force")):
        if key in init:
            s.field(field).copy_(t(init[key]))
    if "ctrl_mode" in init:
        s.field("ctrl_mode").copy_(t(init["ctrl_mode"], torch.int32))
    s.set_qpos(t(init["qpos"]))
    if "vel" in init:
        s.set_dofs_velocity(t(init["vel"]))
    loose = ref_cases.GJK_PATH_TOL.get(name)
    tol_post = loose["post"] if loose else TOL_SOLVER
    for k in range(loose["steps"] if loose else K):
        scene.step()
        for key, _, kind in ref_cases.PRE_FIELDS:
            a = s.rows(GPU_FIELD.get(key, key)).cpu().numpy().astype(np.float64)
            b = ref["pre_" + key][k].reshape(B, -1)
            tol = tol_post if (key in POST_SOLVER or k > 0) else TOL_SINGLE
            close(a, b, tol, f"{name} step {k} {key}")
        for key, _, kind in ref_cases.POST_FIELDS:
            a = s.rows(key).cpu().numpy().astype(np.float64)
            close(a, ref["post_" + key][k].reshape(B, -1), tol_post, f"{name} step {k} {key}")
        np.testing.assert_array_equal(s.field("n_constraints")[:, 0].cpu().numpy(), ref["n_constraints"][k], err_msg=f"{name} step {k}: rows")
        np.testing.assert_array_equal(s.field("n_contacts")[:, 0].cpu().numpy(), ref["n_contacts"][k], err_msg=f"{name} step {k}: contacts")
        ga, gb = s.field("contact_geom_a").cpu().numpy(), s.field("contact_geom_b").cpu().numpy()
        for b_ in range(B):
            n = int(ref["n_contacts"][k, b_])
            mine = sorted(zip(ga[b_, :n].tolist(), gb[b_, :n].tolist()))
            theirs = sorted(zip(ref["contact_geom_a"][k, b_, :n].astype(int).tolist(), ref["contact_geom_b"][k, b_, :n].astype(int).tolist()))
            assert mine == theirs, f"{name} step {k} env {b_}: contact pairs differ"
    assert int(s.field("errno_").max()) == 0
