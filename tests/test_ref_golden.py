"""Pins the oracle to the REFERENCE'S OWN CODE: tests/golden/ref_*.npz were produced by executing the reference's Python kernel
sources (imported in place from /root/reference, run through the `quadrants` stand-in of oracle/qd_shim; generator:
tools/make_ref_golden.py, inputs: tests/ref_cases.py).  Here the fp64 instantiation of oracle/rigid_oracle.hpp replays the same
inputs; every recorded quantity of every step must agree.

Tolerances (both sides are fp64 with the same fp32 constants, so only the order of floating-point operations differs):
  * quantities computed before the constraint solver (mass matrix, L^T D L, bias / passive / applied forces, smooth acceleration,
    AABBs, kinematics of the previous step): 1e-9 * max(1, |ref|)
  * quantities downstream of the iterative solver (acc, qf_constraint, contact forces, vel, qpos, FK of the new state): 1e-6 *
    max(1, |ref|) -- the Newton / CG iteration stops at tolerance * meaninertia * nD = 1e-6-ish, and a different summation
    order may stop one line-search evaluation earlier
  * contacts: same count, same (geom_a, geom_b) multiset; positions / normals / penetrations 1e-7 after matching.
"""

import os

import numpy as np
import pytest

import ref_cases
from helpers import oracle_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_PRE, TOL_POST = 1e-9, 1e-6
POST_SOLVER = {"qf_constraint", "acc", "links_contact_force", "force"}  # dofs_state.force = qf_smooth + qf_constraint after the solve


def model_digest(m):
    import hashlib

    h = hashlib.sha256()
    for k in sorted(m.__dict__):
        v = m.__dict__[k]
        if isinstance(v, np.ndarray) and k not in ("support_v", "support_vid"):
            h.update(k.encode()); h.update(np.ascontiguousarray(np.round(np.asarray(v, np.float64), 9)).tobytes())
    return h.hexdigest()[:16]


def shape_of(kind, m):
    return {"dof": (m.n_dofs,), "q": (m.n_qs,), "mat": (m.n_dofs * m.n_dofs,), "link3": (m.n_links, 3), "link4": (m.n_links, 4),
            "geom3": (m.n_geoms, 3), "geom4": (m.n_geoms, 4), "dof3": (m.n_dofs, 3)}[kind]


def available():
    return sorted(n for n in ref_cases.CASES if os.path.exists(os.path.join(GOLDEN, f"ref_{n}.npz")))


def close(a, b, tol, what):
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= tol * scale, f"{what}: max|oracle - reference| = {err:.3e} > {tol:.0e} * {scale:.3e}"
    return err


def match_contacts(oc, ref, b, k, tol_geom=1e-7, tol_force=1e-5):
    """Pairs every oracle contact of env b with a reference contact of the same geoms at the closest position."""
    n = int(ref["n_contacts"][k, b])
    assert len(oc["geom_a"]) == n, f"step {k} env {b}: {len(oc['geom_a'])} oracle contacts vs {n} reference contacts"
    used, worst = set(), 0.0
    for i in range(n):
        ga, gb = int(oc["geom_a"][i]), int(oc["geom_b"][i])
        cand = [j for j in range(n) if j not in used and int(ref["contact_geom_a"][k, b, j]) == ga and int(ref["contact_geom_b"][k, b, j]) == gb]
        assert cand, f"step {k} env {b}: oracle contact ({ga}, {gb}) has no reference counterpart"
        d = [np.abs(ref["contact_pos"][k, b, j] - oc["position"][i]).max() for j in cand]
        j = cand[int(np.argmin(d))]
        used.add(j)
        for key, okey in (("contact_pos", "position"), ("contact_normal", "normal"), ("contact_penetration", "penetration"),
                          ("contact_friction", "friction")):
            worst = max(worst, float(np.abs(ref[key][k, b, j] - oc[okey][i]).max()))
        worst_f = float(np.abs(ref["contact_force"][k, b, j] - oc["force"][i]).max())
        assert worst_f <= tol_force * max(1.0, float(np.abs(ref["contact_force"][k, b]).max())), f"step {k} env {b}: contact force differs by {worst_f:.2e}"
    assert worst <= tol_geom, f"step {k} env {b}: contact geometry differs by {worst:.2e}"


@pytest.mark.parametrize("name", available())
def test_oracle_matches_reference_fixture(name):
    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    solver, m, init, K, opts = ref_cases.CASES[name]()
    assert model_digest(m) == str(ref["model_digest"]), "the model compiled now is not the model the fixture was generated for (re-run tools/make_ref_golden.py)"
    assert K == int(ref["K"])
    B = len(init["qpos"])
    loose = ref_cases.GJK_PATH_TOL.get(name)
    tol_post = loose["post"] if loose else TOL_POST
    orc = oracle_for(solver, m, B, double=True, sort_contacts=0, broadphase_sap=int(opts.get("_broadphase") == "sap"))
    for k_ in ("qpos", "vel", "ctrl_pos", "ctrl_vel", "ctrl_force", "ctrl_mode"):
        if k_ in init:
            np.testing.assert_array_equal(init[k_], ref["init_" + k_])
            orc.set(k_, np.asarray(init[k_], np.float64))
    for k in range(loose["steps"] if loose else K):
        orc.step(1)
        for key, oname, kind in ref_cases.PRE_FIELDS:
            shp = shape_of(kind, m)
            a = orc.get(oname, shp).reshape(B, -1)
            pre_tol = TOL_PRE if (k == 0 or not loose) else tol_post  # after step 0 the state itself carries the deviation
            close(a, ref["pre_" + key][k].reshape(B, -1), tol_post if key in POST_SOLVER else pre_tol, f"{name} step {k} {key}")
        for key, oname, kind in ref_cases.POST_FIELDS:
            shp = shape_of(kind, m)
            a = orc.get(oname, shp).reshape(B, -1)
            b = ref["post_" + key][k].reshape(B, -1)
            if kind in ("link4", "geom4") or (key == "qpos"):
                pass  # quaternions compared as they are: both sides normalise the same way
            close(a, b, tol_post, f"{name} step {k} {key}")
        nc = orc.get("n_constraints", 3)
        np.testing.assert_array_equal(nc[:, 0].astype(int), ref["n_constraints"][k], err_msg=f"{name} step {k}: number of constraint rows")
        np.testing.assert_array_equal(nc[:, 1].astype(int), ref["n_constraints_equality"][k])
        np.testing.assert_array_equal(nc[:, 2].astype(int), ref["n_constraints_frictionloss"][k])
        assert int(orc.get("errno", 1).max()) == int(ref["errno"][k].max())
        # scope row a13: the broad phase hands the same number of candidate pairs to the narrow phase
        np.testing.assert_array_equal(orc.get("n_broad", 1)[:, 0].astype(int), ref["n_broad_pairs"][k], err_msg=f"{name} step {k}: broad-phase candidates")
        for b in range(B):
            match_contacts(orc.contacts(b), ref, b, k, *((loose["contact"], loose["post"]) if loose else ()))


@pytest.mark.parametrize("name", available())
def test_invweight_and_meaninertia_match_reference_method(name):
    """Scope row a27: links / dofs invweight and meaninertia as the reference's own RigidSolver._init_invweight_and_meaninertia
    (rigid_solver.py:492-655, run on this model through the shim) computes them vs genesis_b200.model.refdyn -- values, not just
    finiteness: a wrong invweight silently rescales every diag / aref downstream."""
    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    _, m, *_ = ref_cases.CASES[name]()
    if m.n_dofs == 0:
        pytest.skip("no dofs")
    free = ~np.asarray(m.links_is_fixed, bool)
    lw = np.asarray(m.links_invweight, np.float64).reshape(-1, 2)
    np.testing.assert_allclose(lw[free], ref["ref_links_invweight"][free], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.asarray(m.dofs_invweight, np.float64), ref["ref_dofs_invweight"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(float(m.meaninertia), ref["ref_meaninertia"][0], rtol=1e-9)


def test_enough_reference_fixtures_are_committed():
    """at least 8 scope rows pinned by fixtures generated from imported reference source."""
    assert len(available()) >= 8, available()
