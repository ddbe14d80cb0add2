"""Writes tests/golden/*.npz: trajectories of the fp64 CPU oracle on the benchmark scenes.

The reference itself (Quadrants 0.5.2 JIT + MuJoCo model compiler) is not installable in this container, so these are
REGRESSION vectors of the restatement (oracle/), not outputs of the reference: they pin the oracle against accidental
change and let the GPU parity tests run without re-deriving anything.  Regenerate with `python tools/make_golden.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import FRANKA_Q0, GO2_CTRL, GO2_Q0, make_franka, make_go2
from oracle.pyoracle import Oracle

gs.init(backend=gs.gpu, seed=0)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def franka_case(B=8, K=60, seed=0):
    scene, _ = make_franka(0, build=False)
    m = scene.rigid_solver.build_model()
    rng = np.random.default_rng(seed)
    q = np.tile(np.array(FRANKA_Q0), (B, 1))
    v = 0.2 * np.clip(rng.normal(size=(B, 9)), -1, 1); v[:, 7:] = 0
    mode = np.full((B, 9), int(gs.CTRL_MODE.POSITION))
    return scene, m, dict(qpos=q, vel=v, ctrl_pos=q.copy(), ctrl_mode=mode), K


def go2_case(B=8, K=80, seed=1):
    scene, _ = make_go2(0, build=False)
    m = scene.rigid_solver.build_model()
    rng = np.random.default_rng(seed)
    lim = np.asarray(m.dofs_limit).reshape(-1, 2)[6:]
    q = np.tile(np.array(GO2_Q0), (B, 1))
    q[:, 7:] = lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((B, 12))
    ctrl = np.zeros((B, 18)); ctrl[:, 6:] = GO2_CTRL
    mode = np.full((B, 18), int(gs.CTRL_MODE.FORCE)); mode[:, 6:] = int(gs.CTRL_MODE.POSITION)
    return scene, m, dict(qpos=q, vel=np.zeros((B, 18)), ctrl_pos=ctrl, ctrl_mode=mode), K


def default_state_case(make, K, **kw):
    """Scenes that start from the model's own qpos0 with zero velocity and the build-time control state."""
    def case():
        scene = make(0, build=False, **kw)[0]
        m = scene.rigid_solver.build_model()
        B = 4
        q = np.tile(np.asarray(m.qpos0, dtype=np.float64), (B, 1))
        mode = np.full((B, m.n_dofs), int(gs.CTRL_MODE.FORCE))
        return scene, m, dict(qpos=q, vel=np.zeros((B, m.n_dofs)), ctrl_pos=np.zeros((B, m.n_dofs)), ctrl_mode=mode), K
    return case


def go2_noslip_case():
    scene, m, init, K = go2_case(B=4, K=60, seed=2)
    scene, _ = make_go2(0, build=False, noslip_iterations=5)
    return scene, scene.rigid_solver.build_model(), init, K


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import make_weld

    os.makedirs(OUT, exist_ok=True)
    for name, case in (("franka_random", franka_case), ("go2_random", go2_case), ("weld", default_state_case(make_weld, 100)), ("go2_noslip", go2_noslip_case)):
        scene, m, init, K = case()
        B = init["qpos"].shape[0]
        orc = Oracle(m, B, double=True, options=scene.rigid_solver.model_options())
        for k in ("qpos", "vel", "ctrl_pos", "ctrl_mode"):
            orc.set(k, init[k])
        traj_q, traj_v, traj_nc = [], [], []
        for k in range(K):
            orc.step(1)
            if (k + 1) % 20 == 0:
                traj_q.append(orc.qpos()); traj_v.append(orc.vel()); traj_nc.append(orc.get("n_contacts", 1)[:, 0])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), steps=np.arange(20, K + 1, 20), qpos=np.array(traj_q), vel=np.array(traj_v),
                            n_contacts=np.array(traj_nc), **{"init_" + k: v for k, v in init.items()})
        print(name, "->", np.array(traj_q).shape, "contacts at end", traj_nc[-1])
