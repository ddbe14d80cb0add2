"""First GPU bring-up: steps the CUDA path next to the CPU oracle and prints per-quantity max-abs differences."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka, make_go2, FRANKA_Q0
from oracle.pyoracle import Oracle

gs.init(backend=gs.gpu)


def compare(name, scene, orc, fields):
    s = scene.rigid_solver
    for fname, oname, shape in fields:
        a = s.rows(fname).cpu().numpy().astype(np.float64)  # (B, rows)
        o = orc.get(oname, shape).reshape(a.shape[0], -1)
        if o.shape != a.shape:
            print(f"  {name}:{fname} shape mismatch {a.shape} vs {o.shape}")
            continue
        d = np.abs(a - o)
        print(f"  {name}:{fname:22s} max|diff|={d.max():.3e}  max|ref|={np.abs(o).max():.3e}")


def soa3(x, n):  # oracle returns (B, n, 3) AoS; device rows are item-major (3n, B) -> after .T (B, 3n) in [item][comp] order
    return x


def run(name, make, n_envs, steps):
    scene, robot = make(n_envs)
    s = scene.rigid_solver
    m = s.model
    print("launch_info", s.launch_info, "state MiB", s._F.numel() * 4 / 2**20)
    orc = Oracle(m, n_envs, double=False, options=s.model_options())
    orc.set("qpos", s.rows("qpos").cpu().numpy())
    orc.set("vel", s.rows("vel").cpu().numpy())
    orc.set("ctrl_pos", s.rows("ctrl_pos").cpu().numpy())
    orc.set("ctrl_mode", s.rows("ctrl_mode").cpu().numpy())
    nD, nL, nG, nQ = m.n_dofs, m.n_links, m.n_geoms, m.n_qs
    fields0 = [("links_pos", "links_pos", (nL, 3)), ("links_quat", "links_quat", (nL, 4)), ("root_COM", "root_COM", (nL, 3)),
               ("cdof_ang", "cdof_ang", (nD, 3)), ("cdof_vel", "cdof_vel", (nD, 3)), ("geoms_pos", "geoms_pos", (nG, 3)),
               ("aabb_min", "aabb_min", (nG, 3)), ("dofs_pos", "dofs_pos", nD)]
    print(f"== {name}: after init/FK")
    orc.detection(0) if False else None
    compare(name, scene, orc, fields0[:-2] + fields0[-1:])
    fields = [("qpos", "qpos", nQ), ("vel", "vel", nD), ("acc", "acc", nD), ("acc_smooth", "acc_smooth", nD), ("qf_bias", "qf_bias", nD),
              ("qf_applied", "qf_applied", nD), ("force", "force", nD), ("mass_mat", "M", nD * nD), ("mass_Dinv", "Dinv", nD),
              ("links_pos", "links_pos", (nL, 3)), ("cd_vel", "cd_vel", (nL, 3)), ("aabb_min", "aabb_min", (nG, 3)),
              ("aabb_max", "aabb_max", (nG, 3)), ("qf_constraint", "qf_constraint", nD), ("links_contact_force", "links_contact_force", (nL, 3))]
    for k in range(steps):
        scene.step()
        orc.step(1)
        if k in (0, 1, 9, steps - 1):
            torch.cuda.synchronize()
            print(f"== {name}: after step {k + 1}")
            compare(name, scene, orc, fields)
            nc_g = s.field("n_contacts")[:, 0].cpu().numpy()
            nc_o = orc.get("n_contacts", 1)[:, 0]
            ncon_g = s.field("n_constraints")[:, 0].cpu().numpy()
            ncon_o = orc.get("n_constraints", 3)[:, 0]
            it_g = s.field("solver_iters")[:, 0].cpu().numpy()
            it_o = orc.get("n_iter", 1)[:, 0]
            print(f"  n_contacts gpu={nc_g[:8]} orc={nc_o[:8].astype(int)}  mismatch envs={int((nc_g != nc_o).sum())}")
            print(f"  n_constraints gpu={ncon_g[:8]} orc={ncon_o[:8].astype(int)} mismatch envs={int((ncon_g != ncon_o).sum())}")
            print(f"  solver iters gpu={it_g[:8]} orc={it_o[:8].astype(int)}  errno={s.field('errno_')[:, 0].cpu().numpy()[:8]}")
    return scene


def bench(name, make, n_envs, steps=200, warm=20):
    scene, robot = make(n_envs)
    for _ in range(warm):
        scene.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        scene.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"## {name} n_envs={n_envs}: {ms:.4f} ms/step -> {n_envs / ms * 1e3 / 1e6:.2f} M env-steps/s  errno_max={int(scene.rigid_solver.field('errno_').max())}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run("franka", lambda n: make_franka(n, randomized=True), 64, 50)
    run("franka_fall", lambda n: make_franka(n, controlled=False), 32, 150)
    run("go2", lambda n: make_go2(n), 64, 100)
    for n in (4096, 30000):
        bench("franka_free", lambda n_: make_franka(n_), n)
        bench("franka_random", lambda n_: make_franka(n_, randomized=True), n)
    for n in (4096, 30000):
        bench("go2", lambda n_: make_go2(n_), n)
