"""Timing sweep over launch tunings (GSR_GROUP_SIZE / GSR_SMEM_ROWS / GSR_DIAGNOSTICS are read at scene build)."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka, make_go2

gs.init(backend=gs.gpu)


def bench(name, make, n_envs, steps=100, warm=20):
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
    s = scene.rigid_solver
    print(f"## {name} n_envs={n_envs} G={s.launch_info['group_size']} smem/env={s.launch_info['smem_bytes_per_env']} rows={s.launch_info['smem_rows']} "
          f"diag={s.tuning['diagnostics']}: {ms:.4f} ms/step -> {n_envs / ms * 1e3 / 1e6:.2f} M env-steps/s errno={int(s.field('errno_').max())} "
          f"ncon_max={int(s.field('n_constraints').max())}", flush=True)
    scene.rigid_solver.destroy()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for G in (8, 16, 32):
        os.environ["GSR_GROUP_SIZE"] = str(G)
        if which in ("all", "franka"):
            bench("franka_random", lambda n: make_franka(n, randomized=True), 30000)
        if which in ("all", "go2"):
            bench("go2", lambda n: make_go2(n), 4096)
    os.environ["GSR_DIAGNOSTICS"] = "0"
    for G in (8, 32):
        os.environ["GSR_GROUP_SIZE"] = str(G)
        if which in ("all", "franka"):
            bench("franka_random", lambda n: make_franka(n, randomized=True), 30000)
        if which in ("all", "go2"):
            bench("go2", lambda n: make_go2(n), 4096)
