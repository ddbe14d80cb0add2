"""Timing of small-model scenes (free box / sphere on a plane) at large batch: lane-per-env kernel (G=1) vs lane groups."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs

gs.init(backend=gs.gpu)


def run(G, n_envs, kind, steps=200, warm=50):
    os.environ["GSR_GROUP_SIZE"] = str(G)
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01))
    scene.add_entity(gs.morphs.Plane())
    if kind == "box":
        scene.add_entity(gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0.0, 0.0, 0.3)))
    elif kind == "boxes3":
        for i in range(3):
            scene.add_entity(gs.morphs.Box(size=(0.1, 0.1, 0.1), pos=(0.3 * i, 0.0, 0.3)))
    else:
        scene.add_entity(gs.morphs.Sphere(radius=0.05, pos=(0.0, 0.0, 0.3)))
    scene.build(n_envs=n_envs)
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
    print(f"## {kind} n_envs={n_envs} G={s.launch_info['group_size']} smem/env={s.launch_info['smem_bytes_per_env']}: {ms:.4f} ms/step -> "
          f"{n_envs / ms * 1e3 / 1e6:.1f} M env-steps/s errno={int(s.field('errno_').max())} contacts={float(s.field('n_contacts').float().mean()):.2f}", flush=True)
    s.destroy()


for kind in ("sphere", "box", "boxes3"):
    for G in (1, 8, 32):
        run(G, 30000, kind)
