"""Host cost of one Scene.step(): a scene so small that the GPU is never the bottleneck (developer tool)."""
import cProfile
import pstats
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import genesis_b200 as gs

gs.init(backend=gs.gpu, seed=0)
scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01))
scene.add_entity(gs.morphs.Plane())
scene.add_entity(gs.morphs.Sphere(radius=0.1, pos=(0, 0, 0.3)))
scene.build(n_envs=int(sys.argv[1]) if len(sys.argv) > 1 else 64)
for _ in range(200):
    scene.step()
torch.cuda.synchronize()
N = 5000
t0 = time.perf_counter()
for _ in range(N):
    scene.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host: {1e6 * (t1 - t0) / N:.1f} us per Scene.step() call (issue), {1e6 * (t2 - t0) / N:.1f} us per step incl. drain")
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    scene.step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
