"""Where the host-buffer call (gsr_step_host) spends its time beyond the step kernel: Franka 30000, 300 calls per variant.
Usage: python tools/e2e_breakdown.py [n_envs]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka
gs.init(backend=gs.gpu)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
scene, robot = make_franka(n, randomized=True)
s = scene.rigid_solver
for _ in range(64):
    scene.step()
m = s.model
pin = lambda k: torch.empty((n, k), dtype=torch.float32).pin_memory().numpy()
ctrl, qpos, vel = pin(m.n_dofs), pin(m.n_qs), pin(m.n_dofs)
ctrl[:] = s.field("ctrl_pos").cpu().numpy()
def run(label, c, q, v, K=300):
    for _ in range(20):
        s.step_host(c, q, v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        s.step_host(c, q, v)
    dt = (time.perf_counter() - t0) / K
    print("%-52s %.1f us per call = %.1f M env-steps/s" % (label, 1e6 * dt, n / dt / 1e6))
run("targets in, qpos + vel out (the e2e leg)", ctrl, qpos, vel)
run("no targets (no H2D copy)", None, qpos, vel)
run("targets in, nothing out (no posted writes)", ctrl, None, None)
run("nothing in, nothing out (launch + synchronise)", None, None, None)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(300):
    scene.step()
ev1.record(); torch.cuda.synchronize()
print("%-52s %.1f us per step = %.1f M env-steps/s" % ("device-resident steps, no synchronisation", 1e3 * ev0.elapsed_time(ev1) / 300, n / (ev0.elapsed_time(ev1) / 300 * 1e-3) / 1e6))
