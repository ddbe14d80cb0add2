"""Per-environment timeline of one step-kernel launch (gsr_debug_trace): how long each environment's step takes, when it
starts, and how busy the SMs are over the launch.  Usage: python tools/step_trace.py go2|franka [n_envs]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200 import _lib
from genesis_b200.bench_scenes import make_go2, make_franka, make_franka_grasp
gs.init(backend=gs.gpu)
which = sys.argv[1] if len(sys.argv) > 1 else "go2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else {"go2": 4096, "grasp": 8192}.get(which, 30000)
scene = {"go2": lambda: make_go2(n), "grasp": lambda: make_franka_grasp(n)}.get(which, lambda: make_franka(n, randomized=True))()[0]
s = scene.rigid_solver
for _ in range(int(os.environ.get("TRACE_WARMUP", 400))):
    scene.step()
L = s._L
_lib.check(s._handle, L.gsr_debug_trace(s._handle, 1))
for _ in range(3):
    scene.step()
torch.cuda.synchronize()
out = np.zeros((n, 8), dtype=np.uint64)
_lib.check(s._handle, L.gsr_debug_trace_read(s._handle, out.ctypes.data))
t0 = out[:, 0].min()
start, end, sm = (out[:, 0] - t0).astype(np.float64) / 1e3, (out[:, 5] - t0).astype(np.float64) / 1e3, out[:, 6].astype(int)
phases = np.diff(out[:, :6].astype(np.int64), axis=1) / 1e3  # dynamics, collision, constraints, integration, epilogue + kinematics
dur = end - start
print(f"{which} n_envs={n} G={s.launch_info['group_size']}: kernel span {end.max():.1f} us")
print("env step duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur.mean(), *np.percentile(dur, [50, 90, 99, 100])))
names = ["dynamics", "collision", "constraints", "integration", "epilogue+FK"]
for k, nm in enumerate(names):
    p = phases[:, k]
    print("  phase %-12s mean %6.1f p50 %6.1f p99 %6.1f max %6.1f us (incl. waiting at the CTA barrier of the previous phase)" % (nm, p.mean(), *np.percentile(p, [50, 99, 100])))
print("start time us: p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(start, [50, 90, 100])))
span = end.max()
edges = np.linspace(0, span, 11)
busy = []
for a, b in zip(edges[:-1], edges[1:]):
    mid = 0.5 * (a + b)
    act = (start <= mid) & (end >= mid)
    busy.append((int(act.sum()), len(set(sm[act]))))
print("resident envs / busy SMs at 10 instants:", busy)
nc = s.field("n_constraints")[:, 0].cpu().numpy()
print("corr(duration, n_constraints) = %.2f" % np.corrcoef(dur, nc)[0, 1])
order = np.argsort(-dur)[:5]
print("slowest envs (env, us, rows, iters, phases):", [(int(e), round(float(dur[e]), 1), int(nc[e]), int(s.field('solver_iters')[e, 0]), phases[e].round(0).tolist()) for e in order])
nb = s.field("n_broad")[:, 0].cpu().numpy()
print("n_broad mean %.1f max %d; corr(collision phase, n_broad) = %.2f" % (nb.mean(), nb.max(), np.corrcoef(phases[:, 1], nb)[0, 1]))
