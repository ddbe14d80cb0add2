"""Distribution of the per-environment solver load of the Go2 benchmark scene at steady state (which envs are the long pole)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_go2, make_franka
gs.init(backend=gs.gpu)
which = sys.argv[1] if len(sys.argv) > 1 else "go2"
scene, robot = (make_go2(4096) if which == "go2" else make_franka(30000, randomized=True))
s = scene.rigid_solver
for k in range(400):
    scene.step()
its, ncs, nct = [], [], []
for k in range(20):
    scene.step()
    its.append(s.field("solver_iters")[:, 0].cpu().numpy()); ncs.append(s.field("n_constraints")[:, 0].cpu().numpy()); nct.append(s.field("n_contacts")[:, 0].cpu().numpy())
its, ncs, nct = np.array(its), np.array(ncs), np.array(nct)
for name, a in (("solver_iters", its), ("n_constraints", ncs), ("n_contacts", nct)):
    print(name, "mean %.2f" % a.mean(), "p50 %d p90 %d p99 %d max %d" % tuple(np.percentile(a, [50, 90, 99, 100])))
cost = its * ncs
print("iters*rows mean %.1f p99 %.1f max %.1f" % (cost.mean(), np.percentile(cost, 99), cost.max()))
print("rows > on-chip capacity (%d): %.2f%% of env-steps" % (s.launch_info["smem_rows"], 100 * (ncs > s.launch_info["smem_rows"]).mean()))
print("hist iters", np.bincount(its.ravel())[:25])
