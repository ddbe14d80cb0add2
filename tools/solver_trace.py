"""Where the constraint phase of the slow environments goes (developer build: make -C genesis_b200/csrc trace).
Usage: GSR_LIB=genesis_b200/csrc/libgsrigid_trace.so python tools/solver_trace.py go2|grasp [n_envs] [warmup]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200 import _lib
from genesis_b200.bench_scenes import make_go2, make_franka, make_franka_grasp
gs.init(backend=gs.gpu)
which = sys.argv[1] if len(sys.argv) > 1 else "go2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else {"go2": 4096, "grasp": 8192}.get(which, 30000)
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 400
scene = {"go2": lambda: make_go2(n), "grasp": lambda: make_franka_grasp(n)}.get(which, lambda: make_franka(n, randomized=True))()[0]
s = scene.rigid_solver
for _ in range(warm):
    scene.step()
torch.cuda.synchronize()
out = np.zeros((n, 16), dtype=np.uint64)
L = s._L
L.gsr_debug_solver_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
_lib.check(s._handle, L.gsr_debug_solver_read(s._handle, out.ctypes.data, n))
d = out.astype(np.float64)
cyc = d[:, :8] / 1965.0  # us at 1965 MHz
names = ["build_rows", "init: first gradient", "linesearch", "update_constraint", "hessian_incremental", "init: Cholesky of H", "update_gradient", "termination",
         "init: Ma, Jaref", "init: update_constraint", "init: H = M + J'DJ"]
cyc = np.concatenate([cyc, d[:, 13:16] / 1965.0], axis=1)
tot = cyc.sum(axis=1)
print(f"{which} n_envs={n} after {warm} steps: constraint phase own time us: mean {tot.mean():.1f} p50 {np.median(tot):.1f} p99 {np.percentile(tot, 99):.1f} max {tot.max():.1f}")
for k, nm in enumerate(names):
    print(f"  {nm:26s} mean {cyc[:, k].mean():7.2f}  p99 {np.percentile(cyc[:, k], 99):7.2f}  max {cyc[:, k].max():7.2f}")
print("  newton iterations mean %.2f max %d; line-search evaluations per iteration mean %.2f; flipped rows per iteration mean %.2f; degenerate downdates %d" % (
    d[:, 9].mean(), d[:, 9].max(), d[:, 8].sum() / max(1, d[:, 9].sum()), d[:, 10].sum() / max(1, d[:, 9].sum()), d[:, 11].sum()))
for e in np.argsort(-tot)[:8]:
    print(f"  env {e}: rows {int(d[e, 12])} iters {int(d[e, 9])} ls evals {int(d[e, 8])} flips {int(d[e, 10])} fallbacks {int(d[e, 11])}  us: " + " ".join(f"{x:.0f}" for x in cyc[e]))
