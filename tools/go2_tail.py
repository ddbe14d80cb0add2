"""What makes the straggler of a Go2 launch slow?  Per launch (gsr_debug_trace): kernel span, the balanced bound, and for the
environments with the longest collision / constraint phases their broad-phase pairs, contacts (geoms, penetration) and rows.
Usage: python tools/go2_tail.py [n_envs] [n_launches]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200 import _lib
from genesis_b200.bench_scenes import make_go2
gs.init(backend=gs.gpu)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene, robot = make_go2(n)
s = scene.rigid_solver
for _ in range(int(os.environ.get("TRACE_WARMUP", 400))):
    scene.step()
L = s._L
_lib.check(s._handle, L.gsr_debug_trace(s._handle, 1))
scene.step()
try:
    names = {robot.link_start + i: l.name for i, l in enumerate(robot.links)}
    geom_link = [names.get(int(l), "link %d" % int(l)) for l in np.asarray(s.model.geoms_link)]
except Exception as ex:  # names are a convenience only
    geom_link = [repr(ex)]
for it in range(n_launch):
    scene.step()
    torch.cuda.synchronize()
    out = np.zeros((n, 8), dtype=np.uint64)
    _lib.check(s._handle, L.gsr_debug_trace_read(s._handle, out.ctypes.data))
    t0 = out[:, 0].min()
    start, end = (out[:, 0] - t0).astype(np.float64) / 1e3, (out[:, 5] - t0).astype(np.float64) / 1e3
    ph = np.diff(out[:, :6].astype(np.int64), axis=1) / 1e3
    dur = end - start
    print("launch %d: span %.1f us, env step mean %.1f p99 %.1f max %.1f; collision > 2x median: %d envs, constraints > 2x median: %d envs" % (
        it, end.max(), dur.mean(), np.percentile(dur, 99), dur.max(), int((ph[:, 1] > 2 * np.median(ph[:, 1])).sum()), int((ph[:, 2] > 2 * np.median(ph[:, 2])).sum())))
    c = s.get_contacts(as_tensor=True, to_torch=False)
    nc = s.field("n_contacts")[:, 0].cpu().numpy(); nb = s.field("n_broad")[:, 0].cpu().numpy(); nr = s.field("n_constraints")[:, 0].cpu().numpy()
    its = s.field("solver_iters")[:, 0].cpu().numpy()
    # own work of an environment: phases of envs sharing a CTA are barrier-coupled, so rank by collision phase and de-duplicate by CTA start time
    for e in np.argsort(-ph[:, 1])[:3]:
        k = int(nc[e])
        pairs = sorted(set((int(a), int(b)) for a, b in zip(c["geom_a"][e][:k], c["geom_b"][e][:k])))
        pens = np.asarray(c["penetration"][e][:k])
        print("   env %5d collision %.0f us constraints %.0f us: n_broad %d contacts %d rows %d iters %d; pairs %s; pen max %.4f" % (
            e, ph[e, 1], ph[e, 2], nb[e], k, nr[e], its[e], pairs, pens.max() if k else 0.0))
if hasattr(L, "gsr_debug_narrow") or True:
    import ctypes
    try:
        fn = L.gsr_debug_narrow
    except AttributeError:
        fn = None
    if fn is not None:
        buf = (ctypes.c_ulonglong * 16)()
        fn(buf)  # clear
        scene.step(); torch.cuda.synchronize()
        fn(buf)
        v = list(buf)
        calls, inside = max(v[0], 1), max(v[6], 1)
        print("MPR queries of ONE launch: %d; cycles mean %.0f max %d; portal discovery cycles mean %.0f max %d" % (v[0], v[1] / calls, v[2], v[7] / calls, v[8]))
        print("  refinement iterations mean %.2f max %d; penetration iterations sum %d max %d" % (v[3] / calls, v[4], v[5], v[6]))
        print("  exact-penetration (GJK + EPA) queries %d, GJK cycles sum %d, vertex-pair candidates walked %d" % (v[9], v[10], v[11]))
        print("  narrow_batch: unperturbed pass max %d cycles (sum %d), perturbed pass max %d cycles (sum %d)" % (v[12], v[14], v[13], v[15]))
print("geoms:", {i: nm for i, nm in enumerate(geom_link)})
