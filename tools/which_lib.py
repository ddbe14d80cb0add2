"""Prints the path of the library copy a benchmark workload runs with (generic or model-specialised): for the profiling tools."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genesis_b200 import _lib
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka, make_go2
gs.init(backend=gs.gpu, seed=0)
w = sys.argv[1]
scene = (make_go2(64)[0] if w == "go2" else make_franka(64, randomized=True)[0])
s = scene.rigid_solver
print(s._lib_instance if isinstance(s._lib_instance, str) else _lib.instance_path(s._lib_instance))
