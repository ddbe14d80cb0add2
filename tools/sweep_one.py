"""One timing point: python tools/sweep_one.py franka|go2 N_ENVS (launch tuning through GSR_* environment variables)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import bench, make_go2, make_franka
which, n = sys.argv[1], int(sys.argv[2])
if which == "franka":
    bench("franka_random", lambda k: make_franka(k, randomized=True), n, steps=200, warm=30)
else:
    bench("go2", lambda k: make_go2(k), n, steps=200, warm=60)
