import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import bench, make_go2, make_franka
for G, W in ((32, 16), (32, 8), (32, 4), (16, 16), (16, 8)):
    os.environ["GSR_GROUP_SIZE"] = str(G); os.environ["GSR_WARPS"] = str(W)
    print("warps", W)
    bench("go2", lambda n: make_go2(n), 4096, steps=150, warm=60)
for G, W in ((16, 16), (16, 8), (16, 4), (8, 16), (8, 8), (8, 4), (32, 16)):
    os.environ["GSR_GROUP_SIZE"] = str(G); os.environ["GSR_WARPS"] = str(W)
    print("warps", W)
    bench("franka_random", lambda n: make_franka(n, randomized=True), 30000)
