import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import bench, make_go2, make_franka
for rows in (32, 64, 96, 128):
    os.environ["GSR_SMEM_ROWS"] = str(rows)
    os.environ["GSR_GROUP_SIZE"] = "32"
    bench("go2", lambda n: make_go2(n), 4096, steps=150, warm=60)
os.environ["GSR_SMEM_ROWS"] = "96"
bench("go2-30000", lambda n: make_go2(n), 30000, steps=50, warm=60)
