"""Small scenes for compute-sanitizer (memcheck / racecheck / synccheck): a few steps of Go2 (contacts, MPR, multi-contact
detections over lanes), the Franka grasp cluster (box-mesh MPR, SPILL rows) and capsules / weld / noslip, for the lane-group
kernels and the lane kernel.   compute-sanitizer --tool racecheck python tools/sanitize.py [group|lane]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
which = sys.argv[1] if len(sys.argv) > 1 else "group"
if which == "lane":
    os.environ["GSR_GROUP_SIZE"] = "1"
os.environ["GSR_REORDER_PERIOD"] = "2"
import torch
import genesis_b200 as gs
from helpers import make_capsules, make_franka, make_franka_grasp, make_go2, make_weld

for name, make, steps in (("go2", lambda: make_go2(9)[0], 30), ("franka-fall", lambda: make_franka(6, self_collision=True, controlled=False)[0], 60),
                          ("grasp", lambda: make_franka_grasp(3)[0], 12), ("capsules", lambda: make_capsules(5)[0], 25),
                          ("weld", lambda: make_weld(5)[0], 5), ("go2-noslip", lambda: make_go2(5, noslip_iterations=3)[0], 30)):
    scene = make()
    for _ in range(steps):
        scene.step()
    torch.cuda.synchronize()
    s = scene.rigid_solver
    print(f"[sanitize:{which}] {name}: G={s.launch_info['group_size']} contacts={int(s.field('n_contacts').sum())} errno={int(s.field('errno_').max())}", flush=True)
    s.destroy()
