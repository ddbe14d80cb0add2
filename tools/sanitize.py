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
import numpy as np
import ref_cases
from helpers import make_box_box, make_capsules, make_franka, make_franka_grasp, make_go2, make_jumble, make_weld, write_tmp


def make_pile():
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, box_box_detection=True))
    scene.add_entity(gs.morphs.MJCF(file=write_tmp(ref_cases.BOXBOX_PILE_XML, ".xml")))
    scene.build(n_envs=5)
    return scene


def make_terrain():
    hf = ref_cases.terrain_height_field(rows=24, cols=24, seed=5)
    scene, robot = make_go2(5, seed=9, terrain=(hf, 0.25, 0.004, (-3.0, -3.0, 0.0)))
    q = scene.rigid_solver.field("qpos").clone()
    q[:, 2] += 0.05
    scene.rigid_solver.set_qpos(q)
    return scene


def host_io_and_reset(scene):
    """gsr_step_host with pinned buffers after device steps, masked set_state(partial) + deferred FK, on-demand diagnostics"""
    s = scene.rigid_solver
    m = s.model
    B = s._B
    ctrl = torch.zeros((B, m.n_dofs)).pin_memory().numpy()
    ctrl[:] = s.field("ctrl_pos").cpu().numpy()
    qh, vh = torch.zeros((B, m.n_qs)).pin_memory().numpy(), torch.zeros((B, m.n_dofs)).pin_memory().numpy()
    st = s.get_state()
    for _ in range(4):
        s.step_host(ctrl, qh, vh)
    mask = torch.zeros(B, dtype=torch.bool, device=gs.device)
    mask[::2] = True
    s.set_state(0, st, envs_idx=mask, partial=True)
    scene.step()
    s.get_mass_mat()
    s.get_links_pos()


for name, make, steps in (("go2", lambda: make_go2(9)[0], 30), ("franka-fall", lambda: make_franka(6, self_collision=True, controlled=False)[0], 60),
                          ("grasp", lambda: make_franka_grasp(3)[0], 12), ("capsules", lambda: make_capsules(5)[0], 25),
                          ("weld", lambda: make_weld(5)[0], 5), ("go2-noslip", lambda: make_go2(5, noslip_iterations=3)[0], 30),
                          ("boxbox-pile", make_pile, 10), ("boxbox-stack", lambda: make_box_box(4, box_box_detection=True)[0], 45),
                          ("terrain-go2", make_terrain, 25), ("jumble-gjk-first", lambda: make_jumble(4, use_gjk_collision=True)[0], 8),
                          ("franka-cg-euler", lambda: make_franka(6, randomized=True, constraint_solver=gs.constraint_solver.CG,
                                                                  integrator=gs.integrator.Euler)[0], 10),
                          ("franka-host-io", lambda: make_franka(6, randomized=True)[0], 6)):
    if which == "lane" and name in ("boxbox-pile", "boxbox-stack", "terrain-go2", "jumble-gjk-first"):
        continue  # rows that exist in the lane-group kernels only
    scene = make()
    for _ in range(steps):
        scene.step()
    if name == "franka-host-io":
        host_io_and_reset(scene)
    torch.cuda.synchronize()
    s = scene.rigid_solver
    print(f"[sanitize:{which}] {name}: G={s.launch_info['group_size']} contacts={int(s.field('n_contacts').sum())} errno={int(s.field('errno_').max())}", flush=True)
    s.destroy()
