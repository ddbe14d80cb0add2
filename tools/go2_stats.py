import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genesis_b200 as gs
from genesis_b200.bench_scenes import make_go2
gs.init(backend=gs.gpu)
scene, robot = make_go2(4096)
s = scene.rigid_solver
for k in range(121):
    scene.step()
    if k in (0, 5, 20, 60, 120):
        nc = s.field("n_constraints")[:, 0].float(); it = s.field("solver_iters")[:, 0].float(); ct = s.field("n_contacts")[:, 0].float(); nb = s.field("n_broad")[:, 0].float()
        q = torch.tensor([0.5, 0.9, 0.99, 1.0], device=nc.device)
        print(f"step {k}: n_con q50/90/99/max={torch.quantile(nc, q).tolist()} iters={torch.quantile(it, q).tolist()} contacts={torch.quantile(ct, q).tolist()} broad={torch.quantile(nb, q).tolist()} frac>64rows={(nc > 64).float().mean().item():.4f}")
