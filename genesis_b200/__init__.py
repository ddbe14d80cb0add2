"""genesis_b200 -- B200-native batched rigid-body step behind the Genesis Python API surface.

Usage mirrors the reference (`import genesis as gs`, genesis/__init__.py:57-362):

    import genesis_b200 as gs
    gs.init(backend=gs.gpu)
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(dt=0.01, enable_self_collision=False))
    scene.add_entity(gs.morphs.Plane())
    franka = scene.add_entity(gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml"))
    scene.build(n_envs=30000)
    franka.control_dofs_position(q)
    scene.step()
"""

import numpy as np

from . import constants as _c
from . import morphs, options
from .constants import CTRL_MODE, EQUALITY_TYPE, GEOM_TYPE, JOINT_TYPE, constraint_solver, integrator
from .rigid_entity import Material
from .rigid_solver import GenesisException, RigidSolver
from .scene import Scene

__version__ = "0.1.0"

cpu, gpu, cuda = _c.backend.cpu, _c.backend.gpu, _c.backend.cuda
_initialized = False
backend = None
device = None
tc_float = None
tc_int = None
np_float = np.float32
np_int = np.int32
EPS = float(np.finfo(np.float32).eps)


class materials:
    Rigid = Material


def init(backend=None, precision="32", seed=None, debug=False, performance_mode=False, logging_level=None, eps=1e-15, **_ignored):
    """genesis/__init__.py:57: only the CUDA backend in fp32 exists here (no CPU fallback, no multi-backend dispatch)."""
    global _initialized, device, tc_float, tc_int, EPS
    import torch

    if backend is not None and backend == cpu:
        raise GenesisException("genesis_b200 only implements the CUDA (B200, sm_100a) backend; use the reference for gs.cpu.")
    if str(precision) != "32":
        raise GenesisException("genesis_b200 computes in fp32 only (the reference's GPU default).")
    globals()["backend"] = gpu
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    tc_float, tc_int = torch.float32, torch.int32
    EPS = float(max(eps, np.finfo(np.float32).eps))
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    _initialized = True


def destroy():
    global _initialized
    _initialized = False
