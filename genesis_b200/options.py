"""Option containers of the rigid path.  Field names and defaults follow the reference's pydantic models
(genesis/options/solvers.py:17-69 SimOptions, :396-545 RigidOptions); unknown keywords are rejected like the
reference's `extra=forbid` (genesis/options/options.py:29)."""

from dataclasses import dataclass, fields

from .constants import constraint_solver as _cs
from .constants import integrator as _integ


class _Options:
    def __init__(self, **kwargs):
        names = {f.name for f in fields(self)}
        for k in kwargs:
            if k not in names:
                raise TypeError(f"{type(self).__name__}: unrecognized option '{k}'")
        for f in fields(self):
            setattr(self, f.name, kwargs.get(f.name, f.default))


@dataclass(init=False)
class SimOptions(_Options):
    dt: float = 1e-2
    substeps: int = 1
    gravity: tuple = (0.0, 0.0, -9.81)
    requires_grad: bool = False


@dataclass(init=False)
class RigidOptions(_Options):
    dt: float | None = None
    gravity: tuple | None = None
    enable_collision: bool = True
    enable_joint_limit: bool = True
    enable_self_collision: bool = True
    enable_neutral_collision: bool = False
    enable_adjacent_collision: bool = False
    disable_constraint: bool = False
    max_collision_pairs: int = 150
    multiplier_collision_broad_phase: int = 8
    integrator: int = _integ.approximate_implicitfast
    constraint_solver: int = _cs.Newton
    iterations: int = 50
    tolerance: float = 1e-6
    ls_iterations: int = 50
    ls_tolerance: float = 1e-2
    noslip_iterations: int = 0
    noslip_tolerance: float = 1e-6
    sparse_solve: bool = False
    constraint_timeconst: float = 0.01
    use_contact_island: bool = False
    box_box_detection: bool = False
    use_hibernation: bool = False
    max_dynamic_constraints: int = 8
    enable_multi_contact: bool = True
    enable_mujoco_compatibility: bool = False
    use_gjk_collision: bool | None = None
    batch_links_info: bool = False
    batch_joints_info: bool = False
    batch_dofs_info: bool = False
