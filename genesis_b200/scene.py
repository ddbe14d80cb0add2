"""Scene / Simulator facade for the rigid-only path (genesis/engine/scene.py:348-510 add_entity, :841-905 build,
:1017-1034 step, :958-996 reset; genesis/engine/simulator.py:272-296 step)."""

import torch

from . import morphs as _morphs
from .options import RigidOptions, SimOptions
from .rigid_solver import RATE_CHECK_ERRNO, GenesisException, RigidSolver


class Simulator:
    def __init__(self, scene, rigid_solver, options):
        self.scene, self.rigid_solver, self.options = scene, rigid_solver, options
        self.cur_substep_global = 0

    @property
    def dt(self):
        return self.options.dt

    @property
    def substeps(self):
        return self.options.substeps

    def step(self):
        # simulator.py:277-283: errno poll every RATE_CHECK_ERRNO substeps, then `substeps` rigid substeps
        if self.cur_substep_global > 0 and self.cur_substep_global % RATE_CHECK_ERRNO < self.options.substeps and self.scene._check_errno:
            self.rigid_solver.check_errno(deferred=True)
        self.rigid_solver.substep(self.options.substeps)
        self.cur_substep_global += self.options.substeps


class Scene:
    def __init__(self, sim_options=None, rigid_options=None, show_viewer=False, show_FPS=False, check_errno=True, **_ignored):
        import genesis_b200 as gs

        if not gs._initialized:
            raise GenesisException("Genesis hasn't been initialized. Did you call `gs.init()`?")
        self.sim_options = sim_options or SimOptions()
        self.rigid_options = rigid_options or RigidOptions()
        if self.rigid_options.dt is not None and sim_options is None:
            self.sim_options.dt = self.rigid_options.dt  # options.py:103-110 model_copy_from semantics
        self.rigid_solver = RigidSolver(self, self.sim_options, self.rigid_options)
        self.sim = Simulator(self, self.rigid_solver, self.sim_options)
        self._entities = []
        self._is_built = False
        self._check_errno = check_errno
        self._t = 0

    @property
    def entities(self):
        return self._entities

    @property
    def is_built(self):
        return self._is_built

    @property
    def dt(self):
        return self.sim_options.dt

    @property
    def t(self):
        return self._t

    def add_entity(self, morph, material=None, surface=None, visualize_contact=False, vis_mode=None, name=None):
        if self._is_built:
            raise GenesisException("Scene is already built.")
        if not isinstance(morph, _morphs.Morph):
            raise GenesisException(f"Unsupported morph type {type(morph)} (rigid path only).")
        entity = self.rigid_solver.add_entity(morph, material)
        self._entities.append(entity)
        return entity

    def build(self, n_envs=0, env_spacing=(0.0, 0.0), device=None, **_ignored):
        import genesis_b200 as gs

        self.n_envs = n_envs
        self._B = max(1, n_envs)
        device = device if device is not None else gs.device
        self.rigid_solver.build(n_envs, device)
        self._envs_idx = torch.arange(self._B, device=self.rigid_solver.device)
        self._is_built = True

    def step(self):
        if not self._is_built:
            raise GenesisException("Scene is not built yet.")
        self.sim.step()
        self._t += 1

    def reset(self, state=None, envs_idx=None):
        if state is not None:
            self.rigid_solver.set_state(0, state, envs_idx)
        else:
            self.rigid_solver.reset(envs_idx)
        if envs_idx is None:
            self._t = 0

    def get_state(self):
        return self.rigid_solver.get_state()

    def _sanitize_envs_idx(self, envs_idx):
        return self._envs_idx if envs_idx is None else torch.as_tensor(envs_idx, device=self._envs_idx.device)
