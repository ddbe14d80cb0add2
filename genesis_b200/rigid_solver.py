"""Host-side RigidSolver: the drop-in for genesis/engine/solvers/rigid/rigid_solver.py on the batched step path.

All state lives in two torch tensors (float32 / int32) on the CUDA device whose layout is dictated by libgsrigid
(include/gsrigid.h); this class only creates views, forwards `substep` to the C ABI and implements the accessor
surface (get_*/set_*/control_dofs_*) with the reference's argument conventions (rigid_solver.py:1902-2580,
kinematic_solver.py:940-1003): inputs broadcast to (len(envs_idx), len(idx)), `envs_idx=None` means all envs,
returned tensors are (B, n) copies on `gs.device`.  There is no CPU path: building a scene requires CUDA.
"""

import ctypes
import os

import numpy as np
import torch

from . import _lib
from .constants import CTRL_MODE, ErrorCode
from .model import ModelBuilder, box_l_infos, parse_mjcf, parse_urdf, plane_l_infos, sphere_l_infos, terrain_l_infos, cylinder_l_infos, capsule_l_infos
from .model.compiled import load_compiled, resolve_asset
from . import morphs as _morphs

RATE_CHECK_ERRNO = 10  # genesis/engine/simulator.py:47


class GenesisException(Exception):
    pass


def _with_masks(l_infos, morph):
    for l in l_infos:
        for g in l["geoms"]:
            g["contype"], g["conaffinity"] = int(morph.contype), int(morph.conaffinity)
    return l_infos


def _morph_to_infos(morph):
    """Returns (l_infos, eqs_info, is_robot, local_collision_mask)."""
    if isinstance(morph, (_morphs.Plane, _morphs.Box, _morphs.Sphere, _morphs.Terrain, _morphs.Cylinder, _morphs.Capsule)) and (morph.contype != 0xFFFF or morph.conaffinity != 0xFFFF):
        import copy

        plain = copy.copy(morph)
        plain.contype = plain.conaffinity = 0xFFFF
        l_infos, eqs, is_robot, local = _morph_to_infos(plain)
        return _with_masks(l_infos, morph), eqs, is_robot, local
    if isinstance(morph, _morphs.Plane):
        return plane_l_infos(morph.pos or (0, 0, 0), morph.quat or (1, 0, 0, 0), morph.normal, morph.plane_size), [], False, False
    if isinstance(morph, _morphs.Box):
        return box_l_infos(morph.size, morph.pos or (0, 0, 0), morph.quat or (1, 0, 0, 0), morph.fixed), [], False, False
    if isinstance(morph, _morphs.Terrain):
        return terrain_l_infos(morph.height_field, morph.horizontal_scale, morph.vertical_scale, morph.pos or (0, 0, 0),
                               morph.quat or (1, 0, 0, 0)), [], False, False
    if isinstance(morph, _morphs.Cylinder):
        return cylinder_l_infos(morph.radius, morph.height, morph.pos or (0, 0, 0), morph.quat or (1, 0, 0, 0), morph.fixed), [], False, False
    if isinstance(morph, _morphs.Capsule):
        return capsule_l_infos(morph.radius, morph.length, morph.pos or (0, 0, 0), morph.quat or (1, 0, 0, 0), morph.fixed), [], False, False
    if isinstance(morph, _morphs.Sphere):
        return sphere_l_infos(morph.radius, morph.pos or (0, 0, 0), morph.fixed), [], False, False
    if isinstance(morph, (_morphs.MJCF, _morphs.URDF)):
        is_mjcf = isinstance(morph, _morphs.MJCF)
        path = resolve_asset(morph.file)
        if path is not None:
            if is_mjcf:
                l_infos, eqs = parse_mjcf(path, default_armature=morph.default_armature, scale=morph.scale)
            else:
                l_infos, eqs = parse_urdf(path, merge_fixed_links=morph.merge_fixed_links, links_to_keep=morph.links_to_keep,
                                          default_armature=morph.default_armature, scale=morph.scale, fixed=morph.fixed)
        else:
            loaded = load_compiled(morph.file)
            if loaded is None:
                raise GenesisException(f"Asset '{morph.file}' not found and no pre-compiled model is bundled for it.")
            l_infos, eqs = loaded
        return l_infos, eqs, True, is_mjcf
    raise GenesisException(f"Unsupported morph: {morph!r}")


class RigidSolverState:
    """Dynamic state queried from a RigidSolver (genesis/engine/states/solvers.py:83-106): (B, n) tensors on the device."""

    FIELDS = ("qpos", "dofs_vel", "dofs_acc", "links_pos", "links_quat", "i_pos_shift", "mass_shift", "friction_ratio")

    def __init__(self, scene=None, s_global=0):
        self.scene, self._s_global = scene, s_global
        for k in self.FIELDS:
            setattr(self, k, None)

    @property
    def s_global(self):
        return self._s_global

    def __getitem__(self, key):  # the dict-style access of round 1 keeps working
        return getattr(self, key)

    def serializable(self):
        self.scene = None
        for k in self.FIELDS:
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.detach())


class RigidSolver:
    def __init__(self, scene, sim_options, rigid_options):
        self._scene = scene
        self._sim_options = sim_options
        self._options = rigid_options
        self._entities = []
        self._morphs = []
        self._is_built = False
        self._substep_dt = (rigid_options.dt or sim_options.dt) / max(1, sim_options.substeps)
        self._gravity = rigid_options.gravity or sim_options.gravity
        self._handle = None
        self._n_substeps_total = 0
        self._errno_host = self._errno_event = None
        self._obs_valid, self._obs, self._obs_have = False, None, set()
        self._fk_dirty_any = False
        self._kin_valid, self._kin_requested, self._eager_fk = True, False, True
        import os as _os
        # Adaptive kinematics (see _apply_kinematics_policy): a pure stepping loop keeps the kinematic cache out of HBM.  Default:
        # on with a model-specialised library copy (Franka 30000: 103 -> 118 M env-steps/s, a third of the DRAM traffic), off with
        # the generic library (measured slower there: profiles/README.md items 16 and 25); GSR_LAZY_FK=0 / 1 forces either.
        self._adaptive_fk = _os.environ.get("GSR_LAZY_FK", "0") == "1"
        self._errno_pending = False
        # launch tuning of libgsrigid (0 = library default): lanes per environment, constraint rows kept in shared
        # memory, and whether the diagnostic fields (mass_mat, qf_*, acc_smooth, geom poses) are written every step
        import os

        self.tuning = dict(group_size=int(os.environ.get("GSR_GROUP_SIZE", 0)), diagnostics=int(os.environ.get("GSR_DIAGNOSTICS", 0)))
        if os.environ.get("GSR_SMEM_ROWS"):
            self.tuning["smem_rows"] = int(os.environ["GSR_SMEM_ROWS"])
        if os.environ.get("GSR_REORDER_PERIOD"):
            self.tuning["reorder_period"] = int(os.environ["GSR_REORDER_PERIOD"])
        if os.environ.get("GSR_WARPS"):
            self.tuning["warps_per_cta"] = int(os.environ["GSR_WARPS"])
        if "GSR_HOST_IO" in os.environ:  # gsr_step_host: 0 staged + scatter / gather kernels, 1 fused device staging, 2 posted writes to pinned buffers
            self.tuning["host_io"] = int(os.environ["GSR_HOST_IO"])
        if "GSR_PHASE_BARRIERS" in os.environ:  # bit mask of the CTA barriers between the phases of k_step (sweeps)
            self.tuning["phase_barriers"] = int(os.environ["GSR_PHASE_BARRIERS"])
        if "GSR_GJK_FALLBACK" in os.environ:  # A/B measurements only: 0 keeps MPR's answer for deep / un-cached pairs
            self.tuning["gjk_fallback"] = int(os.environ["GSR_GJK_FALLBACK"])
        self.model = None

    # ------------------------------------------------------------------ build
    def add_entity(self, morph, material=None):
        from .rigid_entity import RigidEntity

        entity = RigidEntity(self, morph, material, len(self._entities))
        self._entities.append(entity)
        return entity

    def build_model(self):
        o = self._options
        if (o.batch_dofs_info or o.batch_links_info) and self.tuning.get("group_size") == 1:
            raise GenesisException("batch_dofs_info / batch_links_info need the lane-group step kernel (unset GSR_GROUP_SIZE=1).")
        if self.tuning.get("group_size") == 1 and (o.box_box_detection or any(isinstance(e.morph, _morphs.Terrain) for e in self._entities)):
            raise GenesisException("box_box_detection / Terrain need the lane-group step kernel (unset GSR_GROUP_SIZE=1).")
        if o.use_contact_island or o.use_hibernation or o.enable_mujoco_compatibility:
            raise GenesisException(
                "contact islands, hibernation and mujoco-compatibility are outside the hot-path scope of this build "
                "(DESIGN.md 'Out of scope')."
            )
        if o.use_gjk_collision and self.tuning.get("group_size") == 1:
            raise GenesisException("use_gjk_collision needs the lane-group step kernel (unset GSR_GROUP_SIZE=1).")
        b = ModelBuilder(
            dt=self._substep_dt, gravity=self._gravity, constraint_timeconst=o.constraint_timeconst,
            enable_self_collision=o.enable_self_collision, enable_adjacent_collision=o.enable_adjacent_collision,
            max_collision_pairs=o.max_collision_pairs, multiplier_collision_broad_phase=o.multiplier_collision_broad_phase,
            max_dynamic_constraints=o.max_dynamic_constraints,
        )
        b.enable_neutral_collision = o.enable_neutral_collision
        for ent in self._entities:
            l_infos, eqs, is_robot, local_mask = _morph_to_infos(ent.morph)
            rho = getattr(ent.material, "rho", None)
            friction = getattr(ent.material, "friction", None)
            file_morph = isinstance(ent.morph, (_morphs.MJCF, _morphs.URDF))  # primitives bake pos / quat into their l_infos
            ent._info = b.add_entity(type(ent.morph).__name__.lower(), l_infos, eqs, is_robot=is_robot, rho=rho,
                                     friction=friction, local_collision_mask=local_mask,
                                     morph_pos=ent.morph.pos if file_morph else None, morph_quat=ent.morph.quat if file_morph else None)
            ent._info.gravity_compensation = getattr(ent.material, "gravity_compensation", 0.0)
        self.model = b.build()
        return self.model

    def model_options(self):
        o = self._options
        return dict(iterations=o.iterations, ls_iterations=o.ls_iterations, tolerance=o.tolerance, ls_tolerance=o.ls_tolerance,
                    integrator=int(o.integrator), solver_type=int(o.constraint_solver), enable_collision=int(o.enable_collision),
                    enable_joint_limit=int(o.enable_joint_limit), enable_multi_contact=int(o.enable_multi_contact),
                    noslip_iterations=int(o.noslip_iterations), noslip_tolerance=float(o.noslip_tolerance),
                    batch_dofs_info=int(bool(o.batch_dofs_info)), disable_constraint=int(bool(o.disable_constraint)),
                    box_box_detection=int(bool(o.box_box_detection)), use_gjk_collision=int(bool(o.use_gjk_collision)),
                    sort_contacts=1, **self.tuning)

    def build(self, n_envs, device):
        if not torch.cuda.is_available():
            raise GenesisException("genesis_b200 has no CPU fallback: a CUDA device (B200, sm_100a) is required to build a scene.")
        self.n_envs = n_envs
        self._B = max(1, n_envs)
        self.device = torch.device(device)
        m = self.model if self.model is not None else self.build_model()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._lib_instance = _lib.acquire_instance(dev_index, self)  # its own CUDA module (constant memory) when another scene is alive
        self._lib_device = dev_index
        L = self._L = _lib.lib(self._lib_instance)

        def describe(lib_):
            h_ = ctypes.c_void_p()
            _lib.check(None, lib_.gsr_model_begin(ctypes.byref(h_)))
            self._handle, self._L = h_, lib_
            self._upload_model_arrays()
            for k, v in self.model_options().items():
                lib_.gsr_model_set_scalar(h_, f"cfg_{k}".encode(), float(v))
            return h_

        h = describe(L)
        # a copy of the library compiled for exactly this model and option set (tools/specialize.py; layouts and counts are literals
        # there) is used when it was built: the generic library describes the model, the description names the file
        self.specialised = False
        if os.environ.get("GSR_NO_SPEC", "0") != "1" and not os.environ.get("GSR_LIB"):
            n = L.gsr_model_spec_header(h, self._B, None, 0)
            if n > 0:
                buf = ctypes.create_string_buffer(n)
                L.gsr_model_spec_header(h, self._B, buf, n)
                path = _lib.spec_path(buf.value.decode())
                if os.path.exists(path):
                    Ls = _lib.lib_at(path)
                    L.gsr_destroy(h)
                    _lib.release_instance(dev_index, self._lib_instance, self)
                    self._lib_instance = path  # a specialised copy is its own CUDA module as well
                    h = describe(Ls)
                    L = Ls
                    self.specialised = True
        if "GSR_LAZY_FK" not in os.environ:
            self._adaptive_fk = self.specialised
        _lib.check(h, L.gsr_model_commit(h, dev_index, self._B))
        nw = ctypes.c_size_t()
        _lib.check(h, L.gsr_state_words(h, ctypes.byref(nw)))
        with torch.cuda.device(self.device):
            self._F = torch.zeros(nw.value, dtype=torch.float32, device=self.device)
        self._I = self._F.view(torch.int32)
        _lib.check(h, L.gsr_bind(h, self._F.data_ptr()))
        self._fields = {}
        name, is_int = ctypes.c_char_p(), ctypes.c_int()
        v = [ctypes.c_size_t() for _ in range(6)]
        for i in range(L.gsr_field_count(h)):
            _lib.check(h, L.gsr_field_info(h, i, ctypes.byref(name), ctypes.byref(is_int), *[ctypes.byref(x) for x in v]))
            seg, stride, off, items, comps, istride = (x.value for x in v)
            buf = self._I if is_int.value else self._F
            shape, strides = ((self._B, items), (stride, istride)) if comps == 1 else ((self._B, items, comps), (stride, istride, 1))
            self._fields[name.value.decode()] = torch.as_strided(buf, shape, strides, seg + off)
        _lib.check(h, L.gsr_init_state(h, self._stream()))
        self._is_built = True
        self.launch_info_cached = self.launch_info
        self._init_qpos = self._fields["qpos"].clone()  # (B, nQ)
        self._envs_all = torch.arange(self._B, device=self.device)
        m_ = self.model
        self._dofs_limit = torch.tensor(m_.dofs_limit, dtype=torch.float32, device=self.device).reshape(-1, 2)

    def _upload_model_arrays(self):
        L, h, m = self._L, self._handle, self.model
        self._model_keepalive = []
        for k, v in m.__dict__.items():
            if isinstance(v, np.ndarray):
                if v.dtype.kind == "f":
                    a = np.ascontiguousarray(v, dtype=np.float32)
                    L.gsr_model_set_f32(h, k.encode(), a.ctypes.data, a.size)
                elif v.dtype.kind in "iub":
                    a = np.ascontiguousarray(v, dtype=np.int32)
                    L.gsr_model_set_i32(h, k.encode(), a.ctypes.data, a.size)
        for k in ("n_links", "n_joints", "n_dofs", "n_qs", "n_geoms", "n_entities", "n_equalities", "n_pairs", "n_verts",
                  "max_contact_pairs", "max_broad_pairs", "n_contacts_per_pair", "len_constraints", "meaninertia", "dt"):
            L.gsr_model_set_scalar(h, k.encode(), float(getattr(m, k)))
        for i, ax in enumerate("xyz"):
            L.gsr_model_set_scalar(h, f"gravity_{ax}".encode(), float(m.gravity[i]))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # fields that are a function of (qpos, vel): valid in the records only after an eager step or a forward-kinematics pass
    KIN_FIELDS = frozenset(("dofs_pos", "root_COM", "cdof_ang", "cdof_vel", "cdofd_ang", "cdofd_vel", "links_pos", "links_quat", "cinr_I",
                            "cinr_pos", "cd_vel", "cd_ang"))

    def _ensure_kinematics(self):
        """Lazy kinematics (see `substep`): the records hold no link poses / velocities after a lazy step; the first accessor
        that needs them runs one forward-kinematics pass over all environments and makes the next steps eager."""
        self._kin_requested = True
        if not self._kin_valid:
            _lib.check(self._handle, self._L.gsr_forward_kinematics(self._handle, None, self._stream()))
            self._kin_valid = True
            self._fk_dirty_any = False

    def field(self, name):
        """Zero-copy torch view of a state field, (B, items) or (B, items, comps): the reference's transposed zero-copy
        qd_to_torch views (genesis/utils/misc.py:415-536) are plain strided views of the env-major records here."""
        if name in self.KIN_FIELDS:
            self._ensure_kinematics()
        return self._fields[name]

    def rows(self, name):
        """(B, items * comps) copy-free when possible: every scalar of a field, one row per environment."""
        return self.field(name).reshape(self._B, -1)

    # ------------------------------------------------------------------ stepping
    def _flush_deferred_fk(self):
        if self._fk_dirty_any and self._kin_valid:
            _lib.check(self._handle, self._L.gsr_forward_kinematics_dirty(self._handle, self._stream()))
        self._fk_dirty_any = False  # with a stale cache the step recomputes the kinematics of every environment anyway

    def _apply_kinematics_policy(self):
        """Adaptive kinematics: a step leaves link poses / velocities in the records (eager, what the reference does at the end of
        kernel_step_2) only if an accessor asked for them since the previous step -- an RL loop reading observations every step
        stays eager, a pure stepping loop runs lazy: the same arithmetic, but the kinematic cache never travels through HBM.
        Default with a model-specialised library copy; the generic library keeps the reference's eager behaviour (GSR_LAZY_FK=0 / 1 forces either)."""
        eager = self._kin_requested or not self._adaptive_fk or bool(self.tuning.get("diagnostics")) or self.launch_info_cached["group_size"] == 1
        if eager != self._eager_fk:
            _lib.check(self._handle, self._L.gsr_set_kinematics_policy(self._handle, int(eager)))
            self._eager_fk = eager
        self._kin_requested = False
        return eager

    def substep(self, n=1):
        self._flush_deferred_fk()
        eager = self._apply_kinematics_policy()
        _lib.check(self._handle, self._L.gsr_step(self._handle, n, self._stream()))
        self._kin_valid = eager
        self._n_substeps_total += n
        self._obs_valid = False

    @property
    def launch_info(self):
        """Launch geometry of the step kernel as of now (the on-chip row capacity follows the workload, gsr_api.cu adapt_rows)."""
        gi = [ctypes.c_int() for _ in range(4)]
        _lib.check(self._handle, self._L.gsr_launch_info(self._handle, *[ctypes.byref(x) for x in gi]))
        return dict(group_size=gi[0].value, smem_bytes_per_env=gi[1].value, smem_rows=gi[2].value, smem_contacts=gi[3].value)

    def step_host(self, ctrl_pos_host, qpos_host, vel_host, n=1):
        """Reference-facing host-buffer call (bench.py e2e): numpy (B, n) float32 arrays in / out."""
        p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
        self._flush_deferred_fk()
        eager = self._apply_kinematics_policy()
        self._kin_valid = eager
        _lib.check(self._handle, self._L.gsr_step_host(self._handle, n, p(ctrl_pos_host), p(qpos_host), p(vel_host), self._stream()))
        self._n_substeps_total += n
        self._obs_valid = False

    def launch_count(self):
        return int(self._L.gsr_launch_count(self._handle))

    def kernel_launches(self):
        out = {}
        name, cnt = ctypes.c_char_p(), ctypes.c_longlong()
        i = 0
        while self._L.gsr_kernel_info(self._handle, i, ctypes.byref(name), ctypes.byref(cnt)) == 0:
            out[name.value.decode()] = cnt.value
            i += 1
        return out

    def set_profiling(self, enabled):
        _lib.check(self._handle, self._L.gsr_set_profiling(self._handle, int(bool(enabled))))

    def kernel_times(self):
        """{kernel name: (total ms, timed launches)} recorded with CUDA events while profiling was enabled."""
        out = {}
        name, cnt, ms, n = ctypes.c_char_p(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_longlong()
        i = 0
        while self._L.gsr_kernel_info(self._handle, i, ctypes.byref(name), ctypes.byref(cnt)) == 0:
            _lib.check(self._handle, self._L.gsr_kernel_time_ms(self._handle, i, ctypes.byref(ms), ctypes.byref(n)))
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out

    def get_error_envs_mask(self):
        return self._fields["errno_"][:, 0] > 0

    def check_errno(self, deferred=False):
        """rigid_solver.py:985-1009 (polled every RATE_CHECK_ERRNO substeps by Scene.step).

        `deferred=True` (the poll of Simulator.step) never stalls the stream: it evaluates the OR-reduction enqueued by the
        previous poll (complete long ago: it was recorded RATE_CHECK_ERRNO substeps earlier) and enqueues the next one into a
        pinned word, so an error surfaces one poll period later than in the reference and the step loop has no host sync."""
        if not self._B:
            return
        if deferred:
            if self._errno_host is None:
                self._errno_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._errno_event = torch.cuda.Event()
                self._errno_pending = False
            bits = 0
            if self._errno_pending:
                self._errno_event.synchronize()
                bits = int(self._errno_host[0])
            # errno bits are OR-ed per env; max > 0 flags an error, the exact bits are gathered only on the error path
            self._errno_host.copy_(self._fields["errno_"].max().reshape(1), non_blocking=True)
            self._errno_event.record(torch.cuda.current_stream(self.device))
            self._errno_pending = True
            if bits == 0:
                return
        elif int(self._fields["errno_"].max().item()) == 0:
            return
        bits = int(np.bitwise_or.reduce(self._fields["errno_"][:, 0].cpu().numpy()))
        if bits & ErrorCode.OVERFLOW_CANDIDATE_CONTACTS:
            raise GenesisException(f"Exceeding max number of broad phase candidate contact pairs ({self.model.max_broad_pairs}). "
                                   "Please increase the value of RigidSolver's option 'multiplier_collision_broad_phase'.")
        if bits & ErrorCode.OVERFLOW_COLLISION_PAIRS:
            raise GenesisException(f"Exceeding max number of contact pairs ({self.model.max_contact_pairs}). Please increase the "
                                   "value of RigidSolver's option 'max_collision_pairs'.")
        if bits & ErrorCode.INVALID_FORCE_NAN:
            raise GenesisException("Invalid constraint forces causing 'nan'. Please decrease Rigid simulation timestep.")
        if bits & ErrorCode.INVALID_ACC_NAN:
            raise GenesisException("Invalid accelerations causing 'nan'. Please decrease Rigid simulation timestep.")

    # ------------------------------------------------------------------ index helpers
    def _envs(self, envs_idx):
        if envs_idx is None:
            return None
        if isinstance(envs_idx, torch.Tensor) and envs_idx.dtype == torch.bool:
            return torch.nonzero(envs_idx.to(self.device), as_tuple=False)[:, 0]
        return torch.as_tensor(envs_idx, device=self.device, dtype=torch.long).reshape(-1)

    def _idx(self, idx, n):
        if idx is None:
            return torch.arange(n, device=self.device)
        if isinstance(idx, slice):
            return torch.arange(n, device=self.device)[idx]
        if isinstance(idx, range):
            return torch.as_tensor(list(idx), device=self.device, dtype=torch.long)
        return torch.as_tensor(idx, device=self.device, dtype=torch.long).reshape(-1)

    # fields the observation getters of an RL loop read: gathered into dense (B, n) tensors by ONE launch the first time any of them
    # is asked for after a step / setter (k_gather_fields); the getters then return copies of those dense tensors
    OBS_STATE = ("qpos", "vel", "acc", "force", "qf_applied", "cacc_lin", "cacc_ang", "links_contact_force")
    OBS_KIN = ("dofs_pos", "links_pos", "links_quat", "cd_vel", "cd_ang", "root_COM")
    OBS_FIELDS = OBS_STATE + OBS_KIN

    def _observed(self, name):
        """Dense (B, n) snapshot of an observation field.  Two groups, each gathered by ONE launch the first time one of its fields
        is asked for after a step / setter: plain state (never forces kinematics) and kinematic quantities (forces a valid cache)."""
        if name not in self.OBS_FIELDS:
            return self.field(name)
        if not self._obs_valid:
            self._obs_have = set()
            self._obs_valid = True
        if name not in self._obs_have:
            if self._kin_valid and not self._obs_have:
                group = self.OBS_FIELDS  # the records hold everything: one launch serves every getter of this step
            else:
                group = self.OBS_KIN if name in self.OBS_KIN else self.OBS_STATE
            self._obs = {**(self._obs or {}), **self.gather_fields(list(group), out=self._obs)}
            self._obs_have.update(group)
        return self._obs[name]

    def _get(self, name, idx, envs_idx, n, comps=1):
        """(B, len(idx)[, comps]) copy of a field."""
        f = self._observed(name)
        view = idx is None or isinstance(idx, slice)
        out = f if idx is None else (f[:, idx] if view else f[:, self._idx(idx, n)])
        if envs_idx is None:
            out = out.clone() if view else out  # an index_select result is a copy already
        elif self._is_mask(envs_idx):
            out = out[envs_idx.to(self.device)]
        else:
            out = out[self._envs(envs_idx)]
        if self.n_envs == 0:
            out = out[0]
        return out

    def _is_mask(self, envs_idx):
        return isinstance(envs_idx, torch.Tensor) and envs_idx.dtype == torch.bool

    def _set(self, name, value, idx, envs_idx, n, dtype=torch.float32):
        f = self._fields[name]
        full = idx is None or (isinstance(idx, slice) and len(range(*idx.indices(n))) == n)
        if isinstance(idx, slice) and self._envs_is_all_or_mask(envs_idx):
            n_sel = len(range(*idx.indices(n)))  # contiguous columns: plain strided writes, no index tensor
        else:
            idx = self._idx(idx, n)
            n_sel = len(idx)
        value = torch.as_tensor(value, dtype=dtype, device=self.device)
        if self._is_mask(envs_idx):  # boolean env mask: a masked select on the device, no index extraction (no host sync)
            mask = envs_idx.to(self.device)
            cur = f if full else f[:, idx]
            value = torch.broadcast_to(value, cur.shape).to(f.dtype)
            new = torch.where(mask.reshape((-1,) + (1,) * (cur.dim() - 1)), value, cur)
            if full:
                f.copy_(new)
            else:
                f[:, idx] = new
            return
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        value = torch.broadcast_to(value, (nb, n_sel)).to(f.dtype)
        if envs is None:
            if full:
                f.copy_(value)
            else:
                f[:, idx] = value
        else:
            f[envs[:, None], idx[None, :]] = value

    def _envs_is_all_or_mask(self, envs_idx):
        return envs_idx is None or self._is_mask(envs_idx)

    def _mark_fk_dirty(self, envs_idx):
        """qpos / vel written with skip_forward=True: the next step runs the forward kinematics of these environments first
        (the reference's is_forward_pos_updated = False, rigid_solver.py:2738-2765)."""
        d = self._fields["fk_dirty"]
        self._fk_dirty_any = True
        if envs_idx is None:
            d.fill_(1)
        elif self._is_mask(envs_idx):
            d[:, 0] = torch.where(envs_idx.to(self.device), torch.ones_like(d[:, 0]), d[:, 0])
        else:
            d[self._envs(envs_idx)] = 1

    def _env_mask(self, envs_idx):
        if self._is_mask(envs_idx):
            return envs_idx.to(self.device).to(torch.uint8)
        envs = self._envs(envs_idx)
        if envs is None:
            return None
        mask = torch.zeros(self._B, dtype=torch.uint8, device=self.device)
        mask[envs] = 1
        return mask

    def _forward(self, envs_idx=None):
        if envs_idx is None:
            self._kin_valid = True
        elif not self._kin_valid:
            return  # the cache is stale for every environment: the next step (or the next accessor) recomputes all of it
        mask = self._env_mask(envs_idx)
        ptr = ctypes.c_void_p(mask.data_ptr()) if mask is not None else None
        _lib.check(self._handle, self._L.gsr_forward_kinematics(self._handle, ptr, self._stream()))

    # ------------------------------------------------------------------ setters (rigid_solver.py:1902-2287)
    def set_qpos(self, qpos, qs_idx=None, envs_idx=None, *, skip_forward=False):
        self._set("qpos", qpos, qs_idx, envs_idx, self.model.n_qs)
        self._reset_warmstart(envs_idx)
        self._obs_valid = False
        if not skip_forward:
            self._forward(envs_idx)
        else:
            self._mark_fk_dirty(envs_idx)

    def set_dofs_velocity(self, velocity, dofs_idx=None, envs_idx=None, *, skip_forward=False):
        if velocity is None:
            velocity = 0.0
        self._set("vel", velocity, dofs_idx, envs_idx, self.model.n_dofs)
        self._obs_valid = False
        if not skip_forward:
            self._forward(envs_idx)
        else:
            self._mark_fk_dirty(envs_idx)

    def set_base_link_pose(self, link_idx, value, field, envs_idx=None):
        """kernel_set_links_pos / kernel_set_links_quat for a FIXED root link (abd/accessor.py): its world pose is state, FK of
        the subtree follows from it."""
        f = self._fields[field]
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        v = torch.broadcast_to(torch.as_tensor(value, dtype=torch.float32, device=self.device), (nb, f.shape[-1]))
        if field == "links_quat":
            v = v / v.norm(dim=-1, keepdim=True)
        if envs is None:
            f[:, link_idx] = v
        else:
            f[envs, link_idx] = v
        self._forward(envs_idx)

    def set_dofs_position(self, position, dofs_idx=None, envs_idx=None):
        """dofs_state.pos -> qpos (rigid_solver.py set_dofs_position / kernel_set_dofs_position)."""
        from .constants import JOINT_TYPE

        m = self.model
        dofs = self._idx(dofs_idx, m.n_dofs)
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        pos = torch.broadcast_to(torch.as_tensor(position, dtype=torch.float32, device=self.device), (nb, len(dofs)))
        cur = self.field("dofs_pos").clone()  # (B, nD)
        sel = cur if envs is None else cur[envs]
        sel[:, dofs] = pos
        qpos = self._fields["qpos"].clone()
        qsel = qpos if envs is None else qpos[envs]
        for j in range(m.n_joints):
            t, qs, ds = int(m.joints_type[j]), int(m.joints_q_start[j]), int(m.joints_dof_start[j])
            if t in (JOINT_TYPE.REVOLUTE, JOINT_TYPE.PRISMATIC):
                qsel[:, qs] = sel[:, ds] + float(m.qpos0[qs])
            elif t == JOINT_TYPE.FREE:
                qsel[:, qs: qs + 3] = sel[:, ds: ds + 3]
                qsel[:, qs + 3: qs + 7] = _xyz_to_quat_torch(sel[:, ds + 3: ds + 6])
        self.set_qpos(qsel, None, envs_idx)

    def _reset_warmstart(self, envs_idx):
        mask = self._env_mask(envs_idx)
        ptr = ctypes.c_void_p(mask.data_ptr()) if mask is not None else None
        _lib.check(self._handle, self._L.gsr_reset_envs(self._handle, ptr, self._stream()))

    def control_dofs_position(self, position, dofs_idx=None, envs_idx=None):
        n = self.model.n_dofs
        self._set("ctrl_mode", int(CTRL_MODE.POSITION), dofs_idx, envs_idx, n, dtype=torch.int32)
        self._set("ctrl_pos", position, dofs_idx, envs_idx, n)
        self._set("ctrl_vel", 0.0, dofs_idx, envs_idx, n)

    def control_dofs_position_velocity(self, position, velocity, dofs_idx=None, envs_idx=None):
        n = self.model.n_dofs
        self._set("ctrl_mode", int(CTRL_MODE.POSITION), dofs_idx, envs_idx, n, dtype=torch.int32)
        self._set("ctrl_pos", position, dofs_idx, envs_idx, n)
        self._set("ctrl_vel", velocity, dofs_idx, envs_idx, n)

    def control_dofs_velocity(self, velocity, dofs_idx=None, envs_idx=None):
        n = self.model.n_dofs
        self._set("ctrl_mode", int(CTRL_MODE.VELOCITY), dofs_idx, envs_idx, n, dtype=torch.int32)
        self._set("ctrl_vel", velocity, dofs_idx, envs_idx, n)

    def control_dofs_force(self, force, dofs_idx=None, envs_idx=None):
        n = self.model.n_dofs
        self._set("ctrl_mode", int(CTRL_MODE.FORCE), dofs_idx, envs_idx, n, dtype=torch.int32)
        self._set("ctrl_force", force, dofs_idx, envs_idx, n)

    def apply_links_external_force(self, force, links_idx=None, envs_idx=None, *, ref="link_origin", local=False):
        """External linear force on a set of links (rigid_solver.py:1110-1150, abd/misc.py:692-712): accumulated into the c-frame
        wrench about the subtree COM, `cfrc_applied_vel -= F`, `cfrc_applied_ang -= r x F` with r = the link origin
        (`ref="link_origin"`, the default), the link COM (`"link_com"`) or zero (`"root_com"`), all relative to root_COM.
        `local=True` expresses F in the link's inertial frame (the reference rotates by links_state.i_quat for both refs)."""
        if ref not in ("link_origin", "link_com", "root_com"):
            raise GenesisException("'ref' must be one of 'link_origin', 'link_com', 'root_com'.")
        if ref == "root_com" and local:
            raise ValueError("'local=True' not compatible with ref='root_com'.")
        m = self.model
        links = self._idx(links_idx, m.n_links)
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        force = torch.broadcast_to(torch.as_tensor(force, dtype=torch.float32, device=self.device), (nb, len(links), 3))
        sel = (lambda f: f[:, links]) if envs is None else (lambda f: f[envs][:, links])
        if local:
            iq = torch.tensor(m.links_inertial_quat, dtype=torch.float32, device=self.device)[links]
            quat = sel(self.field("links_quat"))
            force = _rotate_torch(force, _quat_mul_torch(quat, iq[None].expand(quat.shape[0], -1, -1)))
        root = sel(self.field("root_COM"))
        if ref == "link_com":
            arm = self._get_links_com(links, envs) - root
        elif ref == "link_origin":
            arm = sel(self.field("links_pos")) - root
        else:
            arm = torch.zeros_like(force)
        torque = torch.cross(arm, force, dim=-1)
        e = self._envs_all if envs is None else envs
        self._fields["cfrc_applied_vel"][e[:, None], links[None, :]] -= force
        self._fields["cfrc_applied_ang"][e[:, None], links[None, :]] -= torque
        self._fields["has_ext_force"][e] = 1

    def apply_links_external_torque(self, torque, links_idx=None, envs_idx=None, *, ref="link_origin", local=False):
        """World-frame (or, with local=True, link-frame) torque on each link (rigid_solver.py:1152-1193, abd/misc.py:720-734):
        cfrc_applied_ang -= torque."""
        m = self.model
        links = self._idx(links_idx, m.n_links)
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        torque = torch.broadcast_to(torch.as_tensor(torque, dtype=torch.float32, device=self.device), (nb, len(links), 3))
        if local:
            quat = self.field("links_quat")[:, links]
            quat = quat if envs is None else quat[envs]
            if ref == "link_com":  # inertial frame: link quat (x) inertial quat
                iq = torch.tensor(m.links_inertial_quat, dtype=torch.float32, device=self.device)[links]
                quat = _quat_mul_torch(quat, iq[None].expand(quat.shape[0], -1, -1))
            if ref in ("link_origin", "link_com"):
                torque = _rotate_torch(torque, quat)
        e = self._envs_all if envs is None else envs
        self._fields["cfrc_applied_ang"][e[:, None], links[None, :]] -= torque
        self._fields["has_ext_force"][e] = 1

    def _get_links_com(self, links, envs):
        m = self.model
        pos = self.field("links_pos")[:, links]
        quat = self.field("links_quat")[:, links]
        if envs is not None:
            pos, quat = pos[envs], quat[envs]
        ipos = torch.tensor(m.links_inertial_pos, dtype=torch.float32, device=self.device)[links]
        return pos + _rotate_torch(ipos[None].expand(pos.shape[0], -1, -1), quat)

    # per-dof model parameters (not batched): refresh the constant-memory model
    def _set_model_array(self, name, values, dofs_idx, width=None):
        m = self.model
        arr = getattr(m, name)
        idx = np.arange(m.n_dofs)[dofs_idx] if isinstance(dofs_idx, slice) else self._idx(dofs_idx, m.n_dofs).cpu().numpy()
        vals = np.asarray(torch.as_tensor(values).detach().cpu().numpy(), dtype=np.float64)
        if vals.ndim >= 1 and self.n_envs > 0 and vals.ndim == (2 if width is None else 3):
            vals = vals[0]  # unbatched parameters: envs must agree
        cur = arr[idx] if width is None else arr[idx, width]
        if np.array_equal(np.broadcast_to(vals, np.shape(cur)).astype(np.float32), np.asarray(cur, np.float32)):
            return  # unchanged (an RL loop re-applying its gains every step): no model upload
        if width is None:
            arr[idx] = vals
        else:
            arr[idx, width] = vals
        a = np.ascontiguousarray(arr, dtype=np.float32)
        L = self._L
        L.gsr_model_set_f32(self._handle, name.encode(), a.ctypes.data, a.size)
        _lib.check(self._handle, L.gsr_model_refresh(self._handle))

    def set_geoms_sol_params(self, sol_params, geoms_idx=None):
        """RigidGeom.set_sol_params (rigid_geom.py; sanitised like rigid_solver.py:187-218: timeconst >= 2 dt, impedances in
        [1e-4, 1 - 1e-4], width >= 0, power >= 1).  Contacts take the mean of both geoms' parameters (contact.py:181-222)."""
        from .model.builder import sanitize_sol_params

        m = self.model
        idx = np.arange(m.n_geoms) if geoms_idx is None else np.atleast_1d(np.asarray(geoms_idx, dtype=np.int64))
        vals = np.asarray(torch.as_tensor(sol_params).detach().cpu().numpy(), dtype=np.float64)
        sp = sanitize_sol_params(np.broadcast_to(vals, (len(idx), 7)), 2.0 * self._substep_dt,
                                 max(self._options.constraint_timeconst, 2.0 * self._substep_dt))
        m.geoms_sol_params[idx] = sp
        a = np.ascontiguousarray(m.geoms_sol_params, dtype=np.float32)
        L = self._L
        L.gsr_model_set_f32(self._handle, b"geoms_sol_params", a.ctypes.data, a.size)
        _lib.check(self._handle, L.gsr_model_refresh(self._handle))

    def get_geoms_sol_params(self, geoms_idx=None):
        m = self.model
        idx = np.arange(m.n_geoms) if geoms_idx is None else np.atleast_1d(np.asarray(geoms_idx, dtype=np.int64))
        return torch.tensor(m.geoms_sol_params[idx], dtype=torch.float32, device=self.device)

    # per-dof parameters: model constants shared by every environment (default), or -- RigidOptions.batch_dofs_info -- fields of each
    # environment's record (dofs_info of shape (n_dofs, B) in the reference, array_class.py:1347-1364)
    @property
    def _batch_dofs(self):
        return bool(self._options.batch_dofs_info)

    def _set_dof_param(self, field, model_name, values, dofs_idx, envs_idx, width=None, comps=1):
        if not self._batch_dofs:
            if envs_idx is not None:
                raise GenesisException("Per-environment dof parameters need RigidOptions(batch_dofs_info=True).")
            return self._set_model_array(model_name, values, dofs_idx, width)
        f = self._fields[field]
        n = self.model.n_dofs
        cols = self._idx(dofs_idx, n) if not isinstance(dofs_idx, slice) else torch.arange(n, device=self.device)[dofs_idx]
        if comps > 1:
            cols = cols * comps + width
        tgt = f.reshape(self._B, -1)
        v = torch.as_tensor(values, dtype=torch.float32, device=self.device)
        if self._is_mask(envs_idx):
            cur = tgt[:, cols]
            tgt[:, cols] = torch.where(envs_idx.to(self.device)[:, None], torch.broadcast_to(v, cur.shape), cur)
            return
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        v = torch.broadcast_to(v, (nb, len(cols)))
        if envs is None:
            tgt[:, cols] = v
        else:
            tgt[envs[:, None], cols[None, :]] = v

    def _get_dof_param(self, field, model_values, dofs_idx, envs_idx, width=None, comps=1):
        n = self.model.n_dofs
        if not self._batch_dofs:
            return torch.tensor(np.asarray(model_values), dtype=torch.float32, device=self.device)[self._idx(dofs_idx, n)]
        cols = self._idx(dofs_idx, n) if not isinstance(dofs_idx, slice) else torch.arange(n, device=self.device)[dofs_idx]
        if comps > 1:
            cols = cols * comps + width
        out = self._fields[field].reshape(self._B, -1)[:, cols]
        if envs_idx is not None:
            out = out[envs_idx.to(self.device)] if self._is_mask(envs_idx) else out[self._envs(envs_idx)]
        return out[0] if self.n_envs == 0 else out

    def set_dofs_kp(self, kp, dofs_idx=None, envs_idx=None):
        """abd/accessor.py:566-583: act_gain = kp, act_bias[1] = -kp."""
        kp = torch.as_tensor(kp, dtype=torch.float32)
        self._set_dof_param("bd_kp", "dofs_kp", kp, dofs_idx, envs_idx)
        self._set_dof_param("bd_bias", "dofs_act_bias", -kp, dofs_idx, envs_idx, width=1, comps=3)

    def set_dofs_kv(self, kv, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_bias", "dofs_act_bias", -torch.as_tensor(kv, dtype=torch.float32), dofs_idx, envs_idx, width=2, comps=3)

    def set_dofs_force_range(self, lower, upper, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_force_range", "dofs_force_range", lower, dofs_idx, envs_idx, width=0, comps=2)
        self._set_dof_param("bd_force_range", "dofs_force_range", upper, dofs_idx, envs_idx, width=1, comps=2)

    def set_dofs_frictionloss(self, frictionloss, dofs_idx=None, envs_idx=None):
        """rigid_solver.py:2100-2120: dofs with a positive friction loss get one constraint row each (solver.py:1236-1286)."""
        self._set_dof_param("bd_frictionloss", "dofs_frictionloss", frictionloss, dofs_idx, envs_idx)

    def set_dofs_limit(self, lower, upper, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_limit", "dofs_limit", lower, dofs_idx, envs_idx, width=0, comps=2)
        self._set_dof_param("bd_limit", "dofs_limit", upper, dofs_idx, envs_idx, width=1, comps=2)
        self._dofs_limit = torch.tensor(self.model.dofs_limit, dtype=torch.float32, device=self.device).reshape(-1, 2)

    def set_dofs_armature(self, armature, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_armature", "dofs_armature", armature, dofs_idx, envs_idx)

    def set_dofs_damping(self, damping, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_damping", "dofs_damping", damping, dofs_idx, envs_idx)

    def set_dofs_stiffness(self, stiffness, dofs_idx=None, envs_idx=None):
        self._set_dof_param("bd_stiffness", "dofs_stiffness", stiffness, dofs_idx, envs_idx)

    @property
    def dofs_info(self):
        """`rigid_solver.dofs_info.<field>.to_numpy() / from_numpy()` as the reference's tests use it (shape (n_dofs, B) when
        batch_dofs_info, else (n_dofs,); array_class.py:1332-1364)."""
        return _DofsInfoFacade(self)

    # ------------------------------------------------------------------ getters (kinematic_solver.py:940-1003, rigid_solver.py:2323-2580)
    def get_qpos(self, qs_idx=None, envs_idx=None):
        return self._get("qpos", qs_idx, envs_idx, self.model.n_qs)

    def get_dofs_position(self, dofs_idx=None, envs_idx=None):
        return self._get("dofs_pos", dofs_idx, envs_idx, self.model.n_dofs)

    def get_dofs_velocity(self, dofs_idx=None, envs_idx=None):
        return self._get("vel", dofs_idx, envs_idx, self.model.n_dofs)

    def get_dofs_force(self, dofs_idx=None, envs_idx=None):
        return self._get("force", dofs_idx, envs_idx, self.model.n_dofs)

    def get_dofs_control_force(self, dofs_idx=None, envs_idx=None):
        return self._get("qf_applied", dofs_idx, envs_idx, self.model.n_dofs)

    def get_dofs_limit(self, dofs_idx=None, envs_idx=None):
        idx = self._idx(dofs_idx, self.model.n_dofs)
        return self._dofs_limit[idx, 0].clone(), self._dofs_limit[idx, 1].clone()

    def get_dofs_kp(self, dofs_idx=None, envs_idx=None):
        return self._get_dof_param("bd_kp", self.model.dofs_kp, dofs_idx, envs_idx)

    def get_dofs_kv(self, dofs_idx=None, envs_idx=None):
        return -self._get_dof_param("bd_bias", np.asarray(self.model.dofs_act_bias).reshape(-1, 3)[:, 2], dofs_idx, envs_idx, width=2, comps=3)

    def get_links_pos(self, links_idx=None, envs_idx=None):
        return self._get("links_pos", links_idx, envs_idx, self.model.n_links, comps=3)

    def get_links_quat(self, links_idx=None, envs_idx=None):
        return self._get("links_quat", links_idx, envs_idx, self.model.n_links, comps=4)

    def get_links_ang(self, links_idx=None, envs_idx=None):
        return self._get("cd_ang", links_idx, envs_idx, self.model.n_links, comps=3)

    def get_links_vel(self, links_idx=None, envs_idx=None):
        """Velocity of the link frame origin: cd_vel + cd_ang x (pos - root_COM) (forward_dynamics.py:281-288)."""
        n = self.model.n_links
        vel = self._get("cd_vel", links_idx, envs_idx, n, comps=3)
        ang = self._get("cd_ang", links_idx, envs_idx, n, comps=3)
        pos = self._get("links_pos", links_idx, envs_idx, n, comps=3)
        com = self._get("root_COM", links_idx, envs_idx, n, comps=3)
        return vel + torch.cross(ang, pos - com, dim=-1)

    def get_links_acc(self, links_idx=None, envs_idx=None):
        """Classical linear acceleration of the link frame origin in world axes (abd/accessor.py:960-983): spatial acceleration
        moved from the subtree COM to the origin, plus the centripetal term ang x vel."""
        n = self.model.n_links
        cpos = self._get("links_pos", links_idx, envs_idx, n, comps=3) - self._get("root_COM", links_idx, envs_idx, n, comps=3)
        acc_ang = self._get("cacc_ang", links_idx, envs_idx, n, comps=3)
        acc_lin = self._get("cacc_lin", links_idx, envs_idx, n, comps=3) + torch.cross(acc_ang, cpos, dim=-1)
        ang = self._get("cd_ang", links_idx, envs_idx, n, comps=3)
        vel = self._get("cd_vel", links_idx, envs_idx, n, comps=3) + torch.cross(ang, cpos, dim=-1)
        return acc_lin + torch.cross(ang, vel, dim=-1)

    def get_links_acc_ang(self, links_idx=None, envs_idx=None):
        return self._get("cacc_ang", links_idx, envs_idx, self.model.n_links, comps=3)

    def get_links_root_COM(self, links_idx=None, envs_idx=None):
        return self._get("root_COM", links_idx, envs_idx, self.model.n_links, comps=3)

    def get_dofs_armature(self, dofs_idx=None, envs_idx=None):
        return self._get_dof_param("bd_armature", self.model.dofs_armature, dofs_idx, envs_idx)

    def get_dofs_frictionloss(self, dofs_idx=None, envs_idx=None):
        return self._get_dof_param("bd_frictionloss", self.model.dofs_frictionloss, dofs_idx, envs_idx)

    def get_dofs_damping(self, dofs_idx=None, envs_idx=None):
        return self._get_dof_param("bd_damping", self.model.dofs_damping, dofs_idx, envs_idx)

    def get_dofs_stiffness(self, dofs_idx=None, envs_idx=None):
        return self._get_dof_param("bd_stiffness", self.model.dofs_stiffness, dofs_idx, envs_idx)

    def get_dofs_force_range(self, dofs_idx=None):
        fr = torch.tensor(self.model.dofs_force_range, dtype=torch.float32, device=self.device).reshape(-1, 2)[self._idx(dofs_idx, self.model.n_dofs)]
        return fr[:, 0], fr[:, 1]

    def get_dofs_invweight(self, dofs_idx=None):
        return torch.tensor(self.model.dofs_invweight, dtype=torch.float32, device=self.device)[self._idx(dofs_idx, self.model.n_dofs)]

    def get_links_invweight(self, links_idx=None):
        return torch.tensor(self.model.links_invweight, dtype=torch.float32, device=self.device).reshape(-1, 2)[self._idx(links_idx, self.model.n_links)]

    def get_links_inertial_mass(self, links_idx=None):
        return torch.tensor(self.model.links_inertial_mass, dtype=torch.float32, device=self.device)[self._idx(links_idx, self.model.n_links)]

    def get_geoms_friction_ratio(self, geoms_idx=None, envs_idx=None):
        return self._get("friction_ratio", geoms_idx, envs_idx, self.model.n_geoms)

    def set_geoms_friction_ratio(self, friction_ratio, geoms_idx=None, envs_idx=None):
        """Per-environment friction scale of each geom (contact friction = max(mu_a ratio_a, mu_b ratio_b, 1e-2), contact.py:202-211)."""
        self._set("friction_ratio", friction_ratio, geoms_idx, envs_idx, self.model.n_geoms)

    # ---- dynamic weld equalities (rigid_solver.py:2591-2598, constraint/solver.py:1293-1390)
    MAX_DYN_EQ, DYN_EQ_WORDS = 8, 20

    def add_weld_constraint(self, link1_idx, link2_idx, envs_idx=None):
        """Welds link2 to link1 at their current relative pose in the selected environments (kernel_add_weld_constraint:
        anchors = link1's origin expressed in both frames, relpose = inv(q1) (x) q2, torquescale 1, sol_params
        (2 dt, 1, 0.9, 0.95, 0.001, 0.5, 2))."""
        envs = self._envs(envs_idx)
        e = self._envs_all if envs is None else envs
        n = self._fields["n_dyn_eq"][e, 0]
        if int(n.max()) >= self.MAX_DYN_EQ:
            raise GenesisException("Exceeding max number of dynamic equality constraints (max_dynamic_constraints).")
        p1, q1 = self.field("links_pos")[e, link1_idx], self.field("links_quat")[e, link1_idx]
        p2, q2 = self.field("links_pos")[e, link2_idx], self.field("links_quat")[e, link2_idx]
        q2c = torch.cat([q2[:, :1], -q2[:, 1:]], dim=-1)
        q1c = torch.cat([q1[:, :1], -q1[:, 1:]], dim=-1)
        pos1 = torch.zeros_like(p1)                      # the shared point is link1's origin
        pos2 = _rotate_torch(p1 - p2, q2c)
        rel = _quat_mul_torch(q1c, q2)
        rec = torch.zeros((len(e), self.DYN_EQ_WORDS), dtype=torch.float32, device=self.device)
        rec[:, 0] = torch.tensor([link1_idx], dtype=torch.int32, device=self.device).view(torch.float32)
        rec[:, 1] = torch.tensor([link2_idx], dtype=torch.int32, device=self.device).view(torch.float32)
        rec[:, 2:5], rec[:, 5:8], rec[:, 8:12], rec[:, 12] = pos2, pos1, rel, 1.0
        rec[:, 13:20] = torch.tensor([2 * self._substep_dt, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0], device=self.device)
        self._fields["dyn_eq"][e, n.long()] = rec
        self._fields["n_dyn_eq"][e, 0] = n + 1

    def delete_weld_constraint(self, link1_idx, link2_idx, envs_idx=None):
        """Removes the dynamic weld between the two links where it exists (the last record takes the freed slot)."""
        envs = self._envs(envs_idx)
        e = self._envs_all if envs is None else envs
        recs = self._fields["dyn_eq"]
        ids = recs[..., :2].contiguous().view(torch.int32)           # (B, MAX, 2)
        n = self._fields["n_dyn_eq"][:, 0]
        for k in range(self.MAX_DYN_EQ):
            hit = (ids[e, k, 0] == link1_idx) & (ids[e, k, 1] == link2_idx) & (k < n[e])
            he = e[hit]
            if len(he) == 0:
                continue
            last = (n[he] - 1).long()
            recs[he, k] = recs[he, last]
            self._fields["n_dyn_eq"][he, 0] = n[he] - 1
            ids = recs[..., :2].contiguous().view(torch.int32)
            n = self._fields["n_dyn_eq"][:, 0]

    def get_weld_constraints(self):
        """{"link_a", "link_b"}: (B, MAX_DYN_EQ) int tensors, -1 where no dynamic weld is stored."""
        ids = self._fields["dyn_eq"][..., :2].contiguous().view(torch.int32).clone()
        n = self._fields["n_dyn_eq"][:, :1]
        valid = torch.arange(self.MAX_DYN_EQ, device=self.device)[None, :] < n
        ids[~valid] = -1
        return {"link_a": ids[..., 0], "link_b": ids[..., 1], "n": n[:, 0].clone()}

    # ---- per-environment domain randomisation of the inertial parameters (rigid_solver.py:1825-1860, accessor.py:400-430)
    def set_links_mass_shift(self, mass_shift, links_idx=None, envs_idx=None):
        """links_state.mass_shift: added to the link mass in the COM and c-frame inertia (forward_kinematics.py:377, 437)."""
        self._set("mass_shift", mass_shift, links_idx, envs_idx, self.model.n_links)
        self._mark_shifted(envs_idx)

    def set_links_COM_shift(self, com_shift, links_idx=None, envs_idx=None):
        """links_state.i_pos_shift: added to the inertial frame origin of the link (forward_kinematics.py:382)."""
        f = self._fields["com_shift"]
        idx = self._idx(links_idx, self.model.n_links)
        envs = self._envs(envs_idx)
        nb = self._B if envs is None else len(envs)
        v = torch.broadcast_to(torch.as_tensor(com_shift, dtype=torch.float32, device=self.device), (nb, len(idx), 3))
        if envs is None:
            f[:, idx] = v
        else:
            f[envs[:, None], idx[None, :]] = v
        self._mark_shifted(envs_idx)

    def _mark_shifted(self, envs_idx):
        envs = self._envs(envs_idx)
        e = self._envs_all if envs is None else envs
        self._fields["has_shift"][e] = 1
        self._forward(envs_idx)  # the kinematic cache (subtree COM, c-frame inertias) depends on the masses

    def get_links_mass_shift(self, links_idx=None, envs_idx=None):
        return self._get("mass_shift", links_idx, envs_idx, self.model.n_links)

    def get_links_COM_shift(self, links_idx=None, envs_idx=None):
        return self._get("com_shift", links_idx, envs_idx, self.model.n_links, comps=3)

    def get_links_net_contact_force(self, links_idx=None, envs_idx=None):
        return self._get("links_contact_force", links_idx, envs_idx, self.model.n_links, comps=3)

    DIAG_FIELDS = ("acc_smooth", "qf_smooth", "qf_passive", "qf_bias", "mass_mat", "mass_Dinv", "geoms_pos", "geoms_quat")

    def refresh_diagnostics(self):
        """With cfg_diagnostics = 0 (default) the step kernel does not write the diagnostic fields (mass_mat, mass_Dinv, qf_*,
        acc_smooth, geoms_pos / quat); the accessors that expose them recompute them for the CURRENT state on demand (k_diag).
        With GSR_DIAGNOSTICS=1 they are written every step with the reference's timing and this is a no-op."""
        if not self.tuning.get("diagnostics"):
            self._ensure_kinematics()
            _lib.check(self._handle, self._L.gsr_compute_diagnostics(self._handle, self._stream()))

    def get_geoms_pos(self, geoms_idx=None, envs_idx=None):
        self.refresh_diagnostics()
        return self._get("geoms_pos", geoms_idx, envs_idx, self.model.n_geoms, comps=3)

    def get_geoms_quat(self, geoms_idx=None, envs_idx=None):
        self.refresh_diagnostics()
        return self._get("geoms_quat", geoms_idx, envs_idx, self.model.n_geoms, comps=4)

    def get_mass_mat(self, dofs_idx=None, envs_idx=None):
        self.refresh_diagnostics()
        nD = self.model.n_dofs
        M = self._fields["mass_mat"].reshape(self._B, nD, nD)
        envs = self._envs(envs_idx)
        M = M if envs is None else M[envs]
        idx = self._idx(dofs_idx, nD)
        out = M[:, idx][:, :, idx].clone()
        return out[0] if self.n_envs == 0 else out

    def get_AABB(self, geoms_idx=None, envs_idx=None):
        n = self.model.n_geoms
        lo = self._get("aabb_min", geoms_idx, envs_idx, n, comps=3)
        hi = self._get("aabb_max", geoms_idx, envs_idx, n, comps=3)
        return torch.stack([lo, hi], dim=-2)

    def get_contacts(self, as_tensor=True, to_torch=True, keep_batch_dim=False):
        """collider.py:810-913: dict of (B, n_contacts_max[, 3]) tensors, cleared slots hold -1 / 0 (`as_tensor=True`), or of
        per-environment tuples cut to each environment's own contact count (`as_tensor=False`, collider.py:834-838)."""
        n = self._fields["n_contacts"][:, 0]
        n_max = int(n.max().item())
        f = self._fields
        counts = n.tolist() if not as_tensor else None

        def cut(name):
            t = f[name][:, :n_max].clone()
            if not to_torch:
                t = t.cpu().numpy()
            if self.n_envs == 0:
                return t if keep_batch_dim else t[0]
            if not as_tensor:
                return tuple(t[i: i + 1, :j] if keep_batch_dim else t[i, :j] for i, j in enumerate(counts))
            return t

        return dict(link_a=cut("contact_link_a"), link_b=cut("contact_link_b"), geom_a=cut("contact_geom_a"),
                    geom_b=cut("contact_geom_b"), penetration=cut("contact_penetration"), position=cut("contact_pos"),
                    normal=cut("contact_normal"), force=cut("contact_force"))

    @property
    def n_contacts(self):
        return self._fields["n_contacts"][:, 0]

    @property
    def collider(self):
        """`rigid_solver.collider.get_contacts(...)` like in the reference (collider/collider.py:810)."""
        return _ColliderFacade(self)

    @property
    def constraint_solver(self):
        """The attributes of the reference's ConstraintSolver that user code and tests read (constraint/solver.py:27-112)."""
        return _ConstraintSolverFacade(self)

    # ------------------------------------------------------------------ state (rigid_solver.py:1430-1579)
    STATE_FIELDS = (("qpos", "qpos"), ("dofs_vel", "vel"), ("dofs_acc", "acc"), ("links_pos", "links_pos"), ("links_quat", "links_quat"),
                    ("i_pos_shift", "com_shift"), ("mass_shift", "mass_shift"), ("friction_ratio", "friction_ratio"))

    def gather_fields(self, names, out=None):
        """{name: dense (B, items[, comps]) tensor} of several record fields in ONE launch (gsr_gather_fields)."""
        out = out or {}
        bufs = []
        if any(nme in self.KIN_FIELDS for nme in names):
            self._ensure_kinematics()
        for nme in names:
            f = self._fields[nme]
            t = out.get(nme)
            if t is None:
                t = torch.empty(f.shape, dtype=torch.float32, device=self.device)
            bufs.append(t)
        for k in range(0, len(names), 16):
            chunk, cb = names[k: k + 16], bufs[k: k + 16]
            arr_n = (ctypes.c_char_p * len(chunk))(*[c.encode() for c in chunk])
            arr_p = (ctypes.c_void_p * len(chunk))(*[b.data_ptr() for b in cb])
            _lib.check(self._handle, self._L.gsr_gather_fields(self._handle, len(chunk), arr_n, arr_p, self._stream()))
        return dict(zip(names, bufs))

    def get_state(self, f=None):
        """RigidSolverState of the current step (kernel_get_state, abd/accessor.py): one fused gather."""
        got = self.gather_fields([src for _, src in self.STATE_FIELDS])
        st = RigidSolverState(self._scene, self._n_substeps_total)
        for dst, src in self.STATE_FIELDS:
            setattr(st, dst, got[src])
        return st

    def set_state(self, f, state=None, envs_idx=None, *, partial=False):
        """rigid_solver.py:1458-1579 `set_state(f, state, envs_idx, *, partial)`: one fused launch (k_set_state) does the collider
        / constraint-solver reset (partial) or clear, writes the state of the selected environments, zeroes the control forces
        and external wrenches and -- unless partial -- runs the forward kinematics.  `envs_idx`: None, indices, or a boolean mask
        (used as is on the device)."""
        if state is None or isinstance(f, (RigidSolverState, dict)):  # tolerate the old (state, envs_idx) calling convention
            f, state, envs_idx = 0, f, state if envs_idx is None else envs_idx
        get = (lambda k: state.get(k)) if isinstance(state, dict) else (lambda k: getattr(state, k, None))
        ptrs, keep = [], []
        for dst, src in self.STATE_FIELDS:
            v = get(dst)
            if v is None:
                ptrs.append(None)
                continue
            shape = self._fields[src].shape
            t = torch.as_tensor(v, dtype=torch.float32, device=self.device)
            if self.n_envs == 0 and t.dim() == len(shape) - 1:
                t = t[None]
            t = torch.broadcast_to(t, shape).contiguous()
            keep.append(t)
            ptrs.append(ctypes.c_void_p(t.data_ptr()))
        mask = self._env_mask(envs_idx)
        mptr = ctypes.c_void_p(mask.data_ptr()) if mask is not None else None
        _lib.check(self._handle, self._L.gsr_set_state(self._handle, mptr, *ptrs, int(bool(partial)), self._stream()))
        self._obs_valid = False
        if partial:
            self._fk_dirty_any = True

    def _restore(self, values, envs_idx):
        f = self._fields
        envs = self._envs(envs_idx)
        rows = slice(None) if envs is None else envs
        for name, val in values.items():
            f[name][rows] = val[rows] if torch.is_tensor(val) else val
        f["ctrl_force"][rows] = 0.0
        f["ctrl_mode"][rows] = int(CTRL_MODE.FORCE)
        f["cfrc_applied_vel"][rows] = 0.0
        f["cfrc_applied_ang"][rows] = 0.0
        f["has_ext_force"][rows] = 0
        self._reset_warmstart(envs_idx)
        self._forward(envs_idx)
        self._obs_valid = False

    def reset(self, envs_idx=None):
        self._restore(dict(qpos=self._init_qpos, vel=0.0, acc=0.0), envs_idx)

    def destroy(self):
        if self._handle is not None:
            self._L.gsr_destroy(self._handle)
            self._handle = None
            _lib.release_instance(self._lib_device, self._lib_instance, self)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class _QdFieldFacade:
    """to_numpy / from_numpy of one dof parameter in the reference's `[dof, env]` orientation."""

    def __init__(self, solver, field, model_name, comps=1):
        self._s, self._field, self._model_name, self._comps = solver, field, model_name, comps

    def to_numpy(self):
        s = self._s
        if s._batch_dofs:
            a = s.field(self._field).cpu().numpy()  # (B, nD[, comps])
            return np.moveaxis(a, 0, 1).copy()
        return np.asarray(getattr(s.model, self._model_name), np.float32).copy()

    def from_numpy(self, arr):
        s = self._s
        arr = np.asarray(arr, np.float32)
        if s._batch_dofs:
            s.field(self._field).copy_(torch.as_tensor(np.moveaxis(arr, 1, 0).copy(), device=s.device))
        else:
            getattr(s.model, self._model_name)[...] = arr.reshape(np.shape(getattr(s.model, self._model_name)))
            a = np.ascontiguousarray(getattr(s.model, self._model_name), dtype=np.float32)
            s._L.gsr_model_set_f32(s._handle, self._model_name.encode(), a.ctypes.data, a.size)
            _lib.check(s._handle, s._L.gsr_model_refresh(s._handle))


class _DofsInfoFacade:
    def __init__(self, solver):
        f = lambda field, name, comps=1: _QdFieldFacade(solver, field, name, comps)
        self.armature, self.damping, self.stiffness = f("bd_armature", "dofs_armature"), f("bd_damping", "dofs_damping"), f("bd_stiffness", "dofs_stiffness")
        self.frictionloss, self.invweight, self.act_gain = f("bd_frictionloss", "dofs_frictionloss"), f("bd_invweight", "dofs_invweight"), f("bd_kp", "dofs_kp")
        self.act_bias, self.limit, self.force_range = f("bd_bias", "dofs_act_bias", 3), f("bd_limit", "dofs_limit", 2), f("bd_force_range", "dofs_force_range", 2)


class _ColliderFacade:
    def __init__(self, solver):
        self._solver = solver

    def get_contacts(self, as_tensor=True, to_torch=True, keep_batch_dim=False):
        return self._solver.get_contacts(as_tensor, to_torch, keep_batch_dim)

    @property
    def n_contacts(self):
        return self._solver.n_contacts


class _ConstraintSolverFacade:
    def __init__(self, solver):
        self._solver = solver

    qacc_ws = property(lambda self: self._solver.field("qacc_ws"))
    n_constraints = property(lambda self: self._solver.field("n_constraints")[:, 0])
    n_constraints_equality = property(lambda self: self._solver.field("n_constraints_equality")[:, 0])
    n_constraints_frictionloss = property(lambda self: self._solver.field("n_constraints_frictionloss")[:, 0])
    n_iterations = property(lambda self: self._solver.field("solver_iters")[:, 0])

    def add_weld_constraint(self, link1_idx, link2_idx, envs_idx=None):
        return self._solver.add_weld_constraint(link1_idx, link2_idx, envs_idx)

    def delete_weld_constraint(self, link1_idx, link2_idx, envs_idx=None):
        return self._solver.delete_weld_constraint(link1_idx, link2_idx, envs_idx)

    def get_weld_constraints(self, as_tensor=True, to_torch=True):
        return self._solver.get_weld_constraints()


def _rotate_torch(v, q):
    w, x, y, z = q[..., 0:1], q[..., 1:2], q[..., 2:3], q[..., 3:4]
    u = torch.cat([x, y, z], dim=-1)
    t = 2.0 * torch.cross(u, v, dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def _quat_mul_torch(u, v):
    """Hamilton product u (x) v of (w, x, y, z) quaternions (genesis/utils/geom.py:235-242)."""
    uw, ux, uy, uz = u.unbind(-1)
    vw, vx, vy, vz = v.unbind(-1)
    return torch.stack([uw * vw - ux * vx - uy * vy - uz * vz, uw * vx + ux * vw + uy * vz - uz * vy,
                        uw * vy - ux * vz + uy * vw + uz * vx, uw * vz + ux * vy - uy * vx + uz * vw], dim=-1)


def _xyz_to_quat_torch(xyz):
    """genesis/utils/geom.py:45-65 (intrinsic x-y-z Euler -> quaternion)."""
    ai, aj, ak = 0.5 * xyz[..., 2], -0.5 * xyz[..., 1], 0.5 * xyz[..., 0]
    si, sj, sk = torch.sin(ai), torch.sin(aj), torch.sin(ak)
    ci, cj, ck = torch.cos(ai), torch.cos(aj), torch.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return torch.stack([cj * cc + sj * ss, cj * cs - sj * sc, -cj * ss - sj * cc, cj * sc - sj * cs], dim=-1)
