"""RigidEntity: per-entity accessor surface of the reference (genesis/engine/entities/rigid_entity/rigid_entity.py:
1214-1261 index mapping, :1641-1720 getters, :3372-3650 setters / controllers, :3919-4013 contacts).  Everything is
forwarded to RigidSolver with global indices."""

import numpy as np
import torch


class RigidLink:
    def __init__(self, entity, name, idx, idx_local, dof_start, n_dofs, q_start, n_qs, parent_idx, is_fixed):
        self.entity, self.name, self.idx, self.idx_local = entity, name, idx, idx_local
        self.dof_start, self.n_dofs, self.q_start, self.n_qs = dof_start, n_dofs, q_start, n_qs
        self.parent_idx, self.is_fixed = parent_idx, is_fixed

    def get_pos(self, envs_idx=None):
        return self.entity._solver.get_links_pos([self.idx], envs_idx)[..., 0, :]

    def get_quat(self, envs_idx=None):
        return self.entity._solver.get_links_quat([self.idx], envs_idx)[..., 0, :]

    def get_vel(self, envs_idx=None):
        return self.entity._solver.get_links_vel([self.idx], envs_idx)[..., 0, :]

    def get_ang(self, envs_idx=None):
        return self.entity._solver.get_links_ang([self.idx], envs_idx)[..., 0, :]


class RigidJoint:
    def __init__(self, entity, name, idx, dof_start, n_dofs, q_start, n_qs, type):
        self.entity, self.name, self.idx = entity, name, idx
        self.dof_start, self.n_dofs, self.q_start, self.n_qs, self.type = dof_start, n_dofs, q_start, n_qs, type

    @property
    def dofs_idx_local(self):
        return list(range(self.dof_start - self.entity.dof_start, self.dof_start - self.entity.dof_start + self.n_dofs))


class RigidGeom:
    """rigid_geom.py: the part of a geom the step depends on and a user may change after build."""

    def __init__(self, entity, idx, idx_local):
        self.entity, self.idx, self.idx_local = entity, idx, idx_local

    @property
    def sol_params(self):
        return self.entity._solver.get_geoms_sol_params([self.idx])[0]

    def set_sol_params(self, sol_params):
        self.entity._solver.set_geoms_sol_params(sol_params, [self.idx])


class Material:
    """genesis/engine/materials/rigid.py: only the parameters that reach the rigid step."""

    def __init__(self, rho=None, friction=None, gravity_compensation=0.0):
        self.rho, self.friction, self.gravity_compensation = rho, friction, gravity_compensation


class RigidEntity:
    def __init__(self, solver, morph, material, idx):
        self._solver = solver
        self.morph = morph
        self.material = material or Material()
        self.idx = idx
        self._info = None
        self._links = None

    # ---- sizes / indices
    @property
    def n_dofs(self):
        return self._info.n_dofs

    @property
    def n_qs(self):
        return self._info.n_qs

    @property
    def n_links(self):
        return self._info.n_links

    @property
    def n_geoms(self):
        return self._info.n_geoms

    @property
    def dof_start(self):
        return self._info.dof_start

    @property
    def q_start(self):
        return self._info.q_start

    @property
    def link_start(self):
        return self._info.link_start

    @property
    def geom_start(self):
        return self._info.geom_start

    @property
    def base_link_idx(self):
        return self._info.link_start

    @property
    def links(self):
        if self._links is None:
            m, e = self._solver.model, self._info
            self._links = [
                RigidLink(self, e.link_names[i], e.link_start + i, i, int(m.links_dof_start[e.link_start + i]),
                          int(m.links_dof_end[e.link_start + i] - m.links_dof_start[e.link_start + i]),
                          int(m.links_q_start[e.link_start + i]),
                          int(m.links_q_end[e.link_start + i] - m.links_q_start[e.link_start + i]),
                          int(m.links_parent[e.link_start + i]), bool(m.links_is_fixed[e.link_start + i]))
                for i in range(e.n_links)
            ]
        return self._links

    @property
    def joints(self):
        m, e = self._solver.model, self._info
        out = []
        for i in range(e.n_joints):
            j = e.joint_start + i
            nd = int(m.joints_n_dofs[j])
            nq = {0: 0, 1: 1, 2: 1, 3: 4, 4: 7}[int(m.joints_type[j])]
            out.append(RigidJoint(self, e.joint_names[i], j, int(m.joints_dof_start[j]), nd, int(m.joints_q_start[j]), nq, int(m.joints_type[j])))
        return out

    @property
    def geoms(self):
        return [RigidGeom(self, self.geom_start + i, i) for i in range(self.n_geoms)]

    def get_link(self, name):
        for l in self.links:
            if l.name == name:
                return l
        raise KeyError(name)

    def get_joint(self, name):
        for j in self.joints:
            if j.name == name:
                return j
        raise KeyError(name)

    def _global(self, idx_local, n, start):
        """rigid_entity.py:1214-1261 _get_global_idx."""
        if idx_local is None:
            return slice(start, start + n)  # contiguous columns of the record: the solver reads / writes strided views, no gather
        if isinstance(idx_local, slice):
            a, b, st = idx_local.indices(n)
            if st == 1:
                return slice(start + a, start + max(a, b))
            return torch.arange(start, start + n)[idx_local]
        if isinstance(idx_local, range):
            return torch.arange(start, start + n)[idx_local]
        idx = torch.as_tensor(idx_local, dtype=torch.long).reshape(-1).cpu()
        if len(idx) and (idx.min() < -n or idx.max() >= n):
            from .rigid_solver import GenesisException

            raise GenesisException("`idx_local` exceeds valid range.")
        return torch.where(idx < 0, idx + n, idx) + start

    def _dofs(self, idx):
        return self._global(idx, self.n_dofs, self.dof_start)

    def _qs(self, idx):
        return self._global(idx, self.n_qs, self.q_start)

    def _links_idx(self, idx):
        return self._global(idx, self.n_links, self.link_start)

    # ---- getters
    def get_qpos(self, qs_idx_local=None, envs_idx=None):
        return self._solver.get_qpos(self._qs(qs_idx_local), envs_idx)

    def get_dofs_position(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_position(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_velocity(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_velocity(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_force(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_force(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_control_force(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_control_force(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_limit(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_limit(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_kp(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_kp(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_kv(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_kv(self._dofs(dofs_idx_local), envs_idx)

    def get_links_pos(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_pos(self._links_idx(links_idx_local), envs_idx)

    def get_links_quat(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_quat(self._links_idx(links_idx_local), envs_idx)

    def get_links_vel(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_vel(self._links_idx(links_idx_local), envs_idx)

    def get_links_ang(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_ang(self._links_idx(links_idx_local), envs_idx)

    def get_links_net_contact_force(self, envs_idx=None):
        return self._solver.get_links_net_contact_force(self._links_idx(None), envs_idx)

    def get_pos(self, envs_idx=None):
        return self._solver.get_links_pos([self.base_link_idx], envs_idx)[..., 0, :]

    def get_quat(self, envs_idx=None):
        return self._solver.get_links_quat([self.base_link_idx], envs_idx)[..., 0, :]

    def get_vel(self, envs_idx=None):
        return self._solver.get_links_vel([self.base_link_idx], envs_idx)[..., 0, :]

    def get_ang(self, envs_idx=None):
        return self._solver.get_links_ang([self.base_link_idx], envs_idx)[..., 0, :]

    def get_mass_mat(self, envs_idx=None):
        return self._solver.get_mass_mat(self._dofs(None), envs_idx)

    def get_AABB(self, envs_idx=None):
        aabb = self._solver.get_AABB(torch.arange(self.geom_start, self.geom_start + self.n_geoms), envs_idx)
        return torch.stack([aabb[..., 0, :].min(dim=-2).values, aabb[..., 1, :].max(dim=-2).values], dim=-2)

    def get_mass(self):
        m = self._solver.model
        return float(m.links_inertial_mass[self.link_start: self.link_start + self.n_links].sum())

    def get_contacts(self, with_entity=None, exclude_self_contact=False):
        """rigid_entity.py:3919-3999: contacts involving this entity, plus `valid_mask`, `force_a`, `force_b`."""
        c = self._solver.get_contacts(as_tensor=True, to_torch=True)
        lo, hi = self.link_start, self.link_start + self.n_links
        in_a = (c["link_a"] >= lo) & (c["link_a"] < hi)
        in_b = (c["link_b"] >= lo) & (c["link_b"] < hi)
        if with_entity is not None and self.idx == with_entity.idx:
            if exclude_self_contact:
                raise ValueError("`with_entity` is self but `exclude_self_contact` is True.")
            valid = in_a & in_b
        else:
            valid = in_a | in_b
            if exclude_self_contact:
                valid = valid & ~(in_a & in_b)
            if with_entity is not None:
                lo2, hi2 = with_entity.link_start, with_entity.link_start + with_entity.n_links
                valid = valid & (((c["link_a"] >= lo2) & (c["link_a"] < hi2)) | ((c["link_b"] >= lo2) & (c["link_b"] < hi2)))
        out = dict(c)
        out["valid_mask"] = valid
        out["force_a"] = -c["force"]
        out["force_b"] = c["force"]
        return out

    # ---- setters / controllers
    def set_qpos(self, qpos, qs_idx_local=None, envs_idx=None, *, zero_velocity=True, skip_forward=False):
        if zero_velocity:
            self._solver.set_dofs_velocity(None, self._dofs(None), envs_idx, skip_forward=True)
        self._solver.set_qpos(qpos, self._qs(qs_idx_local), envs_idx, skip_forward=skip_forward)

    def set_dofs_position(self, position, dofs_idx_local=None, envs_idx=None, *, zero_velocity=True):
        if zero_velocity:
            self._solver.set_dofs_velocity(None, self._dofs(None), envs_idx, skip_forward=True)
        self._solver.set_dofs_position(position, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_velocity(self, velocity=None, dofs_idx_local=None, envs_idx=None, *, skip_forward=False):
        self._solver.set_dofs_velocity(velocity, self._dofs(dofs_idx_local), envs_idx, skip_forward=skip_forward)

    def _base_kind(self):
        """'free' when the base link hangs off a free joint (its pose IS qpos[0:7]), 'fixed' when it is a fixed root link (its pose is
        a user-settable link state that FK keeps, forward_kinematics.py:732-736)."""
        from .constants import JOINT_TYPE

        m, l0 = self._solver.model, self._info.link_start
        js, je = int(m.links_joint_start[l0]), int(m.links_joint_end[l0])
        if je > js and int(m.joints_type[js]) == JOINT_TYPE.FREE:
            return "free"
        if int(m.links_parent[l0]) == -1 and int(m.links_is_fixed[l0]):
            return "fixed"
        return "articulated"

    def _set_base(self, value, comps, envs_idx, zero_velocity):
        """rigid_entity.py set_pos / set_quat -> set_base_links_pos / quat: a free base writes qpos, a fixed base moves the root link
        (joint coordinates are never touched: writing a pose into qpos[0:7] of a fixed-base arm would overwrite joint angles)."""
        kind = self._base_kind()
        if kind == "articulated":
            from .rigid_solver import GenesisException

            raise GenesisException("set_pos / set_quat: the base link of this entity is neither free-floating nor fixed.")
        if zero_velocity:
            self._solver.set_dofs_velocity(None, self._dofs(None), envs_idx, skip_forward=True)
        if kind == "free":
            self._solver.set_qpos(value, self._qs(comps), envs_idx)
        else:
            self._solver.set_base_link_pose(self._info.link_start, value, "links_pos" if len(comps) == 3 else "links_quat", envs_idx)

    def set_pos(self, pos, envs_idx=None, *, zero_velocity=True):
        self._set_base(pos, [0, 1, 2], envs_idx, zero_velocity)

    def set_quat(self, quat, envs_idx=None, *, zero_velocity=True):
        self._set_base(quat, [3, 4, 5, 6], envs_idx, zero_velocity)

    def control_dofs_position(self, position, dofs_idx_local=None, envs_idx=None):
        self._solver.control_dofs_position(position, self._dofs(dofs_idx_local), envs_idx)

    def control_dofs_position_velocity(self, position, velocity, dofs_idx_local=None, envs_idx=None):
        self._solver.control_dofs_position_velocity(position, velocity, self._dofs(dofs_idx_local), envs_idx)

    def control_dofs_velocity(self, velocity, dofs_idx_local=None, envs_idx=None):
        self._solver.control_dofs_velocity(velocity, self._dofs(dofs_idx_local), envs_idx)

    def control_dofs_force(self, force, dofs_idx_local=None, envs_idx=None):
        self._solver.control_dofs_force(force, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_kp(self, kp, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_kp(kp, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_kv(self, kv, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_kv(kv, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_force_range(self, lower, upper, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_force_range(lower, upper, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_armature(self, armature, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_armature(armature, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_damping(self, damping, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_damping(damping, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_stiffness(self, stiffness, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_stiffness(stiffness, self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_stiffness(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_stiffness(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_damping(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_damping(self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_frictionloss(self, frictionloss, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_frictionloss(frictionloss, self._dofs(dofs_idx_local), envs_idx)

    def set_dofs_limit(self, lower, upper, dofs_idx_local=None, envs_idx=None):
        self._solver.set_dofs_limit(lower, upper, self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_armature(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_armature(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_frictionloss(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_frictionloss(self._dofs(dofs_idx_local), envs_idx)

    def get_dofs_force_range(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_force_range(self._dofs(dofs_idx_local))

    def get_dofs_invweight(self, dofs_idx_local=None, envs_idx=None):
        return self._solver.get_dofs_invweight(self._dofs(dofs_idx_local))

    def get_links_acc(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_acc(self._links_idx(links_idx_local), envs_idx)

    def get_links_acc_ang(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_acc_ang(self._links_idx(links_idx_local), envs_idx)

    def get_links_root_COM(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_root_COM(self._links_idx(links_idx_local), envs_idx)

    def get_links_invweight(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_invweight(self._links_idx(links_idx_local))

    def get_links_inertial_mass(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_inertial_mass(self._links_idx(links_idx_local))

    def apply_links_external_force(self, force, links_idx_local=None, envs_idx=None, *, ref="link_origin", local=False):
        self._solver.apply_links_external_force(force, self._links_idx(links_idx_local), envs_idx, ref=ref, local=local)

    def apply_links_external_torque(self, torque, links_idx_local=None, envs_idx=None, *, ref="link_origin", local=False):
        self._solver.apply_links_external_torque(torque, self._links_idx(links_idx_local), envs_idx, ref=ref, local=local)

    def set_friction_ratio(self, friction_ratio, links_idx_local=None, envs_idx=None):
        """rigid_entity.py: per-environment friction scale, one value per link, applied to all the link's geoms."""
        m = self._solver.model
        links = self._links_idx(links_idx_local)
        links = list(range(links.start, links.stop)) if isinstance(links, slice) else torch.as_tensor(links).reshape(-1).tolist()
        ratio = torch.as_tensor(friction_ratio, dtype=torch.float32)
        geoms, cols = [], []
        for k, l in enumerate(links):
            for g in range(m.n_geoms):
                if int(m.geoms_link[g]) == int(l):
                    geoms.append(g)
                    cols.append(k)
        if geoms:
            self._solver.set_geoms_friction_ratio(ratio[..., cols], geoms, envs_idx)

    def set_mass_shift(self, mass_shift, links_idx_local=None, envs_idx=None):
        self._solver.set_links_mass_shift(mass_shift, self._links_idx(links_idx_local), envs_idx)

    def set_COM_shift(self, com_shift, links_idx_local=None, envs_idx=None):
        self._solver.set_links_COM_shift(com_shift, self._links_idx(links_idx_local), envs_idx)

    def get_mass_shift(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_mass_shift(self._links_idx(links_idx_local), envs_idx)

    def get_COM_shift(self, links_idx_local=None, envs_idx=None):
        return self._solver.get_links_COM_shift(self._links_idx(links_idx_local), envs_idx)

    def add_weld_constraint(self, link_local, other_entity, other_link_local=0, envs_idx=None):
        """Convenience forwarder: weld a link of another entity to a link of this one at their current relative pose."""
        self._solver.add_weld_constraint(self.link_start + link_local, other_entity.link_start + other_link_local, envs_idx)
