"""Flat rigid model ("model dump") consumed by the CUDA library and by the oracle.

This is the B200-native replacement for the reference's model compilation:
KinematicSolver.build / kernel_init_{dof,link,joint}_fields (genesis/engine/solvers/kinematic_solver.py:141-340),
RigidSolver._init_geom_fields / _init_equality_fields / _init_mass_mat / _init_invweight_and_meaninertia
(genesis/engine/solvers/rigid/rigid_solver.py:492-679,768-871) and Collider._compute_collision_pair_idx
(genesis/engine/solvers/rigid/collider/collider.py:258-492), SupportField.activate
(collider/support_field.py:22-91).  Everything here runs once on the host in fp64 numpy; the per-step path
never touches it.
"""

from dataclasses import dataclass, field

import numpy as np

from . import geom_np as gn
from . import meshio
from ..constants import EQUALITY_TYPE, GEOM_TYPE, JOINT_TYPE

RHO_OBJECT, RHO_ROBOT = 600.0, 1500.0  # rigid_link.py:23-24
IMP_MIN, IMP_MAX = 1e-4, 0.9999  # rigid_solver.py:179-181
SUPPORT_RES = 180  # support_field.py:18
EPS32 = float(np.finfo(np.float32).eps)


@dataclass
class EntityInfo:
    name: str
    l_infos: list
    eqs_info: list
    is_robot: bool
    rho: float | None = None
    friction: float | None = None
    gravity_compensation: float = 0.0
    local_collision_mask: bool = False
    link_start: int = 0
    n_links: int = 0
    dof_start: int = 0
    n_dofs: int = 0
    q_start: int = 0
    n_qs: int = 0
    geom_start: int = 0
    n_geoms: int = 0
    joint_start: int = 0
    n_joints: int = 0
    link_names: list = field(default_factory=list)
    joint_names: list = field(default_factory=list)
    dof_names: list = field(default_factory=list)


class RigidModel:
    """Container of flat numpy arrays (fp64 / int32). Field names follow the reference's *_info structs."""

    def to_arrays(self, float_dtype=np.float32):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                out[k] = np.ascontiguousarray(v.astype(float_dtype) if v.dtype.kind == "f" else v.astype(np.int32))
        return out


def sanitize_sol_params(sol_params, min_timeconst, default_timeconst):
    """rigid_solver.py:187-218."""
    sp = np.array(sol_params, dtype=np.float64).reshape(-1, 7)
    tc = sp[:, 0]
    tc[tc < EPS32] = default_timeconst
    sp[:, 0] = np.clip(tc, min_timeconst, None)
    sp[:, 2] = np.clip(sp[:, 2], IMP_MIN, IMP_MAX)
    sp[:, 3] = np.clip(sp[:, 3], IMP_MIN, IMP_MAX)
    sp[:, 5] = np.clip(sp[:, 5], IMP_MIN, IMP_MAX)
    sp[:, 4] = np.clip(sp[:, 4], 0.0, None)
    sp[:, 6] = np.clip(sp[:, 6], 1.0, None)
    return sp


def _geom_inertial(g, rho):
    """Unit mass properties of one geom in its own frame (analytic for primitives, hull integral otherwise)."""
    t = g["type"]
    if t == GEOM_TYPE.BOX:
        ex, ey, ez = g["data"][:3]
        m = rho * ex * ey * ez
        I = m / 12.0 * np.diag([ey * ey + ez * ez, ex * ex + ez * ez, ex * ex + ey * ey])
        return m, np.zeros(3), I
    if t == GEOM_TYPE.SPHERE:
        r = g["data"][0]
        m = rho * 4.0 / 3.0 * np.pi * r**3
        return m, np.zeros(3), 0.4 * m * r * r * np.eye(3)
    if t == GEOM_TYPE.ELLIPSOID:
        a, b, c = g["data"][:3]
        m = rho * 4.0 / 3.0 * np.pi * a * b * c
        return m, np.zeros(3), 0.2 * m * np.diag([b * b + c * c, a * a + c * c, a * a + b * b])
    if t == GEOM_TYPE.CYLINDER:
        r, h = g["data"][:2]
        m = rho * np.pi * r * r * h
        ixy = m * (3 * r * r + h * h) / 12.0
        return m, np.zeros(3), np.diag([ixy, ixy, 0.5 * m * r * r])
    if t == GEOM_TYPE.CAPSULE:  # cylinder + two hemispheres (MuJoCo's closed form)
        r, h = g["data"][:2]
        mc, ms = rho * np.pi * r * r * h, rho * 4.0 / 3.0 * np.pi * r**3
        izz = 0.5 * mc * r * r + 0.4 * ms * r * r
        ixy = mc * (3 * r * r + h * h) / 12.0 + ms * (0.4 * r * r + 0.25 * h * h + 0.375 * h * r)
        return mc + ms, np.zeros(3), np.diag([ixy, ixy, izz])
    vol, com, I = meshio.mesh_mass_properties(g["verts"], g["faces"])
    return rho * vol, com, rho * I


def _compose_inertial(items):
    """items: (mass, com_local, I_local_about_com, pos, quat) -> (mass, com, I about com) in link frame."""
    M, first = 0.0, np.zeros(3)
    parts = []
    for m, com, I, pos, quat in items:
        R = gn.quat_to_R(quat)
        c = np.asarray(pos) + R @ com
        parts.append((m, c, R @ I @ R.T))
        M += m
        first += m * c
    if M <= 0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    com = first / M
    I = np.zeros((3, 3))
    for m, c, Ic in parts:
        d = c - com
        I += Ic + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    return M, com, I


class ModelBuilder:
    def __init__(self, dt=0.01, gravity=(0.0, 0.0, -9.81), constraint_timeconst=0.01, enable_self_collision=True,
                 enable_adjacent_collision=False, max_collision_pairs=150, multiplier_collision_broad_phase=8,
                 n_contacts_per_pair=5, max_dynamic_constraints=8):
        self.dt = dt
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.sol_min_timeconst = 2.0 * dt
        self.sol_default_timeconst = max(constraint_timeconst, self.sol_min_timeconst)
        self.enable_self_collision = enable_self_collision
        self.enable_adjacent_collision = enable_adjacent_collision
        self.max_collision_pairs = max_collision_pairs
        self.multiplier_broad = multiplier_collision_broad_phase
        self.n_contacts_per_pair = n_contacts_per_pair
        self.max_dynamic_constraints = max_dynamic_constraints
        self.entities: list[EntityInfo] = []

    # ------------------------------------------------------------------ entities
    def add_entity(self, name, l_infos, eqs_info=(), is_robot=True, rho=None, friction=None, local_collision_mask=False,
                   morph_pos=None, morph_quat=None):
        # Drop the virtual world link when it carries nothing (rigid_entity.py:623-641)
        l_infos = list(l_infos)
        if (len(l_infos) > 1 and not l_infos[0]["joints"] and not l_infos[0]["geoms"] and l_infos[0]["parent_idx"] == -1
                and l_infos[0]["name"] == "world"):
            child = l_infos[1]
            child_free = any(j["type"] == JOINT_TYPE.FREE for j in child["joints"])
            child_ident = np.abs(child["pos"]).max() < EPS32 and np.abs(child["quat"] - np.array([1.0, 0, 0, 0])).max() < EPS32
            if child_free or child_ident:
                l_infos = l_infos[1:]
                for li in l_infos:
                    li["parent_idx"] = max(li["parent_idx"] - 1, -1)
        # links with a free joint hang off the world (rigid_entity.py:648-651)
        for li in l_infos:
            if li["joints"] and all(j["type"] == JOINT_TYPE.FREE for j in li["joints"]):
                li["parent_idx"] = -1
        # morph.pos / morph.quat offset every base link (rigid_entity.py:744-767) -- AFTER the virtual world link is gone and the free
        # links hang off the world, so that the offset lands on the bodies that actually exist
        if morph_pos is not None or morph_quat is not None:
            pos_off = np.zeros(3) if morph_pos is None else np.asarray(morph_pos, float)
            quat_off = np.array([1.0, 0, 0, 0]) if morph_quat is None else np.asarray(morph_quat, float)
            for li in l_infos:
                if li["parent_idx"] < 0:
                    li["pos"], li["quat"] = gn.transform_pos_quat(li["pos"], li["quat"], pos_off, quat_off)
        for li in l_infos:
            if li["parent_idx"] < 0:
                for j in li["joints"]:
                    if j["type"] == JOINT_TYPE.FREE:
                        j["init_qpos"] = np.concatenate([np.asarray(li["pos"], float), np.asarray(li["quat"], float)])
        ent = EntityInfo(name=name, l_infos=l_infos, eqs_info=list(eqs_info), is_robot=is_robot, rho=rho, friction=friction,
                         local_collision_mask=local_collision_mask)
        self.entities.append(ent)
        return ent

    # ------------------------------------------------------------------ build
    def build(self) -> RigidModel:
        m = RigidModel()
        links, joints, geoms = [], [], []
        nL = nJ = nD = nQ = nG = 0
        for i_e, ent in enumerate(self.entities):
            ent.link_start, ent.joint_start, ent.dof_start, ent.q_start, ent.geom_start = nL, nJ, nD, nQ, nG
            for li in ent.l_infos:
                link = dict(li)
                link["entity_idx"] = i_e
                link["idx"] = nL
                link["parent"] = li["parent_idx"] + ent.link_start if li["parent_idx"] >= 0 else -1
                link["joint_start"], link["dof_start"], link["q_start"], link["geom_start"] = nJ, nD, nQ, nG
                for j in li["joints"]:
                    jj = dict(j)
                    jj["link_idx"], jj["q_start"], jj["dof_start"] = nL, nQ, nD
                    joints.append(jj)
                    ent.joint_names.append(j["name"])
                    ent.dof_names += [j["name"]] * j["n_dofs"]
                    nJ += 1
                    nD += j["n_dofs"]
                    nQ += j["n_qs"]
                for g in li["geoms"]:
                    gg = dict(g)
                    gg["link_idx"] = nL
                    gg["entity_idx"] = i_e
                    if ent.friction is not None:
                        gg["friction"] = ent.friction
                    geoms.append(gg)
                    nG += 1
                link["joint_end"], link["dof_end"], link["q_end"], link["geom_end"] = nJ, nD, nQ, nG
                links.append(link)
                ent.link_names.append(li["name"])
                nL += 1
            ent.n_links, ent.n_joints = nL - ent.link_start, nJ - ent.joint_start
            ent.n_dofs, ent.n_qs, ent.n_geoms = nD - ent.dof_start, nQ - ent.q_start, nG - ent.geom_start

        # derived link topology
        for link in links:
            p = link["parent"]
            n_dofs = link["dof_end"] - link["dof_start"]
            link["is_fixed"] = (n_dofs == 0) and (p == -1 or links[p]["is_fixed"])
            link["root"] = link["idx"] if p == -1 else links[p]["root"]

        # inertials (rigid_link.py:667-790)
        for link in links:
            ent = self.entities[link["entity_idx"]]
            lgeoms = geoms[link["geom_start"]: link["geom_end"]]
            if link["inertial_mass"] is None or link["inertial_pos"] is None or link["inertial_i"] is None:
                rho = ent.rho if ent.rho is not None else (RHO_ROBOT if ent.is_robot else RHO_OBJECT)
                if link.get("geom_density"):  # MJCF body without <inertial>: MuJoCo's geom density, whatever the material says
                    rho = link["geom_density"]
                if not link["is_fixed"] and lgeoms:
                    def one(g):
                        if not link.get("geom_density"):
                            return _geom_inertial(g, rho)
                        mg, cg, Ig = _geom_inertial(g, g.get("density", rho))
                        if g.get("mass") is not None and mg > 0:  # an explicit geom mass overrides its density
                            Ig, mg = Ig * (g["mass"] / mg), g["mass"]
                        return mg, cg, Ig
                    mass, com, I = _compose_inertial([(*one(g), g["pos"], g["quat"]) for g in lgeoms])
                else:
                    mass, com, I = 0.0, np.zeros(3), np.zeros((3, 3))
                if link["inertial_mass"] is None:
                    link["inertial_mass"] = mass
                elif mass > 0:  # only the mass was given: the inferred inertia is rescaled to it (rigid_link.py:737)
                    I = I * (link["inertial_mass"] / mass)
                link["inertial_pos"], link["inertial_i"] = com, I
                link["inertial_quat"] = np.array([1.0, 0, 0, 0])
            link["inertial_mass"] = max(link["inertial_mass"], EPS32)

        m.n_links, m.n_joints, m.n_dofs, m.n_qs, m.n_geoms, m.n_entities = nL, nJ, nD, nQ, nG, len(self.entities)
        I32 = np.int32
        m.links_parent = np.array([l["parent"] for l in links], I32)
        m.links_root = np.array([l["root"] for l in links], I32)
        m.links_entity = np.array([l["entity_idx"] for l in links], I32)
        m.links_is_fixed = np.array([l["is_fixed"] for l in links], I32)
        for key in ("joint_start", "joint_end", "dof_start", "dof_end", "q_start", "q_end", "geom_start", "geom_end"):
            setattr(m, "links_" + key, np.array([l[key] for l in links], I32))
        m.links_pos = np.array([l["pos"] for l in links], np.float64).reshape(nL, 3)
        m.links_quat = np.array([gn.quat_normalize(l["quat"]) for l in links], np.float64).reshape(nL, 4)
        m.links_inertial_pos = np.array([l["inertial_pos"] for l in links], np.float64).reshape(nL, 3)
        m.links_inertial_quat = np.array([gn.quat_normalize(l["inertial_quat"]) for l in links], np.float64).reshape(nL, 4)
        m.links_inertial_i = np.array([np.asarray(l["inertial_i"]).reshape(9) for l in links], np.float64).reshape(nL, 9)
        m.links_inertial_mass = np.array([l["inertial_mass"] for l in links], np.float64)

        m.joints_type = np.array([j["type"] for j in joints], I32).reshape(nJ)
        m.joints_link = np.array([j["link_idx"] for j in joints], I32).reshape(nJ)
        m.joints_q_start = np.array([j["q_start"] for j in joints], I32).reshape(nJ)
        m.joints_dof_start = np.array([j["dof_start"] for j in joints], I32).reshape(nJ)
        m.joints_n_dofs = np.array([j["n_dofs"] for j in joints], I32).reshape(nJ)
        m.joints_pos = np.array([j["pos"] for j in joints], np.float64).reshape(nJ, 3)
        m.joints_sol_params = sanitize_sol_params(
            [j["sol_params"] for j in joints] if joints else np.zeros((0, 7)), self.sol_min_timeconst, self.sol_default_timeconst
        ).reshape(nJ, 7)

        def cat(key, width=None):
            arrs = [np.asarray(j[key], np.float64).reshape(j["n_dofs"], -1) for j in joints if j["n_dofs"] > 0]
            a = np.concatenate(arrs, axis=0) if arrs else np.zeros((0, width or 1))
            return a[:, 0] if (width is None) else a

        m.dofs_motion_ang = cat("dofs_motion_ang", 3)
        m.dofs_motion_vel = cat("dofs_motion_vel", 3)
        m.dofs_limit = cat("dofs_limit", 2)
        m.dofs_stiffness = cat("dofs_stiffness")
        m.dofs_damping = cat("dofs_damping")
        m.dofs_armature = cat("dofs_armature")
        m.dofs_frictionloss = cat("dofs_frictionloss")
        m.dofs_kp = cat("dofs_act_gain")
        m.dofs_act_bias = cat("dofs_act_bias", 3)
        m.dofs_force_range = cat("dofs_force_range", 2)
        m.dofs_link = np.concatenate([[j["link_idx"]] * j["n_dofs"] for j in joints]).astype(I32) if nD else np.zeros(0, I32)
        m.dofs_joint = np.concatenate([[i] * j["n_dofs"] for i, j in enumerate(joints)]).astype(I32) if nD else np.zeros(0, I32)
        m.dofs_entity = m.links_entity[m.dofs_link] if nD else np.zeros(0, I32)
        m.qpos0 = (np.concatenate([np.asarray(j["init_qpos"], np.float64) for j in joints]) if nQ else np.zeros(0))

        m.entities_link_start = np.array([e.link_start for e in self.entities], I32)
        m.entities_link_end = np.array([e.link_start + e.n_links for e in self.entities], I32)
        m.entities_dof_start = np.array([e.dof_start for e in self.entities], I32)
        m.entities_dof_end = np.array([e.dof_start + e.n_dofs for e in self.entities], I32)
        m.entities_geom_start = np.array([e.geom_start for e in self.entities], I32)
        m.entities_geom_end = np.array([e.geom_start + e.n_geoms for e in self.entities], I32)
        m.entities_gravity_compensation = np.array([e.gravity_compensation for e in self.entities], np.float64)

        # mass_parent_mask (rigid_solver.py:666-677), stored as the per-dof "ancestor-or-self" relation
        mask = np.zeros((max(nD, 1), max(nD, 1)), np.float64)
        for l in links:
            a = l["idx"]
            while a != -1:
                mask[l["dof_start"]: l["dof_end"], links[a]["dof_start"]: links[a]["dof_end"]] = 1.0
                a = links[a]["parent"]
        m.mass_parent_mask = mask[:nD, :nD].copy()

        # ---- geoms
        m.geoms_link = np.array([g["link_idx"] for g in geoms], I32).reshape(nG)
        m.geoms_type = np.array([g["type"] for g in geoms], I32).reshape(nG)
        m.geoms_pos = np.array([g["pos"] for g in geoms], np.float64).reshape(nG, 3)
        m.geoms_quat = np.array([gn.quat_normalize(g["quat"]) for g in geoms], np.float64).reshape(nG, 4)
        m.geoms_data = np.array([np.asarray(g["data"], np.float64)[:7] for g in geoms], np.float64).reshape(nG, 7)
        m.geoms_friction = np.array([g["friction"] for g in geoms], np.float64).reshape(nG)
        m.geoms_sol_params = sanitize_sol_params(
            [g["sol_params"] for g in geoms] if geoms else np.zeros((0, 7)), self.sol_min_timeconst, self.sol_default_timeconst
        ).reshape(nG, 7)
        m.geoms_is_fixed = m.links_is_fixed[m.geoms_link] if nG else np.zeros(0, I32)
        centers, aabbs, vstart, vend, verts_all = [], [], [], [], []
        nV = 0
        for g in geoms:
            v = np.asarray(g["verts"], np.float64)
            if g["type"] in (GEOM_TYPE.PLANE, GEOM_TYPE.TERRAIN) or len(v) < 4:
                center = v.mean(axis=0)
            else:
                vol, com, _ = meshio.mesh_mass_properties(v, np.asarray(g["faces"]))
                center = com if vol > 0 else v.mean(axis=0)  # rigid_solver.py:782-789
            centers.append(center)
            lo, hi = v.min(axis=0), v.max(axis=0)
            # corner order (abd/misc.py:514-529): index 0 = min corner, 7 = max corner, x slowest
            aabbs.append([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
            vstart.append(nV)
            nV += len(v)
            vend.append(nV)
            verts_all.append(v)
        m.geoms_center = np.array(centers, np.float64).reshape(nG, 3)
        # height-field terrain (collider.py:541-561): at most one, as in the reference (the tables live in collider_info)
        ter = [g for g in geoms if g["type"] == GEOM_TYPE.TERRAIN]
        if len(ter) > 1:
            raise ValueError("only one Terrain entity per scene (collider.py:541-561 keeps a single height field)")
        if ter:
            hf = np.asarray(ter[0]["terrain"]["hf"], np.float64)
            hs, vs = ter[0]["terrain"]["scale"]
            m.terrain_hf = hf.copy()
            m.terrain_rc = np.array(hf.shape, I32)
            m.terrain_scale = np.array([hs, vs], np.float64)
            m.terrain_xyz_maxmin = np.array([hf.shape[0] * hs, hf.shape[1] * hs, hf.max(), 0.0, 0.0, hf.min() - 1.0], np.float64)
        m.geoms_init_aabb = np.array(aabbs, np.float64).reshape(nG, 8, 3)
        m.geoms_vert_start = np.array(vstart, I32).reshape(nG)
        m.geoms_vert_end = np.array(vend, I32).reshape(nG)
        m.verts = np.concatenate(verts_all, axis=0) if verts_all else np.zeros((0, 3))
        m.n_verts = nV

        # support fields for every geom that resolves through the table (mesh / cylinder), support_field.py:22-91
        theta = np.arange(SUPPORT_RES) / SUPPORT_RES * 2 * np.pi - np.pi
        phi = np.arange(SUPPORT_RES) / SUPPORT_RES * np.pi
        dirs = np.stack(
            [np.sin(phi)[None, :] * np.cos(theta)[:, None], np.sin(phi)[None, :] * np.sin(theta)[:, None],
             np.broadcast_to(np.cos(phi)[None, :], (SUPPORT_RES, SUPPORT_RES))], axis=-1).reshape(-1, 3)
        cell_start, sup_v, sup_vid, n_cells = [], [], [], 0
        for g, v in zip(geoms, verts_all):
            if g["type"] in (GEOM_TYPE.MESH, GEOM_TYPE.CYLINDER):
                idx = np.argmax(dirs @ v.T, axis=1)
                cell_start.append(n_cells)
                sup_v.append(v[idx])
                sup_vid.append(idx)  # support_field.py:62-70 support_vid: vertex index inside the geom (simplex-vertex identity in GJK)
                n_cells += len(idx)
            else:
                cell_start.append(-1)
        m.geoms_support_start = np.array(cell_start, I32).reshape(nG)
        m.support_v = np.concatenate(sup_v, axis=0) if sup_v else np.zeros((0, 3))
        m.support_vid = np.concatenate(sup_vid).astype(I32) if sup_vid else np.zeros(0, I32)
        m.n_support_cells = n_cells

        # ---- equalities (entity-local names -> global joint/link ids)
        eq_rows = []
        for ent in self.entities:
            for eq in ent.eqs_info:
                if eq["type"] == EQUALITY_TYPE.JOINT:
                    ids = [ent.joint_start + ent.joint_names.index(n) for n in eq["objs_name"]]
                else:
                    ids = [ent.link_start + ent.link_names.index(n) for n in eq["objs_name"]]
                eq_rows.append((int(eq["type"]), ids[0], ids[1], np.asarray(eq["data"], np.float64), eq["sol_params"]))
        m.n_equalities = len(eq_rows)
        m.eq_type = np.array([r[0] for r in eq_rows], I32).reshape(-1)
        m.eq_obj1 = np.array([r[1] for r in eq_rows], I32).reshape(-1)
        m.eq_obj2 = np.array([r[2] for r in eq_rows], I32).reshape(-1)
        m.eq_data = np.array([r[3] for r in eq_rows], np.float64).reshape(-1, 11)
        m.eq_sol_params = sanitize_sol_params(
            [r[4] for r in eq_rows] if eq_rows else np.zeros((0, 7)), self.sol_min_timeconst, self.sol_default_timeconst
        ).reshape(-1, 7)

        # ---- collision pairs (collider.py:258-492)
        self._build_pairs(m, links, joints, geoms)

        # ---- global constants
        m.dt = float(self.dt)
        m.gravity = self.gravity.copy()

        # ---- invweight / meaninertia at qpos0 (rigid_solver.py:492-655)
        from .refdyn import invweight_and_meaninertia

        links_invweight, dofs_invweight, meaninertia = invweight_and_meaninertia(m)
        links_invweight[m.links_is_fixed.astype(bool)] = 0.0  # rigid_link.py:792-794
        m.links_invweight = links_invweight
        m.dofs_invweight = dofs_invweight
        m.meaninertia = float(meaninertia)

        # capacities (collider.py:531-539, constraint/solver.py:47-52)
        m.max_collision_pairs = int(min(self.max_collision_pairs, m.n_pairs))
        m.max_contact_pairs = int(m.max_collision_pairs * self.n_contacts_per_pair)
        m.max_broad_pairs = int(m.max_collision_pairs * self.multiplier_broad)
        m.n_contacts_per_pair = int(self.n_contacts_per_pair)
        n_1dof = int(np.sum((m.joints_type == JOINT_TYPE.REVOLUTE) | (m.joints_type == JOINT_TYPE.PRISMATIC)))
        m.len_constraints = int(4 * m.max_contact_pairs + n_1dof + nD + 6 * max(1, m.n_equalities + self.max_dynamic_constraints))
        m.entities = self.entities
        return m

    def _build_pairs(self, m, links, joints, geoms):
        nG = m.n_geoms
        pair_idx = -np.ones((max(nG, 1), max(nG, 1)), np.int32)
        pairs = []
        for a in range(nG):
            for b in range(a + 1, nG):
                ga, gb = geoms[a], geoms[b]
                la, lb = ga["link_idx"], gb["link_idx"]
                if la == lb:
                    continue
                if links[la]["is_fixed"] and links[lb]["is_fixed"]:
                    continue
                same_entity = ga["entity_idx"] == gb["entity_idx"]
                con_match = (ga["contype"] & gb["conaffinity"]) | (gb["contype"] & ga["conaffinity"])
                # FileMorph entities carry "local" collision masks: the contype/conaffinity test only applies to
                # pairs inside one entity (collider.py:340-344)
                has_local = self.entities[ga["entity_idx"]].local_collision_mask or self.entities[gb["entity_idx"]].local_collision_mask
                if (same_entity or not has_local) and con_match == 0:
                    continue
                same_root = links[la]["root"] == links[lb]["root"]
                if same_root and not self.enable_self_collision:
                    continue
                if same_root and not self.enable_adjacent_collision:
                    lo, hi = (la, lb) if la < lb else (lb, la)
                    adjacent, cur = False, hi
                    while links[cur]["parent"] != -1:
                        if links[cur]["parent"] == lo:
                            adjacent = True
                            break
                        if not all(j["type"] == JOINT_TYPE.FIXED for j in links[cur]["joints"]):
                            break
                        cur = links[cur]["parent"]
                    if adjacent:
                        continue
                if same_root and self._neutral_overlap(m, links, geoms, a, b):
                    continue
                pair_idx[a, b] = len(pairs)
                pairs.append((a, b))
        m.n_pairs = len(pairs)
        m.pairs = np.array(pairs, np.int32).reshape(-1, 2)
        m.pair_idx = pair_idx[:nG, :nG].copy()

    def _neutral_overlap(self, m, links, geoms, a, b):
        """Self-collision filter at qpos0.  The reference voxelises both meshes with trimesh
        (collider.py:403-421); here the test is a separating-axis-free approximation: the two convex hulls
        (shrunk by 0.1%) are declared overlapping when a hull vertex of one lies inside the other hull."""
        from .refdyn import neutral_geom_vertices

        va = neutral_geom_vertices(m, a)
        vb = neutral_geom_vertices(m, b)
        va = va.mean(axis=0) + (1 - 1e-3) * (va - va.mean(axis=0))
        vb = vb.mean(axis=0) + (1 - 1e-3) * (vb - vb.mean(axis=0))
        if (va.max(axis=0) < vb.min(axis=0)).any() or (vb.max(axis=0) < va.min(axis=0)).any():
            return False
        try:
            from scipy.spatial import ConvexHull

            ha, hb = ConvexHull(va), ConvexHull(vb)
        except Exception:
            return False
        inside_b = (hb.equations[:, :3] @ va.T + hb.equations[:, 3:4] <= 1e-12).all(axis=0).any()
        inside_a = (ha.equations[:, :3] @ vb.T + ha.equations[:, 3:4] <= 1e-12).all(axis=0).any()
        return bool(inside_a or inside_b)


# ---------------------------------------------------------------------- primitive morphs
def plane_l_infos(pos=(0.0, 0.0, 0.0), quat=(1.0, 0.0, 0.0, 0.0), normal=(0.0, 0.0, 1.0), plane_size=(1e3, 1e3)):
    """Fixed plane entity (rigid_entity.py:315-325; collision mesh = 1 cm thick slab, utils/mesh.py:958-961)."""
    hx, hy, th = 0.5 * plane_size[0], 0.5 * plane_size[1], 1e-2
    verts = np.array([[sx * hx, sy * hy, z] for sx in (-1, 1) for sy in (-1, 1) for z in (-th, 0.0)])
    _, faces = meshio.convex_hull(verts)
    geom = dict(type=GEOM_TYPE.PLANE, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.array([*normal, 0, 0, 0, 0.0]),
                verts=verts, faces=faces, friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]))
    return [dict(name="plane_baselink", pos=np.asarray(pos, float), quat=np.asarray(quat, float), parent_idx=-1, joints=[],
                 geoms=[geom], inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None)]


def box_l_infos(size, pos, quat=(1.0, 0.0, 0.0, 0.0), fixed=False):
    """Free (or fixed) box entity (rigid_entity.py:296-302)."""
    from .urdf import _free_joint

    ext = np.asarray(size, float)
    verts, faces = meshio.box_mesh(ext)
    geom = dict(type=GEOM_TYPE.BOX, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.concatenate([ext, np.zeros(4)]),
                verts=verts, faces=faces, friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]))
    joints = []
    if not fixed:
        j = _free_joint("box_baselink_joint")
        j["init_qpos"] = np.concatenate([np.asarray(pos, float), np.asarray(quat, float)])
        joints = [j]
    return [dict(name="box_baselink", pos=np.asarray(pos, float), quat=np.asarray(quat, float), parent_idx=-1, joints=joints,
                 geoms=[geom], inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None)]


def _primitive_l_infos(name, geom, pos, quat, fixed):
    from .urdf import _free_joint

    joints = []
    if not fixed:
        j = _free_joint(f"{name}_baselink_joint")
        j["init_qpos"] = np.concatenate([np.asarray(pos, float), np.asarray(quat, float)])
        joints = [j]
    return [dict(name=f"{name}_baselink", pos=np.asarray(pos, float), quat=np.asarray(quat, float), parent_idx=-1, joints=joints,
                 geoms=[geom], inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None)]


def cylinder_l_infos(radius, height, pos, quat=(1.0, 0.0, 0.0, 0.0), fixed=False):
    """Free (or fixed) cylinder entity: mesh-backed like the reference's (support field), data = (radius, height)."""
    verts, faces = meshio.cylinder_mesh(radius, height)
    geom = dict(type=GEOM_TYPE.CYLINDER, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.array([radius, height, 0, 0, 0, 0, 0.0]),
                verts=verts, faces=faces, friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]))
    return _primitive_l_infos("cylinder", geom, pos, quat, fixed)


def capsule_l_infos(radius, length, pos, quat=(1.0, 0.0, 0.0, 0.0), fixed=False):
    verts, faces = meshio.capsule_mesh(radius, length)
    geom = dict(type=GEOM_TYPE.CAPSULE, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.array([radius, length, 0, 0, 0, 0, 0.0]),
                verts=verts, faces=faces, friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]))
    return _primitive_l_infos("capsule", geom, pos, quat, fixed)


def terrain_l_infos(height_field, horizontal_scale, vertical_scale, pos, quat=(1.0, 0.0, 0.0, 0.0)):
    """Fixed height-field terrain entity (rigid_entity.py `_load_terrain`; collider.py:541-561 keeps the scaled heights, the
    grid shape and the scales for func_contact_mpr_terrain).  The geom's vertices are the grid points (AABB, centre)."""
    hf = np.asarray(height_field, np.float64)
    if hf.ndim != 2 or hf.shape[0] < 2 or hf.shape[1] < 2:
        raise ValueError("Terrain height_field must be a 2-D array of at least 2 x 2 samples")
    hs, vs = float(horizontal_scale), float(vertical_scale)
    rr, cc = np.meshgrid(np.arange(hf.shape[0]) * hs, np.arange(hf.shape[1]) * hs, indexing="ij")
    verts = np.stack([rr, cc, hf * vs], axis=-1).reshape(-1, 3)
    geom = dict(type=GEOM_TYPE.TERRAIN, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.zeros(7), verts=verts,
                faces=np.zeros((0, 3), np.int64), friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]), terrain=dict(hf=hf * vs, scale=(hs, vs)))
    return [dict(name="terrain_baselink", pos=np.asarray(pos, float), quat=np.asarray(quat, float), parent_idx=-1, joints=[],
                 geoms=[geom], inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None)]


def sphere_l_infos(radius, pos, fixed=False):
    from .urdf import _free_joint

    verts, faces = meshio.icosphere(radius, 2)
    geom = dict(type=GEOM_TYPE.SPHERE, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), data=np.array([radius, 0, 0, 0, 0, 0, 0.0]),
                verts=verts, faces=faces, friction=1.0, contype=0xFFFF, conaffinity=0xFFFF,
                sol_params=np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0]))
    joints = []
    if not fixed:
        j = _free_joint("sphere_baselink_joint")
        j["init_qpos"] = np.concatenate([np.asarray(pos, float), [1.0, 0, 0, 0]])
        joints = [j]
    return [dict(name="sphere_baselink", pos=np.asarray(pos, float), quat=np.array([1.0, 0, 0, 0]), parent_idx=-1, joints=joints,
                 geoms=[geom], inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None)]
