"""URDF -> link/joint/geom info dicts (build time, host only).

Restates the reference pipeline for URDF robots without urdfpy/MuJoCo: parse links and joints
(genesis/utils/urdf.py:131-405), merge links connected by fixed joints including their inertials
(merge_fixed_links :436-472, merge_inertia :573-612, update_subtree :624-643), diagonalise the inertia like
MuJoCo's compiler does (genesis/utils/mjcf.py:240-243 reads body_iquat/body_inertia), order links breadth-first
(_order_links :90-128).  Only primitive collision geometries and mesh files in STL/OBJ are supported.
"""

import os
import xml.etree.ElementTree as ET

import numpy as np

from . import geom_np as gn
from . import meshio
from ..constants import EQUALITY_TYPE, GEOM_TYPE, JOINT_TYPE

DEFAULT_SOL_PARAMS = np.array([0.0, 1.0, 0.9, 0.95, 0.001, 0.5, 2.0])  # genesis/utils/geom.py:2109
DEFAULT_KP, DEFAULT_KV = 100.0, 10.0  # genesis/utils/geom.py:2116-2121


def _origin(elem):
    T = np.eye(4)
    if elem is None:
        return T
    o = elem.find("origin")
    if o is None:
        return T
    xyz = [float(x) for x in o.attrib.get("xyz", "0 0 0").split()]
    rpy = [float(x) for x in o.attrib.get("rpy", "0 0 0").split()]
    T[:3, :3] = gn.rpy_to_R(rpy)
    T[:3, 3] = xyz
    return T


class _Link:
    def __init__(self, name):
        self.name = name
        self.mass = None
        self.inertia = None
        self.inertial_origin = None
        self.collisions = []  # (T, kind, params)


class _Joint:
    pass


def _parse(path):
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(os.path.abspath(path))
    links, joints = [], []
    for le in root.findall("link"):
        link = _Link(le.attrib["name"])
        inertial = le.find("inertial")
        if inertial is not None:
            link.inertial_origin = _origin(inertial)
            link.mass = float(inertial.find("mass").attrib["value"])
            i = inertial.find("inertia").attrib
            ixx, ixy, ixz, iyy, iyz, izz = (float(i.get(k, 0.0)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"))
            link.inertia = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        for ce in le.findall("collision"):
            T = _origin(ce)
            g = ce.find("geometry")
            if g.find("box") is not None:
                link.collisions.append((T, "box", [float(x) for x in g.find("box").attrib["size"].split()]))
            elif g.find("cylinder") is not None:
                c = g.find("cylinder").attrib
                link.collisions.append((T, "cylinder", [float(c["radius"]), float(c["length"])]))
            elif g.find("sphere") is not None:
                link.collisions.append((T, "sphere", [float(g.find("sphere").attrib["radius"])]))
            elif g.find("mesh") is not None:
                m = g.find("mesh").attrib
                fname = m["filename"]
                if fname.startswith("package://"):
                    fname = fname[10:]
                sc = [float(x) for x in m.get("scale", "1 1 1").split()]
                link.collisions.append((T, "mesh", [os.path.join(base_dir, fname), sc]))
        links.append(link)
    for je in root.findall("joint"):
        j = _Joint()
        j.name = je.attrib["name"]
        j.type = je.attrib["type"]
        j.origin = _origin(je)
        j.parent = je.find("parent").attrib["link"]
        j.child = je.find("child").attrib["link"]
        ax = je.find("axis")
        j.axis = np.array([float(x) for x in ax.attrib["xyz"].split()]) if ax is not None else np.array([1.0, 0, 0])
        lim = je.find("limit")
        j.lower = float(lim.attrib["lower"]) if lim is not None and "lower" in lim.attrib else None
        j.upper = float(lim.attrib["upper"]) if lim is not None and "upper" in lim.attrib else None
        j.effort = float(lim.attrib["effort"]) if lim is not None and "effort" in lim.attrib else None
        dyn = je.find("dynamics")
        j.damping = float(dyn.attrib.get("damping", 0.0)) if dyn is not None else 0.0
        j.friction = float(dyn.attrib.get("friction", 0.0)) if dyn is not None else 0.0
        mimic = je.find("mimic")
        j.mimic = (
            (mimic.attrib["joint"], float(mimic.attrib.get("multiplier", 1.0)), float(mimic.attrib.get("offset", 0.0)))
            if mimic is not None else None
        )
        joints.append(j)
    return links, joints


def _translate_inertia(I, m, d):
    return I + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))


def _merge_inertia(l1, l2):
    """genesis/utils/urdf.py:573-612."""
    if l2.mass is None:
        return
    if l1.mass is None:
        l1.mass, l1.inertia, l1.inertial_origin = l2.mass, l2.inertia, l2.inertial_origin
        return
    m1, m2 = l1.mass, l2.mass
    com1, R1 = l1.inertial_origin[:3, 3], l1.inertial_origin[:3, :3]
    com2, R2 = l2.inertial_origin[:3, 3], l2.inertial_origin[:3, :3]
    m = m1 + m2
    com = (m1 * com1 + m2 * com2) / m if m > 1e-15 else com1
    I1 = _translate_inertia(R1 @ l1.inertia @ R1.T, m1, com - com1)
    I2 = _translate_inertia(R2 @ l2.inertia @ R2.T, m2, com - com2)
    l1.mass, l1.inertia = m, I1 + I2
    l1.inertial_origin = np.eye(4)
    l1.inertial_origin[:3, 3] = com


def _merge_fixed_links(links, joints, links_to_keep=()):
    """genesis/utils/urdf.py:436-472."""
    children = {j.child for j in joints}
    base = next(l for l in links if l.name not in children)
    child_map = {}
    for j in joints:
        child_map.setdefault(j.parent, []).append(j)
    ordered, queue, seen = [], [base.name], set()
    while queue:
        name = queue.pop(0)
        if name in seen:
            continue
        seen.add(name)
        for j in child_map.get(name, []):
            ordered.append(j)
            queue.append(j.child)
    links, joints = list(links), list(joints)
    for j in ordered[::-1]:
        if j.type == "fixed" and j.child not in links_to_keep:
            parent = next(l for l in links if l.name == j.parent)
            child = next(l for l in links if l.name == j.child)
            # update_subtree: re-express the child's content and its outgoing joints in the parent frame
            if child.mass is not None:
                child.inertial_origin = j.origin @ (child.inertial_origin if child.inertial_origin is not None else np.eye(4))
            child.collisions = [(j.origin @ T, kind, params) for T, kind, params in child.collisions]
            for jj in joints:
                if jj.parent == child.name:
                    jj.origin = j.origin @ jj.origin
            _merge_inertia(parent, child)
            parent.collisions.extend(child.collisions)
            joints.remove(j)
            links.remove(child)
            for jj in joints:
                if jj.parent == child.name:
                    jj.parent = parent.name
    return links, joints


def parse_urdf(path, merge_fixed_links=True, links_to_keep=(), default_armature=0.1, scale=1.0, fixed=False):
    links, joints = _parse(path)
    if merge_fixed_links:
        links, joints = _merge_fixed_links(links, joints, links_to_keep)
    name_to_idx = {l.name: i for i, l in enumerate(links)}

    l_infos = []
    for link in links:
        l_info = dict(
            name=link.name, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), parent_idx=-1, joints=[], geoms=[],
            invweight=None, inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None,
        )
        if link.mass is not None:
            Rin = link.inertial_origin[:3, :3]
            diag, q = gn.principal_inertia(Rin @ link.inertia @ Rin.T)
            l_info.update(
                inertial_mass=link.mass * scale**3,
                inertial_pos=link.inertial_origin[:3, 3] * scale,
                inertial_quat=q,
                inertial_i=np.diag(diag) * scale**5,
            )
        for T, kind, params in link.collisions:
            g_info = dict(
                pos=T[:3, 3] * scale, quat=gn.R_to_quat(T[:3, :3]), contype=1, conaffinity=1, friction=1.0,
                sol_params=DEFAULT_SOL_PARAMS.copy(),
            )
            if kind == "box":
                ext = np.array(params) * scale
                verts, faces = meshio.box_mesh(ext)
                g_info.update(type=GEOM_TYPE.BOX, data=np.concatenate([ext, np.zeros(4)]), verts=verts, faces=faces)
            elif kind == "cylinder":
                r, h = params[0] * scale, params[1] * scale
                verts, faces = meshio.cylinder_mesh(r, h)
                g_info.update(type=GEOM_TYPE.CYLINDER, data=np.array([r, h, 0, 0, 0, 0, 0.0]), verts=verts, faces=faces)
            elif kind == "sphere":
                r = params[0] * scale
                verts, faces = meshio.icosphere(r, 2)
                g_info.update(type=GEOM_TYPE.SPHERE, data=np.array([r, 0, 0, 0, 0, 0, 0.0]), verts=verts, faces=faces)
            else:
                mpath, msc = params
                verts, faces = meshio.convex_hull(meshio.load_vertices(mpath) * np.array(msc) * scale)
                g_info.update(type=GEOM_TYPE.MESH, data=np.zeros(7), verts=verts, faces=faces)
            l_info["geoms"].append(g_info)
        l_infos.append(l_info)

    eqs_info = []
    for j in joints:
        idx = name_to_idx[j.child]
        l_info = l_infos[idx]
        l_info["parent_idx"] = name_to_idx[j.parent]
        l_info["pos"] = j.origin[:3, 3] * scale
        l_info["quat"] = gn.R_to_quat(j.origin[:3, :3])
        if j.type == "fixed":
            continue
        if j.type in ("revolute", "continuous", "prismatic"):
            axis = j.axis / max(np.linalg.norm(j.axis), 1e-15)
            rev = j.type != "prismatic"
            lim = [-np.inf, np.inf]
            if j.type != "continuous":
                s = 1.0 if rev else scale
                lim = [j.lower * s if j.lower is not None else -np.inf, j.upper * s if j.upper is not None else np.inf]
            armature = default_armature if default_armature is not None else 0.0
            j_info = dict(
                name=j.name, type=JOINT_TYPE.REVOLUTE if rev else JOINT_TYPE.PRISMATIC, n_qs=1, n_dofs=1,
                pos=np.zeros(3), init_qpos=np.zeros(1), sol_params=DEFAULT_SOL_PARAMS.copy(),
                dofs_motion_ang=np.array([axis]) if rev else np.zeros((1, 3)),
                dofs_motion_vel=np.zeros((1, 3)) if rev else np.array([axis]),
                dofs_limit=np.array([lim]), dofs_stiffness=np.zeros(1),
                dofs_damping=np.array([j.damping]), dofs_frictionloss=np.array([j.friction]),
                dofs_armature=np.array([armature]), dofs_invweight=None,
                dofs_act_gain=np.array([DEFAULT_KP]), dofs_act_bias=np.array([[0.0, -DEFAULT_KP, -DEFAULT_KV]]),
                dofs_force_range=np.array([[-j.effort, j.effort]]) if j.effort is not None else np.array([[-np.inf, np.inf]]),
            )
        elif j.type == "floating":
            j_info = _free_joint(j.name)
        else:
            raise ValueError(f"Unsupported URDF joint type: {j.type}")
        l_info["joints"].append(j_info)
        if j.mimic is not None:
            data = np.zeros(11)
            data[0], data[1] = j.mimic[2] * scale, j.mimic[1]
            eqs_info.append(dict(type=EQUALITY_TYPE.JOINT, objs_name=(j.name, j.mimic[0]), data=data, sol_params=DEFAULT_SOL_PARAMS.copy()))

    # breadth-first re-ordering (_order_links)
    n = len(l_infos)
    children = {k: [] for k in range(n)}
    for c in range(n):
        p = l_infos[c]["parent_idx"]
        if p != -1:
            children[p].append(c)
    order, level = [], [c for c in range(n) if l_infos[c]["parent_idx"] == -1]
    while level:
        order += level
        level = [c for p in level for c in children[p]]
    for l_info in l_infos:
        if l_info["parent_idx"] >= 0:
            l_info["parent_idx"] = order.index(l_info["parent_idx"])
    l_infos = [l_infos[i] for i in order]

    # free root joint for floating-base robots (genesis/engine/entities/rigid_entity/rigid_entity.py:654-696)
    if not fixed and sum(j["n_dofs"] for j in l_infos[0]["joints"]) == 0:
        l_infos[0]["joints"] = [_free_joint("root_joint")]
    return l_infos, eqs_info


def _free_joint(name):
    return dict(
        name=name, type=JOINT_TYPE.FREE, n_qs=7, n_dofs=6, pos=np.zeros(3),
        init_qpos=np.array([0.0, 0, 0, 1, 0, 0, 0]), sol_params=DEFAULT_SOL_PARAMS.copy(),
        dofs_motion_ang=np.eye(6, 3, -3), dofs_motion_vel=np.eye(6, 3),
        dofs_limit=np.tile([-np.inf, np.inf], (6, 1)), dofs_stiffness=np.zeros(6), dofs_invweight=None,
        dofs_frictionloss=np.zeros(6), dofs_damping=np.zeros(6), dofs_armature=np.zeros(6),
        dofs_act_gain=np.zeros(6), dofs_act_bias=np.zeros((6, 3)), dofs_force_range=np.tile([-np.inf, np.inf], (6, 1)),
    )
