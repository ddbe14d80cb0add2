"""MJCF -> link/joint/geom info dicts (build time, host only).

The reference compiles MJCF through MuJoCo's own compiler (genesis/utils/mjcf.py:177) and then reads the
compiled arrays (parse_link :231-424, parse_geom :440-640, parse_equalities).  MuJoCo is not available in
this environment, so this module restates the subset of the MJCF semantics the Franka Panda model
(genesis/assets/xml/franka_emika_panda/panda.xml) and similar arm models exercise: nested default classes,
childclass, body tree, explicit inertials (diag/full), hinge/slide/free joints with autolimits,
mesh/box/sphere/capsule/cylinder/plane geoms, <general>/<position>/<motor> actuators on joints or on fixed tendons
(the tendon case is folded onto the joints exactly like mjcf.py:357-367), and <equality><joint>.

Known deviations: MuJoCo re-expresses mesh geoms in the mesh's own inertial frame and the
reference then decimates hulls to 500 faces; here mesh geoms keep the frame written in the XML and the full qhull hull.
"""

import os
import xml.etree.ElementTree as ET

import numpy as np

from . import geom_np as gn
from . import meshio
from ..constants import EQUALITY_TYPE, GEOM_TYPE, JOINT_TYPE

# MuJoCo defaults
_DEFAULT_SOLIMP = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
RHO_MUJOCO = 1000.0  # MuJoCo's default geom density (rigid_link.py:25)


def _floats(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) < n:
        v = np.concatenate([v, np.zeros(n - len(v))])
    return v


class _Defaults:
    """Resolved attribute dictionaries per (class name, element tag)."""

    def __init__(self, root):
        self.classes = {"main": {}}
        for top in root.findall("default"):
            self._walk(top, "main", {})

    def _walk(self, node, name, inherited):
        cur = {tag: dict(attrs) for tag, attrs in inherited.items()}
        for child in node:
            if child.tag == "default":
                continue
            cur.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = cur
        for child in node.findall("default"):
            self._walk(child, child.attrib["class"], cur)

    def get(self, cls, tag):
        return dict(self.classes.get(cls or "main", self.classes["main"]).get(tag, {}))


def _merged(defaults, elem, cls):
    attrs = defaults.get(elem.attrib.get("class", cls), elem.tag)
    attrs.update(elem.attrib)
    return attrs


def _geom_quat(attrs, angle_scale):
    if "quat" in attrs:
        return gn.quat_normalize(_floats(attrs["quat"]))
    if "euler" in attrs:
        e = _floats(attrs["euler"]) * angle_scale
        # MuJoCo default eulerseq "xyz" (intrinsic)
        q = np.array([1.0, 0.0, 0.0, 0.0])
        for ax, a in zip(np.eye(3), e):
            q = gn.quat_mul(q, gn.axis_angle_to_quat(ax, a))
        return gn.quat_normalize(q)
    if "axisangle" in attrs:
        aa = _floats(attrs["axisangle"])
        return gn.axis_angle_to_quat(aa[:3], aa[3] * angle_scale)
    return np.array([1.0, 0.0, 0.0, 0.0])


def parse_mjcf(path, default_armature=0.1, scale=1.0):
    """Returns (l_infos, eqs_info); each l_info carries 'joints' and 'geoms' lists."""
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(os.path.abspath(path))
    compiler = root.find("compiler")
    comp = compiler.attrib if compiler is not None else {}
    angle_scale = 1.0 if comp.get("angle", "degree") == "radian" else np.pi / 180.0
    meshdir = os.path.join(base_dir, comp.get("meshdir", comp.get("assetdir", "")))
    autolimits = comp.get("autolimits", "true") == "true"
    defaults = _Defaults(root)

    meshes = {}
    for asset in root.findall("asset"):
        for m in asset.findall("mesh"):
            fname = m.attrib["file"]
            name = m.attrib.get("name", os.path.splitext(os.path.basename(fname))[0])
            meshes[name] = (os.path.join(meshdir, fname), _floats(m.attrib.get("scale", "1 1 1")))

    l_infos = []
    joint_name_to_ref = {}

    def parse_body(body, parent_idx, cls):
        cls = body.attrib.get("childclass", cls)
        l_info = dict(
            name=body.attrib.get("name", f"body{len(l_infos)}"),
            pos=_floats(body.attrib.get("pos", "0 0 0")) * scale,
            quat=_geom_quat(body.attrib, angle_scale),
            parent_idx=parent_idx,
            joints=[],
            geoms=[],
            invweight=None,
        )
        inertial = body.find("inertial")
        if inertial is not None:
            l_info["inertial_mass"] = float(inertial.attrib["mass"]) * scale**3
            l_info["inertial_pos"] = _floats(inertial.attrib.get("pos", "0 0 0")) * scale
            iq = _geom_quat(inertial.attrib, angle_scale)
            if "fullinertia" in inertial.attrib:
                ixx, iyy, izz, ixy, ixz, iyz = _floats(inertial.attrib["fullinertia"])
                full = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
                diag, q = gn.principal_inertia(full)
                l_info["inertial_quat"] = gn.quat_normalize(gn.quat_mul(iq, q))
                l_info["inertial_i"] = np.diag(diag) * scale**5
            else:
                l_info["inertial_quat"] = iq
                l_info["inertial_i"] = np.diag(_floats(inertial.attrib["diaginertia"])) * scale**5
        else:
            # no <inertial>: MuJoCo's compiler infers mass / COM / inertia from the body's geoms with its default geom density
            # (1000 kg/m^3) and the reference takes body_mass / body_ipos / body_inertia as they are (utils/mjcf.py:240-243),
            # so the entity's material density plays no role for such links
            l_info.update(inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, geom_density=RHO_MUJOCO)

        idx = len(l_infos)
        l_infos.append(l_info)

        joint_elems = list(body.findall("joint")) + list(body.findall("freejoint"))
        for j in joint_elems:
            if j.tag == "freejoint":
                attrs = dict(j.attrib, type="free")
            else:
                attrs = _merged(defaults, j, cls)
            jtype = attrs.get("type", "hinge")
            name = attrs.get("name", l_info["name"])
            sol_params = np.concatenate(
                [
                    # the reference injects solreflimit=eps so the solver's default timeconst applies (mjcf.py:118-129)
                    _floats(attrs["solreflimit"], 2) if "solreflimit" in attrs else np.array([0.0, 1.0]),
                    _floats(attrs["solimplimit"], 5) if "solimplimit" in attrs else _DEFAULT_SOLIMP,
                ]
            )
            if "solimplimit" in attrs:
                s = _floats(attrs["solimplimit"])
                sol_params[2 : 2 + len(s)] = s
                sol_params[2 + len(s) :] = _DEFAULT_SOLIMP[len(s) :]
            j_info = dict(name=name, pos=_floats(attrs.get("pos", "0 0 0")) * scale, sol_params=sol_params)
            if jtype == "free":
                j_info.update(
                    type=JOINT_TYPE.FREE, n_qs=7, n_dofs=6,
                    init_qpos=np.concatenate([l_info["pos"], l_info["quat"]]),
                    dofs_motion_ang=np.eye(6, 3, -3), dofs_motion_vel=np.eye(6, 3),
                    dofs_limit=np.tile([-np.inf, np.inf], (6, 1)), dofs_stiffness=np.zeros(6),
                    dofs_damping=np.zeros(6), dofs_armature=np.zeros(6), dofs_frictionloss=np.zeros(6),
                )
            elif jtype in ("hinge", "slide"):
                axis = _floats(attrs.get("axis", "0 0 1"))
                axis = axis / np.linalg.norm(axis)
                limited = attrs.get("limited", "auto")
                has_range = "range" in attrs
                is_limited = limited == "true" or (limited == "auto" and autolimits and has_range)
                rng = _floats(attrs["range"]) if has_range else np.array([0.0, 0.0])
                if jtype == "hinge":
                    rng = rng * angle_scale
                else:
                    rng = rng * scale
                armature = float(attrs["armature"]) if "armature" in j.attrib else (
                    default_armature if default_armature is not None else float(attrs.get("armature", 0.0))
                )
                j_info.update(
                    type=JOINT_TYPE.REVOLUTE if jtype == "hinge" else JOINT_TYPE.PRISMATIC, n_qs=1, n_dofs=1,
                    init_qpos=np.array([float(attrs.get("ref", 0.0)) * (angle_scale if jtype == "hinge" else scale)]),
                    dofs_motion_ang=np.array([axis]) if jtype == "hinge" else np.zeros((1, 3)),
                    dofs_motion_vel=np.zeros((1, 3)) if jtype == "hinge" else np.array([axis]),
                    dofs_limit=np.array([rng]) if is_limited else np.array([[-np.inf, np.inf]]),
                    dofs_stiffness=np.array([float(attrs.get("stiffness", 0.0))]),
                    dofs_damping=np.array([float(attrs.get("damping", 0.0))]),
                    dofs_armature=np.array([armature]),
                    dofs_frictionloss=np.array([float(attrs.get("frictionloss", 0.0))]),
                )
            elif jtype == "ball":
                # spherical joint: qpos = local quaternion, three body-frame angular dofs (utils/mjcf.py parse_link, mjJNT_BALL)
                armature = float(attrs["armature"]) if "armature" in j.attrib else (
                    default_armature if default_armature is not None else float(attrs.get("armature", 0.0))
                )
                j_info.update(
                    type=JOINT_TYPE.SPHERICAL, n_qs=4, n_dofs=3, init_qpos=np.array([1.0, 0.0, 0.0, 0.0]),
                    dofs_motion_ang=np.eye(3), dofs_motion_vel=np.zeros((3, 3)),
                    dofs_limit=np.tile([-np.inf, np.inf], (3, 1)), dofs_stiffness=np.zeros(3),
                    dofs_damping=np.full(3, float(attrs.get("damping", 0.0))), dofs_armature=np.full(3, armature),
                    dofs_frictionloss=np.full(3, float(attrs.get("frictionloss", 0.0))),
                )
            else:
                raise ValueError(f"Unsupported MJCF joint type '{jtype}'")
            n = j_info["n_dofs"]
            j_info.update(
                dofs_invweight=None,
                dofs_act_gain=np.zeros(n), dofs_act_bias=np.zeros((n, 3)),
                dofs_force_range=np.tile([-np.inf, np.inf], (n, 1)),
            )
            joint_name_to_ref[name] = j_info
            l_info["joints"].append(j_info)

        for g in body.findall("geom"):
            attrs = _merged(defaults, g, cls)
            contype = int(attrs.get("contype", 1))
            conaffinity = int(attrs.get("conaffinity", 1))
            if not (contype or conaffinity):
                continue  # visual-only geoms never reach the rigid step
            gtype = attrs.get("type", "sphere")
            size = _floats(attrs.get("size", "0 0 0"), 3)
            g_info = dict(
                pos=_floats(attrs.get("pos", "0 0 0")) * scale,
                quat=_geom_quat(attrs, angle_scale),
                contype=contype, conaffinity=conaffinity,
                friction=float(_floats(attrs.get("friction", "1 0.005 0.0001"))[0]),
                sol_params=np.concatenate(
                    [_floats(attrs["solref"], 2) if "solref" in attrs else np.array([0.0, 1.0]), _DEFAULT_SOLIMP]
                ),
            )
            if "solimp" in attrs:
                s = _floats(attrs["solimp"])
                g_info["sol_params"][2 : 2 + len(s)] = s
            # per-geom density / mass of MuJoCo's inertia inference (used only for bodies without <inertial>)
            if "density" in attrs:
                g_info["density"] = float(attrs["density"])
            if "mass" in attrs:
                g_info["mass"] = float(attrs["mass"]) * scale**3
            if gtype == "mesh":
                mpath, mscale = meshes[attrs["mesh"]]
                verts, faces = meshio.convex_hull(meshio.load_vertices(mpath) * mscale * scale)
                g_info.update(type=GEOM_TYPE.MESH, data=np.zeros(7), verts=verts, faces=faces)
            elif gtype == "box":
                verts, faces = meshio.box_mesh(2 * size * scale)
                g_info.update(type=GEOM_TYPE.BOX, data=np.concatenate([2 * size * scale, np.zeros(4)]), verts=verts, faces=faces)
            elif gtype == "sphere":
                verts, faces = meshio.icosphere(size[0] * scale, 2)
                g_info.update(type=GEOM_TYPE.SPHERE, data=np.concatenate([[size[0] * scale], np.zeros(6)]), verts=verts, faces=faces)
            elif gtype == "ellipsoid":  # semi-axes in geoms_info.data (support_field.py _func_support_ellipsoid)
                semi = size[:3] * scale
                verts, faces = meshio.icosphere(1.0, 2)
                verts = np.asarray(verts) * semi[None, :]
                g_info.update(type=GEOM_TYPE.ELLIPSOID, data=np.concatenate([semi, np.zeros(4)]), verts=verts, faces=faces)
            elif gtype == "capsule":
                if "fromto" in attrs:  # segment end points instead of pos / quat / half length
                    ft = _floats(attrs["fromto"], 6) * scale
                    a, b = ft[:3], ft[3:]
                    axis = b - a
                    half = 0.5 * np.linalg.norm(axis)
                    g_info["pos"] = 0.5 * (a + b)
                    z = axis / max(2 * half, 1e-12)
                    c = np.cross([0.0, 0.0, 1.0], z)
                    w = 1.0 + z[2]
                    g_info["quat"] = gn.quat_normalize(np.array([w, *c])) if w > 1e-9 else np.array([0.0, 1.0, 0.0, 0.0])
                    radius = size[0] * scale
                else:
                    radius, half = size[0] * scale, size[1] * scale
                verts, faces = meshio.capsule_mesh(radius, 2 * half)
                g_info.update(type=GEOM_TYPE.CAPSULE, data=np.concatenate([[radius, 2 * half], np.zeros(5)]), verts=verts, faces=faces)
            elif gtype == "cylinder":
                verts, faces = meshio.cylinder_mesh(size[0] * scale, 2 * size[1] * scale)
                g_info.update(type=GEOM_TYPE.CYLINDER, data=np.concatenate([[size[0] * scale, 2 * size[1] * scale], np.zeros(5)]), verts=verts, faces=faces)
            elif gtype == "plane":
                L = 1e3
                verts = np.array([[-L, L, 0.0], [L, L, 0.0], [-L, -L, 0.0], [L, -L, 0.0]])
                g_info.update(type=GEOM_TYPE.PLANE, data=np.array([0.0, 0.0, 1.0, 0, 0, 0, 0]), verts=verts, faces=np.array([[0, 2, 3], [0, 3, 1]]))
            else:
                raise ValueError(f"Unsupported MJCF geom type '{gtype}'")
            l_info["geoms"].append(g_info)

        for child in body.findall("body"):
            parse_body(child, idx, cls)

    world = root.find("worldbody")
    world_info = dict(
        name="world", pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), parent_idx=-1, joints=[], geoms=[],
        inertial_mass=None, inertial_pos=None, inertial_quat=None, inertial_i=None, invweight=None,
    )
    l_infos.append(world_info)
    for g in world.findall("geom"):
        attrs = _merged(defaults, g, None)
        if attrs.get("type", "sphere") == "plane" and (int(attrs.get("contype", 1)) or int(attrs.get("conaffinity", 1))):
            L = 1e3
            world_info["geoms"].append(dict(
                pos=_floats(attrs.get("pos", "0 0 0")), quat=_geom_quat(attrs, angle_scale), contype=1, conaffinity=1,
                friction=float(_floats(attrs.get("friction", "1"))[0]),
                sol_params=np.concatenate([np.array([0.0, 1.0]), _DEFAULT_SOLIMP]),
                type=GEOM_TYPE.PLANE, data=np.array([0.0, 0.0, 1.0, 0, 0, 0, 0]),
                verts=np.array([[-L, L, 0.0], [L, L, 0.0], [-L, -L, 0.0], [L, -L, 0.0]]), faces=np.array([[0, 2, 3], [0, 3, 1]]),
            ))
    for body in world.findall("body"):
        parse_body(body, 0, None)

    # ---- tendons (only fixed tendons over joints; used to fold tendon actuators onto joints) ----
    tendons = {}
    for tendon in root.findall("tendon"):
        for fixed in tendon.findall("fixed"):
            tendons[fixed.attrib["name"]] = [j.attrib["joint"] for j in fixed.findall("joint")]

    # ---- actuators (reference: genesis/utils/mjcf.py:345-404) ----
    for actuators in root.findall("actuator"):
        for act in actuators:
            attrs = _merged(defaults, act, None)
            if act.tag == "general":
                gain = _floats(attrs.get("gainprm", "1"), 3)
                biastype = attrs.get("biastype", "none")
                bias = _floats(attrs.get("biasprm", "0 0 0"), 3) if biastype == "affine" else np.zeros(3)
            elif act.tag == "position":
                kp = float(attrs.get("kp", 1.0))
                kv = float(attrs.get("kv", 0.0))
                gain, bias, biastype = np.array([kp, 0, 0]), np.array([0.0, -kp, -kv]), "affine"
            elif act.tag == "velocity":
                kv = float(attrs.get("kv", 1.0))
                gain, bias, biastype = np.array([kv, 0, 0]), np.array([0.0, 0.0, -kv]), "affine"
            elif act.tag == "motor":
                gain, bias, biastype = np.array([1.0, 0, 0]), np.zeros(3), "none"
            else:
                continue
            gear = float(_floats(attrs.get("gear", "1"))[0])
            if "joint" in attrs:
                targets = [attrs["joint"]]
            elif "tendon" in attrs and attrs["tendon"] in tendons:
                targets = tendons[attrs["tendon"]]
            else:
                continue
            forcelimited = attrs.get("forcelimited", "auto")
            has_fr = "forcerange" in attrs
            is_fl = forcelimited == "true" or (forcelimited == "auto" and autolimits and has_fr)
            for jname in targets:
                j_info = joint_name_to_ref.get(jname)
                if j_info is None:
                    continue
                n = j_info["n_dofs"]
                j_info["dofs_act_gain"] = np.full(n, gear * gain[0] * scale**3)
                if biastype != "none":
                    j_info["dofs_act_bias"] = np.tile(gear * bias[:3] * scale**3, (n, 1))
                if is_fl:
                    j_info["dofs_force_range"] = np.tile(_floats(attrs["forcerange"]), (n, 1))
                ctrllimited = attrs.get("ctrllimited", "auto")
                is_cl = ctrllimited == "true" or (ctrllimited == "auto" and autolimits and "ctrlrange" in attrs)
                if is_cl and biastype == "none":
                    j_info["dofs_force_range"] = np.minimum(
                        j_info["dofs_force_range"], np.tile(gear * _floats(attrs["ctrlrange"]), (n, 1))
                    )

    # ---- equalities (reference: genesis/utils/mjcf.py parse_equalities) ----
    eqs_info = []
    for equality in root.findall("equality"):
        for eq in equality:
            attrs = _merged(defaults, eq, None)
            sol_params = np.concatenate([np.array([0.0, 1.0]), _DEFAULT_SOLIMP])
            if "solref" in attrs:
                s = _floats(attrs["solref"])
                sol_params[: len(s)] = s
            if "solimp" in attrs:
                s = _floats(attrs["solimp"])
                sol_params[2 : 2 + len(s)] = s
            if eq.tag == "joint":
                data = np.zeros(11)
                data[:5] = _floats(attrs.get("polycoef", "0 1 0 0 0"), 5)
                eqs_info.append(dict(type=EQUALITY_TYPE.JOINT, objs_name=(attrs["joint1"], attrs.get("joint2")), data=data, sol_params=sol_params))
            elif eq.tag == "connect":
                data = np.zeros(11)
                data[:3] = _floats(attrs["anchor"]) * scale
                eqs_info.append(dict(type=EQUALITY_TYPE.CONNECT, objs_name=(attrs["body1"], attrs.get("body2", "world")), data=data, sol_params=sol_params))
            elif eq.tag == "weld":
                # MuJoCo's compiled eq_data (the layout func_equality_weld reads, constraint/solver.py:1007-1030): [0:3] anchor
                # in body 2, [3:6] the same point in body 1, [6:10] orientation of body 2 relative to body 1, [10] torquescale.
                # Unless `relpose` is given, the relative pose is the one of the reference configuration qpos0.
                names = [li["name"] for li in l_infos]
                b1, b2 = attrs["body1"], attrs.get("body2", "world")

                def world_pose(name):
                    chain, i = [], names.index(name)
                    while i >= 0:
                        chain.append(l_infos[i])
                        i = l_infos[i]["parent_idx"]
                    pos, quat = np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0])
                    for li in reversed(chain):
                        pos, quat = gn.transform_pos_quat(li["pos"], li["quat"], pos, quat)
                    return pos, quat

                anchor2 = _floats(attrs.get("anchor", "0 0 0"), 3) * scale
                relpose = _floats(attrs.get("relpose", "0 1 0 0 0 0 0"), 7)
                if np.abs(relpose[3:]).max() > 0:
                    relquat = gn.quat_normalize(relpose[3:])
                    anchor1 = relpose[:3] * scale + gn.quat_to_R(relquat) @ anchor2
                else:
                    (p1, q1), (p2, q2) = world_pose(b1), world_pose(b2)
                    R1, R2 = gn.quat_to_R(q1), gn.quat_to_R(q2)
                    anchor1 = R1.T @ (p2 + R2 @ anchor2 - p1)
                    relquat = gn.quat_normalize(gn.quat_mul(gn.quat_inv(q1), q2))
                data = np.concatenate([anchor2, anchor1, relquat, [float(attrs.get("torquescale", 1.0))]])
                eqs_info.append(dict(type=EQUALITY_TYPE.WELD, objs_name=(b1, b2), data=data, sol_params=sol_params))

    return l_infos, eqs_info
