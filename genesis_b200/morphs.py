"""Morph descriptions accepted by Scene.add_entity (subset of genesis/options/morphs.py:186,415,795,916)."""

from dataclasses import dataclass


@dataclass
class Morph:
    pos: tuple | None = None
    quat: tuple | None = None
    collision: bool = True
    visualization: bool = False
    contype: int = 0xFFFF       # options/morphs.py:171-183: two geoms collide when contype_a & conaffinity_b or contype_b & conaffinity_a
    conaffinity: int = 0xFFFF


@dataclass
class Plane(Morph):
    normal: tuple = (0.0, 0.0, 1.0)
    plane_size: tuple = (1e3, 1e3)
    fixed: bool = True


@dataclass
class Box(Morph):
    size: tuple = (1.0, 1.0, 1.0)
    fixed: bool = False


@dataclass
class Sphere(Morph):
    radius: float = 0.5
    fixed: bool = False


@dataclass
class Cylinder(Morph):
    """options/morphs.py Cylinder: axis along z; collides through its mesh / support field like the reference's cylinders."""
    radius: float = 0.5
    height: float = 1.0
    fixed: bool = False


@dataclass
class Capsule(Morph):
    """A capsule along z: `length` is the cylindrical part, analytic capsule pairs apply (capsule_contact.py)."""
    radius: float = 0.5
    length: float = 1.0
    fixed: bool = False


@dataclass
class Terrain(Morph):
    """Height-field terrain (genesis/options/morphs.py:916+, the `height_field` form): `height_field[r, c]` in units of
    vertical_scale at x = r * horizontal_scale, y = c * horizontal_scale in the terrain's frame.  The procedural sub-terrain
    generators of the reference (utils/terrain.py) produce such an array and are outside the hot-path scope."""
    height_field: object = None
    horizontal_scale: float = 0.25
    vertical_scale: float = 0.005
    fixed: bool = True


@dataclass
class MJCF(Morph):
    file: str = ""
    scale: float = 1.0
    default_armature: float | None = 0.1
    convexify: bool | None = None
    decimate: bool = True
    decimate_face_num: int = 500


@dataclass
class URDF(Morph):
    file: str = ""
    scale: float = 1.0
    fixed: bool = False
    merge_fixed_links: bool = True
    links_to_keep: tuple = ()
    default_armature: float | None = 0.1
    convexify: bool | None = None
