"""Enumerations of the rigid path; values are identical to the reference (genesis/constants.py:17-99,
genesis/utils/array_class.py:91-97) because they cross the C-ABI as plain ints."""

import enum


class IntEnum(enum.IntEnum):
    def __repr__(self):
        return f"<gs.{self.__class__.__name__}.{self.name}: {self.value}>"


class GEOM_TYPE(IntEnum):
    PLANE = 0
    SPHERE = 1
    ELLIPSOID = 2
    CYLINDER = 3
    CAPSULE = 4
    BOX = 5
    MESH = 6
    TERRAIN = 7


class JOINT_TYPE(IntEnum):
    FIXED = 0
    REVOLUTE = 1
    PRISMATIC = 2
    SPHERICAL = 3
    FREE = 4


class EQUALITY_TYPE(IntEnum):
    CONNECT = 0
    WELD = 1
    JOINT = 2


class CTRL_MODE(IntEnum):
    POSITION = 0
    VELOCITY = 1
    FORCE = 2


class integrator(IntEnum):
    Euler = 0
    implicitfast = 1
    approximate_implicitfast = 2


class constraint_solver(IntEnum):
    CG = 0
    Newton = 1


class backend(IntEnum):
    cpu = 0
    gpu = 1
    cuda = 2


class ErrorCode(IntEnum):
    SUCCESS = 0
    OVERFLOW_CANDIDATE_CONTACTS = 1
    OVERFLOW_COLLISION_PAIRS = 2
    OVERFLOW_HIBERNATION_ISLANDS = 4
    INVALID_FORCE_NAN = 8
    INVALID_ACC_NAN = 16
