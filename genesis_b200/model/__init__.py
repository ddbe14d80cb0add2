from .builder import ModelBuilder, RigidModel, plane_l_infos, box_l_infos, sphere_l_infos, terrain_l_infos, cylinder_l_infos, capsule_l_infos
from .mjcf import parse_mjcf
from .urdf import parse_urdf
