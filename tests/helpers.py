"""Shared helpers of the test-suite: host-only model construction and oracle wrappers."""

import os
import tempfile

import numpy as np

import genesis_b200 as gs
from genesis_b200.bench_scenes import make_franka, make_franka_grasp, make_go2
from oracle.pyoracle import Oracle

gs.init(backend=gs.gpu, seed=0)


def host_model(make, **kw):
    """Compile the model of a benchmark scene on the host (no CUDA needed)."""
    out = make(0, build=False, **kw)
    scene = out[0]
    solver = scene.rigid_solver
    model = solver.build_model()
    return scene, solver, model


def oracle_for(solver, model, n_envs=1, double=False, threads=1, **opt_overrides):
    opts = solver.model_options()
    opts.update(opt_overrides)
    return Oracle(model, n_envs, double=double, options=opts, threads=threads)


def scene_from_morphs(morphs, **rigid_kwargs):
    scene = gs.Scene(rigid_options=gs.options.RigidOptions(**rigid_kwargs))
    ents = [scene.add_entity(m) for m in morphs]
    return scene, ents


def write_tmp(text, suffix):
    fd, path = tempfile.mkstemp(suffix=suffix)
    with os.fdopen(fd, "w") as f:
        f.write(text)
    return path


PENDULUM_XML = """
<mujoco model="pendulum">
  <worldbody>
    <body name="arm" pos="0 0 1">
      <joint name="hinge" type="hinge" axis="0 1 0" pos="0 0 0"/>
      <inertial pos="0 0 -0.5" mass="1.0" diaginertia="0.01 0.01 0.01"/>
    </body>
  </worldbody>
</mujoco>
"""
