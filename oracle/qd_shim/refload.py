"""TEST INFRASTRUCTURE ONLY -- imports the reference's own rigid-solver kernel sources (read in place from /root/reference,
nothing is copied) on top of the `quadrants` stand-in of this directory.

`import genesis` proper cannot run here (quadrants, mujoco, trimesh, pyglet ... are absent), and the package `__init__`s pull
in the whole engine.  So the package TREE is registered as empty namespace modules whose `__path__` points at the reference
directories; `import genesis.utils.geom` etc. then load the real files, while nothing outside the hot path is imported.  The
root module `genesis` carries only what `genesis/__init__.py: init()` would have set for `backend=cpu, precision="64"`.
"""

import importlib
import logging
import os
import sys
import types

REF = os.environ.get("GENESIS_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


class GenesisException(Exception):
    pass


def _raise(msg="Something went wrong."):
    raise GenesisException(msg)


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def load(precision="32"):
    """Returns the fake `genesis` root module after registering the package tree; idempotent."""
    if "genesis" in sys.modules and getattr(sys.modules["genesis"], "__qd_shim__", False):
        return sys.modules["genesis"]
    if not os.path.isdir(os.path.join(REF, "genesis")):
        raise RuntimeError(f"reference sources not found under {REF} (this tool only runs where /root/reference exists)")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import numpy as np
    import quadrants as qd
    import torch

    assert getattr(qd, "__version__", "").endswith("shim")
    G = os.path.join(REF, "genesis")
    gs = _pkg("genesis", G)
    gs.__qd_shim__ = True
    # third-party modules the kernel files import at module scope but never use on the executed path
    fd = types.ModuleType("frozendict")
    fd.frozendict = dict
    sys.modules.setdefault("frozendict", fd)

    # ---- what gs.init(backend=gs.cpu, precision="64") defines (genesis/__init__.py:150-260)
    consts = importlib.import_module("genesis.constants")
    for k, v in vars(consts).items():
        if not k.startswith("_"):
            setattr(gs, k, v)
    for b in consts.backend:
        setattr(gs, b.name, b)
    gs._initialized = True
    gs.backend = consts.backend.cpu
    gs.device = torch.device("cpu")
    gs.use_ndarray = False
    gs.use_fastcache = False
    gs.use_zerocopy = False
    # precision "32": the reference's default.  The stand-in's f32 does not round, so the run is "fp32 constants (EPS, CCD_EPS,
    # branch thresholds), fp64 arithmetic" -- exactly what the fp64 instantiation of oracle/rigid_oracle.hpp computes.
    gs.qd_float, gs.np_float, gs.tc_float = (qd.f64, np.float64, torch.float64) if precision == "64" else (qd.f32, np.float64, torch.float64)
    gs.qd_int, gs.np_int, gs.tc_int = qd.i32, np.int32, torch.int32
    gs.qd_bool, gs.np_bool, gs.tc_bool = qd.u1, np.bool_, torch.bool
    for n in (2, 3, 4, 6, 7, 11):
        setattr(gs, f"qd_vec{n}", qd.types.vector(n, gs.qd_float))
    gs.qd_mat3, gs.qd_mat4 = qd.types.matrix(3, 3, gs.qd_float), qd.types.matrix(4, 4, gs.qd_float)
    gs.qd_ivec2, gs.qd_ivec3, gs.qd_ivec4 = (qd.types.vector(n, gs.qd_int) for n in (2, 3, 4))
    gs.EPS = float(max(1e-15, np.finfo(np.float64 if precision == "64" else np.float32).eps))  # genesis/__init__.py:229
    gs.logger = logging.getLogger("genesis-shim")
    gs.raise_exception = _raise
    gs.GenesisException = GenesisException
    gs.raise_exception_from = lambda msg, e: _raise(msg)

    # ---- package tree (no __init__ executed)
    _pkg("genesis.utils", os.path.join(G, "utils"))
    _pkg("genesis.engine", os.path.join(G, "engine"))
    _pkg("genesis.engine.solvers", os.path.join(G, "engine", "solvers"))
    R = os.path.join(G, "engine", "solvers", "rigid")
    _pkg("genesis.engine.solvers.rigid", R)
    abd = _pkg("genesis.engine.solvers.rigid.abd", os.path.join(R, "abd"))
    _pkg("genesis.engine.solvers.rigid.collider", os.path.join(R, "collider"))
    _pkg("genesis.engine.solvers.rigid.constraint", os.path.join(R, "constraint"))

    # genesis.utils.misc imports pyglet / cpuinfo / quadrants internals at module scope; the kernel files only import names from it
    misc = types.ModuleType("genesis.utils.misc")
    for name in ("qd_to_torch", "qd_to_numpy", "indices_to_mask", "assign_indexed_tensor", "tensor_to_array"):
        setattr(misc, name, lambda *a, **k: _raise("genesis.utils.misc is stubbed in the shim"))
    sys.modules["genesis.utils.misc"] = misc
    gs.utils = sys.modules["genesis.utils"]
    gs.utils.misc = misc
    # genesis.typing needs pydantic validators only for options; geom.py imports one alias from it
    typing_stub = types.ModuleType("genesis.typing")
    typing_stub.Vec3FType = object
    sys.modules["genesis.typing"] = typing_stub

    importlib.import_module("genesis.utils.geom")
    importlib.import_module("genesis.utils.array_class")
    # abd/__init__.py star-imports its submodules: reproduce for the ones on the path
    for sub in ("misc", "forward_kinematics", "forward_dynamics", "diff", "accessor"):
        m = importlib.import_module(f"genesis.engine.solvers.rigid.abd.{sub}")
        for k, v in vars(m).items():
            if not k.startswith("_"):
                setattr(abd, k, v)
    # noslip.py and others reach `rigid_solver.func_solve_mass_batch` through the (huge) rigid_solver module: alias to abd
    rs = types.ModuleType("genesis.engine.solvers.rigid.rigid_solver")
    for k, v in vars(abd).items():
        if not k.startswith("_"):
            setattr(rs, k, v)
    sys.modules["genesis.engine.solvers.rigid.rigid_solver"] = rs
    return gs


def mod(name):
    """import a reference module by dotted name below `genesis.` after load()."""
    load()
    return importlib.import_module("genesis." + name)


def extract_defs(relpath, names, namespace, cls=None):
    """Executes selected top-level functions (or methods of class `cls`, as plain functions) of a reference file that cannot be
    imported as a module (its import block pulls in the whole engine), with their ORIGINAL file name and line numbers so that
    the shim's source-based decorators (and tracebacks) see the reference file.  Nothing is copied: the source is read in place."""
    import ast

    load()
    path = os.path.join(REF, "genesis", relpath)
    tree = ast.parse(open(path).read(), filename=path)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise KeyError(f"{relpath}: {sorted(missing)} not found")
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return {n.name: namespace[n.name] for n in picked}
