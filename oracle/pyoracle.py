"""TEST INFRASTRUCTURE ONLY: ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement of the reference).

Imported only by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).  The product
package genesis_b200 never imports this module.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_model_new.restype = ctypes.c_void_p
        L.orc_model_free.argtypes = [ctypes.c_void_p]
        L.orc_model_set_f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_model_set_i.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_model_set_s.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p]
        L.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p]
        L.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_step_timed.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_step_timed.restype = ctypes.c_double
        L.orc_forward_dynamics.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_detection.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


SCALARS = ("n_links", "n_joints", "n_dofs", "n_qs", "n_geoms", "n_entities", "n_equalities", "n_pairs", "n_verts",
           "max_contact_pairs", "max_broad_pairs", "n_contacts_per_pair", "len_constraints", "meaninertia", "dt")


class Oracle:
    """One oracle instance = one model + B independent environments, fp32 or fp64 arithmetic."""

    def __init__(self, model, n_envs=1, double=False, options=None, threads=1):
        L = lib()
        self.m = model
        self.B = n_envs
        raw = L.orc_model_new()
        for k, v in model.__dict__.items():
            if isinstance(v, np.ndarray):
                if v.dtype.kind == "f":
                    a = np.ascontiguousarray(v, dtype=np.float64)
                    L.orc_model_set_f(raw, k.encode(), a.ctypes.data, a.size)
                elif v.dtype.kind in "iub":
                    a = np.ascontiguousarray(v, dtype=np.int32)
                    L.orc_model_set_i(raw, k.encode(), a.ctypes.data, a.size)
        for k in SCALARS:
            L.orc_model_set_s(raw, k.encode(), float(getattr(model, k)))
        for i, ax in enumerate("xyz"):
            L.orc_model_set_s(raw, f"gravity_{ax}".encode(), float(model.gravity[i]))
        for k, v in (options or {}).items():
            L.orc_model_set_s(raw, f"cfg_{k}".encode(), float(v))
        self.h = L.orc_create(raw, n_envs, int(double))
        L.orc_model_free(raw)
        L.orc_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def get(self, name, shape, env=None):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        envs = range(self.B) if env is None else [env]
        out = np.zeros((len(envs),) + shape, dtype=np.float64)
        for i, b in enumerate(envs):
            if lib().orc_get(self.h, b, name.encode(), out[i].ctypes.data) != 0:
                raise KeyError(name)
        return out if env is None else out[0]

    def set(self, name, value, env=None):
        """value: (n,) broadcast to the selected env(s), or (B, n) one row per env when env is None."""
        envs = range(self.B) if env is None else [env]
        value = np.atleast_1d(np.asarray(value, dtype=np.float64))
        for i, b in enumerate(envs):
            v = np.ascontiguousarray(value[i] if (value.ndim >= 2 and env is None) else value)
            if lib().orc_set(self.h, b, name.encode(), v.ctypes.data) != 0:
                raise KeyError(name)

    def step(self, n=1):
        lib().orc_step(self.h, n)

    def step_timed(self, n=1):
        return lib().orc_step_timed(self.h, n)

    def forward_dynamics(self, env=0):
        lib().orc_forward_dynamics(self.h, env)

    def detection(self, env=0):
        lib().orc_detection(self.h, env)

    # convenience accessors
    def qpos(self):
        return self.get("qpos", self.m.n_qs)

    def vel(self):
        return self.get("vel", self.m.n_dofs)

    def contacts(self, env=0):
        n = int(self.get("n_contacts", 1, env)[0])
        raw = self.get("contacts", (max(n, 1), 16), env)[:n]
        return dict(geom_a=raw[:, 0].astype(int), geom_b=raw[:, 1].astype(int), link_a=raw[:, 2].astype(int),
                    link_b=raw[:, 3].astype(int), penetration=raw[:, 4], position=raw[:, 5:8], normal=raw[:, 8:11],
                    force=raw[:, 11:14], friction=raw[:, 14])
